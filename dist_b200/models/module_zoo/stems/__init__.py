from . import dist_stem  # noqa: F401
