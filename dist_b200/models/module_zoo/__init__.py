from . import branches, stems  # noqa: F401
