"""Importing this package registers every DiST component, as ``models/base/__init__.py`` does in the reference."""
from . import base_blocks, clip, backbone, models, builder  # noqa: F401
from ..module_zoo import branches, stems  # noqa: F401
