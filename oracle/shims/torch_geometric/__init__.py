from . import typing, data, utils, nn, loader  # noqa: F401
