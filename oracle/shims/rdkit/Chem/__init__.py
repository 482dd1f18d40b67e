from . import AllChem  # noqa: F401
