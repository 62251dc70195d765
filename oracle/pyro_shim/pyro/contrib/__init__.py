from . import gp  # noqa: F401
