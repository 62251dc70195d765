from . import indexing  # noqa: F401
