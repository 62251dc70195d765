from . import kernels, util  # noqa: F401
from .parameterized import Parameterized  # noqa: F401
