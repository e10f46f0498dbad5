from . import drop, mlp  # noqa: F401
from .drop import DropPath  # noqa: F401
from .mlp import Mlp  # noqa: F401
