from .base import BaseOps
from .elementwise import Elementwise
from .math import MathOps
from .overload import OverloadOps
from .reduce import Reduce

__all__ = ["BaseOps", "Elementwise", "MathOps", "OverloadOps", "Reduce"]
