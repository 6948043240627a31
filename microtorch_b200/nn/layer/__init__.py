from .linear import Linear

__all__ = ["Linear"]
