"""Layers with parameters."""

from .linear import Linear

__all__ = ("Linear",)
