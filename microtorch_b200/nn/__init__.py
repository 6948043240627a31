"""Modules, losses and the container (same names as the reference's `src.nn`); activations, the optimizer and
`Parameter` live in their own modules like there (`nn.activation`, `nn.optimizer`, `nn.param`)."""

from .module import Module
from .sequential import Sequential
from .layer import Linear
from .loss import BCELoss, CrossEntropyLoss, L1Loss, MSELoss

__all__ = ("Module", "Sequential", "Linear", "L1Loss", "MSELoss", "BCELoss", "CrossEntropyLoss")
