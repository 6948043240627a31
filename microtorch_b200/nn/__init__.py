from .layer import Linear
from .loss import L1Loss, MSELoss, BCELoss
from .module import Module
from .sequential import Sequential

__all__ = ["Linear", "L1Loss", "MSELoss", "BCELoss", "Module", "Sequential"]
