"""microtorch_b200 - B200-native compute backend for MicroTorch's device="cuda" path."""

__version__ = "0.1.0"
