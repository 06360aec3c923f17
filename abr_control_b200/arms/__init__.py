"""Arm models: same import surface as ``abr_control.arms`` for the arms on the batched hot path."""
from . import jaco2, threejoint, twojoint, ur5
from .base_config import BaseConfig, builtin_config

__all__ = ["BaseConfig", "builtin_config", "ur5", "jaco2", "threejoint", "twojoint"]
