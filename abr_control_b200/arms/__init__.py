"""Arm models: same import surface as ``abr_control.arms`` for the arms on the batched hot path."""
from . import jaco2, threejoint, twojoint, ur5
from .base_config import BaseConfig, builtin_config
from .mjcf import MjcfConfig, chain_desc_from_mjcf

__all__ = ["BaseConfig", "builtin_config", "MjcfConfig", "chain_desc_from_mjcf", "ur5", "jaco2", "threejoint", "twojoint"]
