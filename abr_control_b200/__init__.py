"""abr_control_b200 — batched operational-space control for robot arms on NVIDIA B200.

Drop-in for the hot path of abr_control: ``arms.<arm>.Config`` (the robot_config duck type) and
``controllers.{OSC, Damping, RestingConfig, AvoidObstacles}``, evaluated over batches of joint states by
hand-written sm_100a kernels behind the C ABI in include/abrb.h.
"""
from . import arms, controllers
from ._lib import AbrbError

__version__ = "0.1.0"
__all__ = ["arms", "controllers", "AbrbError"]
