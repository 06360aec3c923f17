"""Path planners on the batched hot path — same names as ``abr_control.controllers.path_planners``."""
from .inverse_kinematics import InverseKinematics

__all__ = ["InverseKinematics"]
