"""Controllers on the batched hot path — same names as ``abr_control.controllers``."""
from .avoid_joint_limits import AvoidJointLimits
from .avoid_obstacles import AvoidObstacles
from .controller import Controller
from .damping import Damping
from .floating import Floating
from .joint import Joint
from .osc import OSC
from .resting_config import RestingConfig
from .sliding import Sliding

__all__ = ["Controller", "OSC", "Damping", "RestingConfig", "AvoidObstacles", "Joint", "Floating",
           "AvoidJointLimits", "Sliding"]
