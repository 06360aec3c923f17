"""Controllers on the batched hot path — same names as ``abr_control.controllers``."""
from .avoid_obstacles import AvoidObstacles
from .controller import Controller
from .damping import Damping
from .floating import Floating
from .joint import Joint
from .osc import OSC
from .resting_config import RestingConfig

__all__ = ["Controller", "OSC", "Damping", "RestingConfig", "AvoidObstacles", "Joint", "Floating"]
