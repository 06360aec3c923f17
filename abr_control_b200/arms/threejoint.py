"""threejoint arm model (reference: /root/reference/abr_control/arms/threejoint/config.py:30-223).

The kinematic/inertial constants live in data/threejoint.json, recovered from the reference's SymPy
transforms by tools/extract_chain.py.  Use as the reference: ``from abr_control_b200.arms import threejoint;
robot_config = threejoint.Config()``.
"""
from .. import _abi
from .base_config import BaseConfig


class Config(BaseConfig):
    def __init__(self, **kwargs):
        super().__init__(_abi.load_arm_json("threejoint"), ROBOT_NAME="threejoint", **kwargs)
