"""twojoint arm model (reference: /root/reference/abr_control/arms/twojoint/config.py:30-181).

The kinematic/inertial constants live in data/twojoint.json, recovered from the reference's SymPy
transforms by tools/extract_chain.py.  Use as the reference: ``from abr_control_b200.arms import twojoint;
robot_config = twojoint.Config()``.
"""
from .. import _abi
from .base_config import BaseConfig


class Config(BaseConfig):
    def __init__(self, **kwargs):
        super().__init__(_abi.load_arm_json("twojoint"), ROBOT_NAME="twojoint", **kwargs)
