"""jaco2 arm model (reference: /root/reference/abr_control/arms/jaco2/config.py:37-356).

The kinematic/inertial constants live in data/jaco2.json, recovered from the reference's SymPy
transforms by tools/extract_chain.py.  Use as the reference: ``from abr_control_b200.arms import jaco2;
robot_config = jaco2.Config()``.
"""
from .. import _abi
from .base_config import BaseConfig


class Config(BaseConfig):
    def __init__(self, **kwargs):
        super().__init__(_abi.load_arm_json("jaco2"), ROBOT_NAME="jaco2", **kwargs)
