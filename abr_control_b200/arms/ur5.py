"""ur5 arm model (reference: /root/reference/abr_control/arms/ur5/config.py:35-339).

The kinematic/inertial constants live in data/ur5.json, recovered from the reference's SymPy
transforms by tools/extract_chain.py.  Use as the reference: ``from abr_control_b200.arms import ur5;
robot_config = ur5.Config()``.
"""
from .. import _abi
from .base_config import BaseConfig


class Config(BaseConfig):
    def __init__(self, **kwargs):
        super().__init__(_abi.load_arm_json("ur5"), ROBOT_NAME="ur5", **kwargs)
