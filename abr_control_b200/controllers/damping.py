"""Null-space damping ``M(q) (-kv dq)`` (reference: /root/reference/abr_control/controllers/damping.py:4-32)."""
from .. import _abi
from ._null import NullController


class Damping(NullController):
    def __init__(self, robot_config, kv):
        super().__init__(robot_config)
        self.kv = kv

    def _params(self):
        return _abi.null_params("Damping", self.robot_config.N_JOINTS, kv=self.kv)
