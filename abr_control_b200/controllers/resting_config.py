"""Null-space PD towards resting joint angles.

Reference: /root/reference/abr_control/controllers/resting_config.py:8-42 on top of
controllers/joint.py:104-131 with ``account_for_gravity=False``:
``u = M (kp q_tilde - kv dq)``, ``q_tilde = ((rest - q + pi) mod 2 pi) - pi`` on the joints whose rest angle is
not ``None`` and 0 elsewhere.
"""
import numpy as np

from .. import _abi
from ._null import NullController


class RestingConfig(NullController):
    def __init__(self, robot_config, rest_angles, kp=1, kv=None, **kwargs):
        if kwargs.get("quaternions") is not None:
            raise NotImplementedError("ball-joint (quaternion) states are a MuJoCo-only branch of the reference")
        super().__init__(robot_config)
        self.kp = kp
        self.kv = np.sqrt(self.kp) if kv is None else kv  # joint.py:33
        self.rest_angles = np.asarray(rest_angles)
        self.rest_indices = [val is not None for val in rest_angles]

    def _params(self):
        return _abi.null_params("RestingConfig", self.robot_config.N_JOINTS, kp=self.kp, kv=self.kv,
                                rest_angles=list(self.rest_angles))
