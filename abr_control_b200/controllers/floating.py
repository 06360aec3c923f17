"""Gravity compensation ("floating") controller.

Reference: /root/reference/abr_control/controllers/floating.py:4-71 — joint space ``u = -g`` or task space
``u = J^T (-(M^-1 J^T Mx)^T g)`` with ``J = J("EE")[:3]`` and the reference's ``inv`` / ``pinv(rcond=1e-4)`` switch at
``|det| > 1e-3``; ``dynamic=True`` additionally subtracts ``M dq``.
"""
import numpy as np

from .. import _lib
from . import _batch
from .controller import Controller
from .joint import _device_call


class Floating(Controller):
    def __init__(self, robot_config, dynamic=False, task_space=False):
        super().__init__(robot_config)
        self.dynamic = dynamic
        self.task_space = task_space

    def generate(self, q, dq=None):
        rc = self.robot_config
        if dq is None:
            if self.dynamic:
                raise TypeError("dynamic=True needs dq")
            dq = q * 0 if hasattr(q, "shape") else np.zeros_like(np.asarray(q, dtype=float))
        qa, dqa, single, kind, _ = _batch.prep_state(rc, q, dq)
        L = _lib.lib()
        B = qa.shape[0]

        def args(a, u):
            return (rc.handle, int(bool(self.task_space)), int(bool(self.dynamic)), a[0].data_ptr(), a[1].data_ptr(),
                    u.data_ptr(), B)

        return _device_call(rc, L.abrb_floating_generate_f32, L.abrb_floating_generate_f64, args, [qa, dqa], single)
