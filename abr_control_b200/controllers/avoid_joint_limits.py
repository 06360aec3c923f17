"""Joint-limit avoidance.

Reference: /root/reference/abr_control/controllers/avoid_joint_limits.py:6-142 — a torque that pushes a joint back
once it has passed a limit (``max_torque``), optionally with an exponential approach (``gradient``); ``cross_zero``
marks joints whose working range contains the 0 / 2 pi seam.  A function of ``q`` alone.  ``None`` (or NaN) = no limit
on that side.  Standalone batched ``generate(q, dq)`` or a member of ``OSC(null_controllers=[...])``.
"""
import numpy as np

from .. import _abi
from ._null import NullController


class AvoidJointLimits(NullController):
    def __init__(self, robot_config, min_joint_angles, max_joint_angles, max_torque=None, cross_zero=None,
                 gradient=None):
        super().__init__(robot_config)
        n = robot_config.N_JOINTS
        if len(min_joint_angles) != n or len(max_joint_angles) != n:
            raise Exception("joint angles vector incorrect size")
        self._min_in = [None if (v is None or v != v) else float(v) for v in min_joint_angles]
        self._max_in = [None if (v is None or v != v) else float(v) for v in max_joint_angles]
        self.cross_zero = np.array([False] * n if cross_zero is None else cross_zero)
        self.gradient = np.array([False] * n if gradient is None else gradient)
        self.max_torque = np.ones(n) if max_torque is None else np.asarray(max_torque, dtype=float)
        p = self._params()
        # the reference's attributes after its constructor (shifted by -pi, swapped where cross_zero)
        self.min_joint_angles = np.array([p.limit_min[k] for k in range(n)])
        self.max_joint_angles = np.array([p.limit_max[k] for k in range(n)])
        self.no_limits_min = np.isnan(self.min_joint_angles)
        self.no_limits_max = np.isnan(self.max_joint_angles)

    def _params(self):
        return _abi.null_params("AvoidJointLimits", self.robot_config.N_JOINTS, min_joint_angles=self._min_in,
                                max_joint_angles=self._max_in, max_torque=list(self.max_torque),
                                cross_zero=list(self.cross_zero), gradient=list(self.gradient))
