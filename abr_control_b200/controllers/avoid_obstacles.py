"""Obstacle avoidance in the null space (Khatib 1987).

Reference: /root/reference/abr_control/controllers/avoid_obstacles.py:6-133 — for every obstacle
``[x, y, z, radius]`` and every arm segment: closest point, repulsive potential force, offset Jacobian of that
point and a 3x3 ``pinv(rcond=0.01)`` task-space inertia; output clipped to ``+-maximum``.
"""
import numpy as np

from .. import _abi
from ._null import NullController


class AvoidObstacles(NullController):
    def __init__(self, robot_config, obstacles=None, threshold=0.2, gain=1, maximum=500):
        super().__init__(robot_config)
        self.threshold = threshold
        self.gain = gain
        self.maximum = maximum
        self.obstacles = np.array([] if obstacles is None else obstacles, dtype=float).reshape(-1, 4)

    def set_obstacles(self, obstacles):
        """avoid_obstacles.py:122-133"""
        self.obstacles = np.array(obstacles, dtype=float).reshape(-1, 4)
        self._dirty()

    def _params(self):
        return _abi.null_params("AvoidObstacles", self.robot_config.N_JOINTS, obstacles=self.obstacles,
                                threshold=self.threshold, gain=self.gain, maximum=self.maximum)

    def generate(self, q, dq=None):
        if dq is None:  # the reference ignores dq here (avoid_obstacles.py:38-49)
            dq = q * 0 if hasattr(q, "shape") else np.zeros_like(np.asarray(q, dtype=float))
        return super().generate(q, dq)
