"""Iterative inverse-kinematics path planner, one path per state of a batch.

Reference: /root/reference/abr_control/controllers/path_planners/inverse_kinematics.py:7-180 — ``generate_path``
takes resolved-motion steps from ``position`` towards a task-space target ``(x, y, z, alpha, beta, gamma)`` at the
end-effector (methods 1: pinv(J), 2: damped least squares, 3: position first with the orientation in its null space),
clipping the task-space step to ``max_dx dt`` / ``max_dr dt`` and the joint step to ``max_dq dt``.  The whole
iteration runs in one kernel launch with the joint state in registers.  ``plot`` is accepted and ignored.
"""
import numpy as np

from ... import _lib
from .. import _batch


class InverseKinematics:
    def __init__(self, robot_config, max_dx=0.2, max_dr=2 * np.pi, max_dq=np.pi):
        self.robot_config = robot_config
        self.max_dx = max_dx
        self.max_dr = max_dr
        self.max_dq = max_dq

    def generate_path(self, position, target_position, n_timesteps=200, dt=0.001, plot=False, method=3, axes="rxyz"):
        """``position`` (n,) and ``target_position`` (6,) -> ``(position_path, velocity_path)`` of shape
        (n_timesteps, n) as the reference; ``position`` (B, n) with ``target_position`` (B, 6) or (6,) -> shape
        (B, n_timesteps, n) (NumPy in -> NumPy out, CUDA tensors in -> CUDA tensors out, views of one (T, B, n) buffer).
        ``axes`` is accepted for signature parity: the reference always reads the target angles as "sxyz"
        (inverse_kinematics.py:72-81)."""
        import torch

        if method not in (1, 2, 3):
            raise ValueError("method must be 1, 2 or 3")
        rc = self.robot_config
        n = rc.N_JOINTS
        qa, single, kind = rc._prep(position, np.float64 if (np.ndim(position) == 1 and not _batch.is_torch(position)) else None)
        tgt, tstride = _batch.prep_rows(target_position, qa, kind, 6, "target_position")
        B = qa.shape[0]
        arrays = [qa, tgt]
        if kind == "numpy":
            dev = torch.device("cuda", torch.cuda.current_device())
            arrays = [torch.as_tensor(a).to(dev) for a in arrays]
        tq = arrays[0]
        f32 = tq.dtype == torch.float32
        pos = torch.empty((n_timesteps, B, n), dtype=tq.dtype, device=tq.device)
        vel = torch.empty_like(pos)
        L = _lib.lib()
        fn = L.abrb_ik_path_f32 if f32 else L.abrb_ik_path_f64
        with torch.cuda.device(tq.device):
            _lib.check(fn(rc.handle, float(self.max_dx), float(self.max_dr), float(self.max_dq), int(method), float(dt),
                          int(n_timesteps), tq.data_ptr(), arrays[1].data_ptr(), tstride, pos.data_ptr(), vel.data_ptr(),
                          B, torch.cuda.current_stream(tq.device).cuda_stream))
        pos, vel = pos.permute(1, 0, 2), vel.permute(1, 0, 2)
        if kind == "numpy":
            pos, vel = pos.cpu().numpy(), vel.cpu().numpy()
            if single:
                pos, vel = np.array(pos[0], dtype=np.float64), np.array(vel[0], dtype=np.float64)
        elif single:
            pos, vel = pos[0], vel[0]
        # the reference's iterator state (inverse_kinematics.py:160-166)
        self.n_timesteps = n_timesteps
        self.n = 0
        self.position_path = pos
        self.velocity_path = vel
        return self.position_path, self.velocity_path

    def next(self):
        """Next point along a single generated path (inverse_kinematics.py:168-180, including its quirk of
        returning the position path for the velocity as well)."""
        self.position = self.position_path[self.n] if self.n < self.n_timesteps else self.target
        self.velocity = self.position_path[self.n] if self.n < self.n_timesteps else self.velocity
        self.n = min(self.n + 1, self.n_timesteps)
        return self.position, self.velocity
