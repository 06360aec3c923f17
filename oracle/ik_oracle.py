"""TEST INFRASTRUCTURE — CPU restatement of the reference's iterative inverse-kinematics path planner.

Only tests/ may import this module (see oracle/__init__.py).  It follows
/root/reference/abr_control/controllers/path_planners/inverse_kinematics.py:28-166 line by line with the same NumPy
calls (``numpy.linalg.pinv`` / ``solve``), on top of the RobotOracle duck type of oracle/osc_oracle.py; the plotting
branch (:139-158) is not part of the computation and is left out.
"""
import numpy as np

from . import rbd_oracle as ro


def quaternion_from_euler_sxyz(ai, aj, ak):
    """utils/transformations.py:1096-1147 with axes='sxyz' -> axes tuple (0, 0, 0, 0): i, j, k = 1, 2, 3"""
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si, cj, sj, ck, sk = np.cos(ai), np.sin(ai), np.cos(aj), np.sin(aj), np.cos(ak), np.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc])


class InverseKinematics:
    def __init__(self, rc, max_dx=0.2, max_dr=2 * np.pi, max_dq=np.pi):  # :22-26
        self.rc, self.max_dx, self.max_dr, self.max_dq = rc, max_dx, max_dr, max_dq

    def generate_path(self, position, target_position, n_timesteps=200, dt=0.001, method=3):
        n = position.shape[0]
        path = np.zeros((n_timesteps, n * 2))
        max_dq, max_dx, max_dr = self.max_dq * dt, self.max_dx * dt, self.max_dr * dt  # :66-70
        Qd = ro.unit_vector(quaternion_from_euler_sxyz(target_position[3], target_position[4], target_position[5]))
        q = np.copy(position)
        for ii in range(n_timesteps):
            J = self.rc.J("EE", q)
            Tx = self.rc.Tx("EE", q)
            dx = target_position[:3] - Tx
            Qe = self.rc.quaternion("EE", q)
            dr = Qe[0] * Qd[1:] - Qd[0] * Qe[1:] - np.cross(Qd[1:], Qe[1:])  # :93
            norm_dx, norm_dr = np.linalg.norm(dx, 2), np.linalg.norm(dr, 2)
            if norm_dx > max_dx:
                dx = dx / norm_dx * max_dx
            if norm_dr > max_dr:
                dr = dr / norm_dr * max_dr
            Jx = J[:3]
            pinv_Jx = np.linalg.pinv(Jx)
            if method == 1:
                dq = np.dot(np.linalg.pinv(J), np.hstack([dx, dr]))
            if method == 2:
                dq = np.dot(J.T, np.linalg.solve(np.dot(J, J.T) + np.eye(6) * 0.001, np.hstack([dx, dr * 0.3])))
            if method == 3:
                dq = np.dot(pinv_Jx, dx) + np.dot(np.eye(n) - np.dot(pinv_Jx, Jx), np.dot(np.linalg.pinv(J[3:]), dr))
            if max(abs(dq)) > max_dq:
                dq = dq / max(abs(dq)) * max_dq
            path[ii] = np.hstack([q, dq])
            q = q + dq
        return path[:, :n], path[:, n:]


def run_ik_case(case, positions, targets, mode="fp64"):
    """tests/cases.py::IK_CASES -> (position_path, velocity_path), each (B, n_timesteps, n)"""
    from .osc_oracle import RobotOracle

    rc = RobotOracle(case["arm"], mode)
    ik = InverseKinematics(rc, **case.get("init", {}))
    pos, vel = [], []
    for i in range(len(positions)):
        p, v = ik.generate_path(positions[i], targets[i], **case["path"])
        pos.append(p)
        vel.append(v)
    return np.array(pos), np.array(vel)
