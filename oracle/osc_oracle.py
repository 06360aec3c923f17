"""TEST INFRASTRUCTURE — CPU oracle for ``OSC.generate`` and the secondary controllers (NumPy).

Restates, step for step, the reference's controller arithmetic on top of ``rbd_oracle.ChainOracle``:

* ``OSC.__init__`` gains           /root/reference/abr_control/controllers/osc.py:53-118
* ``OSC._Mx``                      osc.py:120-147   (inv / det / pinv(rcond=1e-4), same NumPy calls)
* ``_calc_orientation_forces``     osc.py:149-196
* ``_velocity_limiting``           osc.py:198-215
* ``OSC.generate``                 osc.py:217-320
* ``Damping.generate``             controllers/damping.py:21-32
* ``RestingConfig`` / ``Joint``    controllers/resting_config.py:25-42, controllers/joint.py:104-131
* ``AvoidObstacles.generate``      controllers/avoid_obstacles.py:38-120

``mode="fp64"`` evaluates everything in float64 (the reference with its float32 casts neutralised —
what ``tests/golden/*__u64`` holds); ``mode="ref32"`` rounds J, M, g, C, R to float32 exactly where the
reference's public API does (base_config.py:223,247,270,285,301,336), which reproduces the reference
as shipped (``*__u32``) including its float32 LAPACK calls.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import this.
"""
import numpy as np

from . import rbd_oracle as ro


class RobotOracle:
    """robot_config duck type (one state per call) backed by ChainOracle; applies the float32 casts."""

    def __init__(self, arm, mode="fp64"):
        self.chain = ro.ChainOracle(arm)
        self.N_JOINTS = self.chain.n
        self.N_LINKS = self.chain.n_links
        self.mode = mode

    def _cast(self, a):
        return a.astype(np.float32) if self.mode == "ref32" else a

    def J(self, name, q, x=None):
        return self._cast(self.chain.J(name, q, x)[0])

    def dJ(self, name, q, dq, x=None):
        return self._cast(self.chain.dJ(name, q, dq, x)[0])

    def M(self, q):
        return self._cast(self.chain.M(q)[0])

    def g(self, q):
        return self._cast(self.chain.g(q)[0])

    def C(self, q, dq):
        return self._cast(self.chain.C(q, dq)[0])

    def R(self, name, q):
        return self._cast(self.chain.R(name, q)[0])

    def Tx(self, name, q, x=None):
        return self.chain.Tx(name, q, x)[0]

    def T_inv(self, name, q, x=None):
        return self.chain.T_inv(name, q)[0]

    def quaternion(self, name, q):
        return ro.unit_vector(ro.quaternion_from_matrix(self.R(name, q)))


class Damping:
    def __init__(self, rc, kv):
        self.rc, self.kv = rc, kv

    def generate(self, q, dq):
        return np.dot(self.rc.M(q), -self.kv * dq)


class RestingConfig:
    def __init__(self, rc, rest_angles, kp=1, kv=None):
        self.rc = rc
        self.kp = kp
        self.kv = np.sqrt(kp) if kv is None else kv
        self.idx = [v is not None for v in rest_angles]
        self.rest = np.array([0.0 if v is None else v for v in rest_angles])

    def generate(self, q, dq):
        q_tilde = np.zeros(len(q))
        q_tilde[self.idx] = (self.rest[self.idx] - q[self.idx] + np.pi) % (np.pi * 2) - np.pi
        return np.dot(self.rc.M(q), self.kp * q_tilde + self.kv * (0.0 - dq))


class Joint:
    """controllers/joint.py:104-131 (angle joints only)"""

    def __init__(self, rc, kp=1, kv=None, account_for_gravity=True):
        self.rc, self.kp, self.kv, self.grav = rc, kp, (np.sqrt(kp) if kv is None else kv), account_for_gravity

    def generate(self, q, dq, target, target_velocity=None):
        tvel = np.zeros(len(q)) if target_velocity is None else target_velocity
        q_tilde = ((target - q + np.pi) % (np.pi * 2)) - np.pi
        u = np.dot(self.rc.M(q), self.kp * q_tilde + self.kv * (tvel - dq))
        return u - self.rc.g(q) if self.grav else u


class Floating:
    """controllers/floating.py:27-71"""

    def __init__(self, rc, dynamic=False, task_space=False):
        self.rc, self.dynamic, self.task_space = rc, dynamic, task_space

    def generate(self, q, dq=None):
        g = self.rc.g(q)
        M = None
        if self.task_space:
            J = self.rc.J("EE", q)[:3]
            M = self.rc.M(q)
            M_inv = np.linalg.inv(M)
            Mx_inv = np.dot(J, np.dot(M_inv, J.T))
            Mx = np.linalg.inv(Mx_inv) if abs(np.linalg.det(Mx_inv)) > 1e-3 else np.linalg.pinv(Mx_inv, rcond=1e-4)
            Jbar = np.dot(M_inv, np.dot(J.T, Mx))
            u = np.dot(J.T, -1 * np.dot(Jbar.T, g))
        else:
            u = -g
        if self.dynamic:
            M = self.rc.M(q) if M is None else M
            u = u - np.dot(M, dq)
        return u


class AvoidJointLimits:
    """controllers/avoid_joint_limits.py:36-142.  NaN marks a joint without a limit (the reference's `None`
    placeholder does not survive its own `np.isnan`, avoid_joint_limits.py:74-75)."""

    def __init__(self, rc, min_joint_angles, max_joint_angles, max_torque=None, cross_zero=None, gradient=None):
        n = rc.N_JOINTS
        lo = np.array([np.nan if v is None else v - np.pi for v in min_joint_angles], dtype=float)  # :46-51
        hi = np.array([np.nan if v is None else v - np.pi for v in max_joint_angles], dtype=float)
        self.cross = np.array([False] * n if cross_zero is None else cross_zero)
        self.grad = np.array([False] * n if gradient is None else gradient)
        self.lo, self.hi = lo.copy(), hi.copy()
        self.hi[self.cross] = lo[self.cross]  # :62-66 (flipped so that the maths matches the normal case)
        self.lo[self.cross] = hi[self.cross]
        self.no_lo, self.no_hi = np.isnan(self.lo), np.isnan(self.hi)
        self.tmax = np.ones(n) if max_torque is None else np.asarray(max_torque, dtype=float)
        self.n = n

    def generate(self, q, dq=None):
        q = np.asarray(q, dtype=float) - np.pi  # :91
        with np.errstate(all="ignore"):
            nearer_hi = abs(q - self.lo) >= abs(q - self.hi)  # the reference's `closer_to_min_index`, :94-96
            nearer_lo = abs(q - self.lo) <= abs(q - self.hi)  # the reference's `closer_to_max_index`, :97-99
            a_lo, a_hi = np.zeros(self.n), np.zeros(self.n)
            gr = self.grad
            a_lo[gr] = np.minimum(np.exp(1.0 / (q[gr] - self.lo[gr])), self.tmax[gr])  # :108-111
            a_hi[gr] = -np.minimum(np.exp(-1.0 / (q[gr] - self.hi[gr])), self.tmax[gr])  # :112-115
            below = (q - self.lo) < 0  # :118-119
            above = (q - self.hi) > 0
            cz = self.cross
            below[cz] = below[cz] * ((q[cz] - self.hi[cz]) > 0) * nearer_lo[cz]  # :124-128
            above[cz] = above[cz] * ((q[cz] - self.lo[cz]) < 0) * nearer_hi[cz]  # :130-134
        a_lo[below] = self.tmax[below]
        a_lo[self.no_lo] = 0.0
        a_hi[above] = -self.tmax[above]
        a_hi[self.no_hi] = 0.0
        return a_lo + a_hi


class Sliding:
    """controllers/sliding.py:27-99 (Slotine & Li sliding control); `s` is the reference's training signal."""

    def __init__(self, rc, kd=160.0, lamb=30.0, cartesian=True):
        self.rc, self.kd, self.lamb, self.cartesian = rc, kd, lamb, cartesian
        self.s = None

    def generate(self, q, dq, target, target_velocity=0, target_acc=0, ref_frame="EE", offset=None):
        rc = self.rc
        offset = np.zeros(3) if offset is None else offset
        if self.cartesian:
            J = rc.J(ref_frame, q, x=offset)[:3]
            xyz = rc.Tx(ref_frame, q, x=offset)
            dxyz = np.dot(J, dq)
            J_inv = np.linalg.pinv(J)
            dJ = rc.dJ(ref_frame, q, dq, x=offset)[:3]
            dq_ref = np.dot(J_inv, target_velocity + self.lamb * (target - xyz))
            ddq_ref = np.dot(J_inv, target_acc + self.lamb * (target_velocity - dxyz) - np.dot(dJ, dq_ref))
        else:
            dq_ref = target_velocity - self.lamb * (q - target)
            ddq_ref = target_acc - self.lamb * (dq - target_velocity)
        self.s = dq - dq_ref
        return np.dot(rc.M(q), ddq_ref) + np.dot(rc.C(q, dq), dq_ref) + rc.g(q) - self.kd * self.s


def _segment_closest(p_a, p_b, centre):
    """closest point of segment [p_a, p_b] to ``centre`` (avoid_obstacles.py:69-83)."""
    seg = p_b - p_a
    s = np.dot(centre - p_a, seg) / np.sum(seg ** 2)
    if s < 0:
        return p_a
    if s > 1:
        return p_b
    return p_a + s * seg


class AvoidObstacles:
    """Khatib repulsion from spherical obstacles, one term per (obstacle, arm segment) pair."""

    ETA = 0.02  # avoid_obstacles.py:92

    def __init__(self, rc, obstacles=None, threshold=0.2, gain=1, maximum=500):
        self.rc, self.threshold, self.gain, self.maximum = rc, threshold, gain, maximum
        self.obstacles = np.array([] if obstacles is None else obstacles, dtype=float)

    def _pair_torque(self, q, M, seg, centre, radius):
        rc, n, thr = self.rc, self.rc.N_JOINTS, self.threshold
        p_a = rc.Tx(f"joint{seg}", q)
        p_b = rc.Tx("EE", q) if seg == n - 1 else rc.Tx(f"joint{seg + 1}", q)
        near = _segment_closest(p_a, p_b, centre)
        rho = max(np.sqrt(np.sum((centre - near) ** 2)) - radius, thr / 50)  # :86-89
        if not rho < thr:
            return 0.0
        force = self.ETA * (1.0 / rho - 1.0 / thr) * 1.0 / rho ** 1.5 * ((centre - near) / rho)  # :93-101
        local = np.dot(rc.T_inv(f"link{seg + 1}", q), np.append(near, 1.0))[:-1]  # :106-107
        Jp = rc.J(f"link{seg + 1}", q, x=local)[:3]  # :109
        Mx_pt = np.linalg.pinv(np.dot(Jp, np.dot(np.linalg.inv(M), Jp.T)), rcond=0.01)  # :113-116
        return -1 * np.dot(Jp.T, np.dot(Mx_pt, force))  # :118

    def generate(self, q, dq=None):
        M = self.rc.M(q)
        total = np.zeros(self.rc.N_JOINTS)
        for ob in self.obstacles:
            for seg in range(self.rc.N_JOINTS):
                total = total + self._pair_torque(q, M, seg, np.array(ob[:3]), ob[3])
        return np.clip(total * self.gain, -self.maximum, self.maximum)  # :120


NULL_KINDS = {"Damping": Damping, "RestingConfig": RestingConfig, "AvoidObstacles": AvoidObstacles,
              "AvoidJointLimits": AvoidJointLimits}


class OSC:
    """Operational-space controller, one state per call, written as a pipeline of small steps."""

    def __init__(self, rc, kp=1, ko=None, kv=None, ki=0, vmax=None, ctrlr_dof=None,
                 null_controllers=None, use_g=True, use_C=False, orientation_algorithm=0):
        self.rc = rc
        self.kp = kp
        self.ko = kp if ko is None else ko  # osc.py:71
        self.kv = np.sqrt(self.kp + self.ko) if kv is None else kv  # osc.py:74
        self.ki = ki
        self.nulls = null_controllers or []
        self.use_g, self.use_C, self.alg = use_g, use_C, orientation_algorithm
        self.mask = np.array([1, 1, 1, 0, 0, 0] if ctrlr_dof is None else ctrlr_dof, dtype=bool)
        self.gains = np.array([self.kp] * 3 + [self.ko] * 3)  # osc.py:89
        self.vmax = vmax
        if vmax is not None:  # osc.py:109-115 (sat_gain_* == scale_*)
            self.lim_xyz = vmax[0] / self.kp * self.kv
            self.lim_abg = vmax[1] / self.ko * self.kv
        self.err_sum = np.zeros(6)

    @staticmethod
    def _Mx(M, J, threshold=1e-3):  # osc.py:120-147, same NumPy calls in the same order
        M_inv = np.linalg.inv(M)
        Mx_inv = np.dot(J, np.dot(M_inv, J.T))
        well_posed = abs(np.linalg.det(Mx_inv)) >= threshold
        Mx = np.linalg.inv(Mx_inv) if well_posed else np.linalg.pinv(Mx_inv, rcond=threshold * 0.1)
        return Mx, M_inv

    def orientation_error(self, abg, q, frame):  # osc.py:149-196
        if self.alg == 0:
            want = ro.unit_vector(ro.quaternion_from_euler_rxyz(abg[0], abg[1], abg[2]))
            have = self.rc.quaternion(frame, q)
            rel = ro.quaternion_multiply(want, ro.quaternion_conjugate(have))
            return -rel[1:] * np.sign(rel[0])
        if self.alg == 1:  # Caccavale et al. 1997, eq. 24 / 34
            R_e = self.rc.R(frame, q)
            rel = ro.unit_vector(ro.quaternion_from_matrix(np.dot(R_e.T, ro.euler_matrix_rxyz(*abg[:3]))))
            return -1 * np.dot(R_e, rel[1:])
        raise Exception(f"Invalid algorithm number {self.alg}")

    def limit_velocity(self, e):  # osc.py:198-215
        s = np.ones(6)
        nx, na = np.linalg.norm(e[:3]), np.linalg.norm(e[3:])
        if nx > self.lim_xyz:
            s[:3] *= self.lim_xyz / nx
        if na > self.lim_abg:
            s[3:] *= self.lim_abg / na
        return self.kv * s * (self.gains / self.kv) * e

    def generate(self, q, dq, target, target_velocity=None, ref_frame="EE", xyz_offset=None):
        rc, sel = self.rc, self.mask
        xdot_des = np.zeros(6) if target_velocity is None else target_velocity
        J = rc.J(ref_frame, q, x=xyz_offset)[sel]  # :242-244
        M = rc.M(q)
        Mx, M_inv = self._Mx(M=M, J=J)  # :246-247
        err = np.zeros(6)
        if sel[:3].any():  # :253-255
            err[:3] = rc.Tx(ref_frame, q, x=xyz_offset) - target[:3]
        if sel[3:].any():  # :258-259
            err[3:] = self.orientation_error(target[3:], q, ref_frame)
        if self.ki != 0:  # :262-264
            self.err_sum += err
            err += self.ki * self.err_sum
        err = self.limit_velocity(err) if self.vmax is not None else err * self.gains  # :267-272
        if np.all(xdot_des == 0):  # :275-282
            u = -1 * self.kv * np.dot(M, dq)
        else:
            u = np.zeros(rc.N_JOINTS)
            xdot = np.zeros(6)
            xdot[sel] = np.dot(J, dq)
            err = err + self.kv * (xdot - xdot_des)
        u = u - np.dot(J.T, np.dot(Mx, err[sel]))  # :285-288
        if self.use_C:  # :291-292
            u = u - np.dot(rc.C(q, dq), dq)
        self.training_signal = np.copy(u)  # :297
        if self.use_g:  # :300-301
            u = u - rc.g(q)
        for nc in self.nulls:  # :310-318
            Jbar = np.dot(M_inv, np.dot(J.T, Mx))
            u = u + np.dot(np.eye(rc.N_JOINTS) - np.dot(J.T, Jbar.T), nc.generate(q, dq))
        return u


def run_case(case, q, dq, target, target_velocity=None, mode="fp64"):
    """Evaluate one entry of tests/cases.py::OSC_CASES over a batch; returns (u, training_signal)."""
    rc = RobotOracle(case["arm"], mode)
    nulls = [NULL_KINDS[k](rc, **kw) for k, kw in case.get("null", [])] or None
    ctrlr = OSC(rc, null_controllers=nulls, **case["osc"])
    kw = {}
    if case.get("ref_frame"):
        kw["ref_frame"] = case["ref_frame"]
    if case.get("xyz_offset") is not None:
        kw["xyz_offset"] = np.array(case["xyz_offset"], dtype=float)
    us, ts = [], []
    for i in range(len(q)):
        if case.get("tv"):
            kw["target_velocity"] = target_velocity[i]
        us.append(np.asarray(ctrlr.generate(q[i], dq[i], target[i], **kw), dtype=np.float64))
        ts.append(np.asarray(ctrlr.training_signal, dtype=np.float64))
    return np.array(us), np.array(ts)


def run_ctrl_case(case, q, dq, target_q, target_dq, mode="fp64"):
    """tests/cases.py::CTRL_CASES -> (B, n)"""
    rc = RobotOracle(case["arm"], mode)
    kind, kw = case["ctrl"]
    out = []
    for i in range(len(q)):
        if kind == "Joint":
            c = Joint(rc, **kw)
            out.append(c.generate(q[i], dq[i], target_q[i], target_dq[i] if case.get("tv") else None))
        else:
            out.append(Floating(rc, **kw).generate(q[i], dq[i]))
    return np.array(out, dtype=np.float64)


def run_sliding_case(case, q, dq, target, tv, ta, mode="fp64"):
    """tests/cases.py::SLIDING_CASES -> (u, s), both (B, n)"""
    rc = RobotOracle(case["arm"], mode)
    ctrl = Sliding(rc, **case["ctrl"])
    kw = {}
    if case.get("ref_frame"):
        kw["ref_frame"] = case["ref_frame"]
    if case.get("offset") is not None:
        kw["offset"] = np.array(case["offset"], dtype=float)
    us, ss = [], []
    for i in range(len(q)):
        us.append(ctrl.generate(q[i], dq[i], target[i], target_velocity=0 if tv is None else tv[i],
                                target_acc=0 if ta is None else ta[i], **kw))
        ss.append(ctrl.s)
    return np.array(us, dtype=np.float64), np.array(ss, dtype=np.float64)


def run_null_case(case, q, dq, mode="fp64"):
    rc = RobotOracle(case["arm"], mode)
    kind, kw = case["ctrl"]
    ctrl = NULL_KINDS[kind](rc, **kw)
    return np.array([np.asarray(ctrl.generate(q[i], dq[i]), dtype=np.float64) for i in range(len(q))])
