"""CPU suite: the kernels' per-state arithmetic (abr_control_b200/csrc/*.cuh), instantiated by g++ through
tests/hostsim (TEST INFRASTRUCTURE, never loaded by the package), against the oracle.

This checks the *formulation* the CUDA path uses (joint-axis operators Omega_k, Cholesky-based task-space solve,
power-iteration quaternion, Jacobi pinv) on a box without a GPU; the real parity tests (`-m gpu`) run the same code
as CUDA kernels through the C ABI.
"""
import ctypes as C

import numpy as np
import pytest

import cases
from abr_control_b200 import _abi
from oracle import osc_oracle as oo
from oracle import rbd_oracle as ro


def P(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hs_rbd(hs, arm, q, dq, frame, xoff=None, f32=0, general=0):
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json(arm))
    n, B = cd.n_joints, len(q)
    fid = hs.hs_frame_id(n, frame.encode())
    shapes = dict(Tx=(3,), T=(4, 4), R=(3, 3), Tinv=(4, 4), quat=(4,), J=(6, n), dJ=(6, n), M=(n, n), g=(n,), C=(n, n))
    out = {k: np.zeros((B,) + s) for k, s in shapes.items()}
    xo = None if xoff is None else np.ascontiguousarray(xoff, dtype=np.float64)
    rc = hs.hs_rbd(C.byref(cd), f32, general, fid, P(xo), P(np.ascontiguousarray(q)), P(np.ascontiguousarray(dq)),
                   C.c_int64(B), *[P(out[k]) for k in ("Tx", "T", "R", "Tinv", "quat", "J", "dJ", "M", "g", "C")])
    assert rc == 0
    return out


def hs_osc(hs, cs, q, dq, target, tv, f32=0, general=0, ierr=None):
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json(cs["arm"]))
    n, B = cd.n_joints, len(q)
    nulls = [_abi.null_params(k, n, **kw) for k, kw in cs.get("null", [])]
    p = _abi.osc_params(n, null=nulls, **cs["osc"])
    fid = hs.hs_frame_id(n, cs.get("ref_frame", "EE").encode())
    xo = None if cs.get("xyz_offset") is None else np.array(cs["xyz_offset"], dtype=np.float64)
    u, tr, acc = np.zeros((B, n)), np.zeros((B, n)), np.zeros((B, n))
    tvv = np.ascontiguousarray(tv) if cs.get("tv") else None
    rc = hs.hs_osc(C.byref(cd), C.byref(p), f32, general, fid, P(xo), P(np.ascontiguousarray(q)),
                   P(np.ascontiguousarray(dq)), P(np.ascontiguousarray(target)), 6, P(tvv), 6, C.c_int64(B), P(u), P(tr),
                   P(acc), P(ierr))
    assert rc == 0
    return u, tr, acc


def _err(a, b):
    return float(np.max(np.abs(a - b)))


@pytest.mark.parametrize("arm,general", [("twojoint", 0), ("threejoint", 0), ("ur5", 0), ("ur5", 1), ("jaco2", 0)])
def test_rbd_math_vs_oracle(hostsim, arm, general):
    """general=1 forces the non-orthonormal operator path on an orthonormal chain (must agree)."""
    c = ro.ChainOracle(arm)
    n = c.n
    q, dq, _, _ = cases.states(arm, 24)
    xoff = np.array(cases.XOFF)
    for fr in cases.frames(n):
        o = hs_rbd(hostsim, arm, q, dq, fr, general=general)
        ox = hs_rbd(hostsim, arm, q, dq, fr, xoff, general=general)
        S = c.walk(q, 1)[fr]
        assert _err(o["Tx"], c.Tx(fr, q)) < 1e-13
        assert _err(o["T"], S.T) < 1e-13
        assert _err(o["R"], c.R(fr, q)) < 1e-13
        assert _err(o["Tinv"], c.T_inv(fr, q)) < 1e-13
        assert _err(o["J"], c.J(fr, q)) < 1e-13
        assert _err(o["dJ"], c.dJ(fr, q, dq)) < 1e-12
        assert _err(ox["Tx"], c.Tx(fr, q, xoff)) < 1e-13
        assert _err(ox["J"], c.J(fr, q, xoff)) < 1e-13
        assert _err(ox["dJ"], c.dJ(fr, q, dq, xoff)) < 1e-12
        qa, qb = o["quat"], c.quaternion(fr, q)
        e = np.minimum(np.abs(qa - qb).max(axis=1), np.abs(qa + qb).max(axis=1))
        assert np.all((np.abs(qa - qb).max(axis=1) < 1e-12) | ((np.abs(qb[:, 0]) < 1e-9) & (e < 1e-12)))
    o = hs_rbd(hostsim, arm, q, dq, "EE", general=general)
    assert _err(o["M"], c.M(q)) < 1e-12 * max(1, np.abs(c.M(q)).max())
    assert _err(o["g"], c.g(q)) < 1e-12 * max(1, np.abs(c.g(q)).max())
    assert _err(o["C"], c.C(q, dq)) < 1e-12 * max(1, np.abs(c.C(q, dq)).max())
    # float32 instantiation against the float64 one
    o32 = hs_rbd(hostsim, arm, q, dq, "EE", f32=1, general=general)
    for k, tol in (("Tx", 5e-6), ("J", 5e-6), ("M", 5e-5), ("g", 2e-4), ("C", 5e-4), ("dJ", 1e-4)):
        assert _err(o32[k], o[k]) < tol * max(1, np.abs(o[k]).max()), k


@pytest.mark.parametrize("name", list(cases.OSC_CASES))
def test_osc_math_vs_oracle(hostsim, name):
    cs = cases.OSC_CASES[name]
    q, dq, target, tvel = cases.states(cs["arm"], 32)
    ref, reft = oo.run_case(cs, q, dq, target, tvel, "fp64")
    u, tr, _ = hs_osc(hostsim, cs, q, dq, target, tvel)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(u - ref) / scale) < 1e-9
    assert np.max(np.abs(tr - reft) / scale) < 1e-9
    u32, _, _ = hs_osc(hostsim, cs, q, dq, target, tvel, f32=1)
    assert np.median(np.max(np.abs(u32 - ref) / scale, axis=1)) < 1e-5


def test_osc_math_general_path_on_ur5(hostsim):
    cs = cases.OSC_CASES["ur5_6dof_C_damp"]
    q, dq, target, tvel = cases.states("ur5", 16)
    a, _, _ = hs_osc(hostsim, cs, q, dq, target, tvel, general=0)
    b, _, _ = hs_osc(hostsim, cs, q, dq, target, tvel, general=1)
    assert np.max(np.abs(a - b)) < 1e-9 * np.abs(a).max()


@pytest.mark.parametrize("name", list(cases.NULL_CASES))
def test_null_math_vs_oracle(hostsim, name):
    cs = cases.NULL_CASES[name]
    q, dq, _, _ = cases.states(cs["arm"], 32)
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json(cs["arm"]))
    z = _abi.null_params(cs["ctrl"][0], cd.n_joints, **cs["ctrl"][1])
    u = np.zeros((len(q), cd.n_joints))
    assert hostsim.hs_null(C.byref(cd), C.byref(z), 0, 0, P(np.ascontiguousarray(q)), P(np.ascontiguousarray(dq)),
                           C.c_int64(len(q)), P(u)) == 0
    ref = oo.run_null_case(cs, q, dq)
    assert _err(u, ref) < 1e-8 * max(1, np.abs(ref).max())


@pytest.mark.parametrize("name", list(cases.CTRL_CASES))
def test_joint_floating_math_vs_oracle(hostsim, name):
    cs = cases.CTRL_CASES[name]
    q, dq, _, _ = cases.states(cs["arm"], 32)
    tq, tdq = cases.joint_targets(cs["arm"], 32)
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json(cs["arm"]))
    kind, kw = cs["ctrl"]
    u = np.zeros((len(q), cd.n_joints))
    if kind == "Joint":
        kp = kw.get("kp", 1)
        kv = kw.get("kv", np.sqrt(kp))
        rc = hostsim.hs_ctrl(C.byref(cd), 0, 0, 0, C.c_double(kp), C.c_double(kv), int(kw.get("account_for_gravity", True)), 0,
                             P(np.ascontiguousarray(q)), P(np.ascontiguousarray(dq)), P(np.ascontiguousarray(tq)),
                             P(np.ascontiguousarray(tdq)) if cs.get("tv") else None, C.c_int64(len(q)), P(u))
    else:
        rc = hostsim.hs_ctrl(C.byref(cd), 0, 0, 1, C.c_double(0), C.c_double(0), int(kw.get("task_space", False)),
                             int(kw.get("dynamic", False)), P(np.ascontiguousarray(q)), P(np.ascontiguousarray(dq)), None, None,
                             C.c_int64(len(q)), P(u))
    assert rc == 0
    ref = oo.run_ctrl_case(cs, q, dq, tq, tdq)
    assert _err(u, ref) < 1e-9 * max(1, np.abs(ref).max())


@pytest.mark.parametrize("name", list(cases.SLIDING_CASES))
def test_sliding_math_vs_oracle(hostsim, name):
    cs = cases.SLIDING_CASES[name]
    q, dq, _, _ = cases.states(cs["arm"], 32)
    tgt, tv, ta = cases.sliding_inputs(cs, 32)
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json(cs["arm"]))
    n = cd.n_joints
    kw = cs["ctrl"]
    xo = None if cs.get("offset") is None else np.array(cs["offset"], dtype=np.float64)
    fid = hostsim.hs_frame_id(n, cs.get("ref_frame", "EE").encode())
    ref, ref_s = oo.run_sliding_case(cs, q, dq, tgt, tv, ta)
    for f32, tol in ((0, 1e-9), (1, 2e-3)):
        u, s = np.zeros((len(q), n)), np.zeros((len(q), n))
        rc = hostsim.hs_sliding(C.byref(cd), f32, 0, C.c_double(kw.get("kd", 160.0)), C.c_double(kw.get("lamb", 30.0)),
                                int(kw.get("cartesian", True)), fid, P(xo), P(np.ascontiguousarray(q)),
                                P(np.ascontiguousarray(dq)), P(tgt), P(tv), P(ta), C.c_int64(len(q)), P(u), P(s))
        assert rc == 0
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert np.max(np.abs(u - ref) / scale) < tol, (name, f32)
        assert np.max(np.abs(s - ref_s) / np.abs(ref_s).max(axis=1, keepdims=True)) < tol


@pytest.mark.parametrize("name", list(cases.IK_CASES))
def test_ik_math_vs_oracle(hostsim, name):
    from oracle import ik_oracle

    cs = cases.IK_CASES[name]
    q = np.ascontiguousarray(cases.states(cs["arm"], cases.N_IK)[0])
    tg = np.ascontiguousarray(cases.ik_targets(cs["arm"]))
    ref_p, ref_v = ik_oracle.run_ik_case(cs, q, tg)
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json(cs["arm"]))
    n, B = cd.n_joints, len(q)
    init, path = cs.get("init", {}), cs["path"]
    steps = path["n_timesteps"]
    pp, vv = np.zeros((steps, B, n)), np.zeros((steps, B, n))
    rc = hostsim.hs_ik(C.byref(cd), 0, 0, C.c_double(init.get("max_dx", 0.2)), C.c_double(init.get("max_dr", 2 * np.pi)),
                       C.c_double(init.get("max_dq", np.pi)), path.get("method", 3), C.c_double(path.get("dt", 0.001)),
                       steps, P(q), P(tg), C.c_int64(B), P(pp), P(vv))
    assert rc == 0
    pp, vv = pp.transpose(1, 0, 2), vv.transpose(1, 0, 2)
    vs = np.abs(ref_v).max()
    assert np.abs(vv - ref_v).max() < 1e-8 * vs, name
    assert np.abs(pp - ref_p).max() < 1e-8


def test_singular_states_pinv_branch(hostsim):
    """rank-deficient J M^-1 J^T -> the reference's pinv(rcond=1e-4) branch (osc.py:143-145), incl. truncation."""
    for arm, qs in (("twojoint", [[0.3, 0.0], [1.0, np.pi], [2.0, 1e-9]]),
                    ("threejoint", [[0.5, 0.0, 0.0], [1.0, np.pi, 0.0], [0.2, 1e-7, -1e-7]])):
        q = np.array(qs)
        dq = np.full_like(q, 0.3)
        target = np.tile([0.5, 0.4, 0, 0, 0, 0.0], (len(q), 1))
        cs = dict(arm=arm, osc=dict(kp=10, ctrlr_dof=[True, True, False, False, False, False]))
        ref, _ = oo.run_case(cs, q, dq, target)
        u, _, _ = hs_osc(hostsim, cs, q, dq, target, None)
        assert np.abs(u - ref).max() < 1e-7 * np.abs(ref).max()
    # UR5 6-DOF: random states include truncated ones (3.6 % have |det| < 1e-3, some with an eigenvalue below rcond)
    rng = np.random.default_rng(0)
    q, dq, target = rng.uniform(0, 2 * np.pi, (64, 6)), rng.uniform(0, 5, (64, 6)), rng.uniform(-1, 1, (64, 6))
    cs = dict(arm="ur5", osc=dict(kp=50, ctrlr_dof=[True] * 6))
    ref, _ = oo.run_case(cs, q, dq, target)
    u, _, _ = hs_osc(hostsim, cs, q, dq, target, None)
    assert np.max(np.abs(u - ref) / np.abs(ref).max(axis=1, keepdims=True)) < 1e-9


def test_jacobi_pinv_route_vs_numpy(hostsim):
    """The one-sided Jacobi SVD of the rows of A (abrb_math.cuh: the sequential walk over the schedule the GPU runs
    warp-cooperatively) against numpy.linalg.pinv(A A^T, rcond) — the call OSC._Mx makes (osc.py:143-145) — on matrices
    with 0, 1, 2, 3 singular values below the cut-off, rank-deficient ones and zero rows (uncontrolled DOF)."""
    rng = np.random.default_rng(3)
    worst = 0.0
    for K in (6, 3):
        for trial in range(200):
            U, _ = np.linalg.qr(rng.normal(size=(K, K)))
            V, _ = np.linalg.qr(rng.normal(size=(6, 6)))
            n_small = trial % 4 if K == 6 else trial % 3
            sv = rng.uniform(0.3, 3.0, K)
            sv[:n_small] = rng.uniform(1e-9, 3e-3, n_small) if trial % 5 else 0.0  # squared: below 1e-4 * max
            A = (U * sv) @ V[:K]
            if trial % 7 == 0:
                A[K - 1] = 0.0  # an uncontrolled row
            y = rng.normal(size=K)
            if trial % 7 == 0:
                y[K - 1] = 0.0
            w = np.zeros(6)  # A^T pinv(A A^T) y: what the controller needs (J^T Mx y = L w)
            assert hostsim.hs_pinv(K, P(np.ascontiguousarray(A)), (1 << K) - 1, C.c_double(1e-4), P(y), P(w), 0) == 1
            S = A @ A.T
            lam = np.linalg.eigvalsh(S)
            if np.min(np.abs(lam / lam.max() - 1e-4)) < 1e-7:
                continue  # an eigenvalue on the cut-off itself: either side is right
            ref = A.T @ (np.linalg.pinv(S, rcond=1e-4, hermitian=True) @ y)
            worst = max(worst, np.abs(w - ref).max() / max(np.abs(ref).max(), 1e-30))
    assert worst < 1e-9, worst


def test_ki_integrator_sequence_vs_oracle(hostsim):
    """ki != 0: every state carries its own integrated task-space error (osc.py:81-82, :262-264); 12 consecutive calls
    on 8 independent state streams against 8 oracle controllers stepped the same way."""
    cs = dict(arm="ur5", osc=dict(kp=30, ki=0.7, ctrlr_dof=[True] * 6, use_C=True), null=[("Damping", dict(kv=10))])
    Bq, T = 8, 12
    rng = np.random.default_rng(9)
    ierr = np.zeros((Bq, 6))
    ctrl = []
    for b in range(Bq):
        rc = oo.RobotOracle("ur5", "fp64")
        ctrl.append(oo.OSC(rc, null_controllers=[oo.Damping(rc, kv=10)], **cs["osc"]))
    for t in range(T):
        q, dq, target = rng.uniform(0, 2 * np.pi, (Bq, 6)), rng.uniform(0, 2, (Bq, 6)), rng.uniform(-1, 1, (Bq, 6))
        u, _, _ = hs_osc(hostsim, cs, q, dq, target, None, ierr=ierr)
        ref = np.array([ctrl[b].generate(q[b], dq[b], target[b]) for b in range(Bq)])
        assert np.max(np.abs(u - ref) / np.abs(ref).max(axis=1, keepdims=True)) < 1e-9, t
        assert np.abs(ierr - np.array([c.err_sum for c in ctrl])).max() < 1e-12


def test_non_finite_states_terminate_and_stay_local(hostsim):
    """NaN / Inf in one state must not hang the iterative routes (bounded loops everywhere) nor touch other states."""
    cs = cases.OSC_CASES["ur5_6dof_C_damp"]
    rng = np.random.default_rng(0)
    q, tg = rng.uniform(0, 6, (8, 6)), rng.uniform(-1, 1, (8, 6))
    dq = 0.1 * q
    clean, _, _ = hs_osc(hostsim, cs, q, dq, tg, None)
    q2, dq2, tg2 = q.copy(), dq.copy(), tg.copy()
    q2[1, 2], q2[2, 0], dq2[3, 1], tg2[4, 4] = np.nan, np.inf, np.nan, np.nan
    u, _, _ = hs_osc(hostsim, cs, q2, dq2, tg2, None)
    bad = [1, 2, 3, 4]
    good = [0, 5, 6, 7]
    assert not np.isfinite(u[bad]).all(axis=1).any()
    assert np.array_equal(u[good], clean[good])


def test_truncating_route_on_a_larger_random_sample(hostsim):
    """1500 uniformly random UR5 states, 6-DOF + use_C + Damping: ~55 of them take the truncating pinv route (one, two
    or occasionally three eigenvalues dropped, some with the dropped one close to the next kept one).  Every state must
    match the oracle; a 60 000-state run of the same comparison is quoted in DESIGN.md S5."""
    rng = np.random.default_rng(31)
    B = 1500
    q, dq, target = rng.uniform(0, 2 * np.pi, (B, 6)), rng.uniform(0, 5, (B, 6)), rng.uniform(-1, 1, (B, 6))
    cs = dict(arm="ur5", osc=dict(kp=50, ctrlr_dof=[True] * 6, use_C=True), null=[("Damping", dict(kv=10))])
    ref, _ = oo.run_case(cs, q, dq, target)
    u, _, _ = hs_osc(hostsim, cs, q, dq, target, None)
    err = np.abs(u - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 1e-9 and np.median(err) < 1e-13


def test_plant_acceleration(hostsim):
    """ddq returned by the rollout variant solves M ddq = u + g - C dq."""
    cs = cases.OSC_CASES["ur5_xyz"]
    q, dq, target, _ = cases.states("ur5", 16)
    u, _, acc = hs_osc(hostsim, cs, q, dq, target, None)
    c = ro.ChainOracle("ur5")
    rhs = u + c.g(q) - np.einsum("bij,bj->bi", c.C(q, dq), dq)
    ref = np.linalg.solve(c.M(q), rhs[..., None])[..., 0]
    assert np.max(np.abs(acc - ref)) < 1e-8 * np.abs(ref).max()


def test_frame_names(hostsim):
    assert hostsim.hs_frame_id(6, b"link0") == 0 and hostsim.hs_frame_id(6, b"link6") == 6
    assert hostsim.hs_frame_id(6, b"joint0") == 7 and hostsim.hs_frame_id(6, b"joint5") == 12
    assert hostsim.hs_frame_id(6, b"EE") == 13
    for bad in (b"link7", b"joint6", b"ee", b"link", b"hand", b"link-1", b"joint1x"):
        assert hostsim.hs_frame_id(6, bad) == _abi.EFRAME


def test_ki_integrator_vs_reference_golden(hostsim):
    """the kernels' per-state code (integrated_error carried from call to call) against the reference's own sequences"""
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ur5_osc_ki.npz"))
    q, dq, target = g["q"], g["dq"], g["target"]
    T, S = q.shape[:2]
    cs = dict(arm="ur5", osc=dict(kp=float(g["kp"]), ki=float(g["ki"]), ctrlr_dof=[True] * 6, use_C=True),
              null=[("Damping", dict(kv=10))])
    ierr = np.zeros((S, 6))
    for t in range(T):
        u, _, _ = hs_osc(hostsim, cs, q[t], dq[t], target[t], None, ierr=ierr)
        ref = g["u64"][t]
        assert np.max(np.abs(u - ref) / np.abs(ref).max(axis=1, keepdims=True)) < 1e-9, t
        assert np.abs(ierr - g["integrated_error64"][t]).max() < 1e-11, t
