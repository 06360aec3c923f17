"""GPU parity tests proper: the CUDA path, called through the C ABI (via the ctypes-based Python mirror of the
reference's interface), against the reference's golden fixtures and against the oracle on seeded inputs.

Tolerances:
  * fp64 kernels vs raw fp64 reference / oracle: rtol 1e-10, atol 1e-12 for Tx, T, R, T_inv, J, dJ, M, g, C;
  * fp32 kernels vs the same: rtol 1e-4, atol 1e-5 (on quantities of O(1)), scaled by the row's magnitude;
  * OSC.generate: fp64 <= 1e-9 * |u|_inf of the fp64-mode reference; fp32 <= 1e-3 * |u|_inf, excluding states whose
    task-space inertia is ill conditioned (cond(J M^-1 J^T) > 1e4, where pinv truncation makes u discontinuous).
"""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def _cfg(arm, **kw):
    import abr_control_b200.arms as arms

    return getattr(arms, arm).Config(**kw)


def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert not bad.any(), f"{what}: max abs err {np.abs(a - b).max():.3e} (rtol {rtol}, atol {atol})"


def _quat_close(a, b, tol, what):
    # q and -q are the same rotation and the reference's sign is arbitrary when w == 0 (rotation by pi)
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    e = np.minimum(np.abs(a - b).max(axis=-1), np.abs(a + b).max(axis=-1))
    flip_ok = np.abs(b[..., 0]) < 1e-7
    direct = np.abs(a - b).max(axis=-1)
    assert np.all((direct < tol) | (flip_ok & (e < tol))), f"{what}: {direct.max():.3e}"


@pytest.mark.parametrize("arm", list(cases.ARMS))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rbd_vs_reference_golden(arm, dtype):
    g = np.load(f"{GOLD}/{arm}_rbd.npz")
    rc = _cfg(arm, dtype=dtype)
    q, dq, xoff = g["q"].astype(dtype), g["dq"].astype(dtype), g["xoff"]
    rtol, atol = (1e-10, 1e-12) if dtype == np.float64 else (1e-4, 2e-5)
    n = rc.N_JOINTS
    for fr in cases.frames(n):
        out = rc.eval(q, dq, name=fr, want=("Tx", "R", "T_inv", "J"))
        assert out["J"].dtype == dtype and out["J"].shape == (len(q), 6, n)
        _close(out["Tx"], g[f"Tx_{fr}"], rtol, atol, f"Tx {fr}")
        _close(out["R"], g[f"R_{fr}"], rtol, atol, f"R {fr}")
        _close(out["T_inv"], g[f"Tinv_{fr}"], rtol, atol, f"T_inv {fr}")
        _close(out["J"], g[f"J_{fr}"], rtol, atol, f"J {fr}")
    for fr in ("EE", cases.ARMS[arm]["mid"]):
        out = rc.eval(q, dq, name=fr, want=("T", "dJ", "quat"))
        _close(out["T"], g[f"T_{fr}"], rtol, atol, f"T {fr}")
        _close(out["dJ"], g[f"dJ_{fr}"], rtol * 10, atol * 50, f"dJ {fr}")
        _quat_close(out["quat"], g[f"quat_{fr}"], 1e-9 if dtype == np.float64 else 1e-5, f"quat {fr}")
        outx = rc.eval(q, dq, name=fr, x=xoff, want=("Tx", "J"))
        _close(outx["Tx"], g[f"Txx_{fr}"], rtol, atol, f"Tx(x) {fr}")
        _close(outx["J"], g[f"Jx_{fr}"], rtol, atol, f"J(x) {fr}")
    _close(rc.dJ("EE", q, dq, x=xoff), g["dJx_EE"], rtol * 10, atol * 50, "dJ(x) EE")
    dyn = rc.eval(q, dq, want=("M", "g") + (("C",) if "C" in g else ()))
    scale_g = max(1.0, np.abs(g["g"]).max())
    _close(dyn["M"], g["M"], rtol, atol * max(1.0, np.abs(g["M"]).max()), "M")
    _close(dyn["g"], g["g"], rtol, atol * scale_g * 10, "g")
    if "C" in g:
        _close(dyn["C"], g["C"], rtol * 10, atol * 50 * max(1.0, np.abs(g["C"]).max()), "C")


def test_jaco2_C_vs_oracle():
    """The reference cannot generate Jaco2's C; pin it through the oracle restatement, which is
    itself pinned on UR5/twojoint/threejoint where the reference C exists, and through finite differences of M."""
    from oracle import rbd_oracle

    rc = _cfg("jaco2")
    q, dq, _, _ = cases.states("jaco2", 24)
    C = rc.C(q, dq)
    _close(C, rbd_oracle.ChainOracle("jaco2").C(q, dq), 1e-10, 1e-11, "jaco2 C vs oracle")
    # Christoffel consistency: dM/dt = C + C^T along dq (finite differences of the GPU's own M)
    h = 1e-6
    Mdot = (rc.M(q + h * dq) - rc.M(q - h * dq)) / (2 * h)
    _close(C + np.swapaxes(C, 1, 2), Mdot, 1e-5, 1e-6, "Mdot = C + C^T")


def _build_ctrl(rc, case):
    from abr_control_b200 import controllers

    nulls = [getattr(controllers, k)(rc, **kw) for k, kw in case.get("null", [])] or None
    return controllers.OSC(rc, null_controllers=nulls, **case["osc"])


def _gen(ctrlr, case, q, dq, target, tvel):
    kw = {}
    if case.get("ref_frame"):
        kw["ref_frame"] = case["ref_frame"]
    if case.get("xyz_offset") is not None:
        kw["xyz_offset"] = case["xyz_offset"]
    if case.get("tv"):
        kw["target_velocity"] = tvel
    return ctrlr.generate(q, dq, target, **kw)


def _well_conditioned(case, q):
    """states whose J M^-1 J^T (controlled rows) has cond <= 1e4 (oracle-side diagnostic)."""
    from oracle import rbd_oracle

    ch = rbd_oracle.ChainOracle(case["arm"])
    mask = np.array(case["osc"].get("ctrlr_dof", [1, 1, 1, 0, 0, 0]), dtype=bool)
    x = case.get("xyz_offset")
    J = ch.J(case.get("ref_frame", "EE"), q, x)[:, mask]
    S = J @ np.linalg.inv(ch.M(q)) @ np.swapaxes(J, 1, 2)
    return np.linalg.cond(S) <= 1e4


@pytest.mark.parametrize("name", list(cases.OSC_CASES))
def test_osc_vs_reference_golden(name):
    case = cases.OSC_CASES[name]
    o = np.load(f"{GOLD}/{case['arm']}_osc.npz")
    q, dq, target, tvel = o["q"], o["dq"], o["target"], o["target_velocity"]
    ref = o[f"{name}__u64"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    ok = _well_conditioned(case, q)
    assert ok.sum() >= len(q) // 2
    # fp64
    rc = _cfg(case["arm"])
    ctrlr = _build_ctrl(rc, case)
    u = _gen(ctrlr, case, q, dq, target, tvel)
    assert u.dtype == np.float64 and u.shape == ref.shape
    err = np.abs(u - ref) / scale
    assert err[ok].max() < 1e-9, f"{name} fp64: {err[ok].max():.3e}"
    assert err.max() < 1e-6, f"{name} fp64 incl. ill-conditioned states: {err.max():.3e}"
    tr = np.abs(ctrlr.training_signal - o[f"{name}__train64"]) / scale
    assert tr[ok].max() < 1e-9
    # fp32 kernel against the fp64-mode reference
    rc32 = _cfg(case["arm"], dtype=np.float32)
    u32 = _gen(_build_ctrl(rc32, case), case, q.astype(np.float32), dq.astype(np.float32),
               target.astype(np.float32), tvel.astype(np.float32))
    assert u32.dtype == np.float32
    e32 = np.abs(u32 - ref) / scale
    assert e32[ok].max() < 1e-3, f"{name} fp32: {e32[ok].max():.3e}"
    # and the reference as shipped (float32-rounded J/M/g) is no closer to its own fp64 mode than we are, x10
    ship = np.abs(o[f"{name}__u32"] - ref) / scale
    assert np.median(e32[ok]) < 10 * max(np.median(ship[ok]), 1e-7)


@pytest.mark.parametrize("name", list(cases.NULL_CASES))
def test_null_controllers_vs_reference_golden(name):
    from abr_control_b200 import controllers

    case = cases.NULL_CASES[name]
    o = np.load(f"{GOLD}/{case['arm']}_osc.npz")
    rc = _cfg(case["arm"])
    kind, kw = case["ctrl"]
    ctrl = getattr(controllers, kind)(rc, **kw)
    u = ctrl.generate(o["q"], o["dq"])
    ref = o[f"{name}__null64"]
    _close(u, ref, 1e-8, 1e-9 * max(1.0, np.abs(ref).max()), name)


@pytest.mark.parametrize("name", list(cases.SLIDING_CASES))
def test_sliding_vs_reference_golden(name):
    import torch

    from abr_control_b200.controllers import Sliding

    cs = cases.SLIDING_CASES[name]
    o = np.load(f"{GOLD}/{cs['arm']}_osc.npz")
    q, dq = o["q"], o["dq"]
    tgt, tv, ta = cases.sliding_inputs(cs)
    ref, ref_s = o[f"{name}__sliding64"], o[f"{name}__s64"]
    kw = {}
    if cs.get("ref_frame"):
        kw["ref_frame"] = cs["ref_frame"]
    if cs.get("offset") is not None:
        kw["offset"] = cs["offset"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    sscale = np.abs(ref_s).max(axis=1, keepdims=True)
    for dtype, tol in ((np.float64, 1e-9), (np.float32, 5e-3)):
        ctrl = Sliding(_cfg(cs["arm"], dtype=dtype), **cs["ctrl"])
        cast = lambda a: None if a is None else a.astype(dtype)  # noqa: E731
        u = ctrl.generate(cast(q), cast(dq), cast(tgt), target_velocity=0 if tv is None else cast(tv),
                          target_acc=0 if ta is None else cast(ta), **kw)
        assert u.dtype == dtype and u.shape == ref.shape
        assert (np.abs(u - ref) / scale).max() < tol, (name, dtype)
        assert (np.abs(ctrl.s - ref_s) / sscale).max() < tol
    # one state: float64 out like the reference; CUDA tensors in -> CUDA tensors out
    ctrl = Sliding(_cfg(cs["arm"]), **cs["ctrl"])
    one = ctrl.generate(q[0], dq[0], tgt[0], target_velocity=0 if tv is None else tv[0],
                        target_acc=0 if ta is None else ta[0], **kw)
    assert one.shape == ref[0].shape and np.abs(one - ref[0]).max() < 1e-9 * max(1.0, np.abs(ref[0]).max())
    assert np.abs(ctrl.s - ref_s[0]).max() < 1e-9 * max(1.0, np.abs(ref_s[0]).max())
    ud = ctrl.generate(torch.as_tensor(q, device="cuda"), torch.as_tensor(dq, device="cuda"),
                       torch.as_tensor(tgt, device="cuda"), **kw)
    assert ud.is_cuda and ud.shape == ref.shape and ctrl.s.is_cuda


@pytest.mark.parametrize("name", list(cases.IK_CASES))
def test_inverse_kinematics_paths_vs_reference_golden(name):
    import torch

    from abr_control_b200.controllers.path_planners import InverseKinematics

    cs = cases.IK_CASES[name]
    o = np.load(f"{GOLD}/{cs['arm']}_ik.npz")
    position, target = o["position"], o["target"]
    ref_p, ref_v = o[f"{name}__pos64"], o[f"{name}__vel64"]
    vs = np.abs(ref_v).max()
    # fp64: the iteration is contracting, errors do not grow over the horizon; fp32 drifts by rounding per step
    for dtype, tol in ((np.float64, 1e-8), (np.float32, 2e-2)):
        ik = InverseKinematics(_cfg(cs["arm"], dtype=dtype), **cs.get("init", {}))
        pos, vel = ik.generate_path(position.astype(dtype), target.astype(dtype), **cs["path"])
        assert pos.shape == ref_p.shape and pos.dtype == dtype
        assert np.abs(vel - ref_v).max() < tol * vs, (name, dtype)
        assert np.abs(pos - ref_p).max() < tol * max(1.0, vs * ref_p.shape[1])
    ik = InverseKinematics(_cfg(cs["arm"]), **cs.get("init", {}))
    one_p, one_v = ik.generate_path(position[0], target[0], **cs["path"])  # the reference's single-path contract
    assert one_p.shape == ref_p[0].shape and np.abs(one_p - ref_p[0]).max() < 1e-8
    assert np.abs(one_v - ref_v[0]).max() < 1e-8 * vs
    p0, _ = ik.next()
    assert np.array_equal(p0, one_p[0]) and ik.n == 1
    dp, dv = ik.generate_path(torch.as_tensor(position, device="cuda"), torch.as_tensor(target[0], device="cuda"),
                              **cs["path"])  # CUDA tensors, one broadcast target
    assert dp.is_cuda and tuple(dp.shape) == ref_p.shape and bool(torch.isfinite(dv).all())


@pytest.mark.parametrize("name", list(cases.CTRL_CASES))
def test_joint_and_floating_vs_reference_golden(name):
    import torch

    from abr_control_b200 import controllers

    cs = cases.CTRL_CASES[name]
    o = np.load(f"{GOLD}/{cs['arm']}_osc.npz")
    q, dq, tq, tdq = o["q"], o["dq"], o["joint_target"], o["joint_target_velocity"]
    ref = o[f"{name}__ctrl64"]
    kind, kw = cs["ctrl"]
    for dtype, tol in ((np.float64, 1e-9), (np.float32, 2e-3)):
        rc = _cfg(cs["arm"], dtype=dtype)
        ctrl = getattr(controllers, kind)(rc, **kw)
        cast = lambda a: a.astype(dtype)  # noqa: E731
        if kind == "Joint":
            u = ctrl.generate(cast(q), cast(dq), cast(tq), cast(tdq) if cs.get("tv") else None)
            one = ctrl.generate(q[0], dq[0], tq[0], tdq[0] if cs.get("tv") else None)
        else:
            u = ctrl.generate(cast(q), cast(dq))
            one = ctrl.generate(q[0], dq[0])
        assert u.dtype == dtype and u.shape == ref.shape
        scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-12
        assert (np.abs(u - ref) / scale).max() < tol, (name, dtype)
        assert one.shape == ref[0].shape and one.dtype == np.float64
        assert np.abs(one - ref[0]).max() < 1e-9 * max(1.0, np.abs(ref[0]).max())
    # CUDA tensors in -> CUDA tensor out
    rc = _cfg(cs["arm"])
    ctrl = getattr(controllers, kind)(rc, **kw)
    tq_, tdq_ = torch.as_tensor(q, device="cuda"), torch.as_tensor(dq, device="cuda")
    if kind == "Joint":
        ud = ctrl.generate(tq_, tdq_, torch.as_tensor(tq, device="cuda"))
    else:
        ud = ctrl.generate(tq_, tdq_)
    assert ud.is_cuda and ud.shape == ref.shape


def test_single_state_contract():
    """One state in -> the reference's shapes/dtypes out (float32 J/M/g/C/R, float64 Tx/T/T_inv, fresh arrays)."""
    from abr_control_b200.controllers import OSC

    rc = _cfg("ur5")
    g = np.load(f"{GOLD}/ur5_rbd.npz")
    q, dq = g["q"][3], g["dq"][3]
    J = rc.J("EE", q)
    assert J.shape == (6, 6) and J.dtype == np.float32
    assert np.allclose(J, g["J_EE"][3].astype(np.float32), rtol=1e-6, atol=1e-7)
    assert rc.M(list(q)).dtype == np.float32 and rc.g(q).shape == (6,) and rc.C(q, dq).dtype == np.float32
    assert rc.Tx("EE", q).dtype == np.float64 and rc.Tx("EE", q).shape == (3,)
    assert rc.T("EE", q).shape == (4, 4) and rc.T_inv("link3", q).shape == (4, 4) and rc.R("joint2", q).shape == (3, 3)
    with pytest.raises(Exception, match="Invalid transformation name"):
        rc.Tx("link9", q)
    with pytest.raises(Exception, match="Invalid transformation name"):
        rc.J("hand", q)
    ctrlr = OSC(rc, kp=10)
    o = np.load(f"{GOLD}/ur5_osc.npz")
    u = ctrlr.generate(o["q"][0], o["dq"][0], o["target"][0])
    assert u.shape == (6,) and u.dtype == np.float64 and u.flags.writeable
    ref = o["ur5_xyz__u64"][0]
    assert np.abs(u - ref).max() < 1e-9 * np.abs(ref).max()
    u *= -1  # CoppeliaSim.send_forces negates in place (interfaces/coppeliasim.py:204)
    assert ctrlr.training_signal.shape == (6,)
    with pytest.raises(NotImplementedError):
        OSC(rc, ki=0.1)
    with pytest.raises(Exception, match="Invalid algorithm number"):
        OSC(rc, orientation_algorithm=2)


@pytest.mark.parametrize("B", [0, 1, 31, 32, 33, 127, 129, 1000])
def test_ragged_batches_and_device_tensors(B):
    """Empty, sub-warp and ragged batch sizes; CUDA-tensor path equals host-buffer path bit for bit."""
    import torch

    from abr_control_b200.controllers import OSC, Damping

    rc = _cfg("jaco2")
    q, dq, target, _ = cases.states("jaco2", max(B, 1))
    q, dq, target = q[:B], dq[:B], target[:B]
    ctrlr = OSC(rc, kp=200, ctrlr_dof=[True] * 5 + [False], null_controllers=[Damping(rc, kv=10)])
    u_host = ctrlr.generate(q, dq, target)
    assert u_host.shape == (B, 6)
    tq, tdq, tt = (torch.as_tensor(a, device="cuda") for a in (q, dq, target))
    u_dev = ctrlr.generate(tq, tdq, tt)
    assert u_dev.is_cuda and u_dev.shape == (B, 6)
    assert np.array_equal(u_dev.cpu().numpy(), u_host)
    if B:
        one = ctrlr.generate(q[B - 1], dq[B - 1], target[B - 1])
        assert np.array_equal(one, u_host[B - 1])
        out = rc.eval(tq, tdq, want=("J", "M", "g", "C", "Tx"))
        out_h = rc.eval(q, dq, want=("J", "M", "g", "C", "Tx"))
        for k in out:
            assert np.array_equal(out[k].cpu().numpy(), out_h[k]), k
        # broadcast target row
        ub = ctrlr.generate(q, dq, target[0])
        assert np.array_equal(ub[0], u_host[0])


def test_singular_states_take_the_pinv_branch():
    """Stretched-out planar arms make J M^-1 J^T singular: the reference's pinv(rcond=1e-4) branch (osc.py:143-145)."""
    from oracle import osc_oracle

    for arm, qs in (("twojoint", [[0.3, 0.0], [1.0, np.pi], [2.0, 1e-9]]),
                    ("threejoint", [[0.5, 0.0, 0.0], [1.0, np.pi, 0.0], [0.2, 1e-7, -1e-7]])):
        q = np.array(qs)
        dq = np.full_like(q, 0.3)
        target = np.tile([0.5, 0.4, 0, 0, 0, 0.0], (len(q), 1))
        case = dict(arm=arm, osc=dict(kp=10, ctrlr_dof=[True, True, False, False, False, False]))
        ref, _ = osc_oracle.run_case(case, q, dq, target)
        u = _gen(_build_ctrl(_cfg(arm), case), case, q, dq, target, None)
        assert np.all(np.isfinite(u))
        assert np.abs(u - ref).max() < 1e-7 * np.abs(ref).max(), (arm, u, ref)


def test_large_batch_deferred_pinv_states():
    """Every state of a 4096-state UR5 6-DOF batch that takes the truncating-pinv branch (|det| < 1e-3, ~3.75 % of
    uniformly random states) must match the oracle, fp64 and fp32, and must not depend on the batch it sits in."""
    from oracle import osc_oracle, rbd_oracle

    rng = np.random.default_rng(11)
    B = 4096
    q, dq, target = rng.uniform(0, 2 * np.pi, (B, 6)), rng.uniform(0, 5, (B, 6)), rng.uniform(-1, 1, (B, 6))
    ch = rbd_oracle.ChainOracle("ur5")
    J = ch.J("EE", q)
    S = J @ np.linalg.inv(ch.M(q)) @ np.swapaxes(J, 1, 2)
    slow = np.where(np.abs(np.linalg.det(S)) < 1e-3)[0]
    assert 80 < len(slow) < 300  # ~3.75 % of uniformly random states
    pick = np.concatenate([slow, np.arange(0, B, 97)])
    case = dict(arm="ur5", osc=dict(kp=50, ctrlr_dof=[True] * 6, use_C=True), null=[("Damping", dict(kv=10))])
    ref, _ = osc_oracle.run_case(case, q[pick], dq[pick], target[pick])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    w = np.linalg.eigvalsh(S[pick])
    # states with an eigenvalue within 0.1 % of the pinv cutoff are genuinely ambiguous: exclude them
    clear = np.all(np.abs(w / (1e-4 * w[:, -1:]) - 1) > 1e-3, axis=1)
    for dtype, tol in ((np.float64, 1e-6), (np.float32, 5e-2)):
        rc = _cfg("ur5", dtype=dtype)
        ctrlr = _build_ctrl(rc, case)
        u = ctrlr.generate(q.astype(dtype), dq.astype(dtype), target.astype(dtype))
        err = (np.abs(u[pick] - ref) / scale).max(axis=1)
        assert err[clear].max() < tol, (dtype, err[clear].max())
        assert np.median(err) < (1e-12 if dtype == np.float64 else 1e-5)
        # and identical to evaluating the same rows in a small batch
        small = ctrlr.generate(q[slow[:64]].astype(dtype), dq[slow[:64]].astype(dtype), target[slow[:64]].astype(dtype))
        assert np.allclose(small, u[slow[:64]], rtol=1e-9 if dtype == np.float64 else 1e-4, atol=0)


def test_two_launch_mode_matches_single_launch():
    """Optional two-launch mode of the 6-row OSC path (everything but the truncating-pinv states, then those states
    from an index queue; abr_control_b200/csrc/kernels.cu, abrb_osc_set_option "two_launch_min").  It must agree with
    the single-launch mode (the same rows in chunks below the threshold) for every row, repeatedly (the queue re-arms
    itself), for batch sizes that grow, shrink and are not multiples of the warp size, with and without the
    training-signal output."""
    import torch

    rng = np.random.default_rng(23)
    case = dict(arm="ur5", osc=dict(kp=50, ctrlr_dof=[True] * 6, use_C=True), null=[("Damping", dict(kv=10))])
    for dtype, tol in ((np.float64, 1e-9), (np.float32, 2e-3)):
        rc = _cfg("ur5", dtype=dtype)
        ctrlr = _build_ctrl(rc, case)
        ctrlr.set_option("two_launch_min", 16384)
        for B in (20000, 40003, 16384):
            q = rng.uniform(0, 2 * np.pi, (B, 6)).astype(dtype)
            dq = rng.uniform(0, 5, (B, 6)).astype(dtype)
            target = rng.uniform(-1, 1, (B, 6)).astype(dtype)
            tv = rng.uniform(-0.5, 0.5, (B, 6)).astype(dtype)
            for kw in ({}, {"target_velocity": tv}):
                tq, tdq, tt = (torch.as_tensor(a, device="cuda") for a in (q, dq, target))
                kwd = {k: torch.as_tensor(v, device="cuda") for k, v in kw.items()}
                big = ctrlr.generate(tq, tdq, tt, **kwd)
                big_train = ctrlr.training_signal.clone()
                again = ctrlr.generate(tq, tdq, tt, **kwd)
                assert torch.equal(big, again)  # deterministic, and the queue was re-armed
                parts, parts_train = [], []
                for s0 in range(0, B, 4096):
                    sl = slice(s0, min(B, s0 + 4096))
                    parts.append(ctrlr.generate(tq[sl], tdq[sl], tt[sl], **{k: v[sl] for k, v in kwd.items()}))
                    parts_train.append(ctrlr.training_signal.clone())
                ref, ref_train = torch.cat(parts), torch.cat(parts_train)
                scale = ref.abs().amax(dim=1, keepdim=True)
                assert float(((big - ref).abs() / scale).max()) < tol
                assert float(((big_train - ref_train).abs() / scale).max()) < tol
                assert bool(torch.isfinite(big).all())


def test_full_size_properties():
    """BASELINE configs at full size, through size-independent properties (the oracle is too slow here):
    config 2 (UR5 {J,M,g,C}, B=65536, fp64) and config 3 (Jaco2 OSC, B=262144, fp32)."""
    import torch

    from abr_control_b200.controllers import OSC, Damping

    rng = np.random.default_rng(7)
    B = 65536
    rc = _cfg("ur5")
    q = torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 6)), device="cuda")
    dq = torch.as_tensor(rng.uniform(0, 5, (B, 6)), device="cuda")
    out = rc.eval(q, dq, want=("J", "M", "g", "C", "Tx"))
    M, Cm = out["M"], out["C"]
    assert torch.equal(M, M.transpose(1, 2).contiguous())  # built from one triangle
    assert int(torch.linalg.cholesky_ex(M).info.abs().max()) == 0  # positive definite
    assert all(torch.isfinite(v).all() for v in out.values())
    # idempotence / determinism and batch-composition independence: a permuted batch gives permuted rows
    perm = torch.randperm(B, device="cuda")
    out_p = rc.eval(q[perm].contiguous(), dq[perm].contiguous(), want=("J", "M", "g", "C", "Tx"))
    for k in out:
        assert torch.equal(out_p[k], out[k][perm]), k
    # the Jacobian is the derivative of Tx: central differences along a random direction
    d = torch.as_tensor(rng.normal(size=(B, 6)), device="cuda")
    h = 1e-6
    fd = (rc.Tx("EE", q + h * d) - rc.Tx("EE", q - h * d)) / (2 * h)
    jd = torch.einsum("bij,bj->bi", out["J"][:, :3], d)
    assert (fd - jd).abs().max() < 1e-7
    # Christoffel identity: dM/dt - 2C is skew  <=>  dM/dt = C + C^T
    Mdot = (rc.M(q + h * dq) - rc.M(q - h * dq)) / (2 * h)
    assert (Cm + Cm.transpose(1, 2) - Mdot).abs().max() < 5e-5
    # gravity is minus the gradient of the potential energy sum_l m_l g z_l: check via energy differences
    # (skipped for brevity of runtime: covered at small size against the reference golden)

    B3 = 262144
    rc3 = _cfg("jaco2", dtype=np.float32)
    ctrlr = OSC(rc3, kp=200, ctrlr_dof=[True] * 5 + [False], null_controllers=[Damping(rc3, kv=10)])
    q3 = torch.as_tensor(rng.uniform(0, 2 * np.pi, (B3, 6)), device="cuda", dtype=torch.float32)
    dq3 = torch.as_tensor(rng.uniform(0, 5, (B3, 6)), device="cuda", dtype=torch.float32)
    t3 = torch.as_tensor(rng.uniform(-1, 1, (B3, 6)), device="cuda", dtype=torch.float32)
    u = ctrlr.generate(q3, dq3, t3)
    assert u.shape == (B3, 6) and torch.isfinite(u).all()
    perm = torch.randperm(B3, device="cuda")
    assert torch.equal(ctrlr.generate(q3[perm].contiguous(), dq3[perm].contiguous(), t3[perm].contiguous()), u[perm])
    # fp32 against the fp64 kernel on the same inputs
    rc3d = _cfg("jaco2")
    ctrlr_d = OSC(rc3d, kp=200, ctrlr_dof=[True] * 5 + [False], null_controllers=[Damping(rc3d, kv=10)])
    ud = ctrlr_d.generate(q3.double(), dq3.double(), t3.double())
    rel = (u.double() - ud).abs().amax(dim=1) / ud.abs().amax(dim=1)
    assert rel.median() < 1e-5 and torch.quantile(rel[:100000], 0.99) < 1e-3


def test_rollout_matches_stepwise():
    """The fused rollout kernel equals calling generate + the plant update step by step."""
    import torch

    from abr_control_b200.controllers import OSC

    rc = _cfg("ur5")
    ctrlr = OSC(rc, kp=10)
    q, dq, target, _ = cases.states("ur5", 64)
    dq = dq * 0.1
    steps, dt = 16, 1e-3
    qf, dqf, traj = ctrlr.rollout(q, dq, target, steps=steps, dt=dt)
    assert traj["q"].shape == (steps, 64, 6)
    qs, dqs = q.copy(), dq.copy()
    for t in range(steps):
        u = ctrlr.generate(qs, dqs, target)
        d = rc.eval(qs, dqs, want=("M", "g", "C"))
        rhs = u + d["g"] - np.einsum("bij,bj->bi", d["C"], dqs)
        ddq = np.linalg.solve(d["M"], rhs[..., None])[..., 0]
        dqs = dqs + ddq * dt
        qs = qs + dqs * dt
        assert np.abs(traj["u"][t] - u).max() < 1e-7 * max(1.0, np.abs(u).max())
    assert np.abs(qf - qs).max() < 1e-9 and np.abs(dqf - dqs).max() < 1e-7
