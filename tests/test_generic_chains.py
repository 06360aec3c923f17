"""CPU suite: chains that are NOT one of the four reference arms — every joint count the library is built for
(N = 1..7), orthonormal and deliberately non-orthonormal constant frames — through the kernels' per-state code
(tests/hostsim) against the oracle built from the same descriptor.  This is what SURVEY.md S8(f) row 4 (generic chain
import) rests on: `BaseConfig(desc)` accepts any such descriptor."""
import ctypes as C

import numpy as np
import pytest

from abr_control_b200 import _abi
from oracle import osc_oracle as oo
from oracle import rbd_oracle as ro


def P(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def random_chain(n, ortho, seed, shear=0.02):
    rng = np.random.default_rng(seed)

    def frame():
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        if not ortho:
            Q = Q @ (np.eye(3) + shear * rng.standard_normal((3, 3)))  # measured-looking frames
        return np.hstack([Q, rng.uniform(-0.3, 0.3, (3, 1))]).tolist()

    return dict(name=f"rand{n}", n_joints=n, n_links=n + 1, gravity=[0, 0, -9.81, 0, 0, 0], L0=frame(),
                A=[frame() for _ in range(n)], B=[frame() for _ in range(n)], E=frame(),
                link_inertia=[list(rng.uniform(0.2, 3.0, 3).repeat(1)[[0, 0, 0]]) + list(rng.uniform(0.01, 0.2, 3))
                              for _ in range(n + 1)])


def hs_rbd(hs, desc, q, dq, frame, f32=0):
    cd = _abi.chain_desc_from_dict(desc)
    n, B = cd.n_joints, len(q)
    fid = hs.hs_frame_id(n, frame.encode())
    shapes = dict(Tx=(3,), T=(4, 4), R=(3, 3), Tinv=(4, 4), quat=(4,), J=(6, n), dJ=(6, n), M=(n, n), g=(n,), C=(n, n))
    out = {k: np.zeros((B,) + s) for k, s in shapes.items()}
    rc = hs.hs_rbd(C.byref(cd), f32, 0, fid, None, P(np.ascontiguousarray(q)), P(np.ascontiguousarray(dq)), C.c_int64(B),
                   *[P(out[k]) for k in ("Tx", "T", "R", "Tinv", "quat", "J", "dJ", "M", "g", "C")])
    assert rc == 0
    return out


@pytest.mark.parametrize("ortho", [True, False])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7])
def test_rigid_body_quantities_of_random_chains(hostsim, n, ortho):
    desc = random_chain(n, ortho, 100 * n + ortho)
    c = ro.ChainOracle(desc)
    cd = _abi.chain_desc_from_dict(desc)
    rng = np.random.default_rng(n)
    q, dq = rng.uniform(0, 2 * np.pi, (12, n)), rng.uniform(-3, 3, (12, n))
    for fr in ("EE", f"link{n}", f"joint{n - 1}", "link0", f"link{(n + 1) // 2}"):
        o = hs_rbd(hostsim, desc, q, dq, fr)
        assert np.abs(o["Tx"] - c.Tx(fr, q)).max() < 1e-12
        assert np.abs(o["R"] - c.R(fr, q)).max() < 1e-12
        assert np.abs(o["J"] - c.J(fr, q)).max() < 1e-12
        assert np.abs(o["dJ"] - c.dJ(fr, q, dq)).max() < 1e-11
    o = hs_rbd(hostsim, desc, q, dq, "EE")
    for k, ref in (("M", c.M(q)), ("g", c.g(q)), ("C", c.C(q, dq))):
        assert np.abs(o[k] - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()), k
    assert bool(cd.n_joints == n)


@pytest.mark.parametrize("n,dof,ortho", [(3, [1, 1, 1, 0, 0, 0], True), (4, [1, 1, 1, 0, 1, 0], True),
                                         (5, [1, 1, 1, 1, 1, 0], True), (6, [1, 1, 1, 1, 1, 1], True),
                                         (6, [1, 1, 1, 1, 1, 1], False), (7, [1, 1, 1, 1, 1, 1], True),
                                         (7, [1, 1, 1, 0, 0, 0], False)])
def test_osc_on_random_chains(hostsim, n, dof, ortho):
    """incl. the redundant 7-joint case (null space of dimension 1 or 4) with a Damping secondary controller"""
    # orientation control reads the frame's quaternion by a 5-step power iteration, exact for rotations and accurate to
    # ~1e-12 for frames as far from orthonormal as 2e-3 (20x the Jaco2's); the 2 % shear of the rbd test above would cost
    # 1e-8 there
    desc = random_chain(n, ortho, 7 * n + sum(dof), shear=2e-3)
    case = dict(arm=desc, osc=dict(kp=25, ko=15, ctrlr_dof=[bool(d) for d in dof], use_C=True),
                null=[("Damping", dict(kv=4))])
    rng = np.random.default_rng(n + 40)
    B = 24
    q, dq, target = rng.uniform(0, 2 * np.pi, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-0.6, 0.6, (B, 6))
    ref, _ = oo.run_case(case, q, dq, target)
    cd = _abi.chain_desc_from_dict(desc)
    nulls = [_abi.null_params(k, n, **kw) for k, kw in case["null"]]
    p = _abi.osc_params(n, null=nulls, **case["osc"])
    u, tr, acc = np.zeros((B, n)), np.zeros((B, n)), np.zeros((B, n))
    rc = hostsim.hs_osc(C.byref(cd), C.byref(p), 0, 0, hostsim.hs_frame_id(n, b"EE"), None, P(np.ascontiguousarray(q)),
                        P(np.ascontiguousarray(dq)), P(np.ascontiguousarray(target)), 6, None, 6, C.c_int64(B), P(u), P(tr),
                        P(acc), None)
    assert rc == 0
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(u - ref) / scale) < 1e-8


@pytest.mark.parametrize("n", [4, 7])
def test_sliding_and_inverse_kinematics_on_random_chains(hostsim, n):
    from oracle import ik_oracle

    desc = random_chain(n, True, 900 + n)
    cd = _abi.chain_desc_from_dict(desc)
    rng = np.random.default_rng(n + 5)
    B = 16
    q, dq = np.ascontiguousarray(rng.uniform(0, 2 * np.pi, (B, n))), np.ascontiguousarray(rng.uniform(-2, 2, (B, n)))
    tg, tv = np.ascontiguousarray(rng.uniform(-0.5, 0.5, (B, 3))), np.ascontiguousarray(rng.uniform(-0.3, 0.3, (B, 3)))
    cs = dict(arm=desc, ctrl=dict(kd=40.0, lamb=12.0), tv=True)
    ref, ref_s = oo.run_sliding_case(cs, q, dq, tg, tv, None)
    u, s = np.zeros((B, n)), np.zeros((B, n))
    assert hostsim.hs_sliding(C.byref(cd), 0, 0, C.c_double(40.0), C.c_double(12.0), 1, hostsim.hs_frame_id(n, b"EE"), None,
                              P(q), P(dq), P(tg), P(tv), None, C.c_int64(B), P(u), P(s)) == 0
    assert np.max(np.abs(u - ref) / np.abs(ref).max(axis=1, keepdims=True)) < 1e-9
    assert np.max(np.abs(s - ref_s)) < 1e-9 * np.abs(ref_s).max()
    targets = np.ascontiguousarray(np.hstack([rng.uniform(-0.4, 0.4, (B, 3)), rng.uniform(-3, 3, (B, 3))]))
    for method in (3, 2):
        case = dict(arm=desc, path=dict(n_timesteps=12, dt=0.05, method=method))
        rp, rv = ik_oracle.run_ik_case(case, q, targets)
        pp, vv = np.zeros((12, B, n)), np.zeros((12, B, n))
        assert hostsim.hs_ik(C.byref(cd), 0, 0, C.c_double(0.2), C.c_double(2 * np.pi), C.c_double(np.pi), method,
                             C.c_double(0.05), 12, P(q), P(targets), C.c_int64(B), P(pp), P(vv)) == 0
        assert np.abs(pp.transpose(1, 0, 2) - rp).max() < 1e-8, method
