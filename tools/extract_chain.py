#!/usr/bin/env python
"""Build-time tool: turn a reference arm ``Config`` into a flat numeric chain descriptor.

Runs ONLY in the development container (needs ``/root/reference`` on PYTHONPATH and SymPy).
Nothing here is imported by the product at run time; the output is the JSON model-data file
``abr_control_b200/arms/data/<arm>.json`` that ``abr_control_b200.chain.ChainDesc`` loads.

The reference builds every frame as a product of constant 4x4 factors and ``Rz(q_i)``
(e.g. /root/reference/abr_control/arms/ur5/config.py:301-339):

    link0      = L0                       (constant)
    joint_i    = link_i   . A_i           (constant A_i)
    link_{i+1} = joint_i  . Rz(q_i) . B_i (constant B_i)
    EE         = link_n   . E             (constant E, identity for UR5)

Instead of reading the per-arm attribute names (which differ between arms) the factors are
recovered from the symbolic frames themselves, in 50-digit arithmetic, and then checked:
each factor must be independent of q, and the re-assembled chain must reproduce every
``_calc_T(name)`` at random q.  Rounded to float64 this yields exactly the float64 constants the
reference's SymPy expressions carry (threejoint's float32 link lengths included).
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import sympy as sp

PREC = 50


def _num(Tsym, qsyms, qvals):
    # SymPy Floats carry 53-bit precision; make them exact rationals first so that the
    # 50-digit evaluation is not polluted by double rounding inside products
    Tsym = sp.Matrix(Tsym)
    Tsym = Tsym.xreplace({f: sp.Rational(f) for f in Tsym.atoms(sp.Float)})
    sub = {s: sp.Float(v, PREC) for s, v in zip(qsyms, qvals)}
    return Tsym.subs(sub).evalf(PREC)


def _rz(q):
    c, s = sp.cos(sp.Float(q, PREC)), sp.sin(sp.Float(q, PREC))
    return sp.Matrix([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def _clean(v):
    """Undo SymPy's 53-bit constant folding: if v is within 3e-15 (relative) of a <=12-significant-digit
    decimal, it *is* that source literal; otherwise (e.g. threejoint's float32 lengths) keep v."""
    v = float(v)
    if abs(v) < 1e-15:  # inverse/product residue of an exact zero
        return 0.0
    r = float(f"{v:.12g}")
    return r if abs(r - v) <= 3e-15 * max(abs(v), 1e-3) else v


def _to_np(M):
    return np.array([[_clean(M[i, j]) for j in range(M.shape[1])] for i in range(M.shape[0])])


def extract(arm_name):
    mod = importlib.import_module(f"abr_control.arms.{arm_name}")
    rc = mod.Config()
    n = rc.N_JOINTS
    if rc.N_LINKS != n + 1:
        raise SystemExit(f"{arm_name}: N_LINKS={rc.N_LINKS} != N_JOINTS+1; chain form not supported")

    def factors(qv):
        link = [_num(rc._calc_T(f"link{i}"), rc.q, qv) for i in range(n + 1)]
        joint = [_num(rc._calc_T(f"joint{i}"), rc.q, qv) for i in range(n)]
        ee = _num(rc._calc_T("EE"), rc.q, qv)
        L0 = link[0]
        A = [link[i].inv() * joint[i] for i in range(n)]
        B = [_rz(qv[i]).inv() * joint[i].inv() * link[i + 1] for i in range(n)]
        E = link[n].inv() * ee
        return L0, A, B, E

    rng = np.random.RandomState(1234)
    f0 = factors([0.0] * n)
    f1 = factors(list(rng.uniform(0, 2 * np.pi, n)))
    flat0 = [f0[0]] + f0[1] + f0[2] + [f0[3]]
    flat1 = [f1[0]] + f1[1] + f1[2] + [f1[3]]
    # SymPy folds float constants in 53-bit arithmetic while multiplying the symbolic factors, so the
    # recovered factors agree between two q only to ~1e-20 (not 1e-50); far below float64 rounding
    for a, b in zip(flat0, flat1):
        err = max(abs(x) for x in (a - b))
        assert err < 1e-15, f"chain factor depends on q (err {err}); arm does not fit the chain form"
    L0, A, B, E = f0
    for Mx in flat0:
        assert max(abs(Mx[3, j] - (1 if j == 3 else 0)) for j in range(4)) < 1e-40

    # inertia: first N_LINKS rows only (jaco2 carries a dead 8th row);
    # the reference supports arbitrary 6x6 but every shipped arm is diagonal with zero-mass joints
    def diag6(Mm):
        Mm = np.array(sp.Matrix(Mm).evalf(PREC).tolist(), dtype=float)
        assert np.allclose(Mm, np.diag(np.diag(Mm))), "non-diagonal link inertia not supported"
        return [float(v) for v in np.diag(Mm)]

    link_inertia = [diag6(rc._M_LINKS[i]) for i in range(rc.N_LINKS)]
    joint_inertia = [diag6(rc._M_JOINTS[i]) for i in range(n)]
    assert not np.any(np.array(joint_inertia)), "non-zero joint inertia not supported"

    # J_orientation[i] must be the z axis of joint i's pre-rotation frame (config.py "J_orientation")
    qv = list(rng.uniform(0, 2 * np.pi, n))
    for i in range(n):
        jo = _num(sp.Matrix(rc.J_orientation[i]), rc.q, qv)
        zi = _num(rc._calc_T(f"joint{i}"), rc.q, qv)[:3, 2]
        assert max(abs(x) for x in (jo - zi)) < 1e-30, "J_orientation is not the joint z axis"

    def m34(Mx):
        return _to_np(Mx)[:3, :].tolist()

    desc = {
        "name": arm_name,
        "n_joints": n,
        "n_links": rc.N_LINKS,
        "gravity": [float(v) for v in rc.gravity],
        "L0": m34(L0),
        "A": [m34(a) for a in A],
        "B": [m34(b) for b in B],
        "E": m34(E),
        "link_inertia": link_inertia,
        "start_angles": [float(v) for v in np.asarray(rc.START_ANGLES, dtype=float)],
        "L": np.asarray(rc.L, dtype=float).tolist(),
        "source": f"abr_control/arms/{arm_name}/config.py (factors recovered from _calc_T, 50-digit)",
    }
    return desc


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("arms", nargs="+")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "abr_control_b200", "arms", "data"))
    a = ap.parse_args()
    sys.path.insert(0, "/root/reference")
    for arm in a.arms:
        d = extract(arm)
        path = os.path.join(a.out, f"{arm}.json")
        with open(path, "w") as fh:
            json.dump(d, fh, indent=1)
        print("wrote", path)
