#!/usr/bin/env python
"""TEST INFRASTRUCTURE (development container only): run the REFERENCE and write golden fixtures.

Usage (one arm per process, twice — SURVEY.md S0.1/S0.2 and Appendix B):

    HOME=<cache home> PYTHONPATH=/root/reference python run_reference.py <arm> warm
    HOME=<cache home> PYTHONPATH=/root/reference python run_reference.py <arm> eval

The first ("warm") process makes the reference generate + compile its C/Cython functions (and, because
of the reference's first-run bug, itself evaluates through ``lambdify``); the second process finds the
cached ``.so`` files and therefore runs the reference's real Cython path.  The "eval" pass writes
``tests/golden/<arm>_rbd.npz`` and ``tests/golden/<arm>_osc.npz``.

Nothing from the reference is copied: the fixtures hold inputs and the numbers the reference returned.
Harness-side shims (they do not modify the reference): the NumPy-2 ``quaternion_from_matrix`` shim and
the "fp64 reference mode" proxy that neutralises the public API's float32 casts
(/root/reference/abr_control/arms/base_config.py:223,247,270,285,301,336).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "tests"))
import cases  # noqa: E402

import abr_control.arms.base_config as bc  # noqa: E402
from abr_control.utils import transformations as tf  # noqa: E402

_orig_qfm = tf.quaternion_from_matrix
tf.quaternion_from_matrix = lambda matrix, isprecise=False: _orig_qfm(
    np.asarray(matrix, dtype=np.float64), isprecise
)


class _NP64:
    """numpy proxy: ``np.array(x, dtype="float32")`` -> float64."""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def array(a, dtype=None, **kw):
        return np.array(a, dtype=(float if dtype == "float32" else dtype), **kw)


def set_mode(fp64):
    bc.np = _NP64() if fp64 else np


def build_null(rc, kind, kw):
    import copy

    from abr_control.controllers import AvoidJointLimits, AvoidObstacles, Damping, RestingConfig

    cls = {"Damping": Damping, "RestingConfig": RestingConfig, "AvoidObstacles": AvoidObstacles,
           "AvoidJointLimits": AvoidJointLimits}[kind]
    return cls(rc, **copy.deepcopy(kw))  # AvoidJointLimits shifts the lists it is given in place


def main(arm, phase):
    import importlib

    from abr_control.controllers import OSC

    spec = cases.ARMS[arm]
    n = spec["n"]
    rc = importlib.import_module(f"abr_control.arms.{arm}").Config()
    q, dq, target, tvel = cases.states(arm)
    N = q.shape[0] if phase == "eval" else 1
    xoff = np.array(cases.XOFF)
    out = {"q": q, "dq": dq, "xoff": xoff}
    t0 = time.time()

    def stack(f):
        return np.array([np.array(f(i), dtype=np.float64) for i in range(N)])

    set_mode(True)
    for fr in cases.frames(n):
        out[f"Tx_{fr}"] = stack(lambda i: rc.Tx(fr, q[i]))
        out[f"R_{fr}"] = stack(lambda i: rc.R(fr, q[i]))
        out[f"Tinv_{fr}"] = stack(lambda i: rc.T_inv(fr, q[i]))
        out[f"J_{fr}"] = stack(lambda i: rc.J(fr, q[i]))
        print(f"[{arm}] frame {fr} done {time.time()-t0:.0f}s", flush=True)
    for fr in ("EE", spec["mid"]):
        out[f"T_{fr}"] = stack(lambda i: rc.T(fr, q[i]))
        out[f"Txx_{fr}"] = stack(lambda i: rc.Tx(fr, q[i], x=xoff))
        out[f"Jx_{fr}"] = stack(lambda i: rc.J(fr, q[i], x=xoff))
        out[f"dJ_{fr}"] = stack(lambda i: rc.dJ(fr, q[i], dq[i]))
        out[f"quat_{fr}"] = stack(lambda i: rc.quaternion(fr, q[i]))
        print(f"[{arm}] extras {fr} done {time.time()-t0:.0f}s", flush=True)
    out["dJx_EE"] = stack(lambda i: rc.dJ("EE", q[i], dq[i], x=xoff))
    out["M"] = stack(lambda i: rc.M(q[i]))
    out["g"] = stack(lambda i: rc.g(q[i]))
    if spec["has_C"]:
        out["C"] = stack(lambda i: rc.C(q[i], dq[i]))
    print(f"[{arm}] rbd done {time.time()-t0:.0f}s", flush=True)

    osc_out = {"q": q, "dq": dq, "target": target, "target_velocity": tvel}
    for name, c in cases.OSC_CASES.items():
        if c["arm"] != arm:
            continue
        for fp64 in (True, False):
            set_mode(fp64)
            nulls = [build_null(rc, k, kw) for k, kw in c.get("null", [])] or None
            ctrlr = OSC(rc, null_controllers=nulls, **c["osc"])
            kw = {}
            if c.get("ref_frame"):
                kw["ref_frame"] = c["ref_frame"]
            if c.get("xyz_offset") is not None:
                kw["xyz_offset"] = np.array(c["xyz_offset"])
            us, ts = [], []
            for i in range(N):
                if c.get("tv"):
                    kw["target_velocity"] = tvel[i]
                u = ctrlr.generate(q[i], dq[i], target[i], **kw)
                us.append(np.array(u, dtype=np.float64))
                ts.append(np.array(ctrlr.training_signal, dtype=np.float64))
            tag = "u64" if fp64 else "u32"
            osc_out[f"{name}__{tag}"] = np.array(us)
            if fp64:
                osc_out[f"{name}__train64"] = np.array(ts)
        print(f"[{arm}] osc case {name} done {time.time()-t0:.0f}s", flush=True)
    set_mode(True)
    for name, c in cases.NULL_CASES.items():
        if c["arm"] != arm:
            continue
        ctrl = build_null(rc, *c["ctrl"])
        osc_out[f"{name}__null64"] = stack(lambda i: ctrl.generate(q[i], dq[i]))
        print(f"[{arm}] null case {name} done {time.time()-t0:.0f}s", flush=True)

    tq, tdq = cases.joint_targets(arm)
    osc_out["joint_target"], osc_out["joint_target_velocity"] = tq, tdq
    from abr_control.controllers import Floating, Joint

    for name, c in cases.CTRL_CASES.items():
        if c["arm"] != arm:
            continue
        kind, kw = c["ctrl"]
        if kind == "Joint":
            ctrl = Joint(rc, **kw)
            osc_out[f"{name}__ctrl64"] = stack(
                lambda i: ctrl.generate(q[i], dq[i], tq[i], tdq[i] if c.get("tv") else None))
        else:
            ctrl = Floating(rc, **kw)
            osc_out[f"{name}__ctrl64"] = stack(lambda i: ctrl.generate(q[i], dq[i]))
        print(f"[{arm}] ctrl case {name} done {time.time()-t0:.0f}s", flush=True)

    from abr_control.controllers import Sliding

    for name, c in cases.SLIDING_CASES.items():
        if c["arm"] != arm:
            continue
        tgt, tv, ta = cases.sliding_inputs(c)
        ctrl = Sliding(rc, **c["ctrl"])
        kw = {}
        if c.get("ref_frame"):
            kw["ref_frame"] = c["ref_frame"]
        if c.get("offset") is not None:
            kw["offset"] = np.array(c["offset"])
        us, ss = [], []
        for i in range(N):
            u = ctrl.generate(q[i], dq[i], tgt[i], target_velocity=0 if tv is None else tv[i],
                              target_acc=0 if ta is None else ta[i], **kw)
            us.append(np.array(u, dtype=np.float64))
            ss.append(np.array(ctrl.s, dtype=np.float64))
        osc_out[f"{name}__sliding64"] = np.array(us)
        osc_out[f"{name}__s64"] = np.array(ss)
        print(f"[{arm}] sliding case {name} done {time.time()-t0:.0f}s", flush=True)

    # inverse-kinematics path planner (its module imports matplotlib for an optional plot; the image has none, so the
    # harness registers an empty stand-in — the plotting branch is never taken)
    ik_out = {}
    if any(c["arm"] == arm for c in cases.IK_CASES.values()):
        import types

        for mod in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
            sys.modules.setdefault(mod, types.ModuleType(mod))
        sys.modules["mpl_toolkits.mplot3d"].axes3d = None  # `from mpl_toolkits.mplot3d import axes3d` (path_planner.py)
        from abr_control.controllers.path_planners.inverse_kinematics import InverseKinematics

        n_ik = cases.N_IK if phase == "eval" else 1
        ik_targets = cases.ik_targets(arm)
        ik_out["position"], ik_out["target"] = q[: cases.N_IK], ik_targets
        for name, c in cases.IK_CASES.items():
            if c["arm"] != arm:
                continue
            planner = InverseKinematics(rc, **c.get("init", {}))
            pos, vel = [], []
            for i in range(n_ik):
                p_path, v_path = planner.generate_path(np.copy(q[i]), ik_targets[i], plot=False, **c["path"])
                pos.append(np.array(p_path, dtype=np.float64))
                vel.append(np.array(v_path, dtype=np.float64))
            ik_out[f"{name}__pos64"], ik_out[f"{name}__vel64"] = np.array(pos), np.array(vel)
            print(f"[{arm}] ik case {name} done {time.time()-t0:.0f}s", flush=True)

    # the reference's own pinned quantities for OSC helpers (controllers/tests/test_osc.py:19-59)
    if phase == "eval":
        gdir = os.path.join(REPO, "tests", "golden")
        os.makedirs(gdir, exist_ok=True)
        np.savez_compressed(os.path.join(gdir, f"{arm}_rbd.npz"), **out)
        np.savez_compressed(os.path.join(gdir, f"{arm}_osc.npz"), **osc_out)
        if ik_out:
            np.savez_compressed(os.path.join(gdir, f"{arm}_ik.npz"), **ik_out)
        loaded = [k for k in ("_M", "_g", "_C") if getattr(rc, k, None) is not None]
        kinds = {k: type(getattr(rc, k)).__name__ for k in loaded}
        print(f"[{arm}] wrote golden; function kinds: {kinds}")
    print(f"[{arm}] phase {phase} finished in {time.time()-t0:.0f}s")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
