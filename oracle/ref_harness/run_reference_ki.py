#!/usr/bin/env python
"""TEST INFRASTRUCTURE (development container only): golden fixture for the integral term of OSC.generate
(/root/reference/abr_control/controllers/osc.py:81-82, :262-264).

    HOME=/tmp/abr_home PYTHONPATH=/root/reference python oracle/ref_harness/run_reference_ki.py

Runs the REFERENCE's OSC with ki != 0 over SEQUENCES of calls — 6 independent streams of 12 consecutive states each, one
fresh controller per stream (the reference keeps one `integrated_error` per controller) — in the fp64 reference mode of
run_reference.py (float32 casts of the public arm API neutralised) and as shipped, on the warm Cython cache, and
writes tests/golden/ur5_osc_ki.npz: inputs, u of every call and the controller's integrated_error after every call.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as rr  # noqa: E402  (installs the NumPy-2 quaternion shim and the fp64-mode proxy)

from abr_control.arms import ur5  # noqa: E402
from abr_control.controllers import OSC, Damping  # noqa: E402

OSC_KW = dict(kp=30, ki=0.7, ctrlr_dof=[True] * 6, use_C=True)
T, S = 12, 6


def main():
    rc = ur5.Config()
    rng = np.random.default_rng(2024)
    q, dq = rng.uniform(0, 2 * np.pi, (T, S, 6)), rng.uniform(0, 2, (T, S, 6))
    target = rng.uniform(-1, 1, (T, S, 6))
    out = {"q": q, "dq": dq, "target": target, "ki": np.array(OSC_KW["ki"]), "kp": np.array(OSC_KW["kp"])}
    for fp64 in (True, False):
        rr.set_mode(fp64)
        us, ies = np.zeros((T, S, 6)), np.zeros((T, S, 6))
        for s in range(S):
            ctrlr = OSC(rc, null_controllers=[Damping(rc, kv=10)], **OSC_KW)
            for t in range(T):
                us[t, s] = ctrlr.generate(q[t, s], dq[t, s], target[t, s])
                ies[t, s] = ctrlr.integrated_error
        tag = "64" if fp64 else "32"
        out[f"u{tag}"], out[f"integrated_error{tag}"] = us, ies
    kinds = type(rc._M).__name__
    np.savez_compressed(os.path.join(rr.REPO, "tests", "golden", "ur5_osc_ki.npz"), **out)
    print("wrote tests/golden/ur5_osc_ki.npz; M is a", kinds)


if __name__ == "__main__":
    main()
