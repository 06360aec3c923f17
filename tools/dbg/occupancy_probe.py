#!/usr/bin/env python
"""Is co-residency of two CTAs per SM worth anything for the fp64 OSC kernel?  Times UR5 6-DOF + use_C at batch sizes
that give 74 ... 592 tiles of 128 states (148 = one CTA per SM in one round, 296 = two per SM in one round, 512 = the
headline, 592 = two full rounds), with the `_Mx` threshold at 0 (no state takes the pseudo-inverse route: the
evaluation alone) and at its default."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from abr_control_b200 import _abi  # noqa: E402
from abr_control_b200.arms import ur5  # noqa: E402
from abr_control_b200.controllers import OSC  # noqa: E402

dev = torch.device("cuda:0")
rc = ur5.Config()
_orig = _abi.osc_params
res = {}
for name, thr in (("noslow", 0.0), ("full", None)):
    _abi.osc_params = (lambda *a, **k: _orig(*a, **dict(k, mx_threshold=thr))) if thr is not None else _orig
    c = OSC(rc, kp=10.0, ctrlr_dof=[True] * 6, use_C=True)
    c._native()
    _abi.osc_params = _orig
    for tiles in (74, 148, 222, 296, 444, 512, 592):
        B = tiles * 128
        S = [tuple(torch.as_tensor(a, device=dev) for a in bench.synth(B, 6, 100 + i)) for i in range(12)]
        u = torch.empty((B, 6), dtype=torch.float64, device=dev)
        res[f"{name}_{tiles}"] = round(bench.time_kernel(lambda s: c.generate_into(s[0], s[1], s[2], u), 100, torch, S) * 1e6, 2)
print(os.environ.get("ABRB_LIBRARY", "default"), json.dumps(res))
