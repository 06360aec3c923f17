#!/usr/bin/env python
"""A/B of the host-buffer path's knobs on ONE box (GPU box only): copy streams per chunk x chunk size, synchronous calls
and the two-slot asynchronous pipeline, same loop as bench.py's `e2e` (UR5 6-DOF fp64, B = 65536, pinned buffers)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from abr_control_b200.arms import ur5  # noqa: E402
from abr_control_b200.controllers import OSC  # noqa: E402

B, n = 65536, 6
c = OSC(ur5.Config(), **bench.OSC_KW)
c.record_training_signal = False
host = [tuple(torch.as_tensor(a).pin_memory().numpy() for a in bench.synth(B, n, 77 + s)) for s in range(4)]


def rates(per_block=40, blocks=5):
    for s in range(3):
        c.generate(*host[s % 4])
    st, pt = [], []
    for _ in range(blocks):
        t0 = time.perf_counter()
        for i in range(per_block):
            c.generate(*host[i % 4])
        st.append(time.perf_counter() - t0)
    for _ in range(blocks):
        t0 = time.perf_counter()
        pend = [None, None]
        for i in range(per_block):
            sl = i & 1
            if pend[sl] is not None:
                pend[sl].wait()
            pend[sl] = c.generate_async(*host[i % 4], slot=sl)
        for p_ in pend:
            if p_ is not None:
                p_.wait()
        pt.append(time.perf_counter() - t0)
    return round(B * per_block / float(np.median(st)) / 1e6, 1), round(B * per_block / float(np.median(pt)) / 1e6, 1)


res = {}
for rep in range(2):
    for streams in (1, 2, 3):
        for chunk in (0, 65536, 16384):
            c.set_option("host_upload_streams", streams)
            c.set_option("host_chunk_states", chunk)
            res[f"rep{rep}_streams{streams}_chunk{chunk or 'auto'}"] = rates()
print(json.dumps(res))
