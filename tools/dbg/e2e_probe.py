#!/usr/bin/env python
"""Where does the end-to-end time of a host-buffer OSC call go?  (tuning probe, GPU box only)

Times, for the bench workload (UR5 6-DOF, fp64, B = 65536): the raw PCIe copies, the library's *_host entry point
(OSC.generate on pinned NumPy buffers) and the same pipeline spelled out with torch streams, so that a gap between the
last two points at the library's host path rather than at the bus."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from abr_control_b200.arms import ur5
from abr_control_b200.controllers import OSC

B, n = 65536, 6
dev = torch.device("cuda", 0)
rc = ur5.Config()
c = OSC(rc, **bench.OSC_KW)
c.record_training_signal = False
hq, hdq, htg = (torch.as_tensor(a).pin_memory() for a in bench.synth(B, n, 77))
hu = torch.empty((B, n), dtype=torch.float64).pin_memory()
nq, ndq, ntg = hq.numpy(), hdq.numpy(), htg.numpy()
dq_, ddq_, dtg_ = (torch.empty_like(t, device=dev) for t in (hq, hdq, htg))
du = torch.empty((B, n), dtype=torch.float64, device=dev)

def timeit(fn, reps=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(reps // 5): fn()
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) / (reps // 5))
    return sorted(ts)[2] * 1e6, min(ts) * 1e6

def h2d():
    dq_.copy_(hq, non_blocking=True); ddq_.copy_(hdq, non_blocking=True); dtg_.copy_(htg, non_blocking=True); torch.cuda.synchronize()
def d2h():
    hu.copy_(du, non_blocking=True); torch.cuda.synchronize()
def kern():
    c.generate_into(dq_, ddq_, dtg_, du); torch.cuda.synchronize()
def torch_pipe():
    dq_.copy_(hq, non_blocking=True); ddq_.copy_(hdq, non_blocking=True); dtg_.copy_(htg, non_blocking=True)
    c.generate_into(dq_, ddq_, dtg_, du); hu.copy_(du, non_blocking=True); torch.cuda.synchronize()
def lib_host():
    c.generate(nq, ndq, ntg)
L = __import__("abr_control_b200._lib", fromlist=["lib"]).lib()
import ctypes as C
h = c._native(); fid = rc.frame_id("EE")
def lib_host_raw():  # the ctypes call alone, result into the pinned buffer
    L.abrb_osc_generate_host_f64(h, fid, None, nq.ctypes.data, ndq.ctypes.data, ntg.ctypes.data, 6, None, 0, hu.numpy().ctypes.data, None, B)
for name, fn in (("h2d 9.4MB (3 copies)+sync", h2d), ("d2h 3.1MB+sync", d2h), ("kernel+sync", kern),
                 ("torch pipeline", torch_pipe), ("lib *_host raw ctypes", lib_host_raw), ("OSC.generate(numpy)", lib_host)):
    med, best = timeit(fn)
    print(f"{name:32s} median {med:8.1f} us   best {best:8.1f} us   -> {B/med:8.1f} M evals/s")
