#!/usr/bin/env python
"""Kernel micro-bench for tuning experiments: CUDA-event time per launch of the main kernels (UR5, B=65536)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abr_control_b200.arms import ur5, jaco2
from abr_control_b200.controllers import OSC, Damping
import bench

dev = torch.device("cuda", 0)
B, n = int(os.environ.get("KB_B", 65536)), 6
res = {}
def sets(dtype, count=24, Bx=B):
    out = []
    for s in range(count):
        q, dq, tg = bench.synth(Bx, n, 100 + s, dtype)
        out.append(tuple(torch.as_tensor(a, device=dev) for a in (q, dq, tg)))
    return out
for name, dt in (("f64", np.float64), ("f32", np.float32)):
    tdt = torch.float64 if dt == np.float64 else torch.float32
    S = sets(dt)
    rc = ur5.Config()
    shp = dict(J=(B, 6, n), M=(B, n, n), g=(B, n), C=(B, n, n))
    for key, want in (("rbd_JMgC", ("J", "M", "g", "C")), ("rbd_JMg", ("J", "M", "g"))):
        # inputs AND outputs rotate over a ring larger than L2, so that every launch's outputs go to HBM
        per_set = B * (12 + sum(int(np.prod(shp[k][1:])) for k in want)) * (8 if dt == np.float64 else 4)
        ring = []
        for s_ in range(max(3, int(np.ceil(320e6 / per_set)))):
            q_, dq_, _ = bench.synth(B, n, 300 + s_, dt)
            ring.append((torch.as_tensor(q_, device=dev), torch.as_tensor(dq_, device=dev),
                         {k: torch.empty(shp[k], dtype=tdt, device=dev) for k in want}))
        res[f"{key}_{name}"] = bench.time_kernel(lambda s: rc.eval_into(s[0], s[1], s[2]), 200, torch, ring) * 1e6
        del ring
    import abr_control_b200._abi as _abi
    _orig = _abi.osc_params
    cfgs = [("osc6C", dict(kp=10.0, ctrlr_dof=[True] * 6, use_C=True), None), ("osc_xyz", dict(kp=10.0), None)]
    if os.environ.get("KB_MORE"):
        cfgs += [("osc6", dict(kp=10.0, ctrlr_dof=[True] * 6), None), ("osc_xyzC", dict(kp=10.0, use_C=True), None),
                 ("osc6C_noslow", dict(kp=10.0, ctrlr_dof=[True] * 6, use_C=True), 0.0),
                 ("osc6_alg1", dict(kp=10.0, ctrlr_dof=[True] * 6, orientation_algorithm=1), None),
                 ("osc5", dict(kp=10.0, ctrlr_dof=[True] * 5 + [False]), None)]
    for key, kw, thr in cfgs:
        _abi.osc_params = (lambda *a, **k: _orig(*a, **dict(k, mx_threshold=thr))) if thr is not None else _orig
        c = OSC(rc, **kw)
        c._native()
        _abi.osc_params = _orig
        u = torch.empty((B, n), dtype=tdt, device=dev)
        res[f"{key}_{name}"] = bench.time_kernel(lambda s: c.generate_into(s[0], s[1], s[2], u), 200, torch, S) * 1e6
rc3 = jaco2.Config()
c3 = OSC(rc3, kp=200, ctrlr_dof=[True] * 5 + [False], null_controllers=[Damping(rc3, kv=10)])
for name, dt in (("f64", np.float64), ("f32", np.float32)):
    tdt = torch.float64 if dt == np.float64 else torch.float32
    S3 = sets(dt, 8, 262144)
    u3 = torch.empty((262144, 6), dtype=tdt, device=dev)
    res[f"jaco2_cfg3_B262144_{name}"] = bench.time_kernel(lambda s: c3.generate_into(s[0], s[1], s[2], u3), 50, torch, S3) * 1e6
# BASELINE config 5 (per-GPU share): Jaco2 OSC x,y,z + vmax + AvoidObstacles + Damping, fp32, B = 131072
from abr_control_b200.controllers import AvoidObstacles
c5 = OSC(rc3, kp=200, vmax=[0.5, 0], ctrlr_dof=[True, True, True, False, False, False],
         null_controllers=[AvoidObstacles(rc3, obstacles=[[0.09596, -0.2661, 0.64204, 0.05]], threshold=0.2), Damping(rc3, kv=10)])
S5 = [tuple(t_[:131072].contiguous() for t_ in s) for s in S3]
u5 = torch.empty((131072, 6), dtype=torch.float32, device=dev)
res["jaco2_cfg5_B131072_f32"] = bench.time_kernel(lambda s: c5.generate_into(s[0], s[1], s[2], u5), 40, torch, S5) * 1e6
# BASELINE config 4: UR5 OSC(kp=10) rollout, 4096 trajectories x 128 steps (us per STEP)
c4 = OSC(ur5.Config(), kp=10.0)
q4, dq4, tg4 = (torch.as_tensor(a, device=dev) for a in bench.synth(4096, 6, 4242))
for _ in range(2):
    c4.rollout(q4, dq4 * 0.1, tg4, steps=128, dt=1e-3, record=())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    c4.rollout(q4, dq4 * 0.1, tg4, steps=128, dt=1e-3, record=())
e1.record()
torch.cuda.synchronize()
res["rollout4096_us_per_step"] = e0.elapsed_time(e1) * 1e3 / 3 / 128
if os.environ.get("KB_MORE"):
    # the same rollouts with the `_Mx` threshold at 0 (no step takes the pseudo-inverse route: the evaluation + plant alone)
    _abi.osc_params = lambda *a, **k: _orig(*a, **dict(k, mx_threshold=0.0))
    c4n = OSC(ur5.Config(), kp=10.0)
    c4n._native()
    _abi.osc_params = _orig
    for _ in range(2):
        c4n.rollout(q4, dq4 * 0.1, tg4, steps=128, dt=1e-3, record=())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        c4n.rollout(q4, dq4 * 0.1, tg4, steps=128, dt=1e-3, record=())
    e1.record()
    torch.cuda.synchronize()
    res["rollout4096_noslow_us_per_step"] = e0.elapsed_time(e1) * 1e3 / 3 / 128
print(os.environ.get("ABRB_LIBRARY", "default"), json.dumps({k: round(v, 1) for k, v in res.items()}))
if os.environ.get("KB_CLOCKS"):
    # SM clock / throttle reasons over a sustained run of the heaviest kernel (is a long FP64 burst power capped?)
    S = sets(np.float64)
    c = OSC(ur5.Config(), kp=10.0, ctrlr_dof=[True] * 6, use_C=True)
    u = torch.empty((B, n), dtype=torch.float64, device=dev)
    sampler = bench.ClockSampler(0, getattr(torch.cuda.get_device_properties(0), "uuid", None))
    dt = bench.time_kernel(lambda s: c.generate_into(s[0], s[1], s[2], u), int(os.environ.get("KB_CLOCKS")), torch, S)
    print("clocks", json.dumps(dict(us_per_launch=round(dt * 1e6, 1), **sampler.stop())))
