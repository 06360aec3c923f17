"""ncu targets for the other OSC configurations.  usage: python tools/dbg/osc_cfg_bench.py cfg3|cfg5|rollout|ur5f32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, bench
from abr_control_b200.arms import jaco2, ur5
from abr_control_b200.controllers import OSC, AvoidObstacles, Damping
which = sys.argv[1]
dev = torch.device("cuda", 0)
if which == "rollout":  # BASELINE config 4
    c = OSC(ur5.Config(), kp=10.0)
    q, dq, tg = (torch.as_tensor(a, device=dev) for a in bench.synth(4096, 6, 4242))
    for _ in range(3):
        c.rollout(q, dq * 0.1, tg, steps=128, dt=1e-3, record=())
else:
    if which == "ur5f32":
        B, rc = 65536, ur5.Config()
        c = OSC(rc, kp=10.0, ctrlr_dof=[True] * 6, use_C=True)
    elif which == "cfg3":  # BASELINE config 3
        B, rc = 262144, jaco2.Config()
        c = OSC(rc, kp=200, ctrlr_dof=[True] * 5 + [False], null_controllers=[Damping(rc, kv=10)])
    else:  # BASELINE config 5, per-GPU share
        B, rc = 131072, jaco2.Config()
        c = OSC(rc, kp=200, vmax=[0.5, 0], ctrlr_dof=[True, True, True, False, False, False],
                null_controllers=[AvoidObstacles(rc, obstacles=[[0.09596, -0.2661, 0.64204, 0.05]], threshold=0.2), Damping(rc, kv=10)])
    S = [tuple(torch.as_tensor(a, device=dev) for a in bench.synth(B, 6, 100 + s, np.float32)) for s in range(6)]
    u = torch.empty((B, 6), dtype=torch.float32, device=dev)
    for i in range(10):
        c.generate_into(S[i % 6][0], S[i % 6][1], S[i % 6][2], u)
torch.cuda.synchronize()
