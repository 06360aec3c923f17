"""ncu target: rbd_kernel over a ring of input AND output sets larger than L2 (steady state: every launch's outputs
go to HBM).  usage: python tools/dbg/rbd_only_bench.py JMg|JMgC [B] [f32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, bench
from abr_control_b200.arms import ur5
want = ("J", "M", "g", "C") if (len(sys.argv) > 1 and sys.argv[1] == "JMgC") else ("J", "M", "g")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
f32 = len(sys.argv) > 3 and sys.argv[3] == "f32"
dt, ndt = (torch.float32, np.float32) if f32 else (torch.float64, np.float64)
dev = torch.device("cuda", 0)
rc = ur5.Config()
shp = dict(J=(B, 6, 6), M=(B, 6, 6), g=(B, 6), C=(B, 6, 6))
per_set = B * (12 + sum(int(np.prod(shp[k][1:])) for k in want)) * (4 if f32 else 8)
count = max(3, int(np.ceil(320e6 / per_set)))
ring = []
for s in range(count):
    q, dq, _ = bench.synth(B, 6, 5000 + s, ndt)
    ring.append((torch.as_tensor(q, device=dev), torch.as_tensor(dq, device=dev), {k: torch.empty(shp[k], dtype=dt, device=dev) for k in want}))
for i in range(2 * count + 4):
    s = ring[i % count]
    rc.eval_into(s[0], s[1], s[2])
torch.cuda.synchronize()
print("ring sets", count, "MB", count * per_set / 1e6)
