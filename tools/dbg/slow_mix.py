import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abr_control_b200.arms import ur5
from abr_control_b200.controllers import OSC
from oracle import rbd_oracle as ro
dev = torch.device('cuda', 0); B = 65536
rng = np.random.default_rng(5)
q = rng.uniform(0, 2*np.pi, (60000, 6)); ch = ro.ChainOracle('ur5')
J = ch.J('EE', q); S = J @ np.linalg.inv(ch.M(q)) @ np.swapaxes(J, 1, 2)
slow = np.abs(np.linalg.det(S)) < 1e-3
qs, qf = q[slow], q[~slow]
print('slow fraction', slow.mean(), len(qs))
rc = ur5.Config(); c = OSC(rc, kp=10.0, ctrlr_dof=[True]*6, use_C=True)
u = torch.empty((B, 6), dtype=torch.float64, device=dev)
def run(qsel, label):
    sets = []
    for s in range(8):
        idx = rng.integers(0, len(qsel), B)
        qq = qsel[idx]; dq = rng.uniform(0, 5, (B, 6)); tg = rng.uniform(-1, 1, (B, 6))
        sets.append(tuple(torch.as_tensor(a, device=dev) for a in (qq, dq, tg)))
    dt = bench.time_kernel(lambda s: c.generate_into(s[0], s[1], s[2], u), 100, torch, sets)
    print(label, round(dt*1e6, 1), 'us')
run(qf, 'all fast states')
run(qs, 'all slow states')
mix = np.concatenate([qf[:len(qf)], qs]); run(q, 'natural mix')
# slow states grouped: first 3.75% of each batch slow (dense warps) 
def run_grouped():
    sets = []
    ns = int(B*0.0375)//32*32
    for s in range(8):
        qq = np.concatenate([qs[rng.integers(0, len(qs), ns)], qf[rng.integers(0, len(qf), B-ns)]])
        dq = rng.uniform(0, 5, (B, 6)); tg = rng.uniform(-1, 1, (B, 6))
        sets.append(tuple(torch.as_tensor(a, device=dev) for a in (qq, dq, tg)))
    dt = bench.time_kernel(lambda s: c.generate_into(s[0], s[1], s[2], u), 100, torch, sets)
    print('grouped (slow states contiguous)', round(dt*1e6, 1), 'us')
run_grouped()
