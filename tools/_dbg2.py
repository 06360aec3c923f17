import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cases, torch
from abr_control_b200.arms import ur5
from abr_control_b200.controllers import OSC, Damping
from oracle import osc_oracle, rbd_oracle
rng = np.random.default_rng(0); B=257
q = rng.uniform(0, 2*np.pi, (B,6)); dq = rng.uniform(0,5,(B,6)); target = rng.uniform(-1,1,(B,6))
rc = ur5.Config()
for label,kw,nul in (("6dof+C+damp",dict(kp=50, ctrlr_dof=[True]*6, use_C=True),True),("6dof",dict(kp=50, ctrlr_dof=[True]*6),False)):
    ctrlr = OSC(rc, null_controllers=[Damping(rc,kv=10)] if nul else None, **kw)
    u = ctrlr.generate(q[:32],dq[:32],target[:32])
    case = dict(arm="ur5", osc=kw, null=[("Damping", dict(kv=10))] if nul else [])
    ref,_ = osc_oracle.run_case(case,q[:32],dq[:32],target[:32])
    err=(np.abs(u-ref)/np.abs(ref).max(axis=1,keepdims=True)).max(axis=1)
    ch=rbd_oracle.ChainOracle('ur5'); J=ch.J('EE',q[:32]); S=J@np.linalg.inv(ch.M(q[:32]))@np.swapaxes(J,1,2)
    print(label)
    for i in range(32):
        w=np.linalg.eigvalsh(S[i])
        if err[i]>1e-9: print(i, f'err {err[i]:.2e} det {np.linalg.det(S[i]):.2e} eig/lmax', np.array2string(w/w[-1],precision=2))
    rc32=ur5.Config(dtype=np.float32); c32=OSC(rc32, null_controllers=[Damping(rc32,kv=10)] if nul else None, **kw)
    u32=c32.generate(q[:32].astype(np.float32),dq[:32].astype(np.float32),target[:32].astype(np.float32))
    e32=(np.abs(u32-ref)/np.abs(ref).max(axis=1,keepdims=True)).max(axis=1)
    print(' fp32 errs >1e-3:', [(i,float(f'{e32[i]:.2e}')) for i in range(32) if e32[i]>1e-3])
