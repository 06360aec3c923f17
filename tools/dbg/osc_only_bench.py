import os,sys
sys.path.insert(0,'/root/repo')
import numpy as np, torch, bench
from abr_control_b200.arms import ur5
from abr_control_b200.controllers import OSC
dev=torch.device('cuda',0); B=65536
rc=ur5.Config(); c=OSC(rc, kp=10.0, ctrlr_dof=[True]*6, use_C=True)
S=[tuple(torch.as_tensor(a,device=dev) for a in bench.synth(B,6,100+s)) for s in range(8)]
u=torch.empty((B,6),dtype=torch.float64,device=dev)
for i in range(12): c.generate_into(S[i%8][0],S[i%8][1],S[i%8][2],u)
torch.cuda.synchronize()
