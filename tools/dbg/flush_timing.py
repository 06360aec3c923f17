"""Reads the per-warp cycle counters an ABRB_DBG_TIMING build of osc_kernel writes over the training-signal buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, bench
from abr_control_b200.arms import ur5
from abr_control_b200.controllers import OSC
dev = torch.device("cuda", 0); B = 65536
for dt in (torch.float64, torch.float32):
    rc = ur5.Config(); c = OSC(rc, kp=10.0, ctrlr_dof=[True] * 6, use_C=True)
    S = [tuple(torch.as_tensor(a, device=dev).to(dt) for a in bench.synth(B, 6, 100 + s)) for s in range(8)]
    u = torch.empty((B, 6), dtype=dt, device=dev); tr = torch.zeros((B, 6), dtype=dt, device=dev)
    for i in range(6):
        c.generate_into(S[i % 8][0], S[i % 8][1], S[i % 8][2], u, training_out=tr)
    torch.cuda.synchronize()
    d = tr.cpu().numpy().reshape(-1, 6)
    grid = int(d[0, 5]); warps = int(os.environ.get("BLK", "64")) // 32
    d = d[: grid * warps]
    tot, fl, wt, nf, rec = d[:, 0], d[:, 1], d[:, 2], d[:, 3], d[:, 4]
    for lo_, hi_ in ((0, 4), (5, 8), (9, 12), (13, 16), (17, 99)):
        m = (rec >= lo_) & (rec <= hi_) & (nf > 0)
        if m.any():
            print("   records %2d-%2d: CTAs(warps) %4d  flush cycles mean %.0f p90 %.0f max %.0f | total mean %.0f max %.0f"
                  % (lo_, hi_, m.sum(), fl[m].mean(), np.quantile(fl[m], 0.9), fl[m].max(), tot[m].mean(), tot[m].max()))
    print("   total cycles percentiles 50/90/99/100:", [int(np.quantile(tot, p)) for p in (0.5, 0.9, 0.99, 1.0)])
    print(dt, "grid", grid, "per-warp cycles: total mean %.0f max %.0f | in flush (incl. barriers) mean %.0f max %.0f | tile-end barrier wait mean %.0f | flushes/CTA %.2f records/flush %.1f"
          % (tot.mean(), tot.max(), fl.mean(), fl.max(), wt.mean(), nf.mean(), (rec.sum() / max(nf.sum(), 1))))
