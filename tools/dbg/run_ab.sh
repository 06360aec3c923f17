# scratch A/B driver for gpurun (tuning only; numbers quoted in DESIGN.md come from bench.py / tools/kbench.py runs)
nvidia-smi topo -m 2>/dev/null | head -20
for loc in "" 1 "" 1; do echo "== PROBE_LOCAL=$loc"; PROBE_LOCAL=$loc python tools/dbg/e2e_probe.py 2>&1 | tail -8; done
python - <<'PY'
import time, numpy as np, torch
from abr_control_b200.arms import ur5
from abr_control_b200.controllers.path_planners import InverseKinematics
rc = ur5.Config(); ik = InverseKinematics(rc)
rng = np.random.default_rng(0)
for B in (4096, 65536):
    q = torch.as_tensor(rng.uniform(0, 6, (B, 6)), device="cuda"); tg = torch.as_tensor(np.hstack([rng.uniform(-.4, .4, (B, 3)) + [0, 0, .45], rng.uniform(-3, 3, (B, 3))]), device="cuda")
    for m in (3, 2, 1):
        ik.generate_path(q, tg, n_timesteps=50, dt=0.05, method=m); torch.cuda.synchronize(); t0 = time.perf_counter()
        ik.generate_path(q, tg, n_timesteps=200, dt=0.05, method=m); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"ik B={B} method {m}: {dt*1e3:.2f} ms for 200 steps -> {B*200/dt/1e6:.1f} M IK steps/s")
PY
