# scratch A/B driver for gpurun (tuning only; numbers quoted in DESIGN.md come from bench.py / tools/kbench.py runs)
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python - <<'PY'
import torch, time
for mb in (3, 9, 64):
    n = mb * 1024 * 1024 // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"pcie {name} {mb} MB: {dt*1e6:.0f} us  {mb*1.048576/dt/1e3:.1f} GB/s")
PY
echo "== kbench"; KB_MORE=1 python tools/kbench.py 2>&1 | tail -1
echo "== kbench B=1M"; KB_B=1048576 python tools/kbench.py 2>&1 | tail -1
echo "== bench"; python bench.py --steps 1000 --warmup 10 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 300 gpurun_out/bench_r1.err; cut -c1-200 gpurun_out/bench_r1.json
