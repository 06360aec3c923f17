# scratch A/B driver for gpurun (tuning only; numbers quoted in DESIGN.md come from bench.py / tools/kbench.py runs)
for ch in 0 32768 22016 16384 8192; do echo "== chunk $ch"; ABRB_HOST_CHUNK=$ch python tools/dbg/e2e_probe.py 2>&1 | tail -2; done
