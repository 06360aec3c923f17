# scratch A/B driver for gpurun (tuning only; numbers quoted in DESIGN.md come from bench.py / tools/kbench.py runs)
for v in "" exp3 exp4; do
  echo "== variant ${v:-default}"
  if [ -n "$v" ]; then export ABRB_LIBRARY=$PWD/abr_control_b200/libabrb_$v.so; else unset ABRB_LIBRARY; fi
  KB_MORE=1 python tools/kbench.py 2>&1 | tail -1
done
unset ABRB_LIBRARY
python tools/dbg/e2e_probe.py 2>&1 | tail -2
