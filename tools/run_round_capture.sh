#!/bin/bash
# Round-end capture on the GPU box (run through gpurun): parity tests, both bench arms, the kernel micro-bench, the
# ncu launch list of the bench command and one `ncu --set full` capture of each main kernel, condensed ON THE BOX
# (the .ncu-rep files exceed gpurun_out's merge-back limit).  tools/make_profiles.py then files them under profiles/.
#   usage: bash tools/run_round_capture.sh <tag>        e.g. r02
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 1000 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 300 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_ref.json 2>/dev/null
KB_MORE=1 python tools/kbench.py 2>/dev/null | tail -1 > gpurun_out/kbench_${TAG}.txt
ABRB_BENCH_QUICK=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/b_ncu_l.log 2>&1
cap() {  # name  kernel-regex  launches-to-skip  command...
  local name=$1 pat=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o /tmp/prof_$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep --json gpurun_out/ncu_${TAG}_$name.json > gpurun_out/ncu_${TAG}_$name.txt 2>/dev/null
  python tools/ncu_source_hot.py /tmp/prof_$name.ncu-rep --top 40 > gpurun_out/ncu_${TAG}_${name}_source.txt 2>/dev/null
}
cap osc osc_kernel 8 python tools/dbg/osc_only_bench.py
cap osc_ur5_f32 osc_kernel 6 python tools/dbg/osc_cfg_bench.py ur5f32
cap osc_cfg3 osc_kernel 6 python tools/dbg/osc_cfg_bench.py cfg3
cap osc_cfg5 osc_kernel 6 python tools/dbg/osc_cfg_bench.py cfg5
cap rollout rollout_kernel 1 python tools/dbg/osc_cfg_bench.py rollout
cap rbd_JMg rbd_kernel 9 python tools/dbg/rbd_only_bench.py JMg 65536
cap rbd_JMgC rbd_kernel 7 python tools/dbg/rbd_only_bench.py JMgC 65536
cap rbd_JMg_B262144 rbd_kernel 5 python tools/dbg/rbd_only_bench.py JMg 262144
cap rbd_JMgC_B262144 rbd_kernel 5 python tools/dbg/rbd_only_bench.py JMgC 262144
ls -la gpurun_out | tail -30
