#!/bin/bash
# Round-end capture on the GPU box (run through gpurun): parity tests, both bench arms, the kernel micro-bench, the
# ncu launch list of the bench command and one `ncu --set full` capture of each main kernel, condensed ON THE BOX
# (the .ncu-rep files exceed gpurun_out's merge-back limit).  tools/make_profiles.py then files them under profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 1000 --warmup 10 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 200 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1_ref.json 2>/dev/null
KB_MORE=1 python tools/kbench.py 2>/dev/null | tail -1 > gpurun_out/kbench_r1.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/b_ncu_l.log 2>&1
for spec in "osc:osc_kernel:8:1:tools/dbg/osc_only_bench.py" "rbd_JMgC:rbd_kernel:2:1:tools/kbench.py" "rbd_JMg:rbd_kernel:210:1:tools/kbench.py"; do
  IFS=: read name pat skip cnt script <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o /tmp/prof_$name python $script > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep --json gpurun_out/ncu_$name.json > gpurun_out/ncu_$name.txt 2>/dev/null
done
ls -la gpurun_out | tail -12
