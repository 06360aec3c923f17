#!/bin/bash
# compute-sanitizer pass over a subset of the GPU parity tests (run through gpurun): memcheck (out-of-bounds /
# misaligned accesses, leaks) and racecheck (shared-memory hazards: the staging tiles with their __syncwarp pairs, the
# warp exchange area and the CTA queue of the cooperative pseudo-inverse).  Summaries land in gpurun_out/.
SEL="ragged or cooperative or unaligned or rollout_matches or ki_integrator_vs_reference or mjcf or singular or test_single_state"
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_${tool}_pytest.txt 2>&1
  echo "== $tool: $(tail -n 1 gpurun_out/sanitizer_${tool}_pytest.txt)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -20
done
