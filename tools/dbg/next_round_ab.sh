#!/bin/bash
# A/B of the prepared (opt-in) variants against the shipped default, N = 6 only.
#   here (CPU):   bash tools/dbg/next_round_ab.sh build      -> abr_control_b200/libabrb_{fastdiv,cert2,both}.so
#   on the GPU:   gpurun -- 'bash tools/dbg/next_round_ab.sh run'
# Remove the libabrb_*.so copies afterwards (they travel with every gpurun snapshot, ~30 MB each).
set -e
cd "$(dirname "$0")/../.."
if [ "$1" = build ]; then
  bash tools/build_variant.sh fastdiv "-DABRB_FAST_DIV=1" &
  bash tools/build_variant.sh cert2 "-DABRB_CERT_PRECHECK=2" &
  bash tools/build_variant.sh both "-DABRB_FAST_DIV=1 -DABRB_CERT_PRECHECK=2" &
  wait
  for v in fastdiv cert2 both; do cp abr_control_b200/variants/libabrb_$v.so abr_control_b200/libabrb_$v.so; done
else
  for v in "" fastdiv cert2 both; do
    echo "== variant ${v:-default}"
    if [ -n "$v" ]; then export ABRB_LIBRARY=$PWD/abr_control_b200/libabrb_$v.so; else unset ABRB_LIBRARY; fi
    KB_MORE=1 python tools/kbench.py 2>&1 | tail -1
    [ -n "$v" ] && python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ur5 and not threejoint" 2>&1 | tail -1
  done
fi
