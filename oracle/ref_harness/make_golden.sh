#!/bin/bash
# TEST INFRASTRUCTURE (development container only). Regenerates tests/golden/*.npz by running the
# reference twice per arm (warm = codegen, eval = Cython path). See run_reference.py.
# usage: oracle/ref_harness/make_golden.sh [arm ...]      (default: all four arms, in parallel)
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
export HOME="${ABR_REF_HOME:-/tmp/abr_home}"
export PYTHONPATH=/root/reference
mkdir -p "$HOME" "$HERE/../_ref/logs"
ARMS=("$@"); [ ${#ARMS[@]} -eq 0 ] && ARMS=(twojoint threejoint ur5 jaco2)
for arm in "${ARMS[@]}"; do
  ( python -W ignore "$HERE/run_reference.py" "$arm" warm && \
    python -W ignore "$HERE/run_reference.py" "$arm" eval ) > "$HERE/../_ref/logs/golden_$arm.log" 2>&1 &
done
wait
tail -n 2 "$HERE"/../_ref/logs/golden_*.log
