#!/bin/bash
# compute-sanitizer on one AAConv2d forward + backward (bf16 path) at small instances of the hot shapes.  ONE tool per gpurun call
# (B200_PROFILING.md): usage  tools/sanitize.sh <racecheck|synccheck|memcheck> <outdir>
set -u
tool=$1; out=$2
mkdir -p $out
for cfg in "T1 1" "T1_512 1" "T2 1"; do
  set -- $cfg
  # plain run first: never put a faulting program under the tool
  python tools/one_step.py --shape $1 --batch $2 --steps 1 > $out/plain_$1.log 2>&1 || { echo "plain run failed for $1"; continue; }
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/one_step.py --shape $1 --batch $2 --steps 1 > $out/${tool}_$1.log 2>&1
  echo "$tool $1 B=$2 rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|error" $out/${tool}_$1.log | tail -3
done
