#!/bin/bash
# All BASELINE.json layer shapes + the reference arm, one JSON file each under gpurun_out/<tag>/
tag=${1:-shapes}
out=gpurun_out/$tag
mkdir -p $out
for s in T1 T2 T3 T1_512; do
  python bench.py --shape $s --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_$s.json 2> $out/bench_$s.err; echo "$s rc=$?"
done
