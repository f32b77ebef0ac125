#!/bin/bash
# One ncu --set full capture of selected kernels of one AAConv2d step + per-kernel summary and source-level stall tables.
# usage (on the GPU box, via gpurun): tools/ncu_capture.sh <outdir> <kernel regex> <skip> <count> [shape]
set -u
out=$1; regex=$2; skip=$3; count=$4; shape=${5:-T1}
mkdir -p $out
python tools/one_step.py --shape $shape --steps 2 > $out/one_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f -o $out/prof python tools/one_step.py --shape $shape --steps 2 > $out/ncu.log 2>&1
tail -2 $out/ncu.log
