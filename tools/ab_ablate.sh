#!/bin/bash
# ablation of aug_build_tc (AACONV_AB_DBG bits: 1 no stores, 2 no skew, 4 no staging writes, 8 no k loads)
out=gpurun_out/${1:-abl}; mkdir -p $out
for m in ${MODES:-0 1 2 4 8 15}; do
  AACONV_AB_DBG=$m timeout 120 python bench.py --steps 5 --warmup 3 --no-model --no-shapes 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dbg', $m, {k: round(v['ms_total_per_step'] * 1e3, 1) for k, v in j['kernels'].items() if 'aug_build' in k})" | tee -a $out/ablate.txt
done
