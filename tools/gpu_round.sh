#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of our kernels (T1 step).
# usage: tools/gpu_round.sh <tag> [skip-ncu]
set -u
tag=${1:-r}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest.log
tail -3 $out/pytest.log
python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
head -c 600 $out/bench.json; echo
if [ "${2:-}" != "skip-ncu" ]; then
  K='regex:(attn_|aug_|pack_|pixel_gemm|wgrad_|in_stats|in_relu|out_proj|out_bwd|out_w_reduce|rel_bwd|simt_|bce_)'
  python tools/one_step.py --steps 2 > $out/one_step.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file $out/launches.csv python tools/one_step.py --steps 2 > $out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  python tools/one_step.py --steps 2 > $out/one_step.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$K" -s ${NCU_SKIP:-17} -c ${NCU_COUNT:-17} -f -o $out/prof python tools/one_step.py --steps 2 > $out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ls -la $out
fi
