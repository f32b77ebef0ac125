#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of our kernels.
# usage: tools/gpu_round.sh <tag> [skip-ncu]
set -u
tag=${1:-r}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest.log
tail -3 $out/pytest.log
python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json | head -c 1500; echo
if [ "${2:-}" != "skip-ncu" ]; then
  K='regex:(attn_|aug_|pack_|pixel_gemm|wgrad_|nchw_to|f32_|splitk_|simt_|bce_|rel_bwd|out_bwd_patch|out_w_reduce)'
  python tools/one_step.py --steps 2 > $out/one_step.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file $out/launches.csv python tools/one_step.py --steps 2 > $out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  python tools/one_step.py --steps 2 > $out/one_step.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$K" -s ${NCU_SKIP:-19} -c ${NCU_COUNT:-19} -f -o $out/prof python tools/one_step.py --steps 2 > $out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ls -la $out
fi
