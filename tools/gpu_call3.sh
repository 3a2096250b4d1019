#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_costa_gpu.py -x -q > gpurun_out/pytest_costa.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_costa.log
tail -4 gpurun_out/pytest_costa.log
rm -f gpurun_out/relayout_sweep.jsonl
for lr in 4 8 16; do
  COSMA_B200_RELAYOUT_LR=$lr timeout 200 python tools/relayout_bench.py --n 8192 --reps 8 --tag "lr$lr" >> gpurun_out/relayout_sweep.jsonl 2>> gpurun_out/relayout_sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/relayout_sweep.jsonl'):
    d=json.loads(l); print(d['tag'], d['dtype'], d['case'], round(d['GBps_mean']), round(d['frac_mean'],3))
PY
