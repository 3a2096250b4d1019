#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_costa_gpu.py -x -q > gpurun_out/pytest_costa.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_costa.log
tail -25 gpurun_out/pytest_costa.log
timeout 300 python tools/relayout_bench.py --n 8192 --block 256 > gpurun_out/relayout_bench.jsonl 2> gpurun_out/relayout_bench.err
cat gpurun_out/relayout_bench.jsonl; tail -3 gpurun_out/relayout_bench.err
