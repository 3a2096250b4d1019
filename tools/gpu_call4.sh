#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_f32_gpu.py -x -q > gpurun_out/pytest_f32.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_f32.log
tail -30 gpurun_out/pytest_f32.log
timeout 300 python tools/gemm_f32_probe.py > gpurun_out/gemm_f32_probe.jsonl 2> gpurun_out/gemm_f32_probe.err; cat gpurun_out/gemm_f32_probe.jsonl; tail -3 gpurun_out/gemm_f32_probe.err
