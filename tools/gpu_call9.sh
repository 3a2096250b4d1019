#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_z_cpp_api.py -m gpu -q --durations=10 > gpurun_out/pytest_cpp_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cpp_n.log
tail -40 gpurun_out/pytest_cpp_n.log
