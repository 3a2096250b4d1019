#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_z_cpp_api.py -m gpu -q --durations=10 > gpurun_out/pytest_cpp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cpp.log
tail -40 gpurun_out/pytest_cpp.log
timeout 600 python -m pytest tests/test_costa_gpu.py -m gpu -q -x > gpurun_out/pytest_costa.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_costa.log
tail -5 gpurun_out/pytest_costa.log
