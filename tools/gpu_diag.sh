#!/bin/bash
# 2-rank diagnosis of the C++ host layer: where does test_multiply stop, and does the NCCL build matter?
mkdir -p gpurun_out
export COSMA_B200_TRACE=ON NCCL_DEBUG=WARN
timeout 30 python -m cosma_b200.launch -np 2 --timeout 22 tests/cpp/bin/test_multiply > gpurun_out/diag_a.log 2>&1; echo "rc=$?" >> gpurun_out/diag_a.log
COSMA_B200_NCCL_LIB=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2 \
timeout 30 python -m cosma_b200.launch -np 2 --timeout 22 tests/cpp/bin/test_multiply > gpurun_out/diag_b.log 2>&1; echo "rc=$?" >> gpurun_out/diag_b.log
tail -25 gpurun_out/diag_a.log; echo ======; tail -25 gpurun_out/diag_b.log
