#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from cosma_b200 import gemm
n = 8192
A = torch.randn(n * n, device="cuda"); B = torch.randn(n * n, device="cuda"); C = torch.empty(n * n, device="cuda")
for _ in range(3):
    gemm.gemm_raw("s", "N", "N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -o gpurun_out/r1_sgemm8192 python /tmp/one.py > gpurun_out/ncu_sgemm.log 2>&1
tail -3 gpurun_out/ncu_sgemm.log
timeout 600 python -m pytest tests/test_costa_gpu.py -q -m gpu -k "pxgemm" 2>&1 | tail -3
