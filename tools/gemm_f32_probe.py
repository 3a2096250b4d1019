#!/usr/bin/env python
"""Throughput of the 3xTF32 SGEMM/CGEMM kernels vs cuBLAS FP32 (torch.matmul with TF32 disabled), one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cosma_b200 import gemm  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for dtype, n in (("s", 4096), ("s", 8192), ("s", 16384), ("c", 4096), ("c", 8192)):
    tdt = torch.float32 if dtype == "s" else torch.complex64
    A = torch.randn(n * n, device="cuda", dtype=torch.float32).to(tdt) if dtype == "s" else torch.view_as_complex(torch.randn(n * n, 2, device="cuda"))
    B = torch.randn(n * n, device="cuda", dtype=torch.float32).to(tdt) if dtype == "s" else torch.view_as_complex(torch.randn(n * n, 2, device="cuda"))
    C = torch.empty(n * n, device="cuda", dtype=tdt)
    ms = timeit(lambda: gemm.gemm_raw(dtype, "N", "N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n))
    Am, Bm = A.view(n, n), B.view(n, n)
    ms_cublas = timeit(lambda: torch.matmul(Bm, Am))  # row-major (B^T A^T)^T == column-major A B
    ref = torch.matmul(Bm.to(torch.complex128 if dtype == "c" else torch.float64), Am.to(torch.complex128 if dtype == "c" else torch.float64))
    got = C.view(n, n)
    err = (torch.linalg.norm(got.to(ref.dtype) - ref) / torch.linalg.norm(ref)).item()
    err_cublas = (torch.linalg.norm(torch.matmul(Bm, Am).to(ref.dtype) - ref) / torch.linalg.norm(ref)).item()
    flops = (8.0 if dtype == "c" else 2.0) * n ** 3
    print(json.dumps({"kernel": "gemm_tf32x3_sm100_kernel", "dtype": dtype, "n": n, "ms": ms, "tflops": flops / ms * 1e-9, "cublas_fp32_ms": ms_cublas,
                      "cublas_fp32_tflops": flops / ms_cublas * 1e-9, "normwise_err_vs_fp64": err, "cublas_normwise_err_vs_fp64": err_cublas}), flush=True)
    del A, B, C, ref
