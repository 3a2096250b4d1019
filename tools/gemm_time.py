#!/usr/bin/env python
"""One local GEMM through the C ABI (cosma_b200_?gemm), timed with CUDA events: any dtype, shape and leading-dimension padding.
Prints one JSON line: the path taken (1 = tensor-pipe / DMMA kernel, 2 = generic kernel), ms and TFLOP/s (2mnk, 8mnk complex)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cosma_b200 import _lib, gemm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="d", choices=list("sdcz"))
    ap.add_argument("--m", type=int, default=8192)
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--k", type=int, default=8192)
    ap.add_argument("--transa", default="N")
    ap.add_argument("--transb", default="N")
    ap.add_argument("--pad", type=int, default=0, help="added to every leading dimension")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    lib = _lib.load()
    tdt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[a.dtype]
    ar, ac = (a.m, a.k) if a.transa == "N" else (a.k, a.m)
    br, bc = (a.k, a.n) if a.transb == "N" else (a.n, a.k)
    lda, ldb, ldc = ar + a.pad, br + a.pad, a.m + a.pad
    mk = lambda ld, cols: torch.randn(ld * cols, device="cuda", dtype=torch.float32).to(tdt) if not tdt.is_complex else \
        torch.complex(torch.randn(ld * cols, device="cuda"), torch.randn(ld * cols, device="cuda")).to(tdt)
    A, B = mk(lda, ac), mk(ldb, bc)
    C = torch.zeros(ldc * a.n, device="cuda", dtype=tdt)
    run = lambda: gemm.gemm_raw(a.dtype, a.transa, a.transb, a.m, a.n, a.k, 1.0, A.data_ptr(), lda, B.data_ptr(), ldb, 0.0, C.data_ptr(), ldc)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    flops = (8.0 if a.dtype in "cz" else 2.0) * a.m * a.n * a.k
    best = min(ms)
    print(json.dumps({"dtype": a.dtype, "transa": a.transa, "transb": a.transb, "m": a.m, "n": a.n, "k": a.k, "lda": lda, "ldb": ldb, "ldc": ldc,
                      "path": lib.cosma_b200_last_gemm_path(), "repack": os.environ.get("COSMA_B200_REPACK_UNALIGNED", ""),
                      "ms_best": best, "ms_all": ms, "tflops": flops / (best * 1e-3) * 1e-12}))


if __name__ == "__main__":
    main()
