"""PCIe rates the host-operand pipeline (cosma_b200/csrc/host_gemm.cu) is designed around: pinned H2D contiguous, H2D as
the 2-D row-chunk copy of A (768 rows x 8 B wide, pitch m x 8 B), D2H contiguous, and H2D + D2H together."""
import json
import sys

import torch


def rate(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-9


def main():
    m = k = 16384
    h = torch.empty(m * k, dtype=torch.float64).pin_memory()
    d = torch.empty(m * k, dtype=torch.float64, device="cuda")
    h2 = torch.empty(m * k // 4, dtype=torch.float64).pin_memory()
    d2 = torch.empty(m * k // 4, dtype=torch.float64, device="cuda")
    out = {}
    out["h2d_contig_gbs"] = rate(lambda: d.copy_(h, non_blocking=True), m * k * 8)
    out["d2h_contig_gbs"] = rate(lambda: h.copy_(d, non_blocking=True), m * k * 8)
    hv, dv = h.view(k, m), d.view(k, m)  # column-major m x k: row chunk = [:, i0:i0+768]

    def chunks():
        for i0 in range(0, m, 768):
            dv[:, i0:i0 + 768].copy_(hv[:, i0:i0 + 768], non_blocking=True)
    out["h2d_rowchunk768_gbs"] = rate(chunks, m * k * 8)
    s2 = torch.cuda.Stream()

    def both():
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)
    out["h2d_with_d2h_gbs"] = rate(both, m * k * 8)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
