#!/usr/bin/env python
"""HBM roofline of the batched relayout kernel (R3/R4) on one GPU: a block-cyclic-shaped piece list (blocks x blocks
tiles of an n x n matrix) moved by ONE launch, as a plain copy and as a (conjugate-)transpose, per dtype.
Algorithmic bytes = 2 * elements * sizeof(T). Prints one JSON line per case."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cosma_b200 import costa  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--dtypes", default="zdcs")
    ap.add_argument("--cases", default="copy,transpose,conj_transpose")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    peak = None
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
    except Exception:
        pass
    n, nb = args.n, args.block
    for dtype in args.dtypes:
        eb = costa.ELEM_BYTES[dtype]
        src = torch.randn(n * n * eb // 4, device="cuda", dtype=torch.float32)
        dst = torch.empty_like(src)
        flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
        flush2 = torch.zeros(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
        for name, tr, cj in (("copy", 0, 0), ("transpose", 1, 0), ("conj_transpose", 1, 1)):
            if (cj and dtype in "sd") or name not in args.cases.split(","):
                continue
            # one n x n column-major matrix described as (n/nb)^2 block pieces: block (i, j) -> block (i, j) or (j, i)
            blocks = range(0, n, nb)
            rs = [0] + [min(b + nb, n) for b in blocks]
            owners = [[0] * (len(rs) - 1) for _ in range(len(rs) - 1)]
            # a block-grid layout over ONE allocation: block (i, j) starts at (i*nb) + (j*nb)*n
            fb = [(i, j, src.data_ptr() + (rs[i] + rs[j] * n) * eb, n) for i in range(len(rs) - 1) for j in range(len(rs) - 1)]
            tb = [(i, j, dst.data_ptr() + (rs[i] + rs[j] * n) * eb, n) for i in range(len(rs) - 1) for j in range(len(rs) - 1)]
            F = costa.custom_layout(rs, rs, owners, fb)
            Tl = costa.custom_layout(rs, rs, owners, tb)
            op = "N" if not tr else ("C" if cj else "T")
            tp = costa.TransformPlan(None, dtype, [(F, Tl, op, 1.0, 0.0)], rank=0, nranks=1)
            for _ in range(3):
                tp.run()
            torch.cuda.synchronize()
            ms = []
            for _ in range(args.reps):
                flush.zero_(); flush2.sum()  # evict L2 (write 256 MB), then read 256 MB so no dirty lines are left to
                # be written back during the timed kernel
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); tp.run(); e1.record(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            tp.destroy()
            best, mean = min(ms), sum(ms) / len(ms)
            gb = 2.0 * n * n * eb * 1e-9
            line = {"tag": args.tag, "kernel": "relayout_kernel", "case": name, "dtype": dtype, "n": n, "block": nb, "pieces": len(fb), "ms_mean": mean, "ms_best": best,
                    "GBps_mean": gb / (mean * 1e-3), "GBps_best": gb / (best * 1e-3), "peak_GBps": peak,
                    "frac_mean": gb / (mean * 1e-3) / peak if peak else None}
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
