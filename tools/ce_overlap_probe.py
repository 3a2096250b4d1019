"""Probe for DESIGN.md 9 item 5 (needs 2 GPUs of one box; ~20 s): does a copy-engine peer copy over NVLink run UNDER the persistent
DGEMM without slowing it? Times (a) the 16384^3 DGEMM alone, (b) a 1 GiB cuda:0 -> cuda:1 copy alone, (c) both together, copy on a
side stream. If (c) ~= max(a, b) the allgather/reduce-scatter of COSMA's ring-2 steps can be hidden behind panel-split GEMMs with
cudaMemcpyPeerAsync; NCCL kernels cannot (they need SMs the persistent kernel holds).
    python tools/ce_overlap_probe.py > gpurun_out/ce_overlap.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosma_b200 import gemm  # noqa: E402


def main():
    if torch.cuda.device_count() < 2:
        print(json.dumps({"error": "needs 2 GPUs"}))
        return 0
    n = 16384
    torch.cuda.set_device(0)
    A = torch.rand(n * n, device="cuda:0", dtype=torch.float64)
    B = torch.rand(n * n, device="cuda:0", dtype=torch.float64)
    C = torch.empty(n * n, device="cuda:0", dtype=torch.float64)
    src = torch.empty(1 << 27, device="cuda:0", dtype=torch.float64)  # 1 GiB
    dst = torch.empty(1 << 27, device="cuda:1", dtype=torch.float64)
    side = torch.cuda.Stream(device=0)

    def run_gemm():
        gemm.gemm_raw("d", "N", "N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)

    def run_copy(reps):
        with torch.cuda.stream(side):
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)

    def timed(fn):
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        torch.cuda.current_stream().wait_stream(side)
        e1.record()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        return e0.elapsed_time(e1)

    for _ in range(2):
        run_gemm(); run_copy(1)
    out = {"gemm_ms": min(timed(run_gemm) for _ in range(3)), "copy_4GiB_ms": min(timed(lambda: run_copy(4)) for _ in range(3))}

    def both():
        run_copy(4)
        run_gemm()
    out["both_ms"] = min(timed(both) for _ in range(3))
    out["copy_GBps_alone"] = 4 * (1 << 30) / (out["copy_4GiB_ms"] * 1e-3) * 1e-9
    out["gemm_slowdown_under_copy"] = out["both_ms"] / max(out["gemm_ms"], out["copy_4GiB_ms"])
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
