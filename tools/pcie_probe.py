"""PCIe rates the host-operand pipeline (cosma_b200/csrc/host_gemm.cu) is designed around: pinned H2D contiguous, H2D as
the 2-D row-chunk copy of A (768 rows x 8 B wide, pitch m x 8 B), D2H contiguous, and H2D + D2H together.

Under torchrun (one rank per GPU) all ranks measure AT THE SAME TIME (barrier before every measurement) and every rank prints its
line: how much of the host link each rank keeps when 2 / 4 / 8 ranks share the host. --bind first moves each rank
to the CPUs local to its GPU (cosma_b200/affinity.py), so a run with and one without show what NUMA placement is worth:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py [--bind]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


BARRIER = [lambda: None]


def rate(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    BARRIER[0]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-9


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bound = None
    if "--bind" in sys.argv:
        from cosma_b200 import affinity
        bound = affinity.bind_to_gpu(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        BARRIER[0] = lambda: (dist.barrier(), torch.cuda.synchronize())
    m = k = 16384
    h = torch.empty(m * k, dtype=torch.float64).pin_memory()
    d = torch.empty(m * k, dtype=torch.float64, device="cuda")
    h2 = torch.empty(m * k // 4, dtype=torch.float64).pin_memory()
    d2 = torch.empty(m * k // 4, dtype=torch.float64, device="cuda")
    out = {"rank": rank, "ranks": world, "affinity": bound}
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
    print(json.dumps(out), flush=True)
    if world > 1:
        BARRIER[0]()
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
