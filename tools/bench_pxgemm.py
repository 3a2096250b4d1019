#!/usr/bin/env python
"""BASELINE configs[4]: pzgemm, 2D block-cyclic 256x256, m=n=k=16384, A conjugate-transposed (COSTA relayout), N GPUs.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_pxgemm.py --steps 5 --warmup 3

Device-resident local arrays for `value` (8mnk real flops / device time, max over ranks), pinned host arrays for `e2e`
(what a ScaLAPACK application passes). Reports the three phases (relayout in, multiply, relayout out) and the relayout
kernels' effective HBM rate. Inputs: complex U[0,1)^2, seed 1234+rank; C pre-filled with NaN (beta = 0 must not read it),
exactly SURVEY 8(d). One JSON line on rank 0."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--dtype", default="z")
    ap.add_argument("--transa", default="C")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cosma_b200 import costa
    from cosma_b200.distributed import init_comm
    comm = init_comm(dev)
    nprow, npcol = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
    grid = costa.Grid(comm, "R", nprow, npcol)
    m = n = k = args.n
    nb = args.block
    tdt = torch.complex128 if args.dtype == "z" else torch.float64
    eb = 16 if args.dtype == "z" else 8
    am, an = (k, m) if args.transa != "N" else (m, k)

    def local(rows, cols, fill=None):
        lr_ = costa.numroc(rows, nb, grid.myrow, 0, nprow); lc_ = costa.numroc(cols, nb, grid.mycol, 0, npcol)
        lld = max(lr_, 1)
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        if fill is None:
            t = torch.rand(lld * max(lc_, 1) * (2 if args.dtype == "z" else 1), device=dev, dtype=torch.float64, generator=gen)
            t = torch.view_as_complex(t.reshape(-1, 2)) if args.dtype == "z" else t
        else:
            t = torch.full((lld * max(lc_, 1),), fill, device=dev, dtype=tdt)
        return t, costa.descinit(rows, cols, nb, nb, 0, 0, lld)

    A, da = local(am, an); B, db = local(k, n); C, dc = local(m, n, float("nan"))

    def step(a=A, b=B, c=C):
        costa.pxgemm(grid, args.dtype, args.transa, "N", m, n, k, 1.0, a.data_ptr(), 1, 1, da, b.data_ptr(), 1, 1, db, 0.0, c.data_ptr(), 1, 1, dc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    stats = costa.last_layout_multiply_stats(comm)
    phases = torch.tensor([stats["ms_relayout_in"], stats["ms_multiply"], stats["ms_relayout_out"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(phases, op=dist.ReduceOp.MAX)
    finite = bool(torch.isfinite(torch.view_as_real(C) if args.dtype == "z" else C).all().item())
    flops = (8.0 if args.dtype == "z" else 2.0) * m * n * k
    e2e = None
    if not args.no_e2e:
        hA, hB, hC = (torch.empty(t.numel(), dtype=tdt).pin_memory() for t in (A, B, C))
        hA.copy_(A); hB.copy_(B)
        step(hA, hB, hC); barrier()
        e0.record()
        reps = max(2, min(args.steps, 3))
        for _ in range(reps):
            step(hA, hB, hC)
        e1.record(); barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": flops / (t.item() * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": t.item(), "h2d_bytes_per_step": (A.numel() + B.numel()) * eb,
               "d2h_bytes_per_step": C.numel() * eb, "api": "cosma_b200_p%sgemm with pinned host local arrays" % args.dtype,
               "matches_device_path": bool(torch.equal(hC[:4096], C[:4096].cpu()))}
    if rank == 0:
        peak = 37.0
        try:
            peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "FP64_PEAK.json")))["fp64_tflops"]
        except Exception:
            pass
        moved_in = (stats["in_local_elements"] + stats["in_remote_elements"]) * eb
        line = {"metric": "GEMM TFLOP/s (device-timed, max over ranks)", "value": flops / (ms.item() * 1e-3) * 1e-12, "unit": "TFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms.item(), "dtype": "c128" if args.dtype == "z" else "f64", "data": "synthetic",
                "config": {"workload": "p%sgemm 2D block-cyclic %dx%d, m=n=k=%d, transa=%s, grid %dx%d R (BASELINE configs[4])" % (args.dtype, nb, nb, n, args.transa, nprow, npcol),
                           "strategy": stats["strategy"], "frac_of_fp64_peak": flops / (ms.item() * 1e-3) * 1e-12 / (peak * world)},
                "phases_ms_max_over_ranks": {"relayout_in": phases[0].item(), "multiply": phases[1].item(), "relayout_out": phases[2].item()},
                "relayout": {"rank0_in_bytes": moved_in, "rank0_in_remote_fraction": stats["in_remote_elements"] / max(1, stats["in_local_elements"] + stats["in_remote_elements"]),
                             "rank0_in_effective_GBps": 2.0 * moved_in / (stats["ms_relayout_in"] * 1e-3) * 1e-9 if stats["ms_relayout_in"] > 0 else None},
                "result_finite": finite, "e2e": e2e, "gpu_launches": stats["launches"] * args.steps}
        print(json.dumps(line))
    grid.destroy(); comm.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
