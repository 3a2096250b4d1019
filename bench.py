#!/usr/bin/env python
"""bench.py -- the headline benchmark of cosma_b200 (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU COSMA (oracle/_ref)

Metric (BASELINE.json): GEMM TFLOP/s, device-timed with CUDA events, max over ranks.
A "step" is one cosma::multiply of the named workload on synthetic U[0,10) matrices (the reference
miniapp's fill, miniapp/cosma_miniapp.cpp:21-25), alpha = 1, beta = 0.
  N = 1 : BASELINE configs[1]  square dgemm m=n=k=16384 (strategy: empty -> one local GEMM)
  N > 1 : BASELINE configs[2]  square dgemm m=n=k=32768 partitioned by COSMA's strategy over N GPUs
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GEMM TFLOP/s (device-timed, max over ranks)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cosma_b200", choices=["cosma_b200", "reference"])
    ap.add_argument("--mnk", type=str, default="", help="override workload: m,n,k")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(args):
    if args.mnk:
        m, n, k = [int(x) for x in args.mnk.split(",")]
        name = "dgemm m=%d n=%d k=%d (override)" % (m, n, k)
    elif args.gpus == 1:
        m = n = k = 16384
        name = "square dgemm m=n=k=16384 on 1 B200 (BASELINE configs[1])"
    else:
        m = n = k = 32768
        name = "square dgemm m=n=k=32768 partitioned at %d B200 (BASELINE configs[2])" % args.gpus
    return name, m, n, k


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1]); power.append(float(f[2]))
            except Exception:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def fp64_peak():
    """FP64 roofline denominator. MEASURED_PEAKS.json (driver-written) carries no FP64 figure, so the
    denominator is this repo's own measurement on the same pool: profiles/FP64_PEAK.json (DMMA.8x8x4 issue-rate
    probe = 37.0 TFLOP/s; nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.24)."""
    try:
        with open(os.path.join(ROOT, "profiles", "FP64_PEAK.json")) as f:
            d = json.load(f)
        return float(d["fp64_tflops"]), "profiles/FP64_PEAK.json (measured DMMA.8x8x4 issue-rate probe, this pool)"
    except Exception:
        return 37.24, "nominal 148 SM x 64 FP64 FMA/clk x 1.965 GHz"


def reference_arm(args):
    """The reference's own CPU implementation of the path: cosma::multiply at P=1 (unmodified sources built into
    oracle/_ref with a single-rank MPI stand-in + OpenBLAS), all host threads, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name, m, n, k = workload(args)
    from oracle import oracle as orc
    if not orc.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcosma_ref.so was not built (no /root/reference at build time)"}))
        return 0
    cores = os.cpu_count() or 1
    # bounded sample: same m, n; k cut so one step is ~2 TFLOP of CPU work (a few seconds on 16 cores)
    ks = min(k, max(256, int(2.2e12 / (2.0 * m * n)) // 256 * 256))
    reps = args.warmup + args.steps
    if args.gpus > 1 and os.path.exists(orc.REF_MINIAPP):
        return reference_arm_ranks(args, name, m, n, k, ks, reps, cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    R = orc.ref()
    R.ref_set_blas_threads(ctypes.c_int(cores))
    times = (ctypes.c_double * reps)()
    cs = ctypes.c_double()
    rc = R.ref_multiply_time_d(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(ks), ctypes.c_int(reps), times, ctypes.byref(cs))
    if rc != 0:
        print(json.dumps({"impl": "reference", "unavailable": "reference multiply threw"}))
        return 0
    timed = list(times)[args.warmup:]
    ms = sum(timed) / len(timed)
    tf = 2.0 * m * n * ks / (ms * 1e-3) * 1e-12
    sample = "reference cosma::multiply P=1 (OpenBLAS 0.3.30, %d threads) on m=%d n=%d k=%d (k cut from %d)" % (cores, m, n, ks, k)
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "m": m, "n": n, "k": k, "sample": sample},
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def reference_arm_ranks(args, name, m, n, k, ks, reps, cores):
    """N > 1: the UNMODIFIED reference miniapp (miniapp/cosma_miniapp.cpp built into oracle/_ref/cosma_miniapp_ref) on N ranks
    -- one per GPU of our arm -- started by oracle/minirun.py over the minimpi stand-in (processes + unix sockets), with
    cores/N OpenBLAS threads each: the reference's own distributed CPU path (Strategy -> Mapper -> allgather / reduce ->
    OpenBLAS dgemm) on the same bounded sample as the N = 1 arm."""
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from minirun import launch
    R = args.gpus
    threads = max(1, cores // R)
    argv = [orc.REF_MINIAPP, "-m", str(m), "-n", str(n), "-k", str(ks), "-r", str(reps)]
    code, outs = launch(R, argv, threads=threads, stdout=subprocess.PIPE, timeout=1500)
    text = outs[0].decode() if outs else ""
    times = []
    for ln in text.splitlines():
        if ln.startswith("COSMA TIMES [ms] ="):
            times = [float(x) for x in ln.split("=")[1].split()]
    if code != 0 or not times:
        print(json.dumps({"impl": "reference", "unavailable": "reference miniapp on %d minimpi ranks failed (exit %s)" % (R, code)}))
        return 0
    steps = []
    for ln in text.splitlines():
        if ln.startswith("parallel") or ln.startswith("sequential"):
            steps.append(ln.strip())
    steps = steps[:max(1, len(steps) // max(1, text.count("Divisions strategy")))]  # printed once per repetition block
    timed = sorted(times)[:args.steps]  # the miniapp reports its repetitions sorted: keep the K fastest of W + K
    ms = sum(timed) / len(timed)
    tf = 2.0 * m * n * ks / (ms * 1e-3) * 1e-12
    sample = ("reference cosma_miniapp on %d minimpi ranks x %d OpenBLAS 0.3.30 threads, m=%d n=%d k=%d (k cut from %d), strategy [%s], "
              "mean of the %d fastest of %d repetitions" % (R, threads, m, n, ks, k, "; ".join(steps), len(timed), reps))
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "m": m, "n": n, "k": k, "sample": sample},
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": threads * R, "kind": "reference", "sample": sample},
            "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_baseline(m, n, k):
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    if orc.have_ref():
        R = orc.ref()
        R.ref_set_blas_threads(ctypes.c_int(cores))
        ks = min(k, max(256, int(2.2e12 / (2.0 * m * n)) // 256 * 256))
        times = (ctypes.c_double * 3)()
        rc = R.ref_multiply_time_d(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(ks), ctypes.c_int(3), times, None)
        if rc == 0:
            ms = min(list(times)[1:])
            return {"value": 2.0 * m * n * ks / (ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "cores": cores, "kind": "reference",
                    "sample": "reference cosma::multiply P=1 (oracle/_ref, OpenBLAS 0.3.30, %d threads), m=%d n=%d k=%d (k cut from %d), best of 2 after 1 warm-up" % (cores, m, n, ks, k)}
    # oracle port (naive triple loop restating local_multiply_cpu) on a small sample
    import numpy as np
    s = 1024
    A = np.random.rand(s * s); B = np.random.rand(s * s); C = np.zeros(s * s)
    t0 = time.time(); orc.gemm("N", "N", s, s, s, 1.0, A, s, B, s, 0.0, C, s); dt = time.time() - t0
    return {"value": 2.0 * s ** 3 / dt * 1e-12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
            "sample": "oracle/gemm_oracle.c naive triple loop (OpenMP over columns) on %d^3" % s}


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = None
    if world > 1:
        # several ranks share the host: keep each rank (and the pinned buffers it allocates) on the NUMA node of its GPU
        try:
            from cosma_b200 import affinity as _aff
            affinity = _aff.bind_to_gpu(local_rank)
        except Exception:
            affinity = None
        dist.init_process_group("nccl", device_id=dev)
    from cosma_b200 import _lib, gemm
    lib = _lib.load()  # raises if the CUDA library is missing: there is no CPU fallback

    name, m, n, k = workload(args)
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))

    if world == 1:
        strategy = ""
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        A = torch.rand(m * k, device=dev, dtype=torch.float64, generator=gen) * 10
        B = torch.rand(k * n, device=dev, dtype=torch.float64, generator=gen) * 10
        C = torch.full((m * n,), float("nan"), device=dev, dtype=torch.float64)

        def step():
            gemm.local_multiply(A, B, C, m, n, k, 1.0, 0.0)
            return 1
        flops_per_kernel = 2.0 * m * n * k
    else:
        from cosma_b200 import distributed
        job = distributed.MultiplyJob(m, n, k, world, rank, dev)
        strategy = job.strategy_string

        def step():
            return job.run()
        flops_per_kernel = job.flops_per_local_gemm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    t0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        launches += step()
        ev[i + 1].record()
    barrier()
    t1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = tt.item()
    ms_per_step = total_ms / args.steps
    value = 2.0 * m * n * k / (ms_per_step * 1e-3) * 1e-12

    # roofline of the dominant kernel (the DMMA GEMM): algorithmic flops per launch / average launch duration
    peak, peak_src = fp64_peak()
    collectives = None
    if world == 1:
        kern_ms = sum(step_ms) / len(step_ms)
    else:
        kern_ms = job.mean_gemm_ms()
        try:
            collectives = job.collectives()
        except Exception as e:  # a reporting extra must never cost the bench line
            collectives = {"error": str(e)}
    achieved = flops_per_kernel / (kern_ms * 1e-3) * 1e-12
    roofline = {"bound": "tensor", "kernel": "gemm_f64_sm100_kernel (FP64 DMMA.8x8x4 pipe)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "flops_per_launch": flops_per_kernel, "launch_ms": kern_ms}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            roofline["traffic"] = json.load(f).get("dgemm_%d" % m)
    except Exception:
        pass

    # end-to-end through the host-pointer C ABI (what a user of the reference's GPU path passes): pinned host
    # A, B in; C out; H2D and D2H inside the timed region
    e2e = None
    if world == 1 and not args.no_e2e:
        hA = torch.empty(m * k, dtype=torch.float64).pin_memory(); hA.copy_(A)
        hB = torch.empty(k * n, dtype=torch.float64).pin_memory(); hB.copy_(B)
        hC = torch.empty(m * n, dtype=torch.float64).pin_memory()
        one = (ctypes.c_double * 1)(1.0); zero = (ctypes.c_double * 1)(0.0)
        lib.cosma_b200_dgemm_host.argtypes = [ctypes.c_void_p] + [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 2 + [ctypes.c_int64] + \
            [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]

        def host_step():
            st = lib.cosma_b200_dgemm_host(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), m, n, k, one, hA.data_ptr(), m,
                                           hB.data_ptr(), k, zero, hC.data_ptr(), m)
            _lib.check(st, "cosma_b200_dgemm_host")
        host_step(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        nrep = max(2, min(args.steps, 3))
        e0.record()
        for _ in range(nrep):
            host_step()
        e1.record(); torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / nrep
        ok = bool(torch.equal(hC[:4096], C[:4096].cpu()))
        e2e = {"value": 2.0 * m * n * k / (e2e_ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": (m * k + k * n) * 8,
               "d2h_bytes_per_step": m * n * 8, "ms_per_step": e2e_ms, "api": "cosma_b200_dgemm_host (pinned host A,B -> C)",
               "matches_device_path": ok}
        lib.cosma_b200_release_workspace()
    elif world > 1 and not args.no_e2e:
        try:
            e2e = job.e2e(max(2, min(args.steps, 3)))
        except Exception as e:  # an error every rank sees (e.g. out of pinned memory) must not cost the device-resident line
            e2e = {"error": str(e)[:300]}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            del A, B, C
            cpu = cpu_baseline(m, n, k)
        line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": name, "m": m, "n": n, "k": k, "strategy": strategy, "alpha": 1, "beta": 0,
                           "l2": "inputs larger than L2 (A+B+C = %.1f GB per job vs 126 MB L2)" % (8e-9 * (m * k + k * n + m * n)),
                           "fp64_peak_per_gpu_tflops": peak, "frac_of_fp64_peak": value / (peak * world)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if affinity is not None:
            line["config"]["host_affinity_rank0"] = affinity
        if collectives is not None:
            line["collectives"] = collectives  # rank 0's allgather / reduce-scatter device time and bus bandwidth in the last timed step
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
