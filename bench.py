#!/usr/bin/env python
"""bench.py -- the headline benchmark of cosma_b200 (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU COSMA (oracle/_ref)

Metric (BASELINE.json): GEMM TFLOP/s, device-timed with CUDA events, max over ranks.
A "step" is one cosma::multiply of the named workload on synthetic U[0,10) matrices (the reference miniapp's fill,
miniapp/cosma_miniapp.cpp:21-25), alpha = 1, beta = 0, in COSMA's native layout, partitioned over the N GPUs by COSMA's strategy.

  --workload cfg3 (default, every N): square dgemm m=n=k=32768 -- the shape north_star's targets are quoted on (>= 80 % of FP64 peak at
                                      1 GPU, >= 70 % aggregate at 8) and BASELINE configs[2] partitions at 2/4/8 GPUs; 25.8 GB, fits one GPU
             cfg2                   : square dgemm m=n=k=16384 (BASELINE configs[1])
             largek                 : dgemm m=n=8192, k=1048576 (BASELINE configs[3], pk8 at 8 GPUs)
             pzgemm                 : ScaLAPACK pzgemm, 256x256 block-cyclic, 16384^3, A conjugate-transposed (BASELINE configs[4])
             sgemm | cgemm | zgemm  : 16384^3 in the other three types (3xTF32 tcgen05 kernels / ZGEMM)
The default line also carries, under "also", short runs of the other named configs that fit the job (N = 1: cfg2; N = 8: largek and
pzgemm), so that the driver's records show them. Prints ONE JSON line on rank 0.
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

WORKLOADS = {
    "cfg3": (32768, 32768, 32768, "d", "square dgemm m=n=k=32768 partitioned over %d B200 by COSMA's strategy (BASELINE configs[2]; north_star's target shape)"),
    "cfg2": (16384, 16384, 16384, "d", "square dgemm m=n=k=16384 on %d B200 (BASELINE configs[1])"),
    "largek": (8192, 8192, 1048576, "d", "large-K dgemm m=n=8192 k=1048576 on %d B200 (BASELINE configs[3])"),
    "sgemm": (16384, 16384, 16384, "s", "square sgemm m=n=k=16384 on %d B200 (3xTF32 tcgen05 kernel)"),
    "cgemm": (16384, 16384, 16384, "c", "square cgemm m=n=k=16384 on %d B200 (3xTF32 tcgen05 kernel)"),
    "zgemm": (16384, 16384, 16384, "z", "square zgemm m=n=k=16384 on %d B200 (DMMA kernel, complex embedding)"),
}
DTYPE_NAME = {"d": "f64", "z": "c128", "s": "f32 (3xTF32)", "c": "c64 (3xTF32)"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cosma_b200", choices=["cosma_b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["pzgemm"])
    ap.add_argument("--mnk", type=str, default="", help="override workload: m,n,k")
    ap.add_argument("--strategy", type=str, default="", help="explicit COSMA strategy, e.g. pm2,pn2,pk2 (default: automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other named configs")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def workload(args):
    m, n, k, dtype, name = WORKLOADS[args.workload if args.workload in WORKLOADS else "cfg3"]
    name = name % args.gpus
    if args.mnk:
        m, n, k = [int(x) for x in args.mnk.split(",")]
        name = "%sgemm m=%d n=%d k=%d (override)" % (dtype, m, n, k)
    return name, m, n, k, dtype


def workload_config(name, m, n, k, strategy, dtype):
    """The `config` of the line: what the workload IS -- the same dict in our arm and in the reference arm, which times the reference's CPU
    COSMA on this very config (a bounded sample of it, described under cpu_baseline.sample). Nothing value-dependent goes in here."""
    eb = 8 if dtype == "d" else (16 if dtype == "z" else (4 if dtype == "s" else 8))
    return {"workload": name, "m": m, "n": n, "k": k, "strategy": strategy, "alpha": 1, "beta": 0,
            "l2": "inputs larger than L2 (A+B+C = %.1f GB per job vs 126 MB L2)" % (eb * 1e-9 * (m * k + k * n + m * n))}


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


class Watchdog:
    """A stage that never returns (a collective some rank does not join, a wedged device) must not cost the line: once the deadline has
    passed, rank 0 prints what has been measured so far -- the headline measurement comes first, the extras (`e2e`, `parity`, `also`,
    `cpu_baseline`) after it -- with the unfinished stage named under "incomplete", and every rank leaves. Never fires in a healthy run
    (N = 1 takes about 100 s, N = 8 about 60 s with the driver's 20 + 5 steps; the limit is 600 s, more when more steps are asked for).
    COSMA_B200_BENCH_DEADLINE_S overrides the limit; 0 switches the watchdog off."""

    def __init__(self, rank, reps=0):
        self.rank, self.line, self.stage, self.t0, self.printed = rank, None, "start-up", time.time(), False
        default = max(600.0, 150.0 + 4.0 * reps)  # a run asked for many steps gets the time they need (a step is ~2 s at N = 1)
        try:
            self.deadline = float(os.environ.get("COSMA_B200_BENCH_DEADLINE_S", default))
        except ValueError:
            self.deadline = default
        if self.deadline > 0:
            t = threading.Timer(self.deadline + (0.0 if rank == 0 else 5.0), self.fire)
            t.daemon = True
            t.start()

    def elapsed(self):
        return time.time() - self.t0

    def fire(self):
        if self.rank == 0 and not self.printed:  # (printed: only the teardown hangs -- the line is out, leave quietly)
            line = dict(self.line) if self.line else {"metric": METRIC, "value": None, "unit": "TFLOP/s"}
            line["incomplete"] = "stage '%s' had not finished %d s after start; the line holds what was measured before it" % (self.stage, int(self.deadline))
            try:
                print(json.dumps(line), flush=True)
            except Exception:
                pass
        os._exit(0 if (self.rank != 0 or self.line or self.printed) else 3)


def gemm_peak(dtype):
    """Roofline denominator of the local GEMM kernels, per GPU. FP64 (d, z): MEASURED_PEAKS.json (driver-written) carries no FP64 figure,
    so the larger of this repo's own DMMA issue-rate probe on the same pool (profiles/FP64_PEAK.json, 37.0) and the nominal
    148 SM x 64 FP64 FMA/clk x 2 x 1.965 GHz = 37.24 is used. FP32 (s, c) through 3xTF32: dense TF32 tensor peak / 3, with the TF32 peak taken
    as half the MEASURED bf16 peak of MEASURED_PEAKS.json (tcgen05 kind::tf32 issues at half the kind::f16 rate)."""
    nominal64 = 148 * 64 * 2 * 1.965e9 * 1e-12
    if dtype in "dz":
        try:
            with open(os.path.join(ROOT, "profiles", "FP64_PEAK.json")) as f:
                probe = float(json.load(f)["fp64_tflops"])
        except Exception:
            probe = 0.0
        if probe > nominal64:
            return probe, "profiles/FP64_PEAK.json (measured DMMA.8x8x4 issue-rate probe, this pool)"
        return nominal64, "nominal 148 SM x 64 FP64 FMA/clk x 2 x 1.965 GHz = 37.24 (MEASURED_PEAKS.json has no FP64 figure; own DMMA probe on this pool: %.1f)" % probe
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            bf16 = float(json.load(f)["bf16_tflops"])
        return bf16 / 2.0 / 3.0, "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 rate) / 3 (three TF32 MMAs per FP32 product)"
    except Exception:
        return 1125.0 / 3.0, "fallback: nominal dense TF32 1125 TFLOP/s / 3"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6540.0, "fallback of B200_PROFILING.md"


def sample_k(m, n, k, reps, budget_s=150.0, cpu_tflops=1.5):
    """The CPU arms time a BOUNDED sample of the workload: same m, n, the k dimension cut so that `reps` steps take about budget_s on
    the host cores (the GEMM rate does not depend on k at these sizes)."""
    per_step = budget_s / max(reps, 1) * cpu_tflops * 1e12
    return int(min(k, max(256, int(per_step / (2.0 * m * n)) // 256 * 256)))


def reference_arm(args):
    """The reference's own CPU implementation of the path: cosma::multiply at P=1 (unmodified sources built into
    oracle/_ref with a single-rank MPI stand-in + OpenBLAS), all host threads, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name, m, n, k, dtype = workload(args)
    if args.workload == "pzgemm" or dtype != "d":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the FP64 multiply workloads (cfg2, cfg3, largek)"}))
        return 0
    from oracle import oracle as orc
    if not orc.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcosma_ref.so was not built (no /root/reference at build time)"}))
        return 0
    cores = os.cpu_count() or 1
    reps = args.warmup + args.steps
    ks = sample_k(m, n, k, reps)
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
    sample = "reference cosma::multiply P=1 (OpenBLAS 0.3.30, %d threads) on m=%d n=%d k=%d (%s)" % (
        cores, m, n, ks, "the whole workload" if ks == k else "k cut from %d to bound the run" % k)
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, m, n, k, args.strategy or orc.ref_strategy(m, n, k, max(1, args.gpus))[0], dtype),
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
    sample = ("reference cosma_miniapp on %d minimpi ranks x %d OpenBLAS 0.3.30 threads, m=%d n=%d k=%d (k cut from %d), strategy of the sample [%s], "
              "mean of the %d fastest of %d repetitions" % (R, threads, m, n, ks, k, "; ".join(steps), len(timed), reps))
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, m, n, k, args.strategy or orc.ref_strategy(m, n, k, R)[0], "d"),
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
        ks = sample_k(m, n, k, 3, budget_s=20.0)
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


class Env:
    """torch / torch.distributed state of this rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.affinity = None
        if self.world > 1:
            # several ranks share the host: keep each rank (and the pinned buffers it allocates) on the NUMA node of its GPU
            try:
                from cosma_b200 import affinity as _aff
                self.affinity = _aff.bind_to_gpu(self.local_rank)
            except Exception:
                self.affinity = None
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()


def run_multiply(env, m, n, k, dtype, steps, warmup, strategy="", with_e2e=True, with_parity=True, sample_clocks=False, comm=None,
                 on_measured=None, stage=None):
    """One multiply workload: W warm-up steps, K timed steps (device events, max over ranks), roofline of the GEMM launches, collectives,
    end to end from pinned host memory, and an exact parity check on integer-valued operands. -> dict of results."""
    torch = env.torch
    from cosma_b200 import distributed
    job = distributed.MultiplyJob(m, n, k, env.world, env.rank, env.dev, steps=strategy, dtype=dtype, comm=comm)
    out = {"strategy": job.strategy_string}
    for _ in range(warmup):
        job.run()
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if sample_clocks and env.rank == 0:
        sampler.start()
        time.sleep(0.3)
    env.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    launches = 0
    t0 = time.time()
    ev[0].record()
    for i in range(steps):
        launches += job.run()
        ev[i + 1].record()
    env.barrier()
    t1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    last_step_ms = ev[-2].elapsed_time(ev[-1])
    if sample_clocks and env.rank == 0:
        out["clocks"] = sampler.stop(t0, t1)
    ms_per_step = env.max_over_ranks(total_ms) / steps
    flop = (8.0 if dtype in "zc" else 2.0) * m * n * k
    out["ms_per_step"] = ms_per_step
    out["value"] = flop / (ms_per_step * 1e-3) * 1e-12
    out["launches"] = launches
    # roofline of the dominant kernel: algorithmic flops of this rank's GEMM launches / their summed duration (CUDA events on the launching stream)
    peak, peak_src = gemm_peak(dtype)
    gflop, gms, glaunches = job.gemm_launch_stats()
    achieved = gflop / (gms * 1e-3) * 1e-12 if gms > 0 else 0.0
    kernel = "gemm_f64_sm100_kernel (FP64 DMMA.8x8x4 pipe)" if dtype in "dz" else "gemm_tf32x3_sm100_kernel (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators)"
    out["roofline"] = {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                       "traffic": None, "peak_source": peak_src, "flops_per_launch": gflop / max(glaunches, 1), "launch_ms": gms / max(glaunches, 1),
                       "launches_per_step": glaunches}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            out["roofline"]["traffic"] = json.load(f).get("%sgemm_%d" % (dtype, m))
    except Exception:
        pass
    if env.world > 1:
        try:
            out["collectives"] = job.collectives(last_step_ms)  # rank 0's last timed step
        except Exception as e:  # a reporting extra must never cost the bench line
            out["collectives"] = {"error": str(e)[:300]}
    if on_measured:
        on_measured(out)  # the device-resident measurement is complete: from here on the watchdog has a line to print
    stage = stage or (lambda name: None)
    if with_e2e:
        stage("e2e")
        try:
            out["e2e"] = job.e2e(max(2, min(steps, 3)))
        except Exception as e:  # an error every rank sees (e.g. out of pinned memory) must not cost the device-resident line
            out["e2e"] = {"error": str(e)[:300]}
    if with_parity:
        stage("parity")
        try:
            out["parity"] = job.parity()
        except Exception as e:
            out["parity"] = {"ok": False, "error": str(e)[:300]}
    job.destroy()
    from cosma_b200 import _lib
    _lib.load().cosma_b200_release_workspace()
    env.barrier()  # every rank has dropped its mappings of the others' arenas before anybody frees memory
    torch.cuda.empty_cache()
    return out


def run_pzgemm(env, steps, warmup, comm, n=16384, nb=256, with_e2e=True):
    """BASELINE configs[4]: pzgemm on a 2D block-cyclic 256 x 256 distribution, A conjugate-transposed -- COSTA relayout in, COSMA multiply,
    COSTA relayout out. Two rooflines: the ZGEMM launches against the FP64 peak, the relayout kernels against the HBM copy peak."""
    torch = env.torch
    from cosma_b200 import costa
    world, rank, dev = env.world, env.rank, env.dev
    nprow, npcol = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
    grid = costa.Grid(comm, "R", nprow, npcol)
    m = k = n
    tdt, eb = torch.complex128, 16

    def local(rows, cols, fill=None):
        lr_ = costa.numroc(rows, nb, grid.myrow, 0, nprow); lc_ = costa.numroc(cols, nb, grid.mycol, 0, npcol)
        lld = max(lr_, 1)
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        if fill is None:
            t = torch.view_as_complex(torch.rand(lld * max(lc_, 1), 2, device=dev, dtype=torch.float64, generator=gen))
        else:
            t = torch.full((lld * max(lc_, 1),), fill, device=dev, dtype=tdt)
        return t, costa.descinit(rows, cols, nb, nb, 0, 0, lld)

    A, da = local(k, m); B, db = local(k, n); C, dc = local(m, n, float("nan"))

    def step(a=A, b=B, c=C):
        costa.pxgemm(grid, "z", "C", "N", m, n, k, 1.0, a.data_ptr(), 1, 1, da, b.data_ptr(), 1, 1, db, 0.0, c.data_ptr(), 1, 1, dc)

    for _ in range(warmup):
        step()
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1) / steps)
    stats = costa.last_layout_multiply_stats(comm)
    phases = [env.max_over_ranks(stats[x]) for x in ("ms_relayout_in", "ms_multiply", "ms_relayout_out")]
    finite = bool(torch.isfinite(torch.view_as_real(C)).all().item())
    flops = 8.0 * m * n * k
    peak, peak_src = gemm_peak("z")
    hbm, hbm_src = hbm_peak()
    in_bytes = 2.0 * (stats["in_local_elements"] + stats["in_remote_elements"]) * eb    # read + write of every element, rank 0
    out_bytes = 2.0 * (stats["out_local_elements"] + stats["out_remote_elements"]) * eb
    res = {"workload": "pzgemm 2D block-cyclic %dx%d, m=n=k=%d, transa=C, grid %dx%d R (BASELINE configs[4])" % (nb, nb, n, nprow, npcol),
           "value": flops / (ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": ms, "strategy": stats["strategy"], "dtype": "c128",
           "frac_of_fp64_peak": flops / (ms * 1e-3) * 1e-12 / (peak * world),
           "phases_ms_max_over_ranks": {"relayout_in": phases[0], "multiply": phases[1], "relayout_out": phases[2]},
           "roofline": {"bound": "tensor", "kernel": "gemm_f64_sm100_kernel<.,.,CPLX> inside the multiply phase (allgathers and reduce included in its time)",
                        "achieved": flops / world / (phases[1] * 1e-3) * 1e-12, "peak": peak, "unit": "TFLOP/s",
                        "frac": flops / world / (phases[1] * 1e-3) * 1e-12 / peak, "peak_source": peak_src},
           "relayout_roofline": {"bound": "hbm", "kernel": "relayout_kernel (pack + local transposes | unpack), exchange included in the phase time",
                                 "rank0_in_GBps": in_bytes / (stats["ms_relayout_in"] * 1e-3) * 1e-9 if stats["ms_relayout_in"] > 0 else None,
                                 "rank0_out_GBps": out_bytes / (stats["ms_relayout_out"] * 1e-3) * 1e-9 if stats["ms_relayout_out"] > 0 else None,
                                 "peak": hbm, "unit": "GB/s", "peak_source": hbm_src,
                                 "rank0_in_remote_fraction": stats["in_remote_elements"] / max(1, stats["in_local_elements"] + stats["in_remote_elements"])},
           "result_finite": finite, "gpu_launches": stats["launches"] * steps}
    if with_e2e:
        try:
            hA, hB, hC = (torch.empty(t.numel(), dtype=tdt).pin_memory() for t in (A, B, C))
            hA.copy_(A); hB.copy_(B)
            step(hA, hB, hC); env.barrier()
            reps = max(2, min(steps, 3))
            e0.record()
            for _ in range(reps):
                step(hA, hB, hC)
            e1.record(); env.barrier()
            t = env.max_over_ranks(e0.elapsed_time(e1) / reps)
            res["e2e"] = {"value": flops / (t * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": t, "h2d_bytes_per_step": (A.numel() + B.numel()) * eb,
                          "d2h_bytes_per_step": C.numel() * eb, "api": "cosma_b200_pzgemm with pinned host local arrays",
                          "matches_device_path": bool(torch.equal(hC[:4096], C[:4096].cpu()))}
        except Exception as e:
            res["e2e"] = {"error": str(e)[:300]}
    grid.destroy()
    del A, B, C
    torch.cuda.empty_cache()
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    dog = Watchdog(int(os.environ.get("RANK", "0")), args.steps + args.warmup)
    env = Env()
    from cosma_b200 import _lib
    from cosma_b200.distributed import init_comm
    _lib.load()  # raises if the CUDA library is missing: there is no CPU fallback
    if env.world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, env.world))
    comm = init_comm(env.dev)
    name, m, n, k, dtype = workload(args)

    if args.workload == "pzgemm":
        res = run_pzgemm(env, args.steps, args.warmup, comm, with_e2e=not args.no_e2e)
        if env.rank == 0:
            line = {"metric": METRIC, "value": res["value"], "unit": "TFLOP/s", "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
                    "data": "synthetic", "config": {"workload": res["workload"], "strategy": res["strategy"]}, "roofline": res["roofline"],
                    "relayout_roofline": res["relayout_roofline"], "phases_ms_max_over_ranks": res["phases_ms_max_over_ranks"],
                    "cpu_baseline": None, "e2e": res.get("e2e"), "gpu_launches": res["gpu_launches"], "result_finite": res["result_finite"]}
            print(json.dumps(line))
        comm.destroy()
        if env.world > 1:
            env.dist.destroy_process_group()
        return 0

    peak, _ = gemm_peak(dtype)

    def assemble(res):
        """The bench line from the results measured so far (built again as the extras complete; key order as in the contract)."""
        line = {"metric": METRIC, "value": res["value"], "unit": "TFLOP/s", "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NAME[dtype],
                "data": "synthetic", "config": workload_config(name, m, n, k, res["strategy"], dtype),
                "peak_per_gpu_tflops": peak, "frac_of_peak": res["value"] / (peak * env.world), "roofline": res["roofline"], "cpu_baseline": res.get("cpu_baseline"), "e2e": res.get("e2e"), "gpu_launches": res["launches"],
                "clocks": res.get("clocks"), "parity": res.get("parity")}
        if env.affinity is not None:
            line["host_affinity_rank0"] = env.affinity
        if "collectives" in res:
            line["collectives"] = res["collectives"]  # rank 0's last timed step
        if res.get("also"):
            line["also"] = dict(res["also"])
        dog.line = line  # one reference assignment: the watchdog thread sees the old line or the new one, never half of one
        return line

    def stage(name):
        dog.stage = name

    stage("device-resident measurement of %s" % args.workload)
    res = run_multiply(env, m, n, k, dtype, args.steps, args.warmup, strategy=args.strategy, with_e2e=not args.no_e2e,
                       with_parity=not args.no_parity, sample_clocks=True, comm=comm, on_measured=assemble, stage=stage)
    assemble(res)

    # the other named configs that fit this job, in short (2 timed steps each): visible in the driver's records
    also = {}
    if not args.no_also and not args.mnk and args.workload == "cfg3":
        extra = []
        if env.world == 1:
            extra.append("cfg2")
        if env.world == 8:
            extra += ["largek", "pzgemm"]
        for w in extra:
            stage("also." + w)
            if dog.deadline > 0 and env.max_over_ranks(dog.elapsed()) > 0.5 * dog.deadline:  # (one verdict for all ranks: the extras are collective)
                also[w] = {"skipped": "more than half of the run's time limit was used before this extra"}
                res["also"] = also
                continue
            try:
                if w == "pzgemm":
                    also[w] = run_pzgemm(env, 2, 2, comm, with_e2e=False)
                else:
                    wm, wn, wk, wd, wname = WORKLOADS[w]
                    r = run_multiply(env, wm, wn, wk, wd, 2, 2, with_e2e=(w == "cfg2"), with_parity=(w != "cfg2"), comm=comm, stage=stage)
                    wpeak, _ = gemm_peak(wd)
                    also[w] = {"workload": wname % env.world, "value": r["value"], "unit": "TFLOP/s", "ms_per_step": r["ms_per_step"],
                               "strategy": r["strategy"], "frac_of_peak": r["value"] / (wpeak * env.world), "roofline": r["roofline"],
                               "collectives": r.get("collectives"), "e2e": r.get("e2e"), "parity": r.get("parity"), "steps": 2, "warmup": 2}
            except Exception as e:
                also[w] = {"error": str(e)[:300]}
            res["also"] = also
            assemble(res)

    if env.rank == 0:
        # rank 0 at N = 1 only (the contract): at N > 1 the reference arm (--impl reference) gives the CPU number of the same job
        if not args.no_cpu_baseline and dtype == "d" and env.world == 1:
            stage("cpu_baseline")
            res["cpu_baseline"] = cpu_baseline(m, n, k)
        line = assemble(res)
        stage("printing")
        print(json.dumps(line), flush=True)
        dog.printed = True
    stage("teardown")
    comm.destroy()
    if env.world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
