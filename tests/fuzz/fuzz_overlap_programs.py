"""Randomised check of the overlap planner (csrc/host/overlap.cpp): random shapes, rank counts, strategies, element types and planner
inputs; for every job whose ranks all lower
  (1) both transports' programs pass the static hazard analysis (tests/overlap_hazards.py) for beta == 0 and beta != 0,
  (2) the panels cover the base-case GEMM exactly once (columns x depth), never beyond its bounds,
  (3) small jobs are executed for all ranks in lock-step with poisoned workspaces and must give the dense product exactly,
  (4) bench-size jobs whose local matrices are never cut along m are replayed numerically with m compressed (tests/full_size_replay.py).
Either every rank of a job lowers or none does.   python tests/fuzz/fuzz_overlap_programs.py SEED N [big]
Last run: see DESIGN.md 5a."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import overlap_hazards as H  # noqa: E402
from schedule_sim import simulate  # noqa: E402
from cosma_b200.distributed import MultiplyPlan  # noqa: E402

ENV = ("COSMA_OVERLAP_COMM_AND_COMP", "COSMA_B200_OVERLAP_GRANULE", "COSMA_B200_OVERLAP_SMS", "COSMA_B200_OVERLAP_GBPS", "COSMA_B200_OVERLAP_ZERO_SM",
       "COSMA_B200_OVERLAP_COVER")


def set_env(**kw):
    for v in ENV:
        os.environ.pop(v, None)
    for k, v in kw.items():
        if v is not None:
            os.environ[k] = str(v)


def ring2_strategy(rnd, P):
    n_steps = {2: 1, 4: 2, 8: 3, 16: 4}[P]
    return ",".join("p%s2" % rnd.choice("mnk") for _ in range(n_steps))


def check_job(m, n, k, P, steps, dtype, copy_engine):
    """-> (ranks lowered, problems). Structural + hazard checks of every rank's program."""
    lowered, problems = 0, []
    for rank in range(P):
        try:
            pl = MultiplyPlan(None, m, n, k, steps, dtype, rank=rank, nranks=P, allocate=False)
        except Exception as e:  # a strategy the library refuses (e.g. more parts than elements) is not this fuzzer's business
            return -1, ["plan refused: %s" % str(e)[:80]]
        ov = pl.overlap()
        if ov["enabled"] and not pl.idle:
            lowered += 1
            micro, sched = ov["ops"], pl.ops()
            g = next(o for o in sched if o["kind"] == "gemm")
            gemms = [o for o in micro if o["kind"] == "gemm"]
            if sum(o["n"] * o["k"] for o in gemms) != g["n"] * g["k"] or any(o["m"] != g["m"] for o in gemms):
                problems.append("rank %d: the panels do not cover the GEMM once" % rank)
            if any(o["n"] <= 0 or o["k"] <= 0 or o["n"] > g["n"] or o["k"] > g["k"] for o in gemms):
                problems.append("rank %d: a panel exceeds the GEMM" % rank)
            if copy_engine and any(o["narrow"] for o in gemms):
                problems.append("rank %d: narrow launch in a copy-engine program" % rank)
            for bz in (True, False):
                for h in H.hazards(micro, sched, bz, copy_engine):
                    problems.append("rank %d beta%s0: %s" % (rank, "=" if bz else "!=", h))
            for i in H.final_order(micro):
                problems.append("rank %d: communication op #%d is never waited for" % (rank, i))
        pl.destroy()
    return lowered, problems


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    rnd = random.Random(seed)
    bad = jobs = lowered_jobs = executed = refused = 0
    for it in range(N):
        P = rnd.choice([2, 4, 8] + ([16] if big else []))
        dtype = rnd.choice("dzsc") if big else rnd.choice("dz")
        if big:
            # bench-like shapes: natural planner decisions (no FORCE), real granule, random SM counts / link rates
            unit = rnd.choice([128, 256, 1000, 1024, 37 * 128])
            m, n, k = (unit * rnd.randint(2, 40) for _ in range(3))
            steps = rnd.choice(["", "", ring2_strategy(rnd, P)])
            env = dict(COSMA_OVERLAP_COMM_AND_COMP=rnd.choice([None, None, "FORCE"]), COSMA_B200_OVERLAP_SMS=rnd.choice([None, 4, 8, 16]),
                       COSMA_B200_OVERLAP_GBPS=rnd.choice([None, None, 50, 300, 900]), COSMA_B200_OVERLAP_COVER=rnd.choice([None, None, 0.5, 2.0]))
        else:
            g = rnd.choice([4, 8, 16])
            m, n, k = (g * rnd.randint(2, 14) * rnd.choice([1, 2]) for _ in range(3))
            steps = ring2_strategy(rnd, P)
            env = dict(COSMA_OVERLAP_COMM_AND_COMP="FORCE", COSMA_B200_OVERLAP_GRANULE=g, COSMA_B200_OVERLAP_SMS=rnd.choice([None, 2, 8, 20]),
                       COSMA_B200_OVERLAP_COVER=rnd.choice([None, None, 0.0, 1.0]))
        for copy_engine in (False, True):
            set_env(**env, COSMA_B200_OVERLAP_ZERO_SM="ON" if copy_engine else None)
            low, problems = check_job(m, n, k, P, steps, dtype, copy_engine)
            if low < 0:
                refused += 1
                continue
            jobs += 1
            pl0 = MultiplyPlan(None, m, n, k, steps, dtype, rank=0, nranks=P, allocate=False)
            used = pl0.P_used
            pl0.destroy()
            if low not in (0, used):
                problems.append("only %d of %d ranks lower" % (low, used))
            if low == used and not problems:
                lowered_jobs += 1
                if not big and dtype in "dz":
                    beta = rnd.choice([0.0, 1.0, -0.5])
                    got, want, _ = simulate(m, n, k, P, steps, alpha=2.0, beta=beta, dtype=dtype, overlapped=True, poison=True)
                    executed += 1
                    if not np.array_equal(got, want):
                        problems.append("lock-step result differs from the dense product (beta = %s)" % beta)
                if big and dtype in "ds" and k * n <= 2 ** 27 and max(m, n, k) * 9 * 9 < 2 ** 24:
                    # (4) numeric replay of the real program with m compressed by the largest factor every row boundary allows
                    import math
                    import full_size_replay as R
                    f = m
                    for r in range(used):
                        pl = MultiplyPlan(None, m, n, k, steps, dtype, rank=r, nranks=P, allocate=False)
                        for label in "AC":
                            for (r0, r1, _, _) in pl.local_blocks(label):
                                f = math.gcd(f, math.gcd(r0, r1 + 1))
                        pl.destroy()
                    if f > 1 and m // f <= 256:
                        beta = rnd.choice([0.0, 2.0])
                        try:
                            got, want, _ = R.replay(m, n, k, P, steps=steps, dtype=dtype, m_small=m // f, beta=beta, seed=it)
                            executed += 1
                            if not np.array_equal(got, want):
                                problems.append("replay of the full-size program differs from the dense product (beta = %s)" % beta)
                        except AssertionError as e:
                            if "whole number" not in str(e):
                                raise
            if problems:
                bad += 1
                print("PROBLEM seed %d it %d: %dx%dx%d P=%d steps '%s' %s %s env %s" % (seed, it, m, n, k, P, steps, dtype, "copy-engine" if copy_engine else "nccl", env))
                for p in problems[:6]:
                    print("   ", p)
    set_env()
    print("seed %d: %d jobs (x transports), %d lowered by every rank, %d executed in lock-step, %d refused, %d with problems" % (seed, jobs, lowered_jobs, executed, refused, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
