"""Static hazard analysis (tests/overlap_hazards.py) of the overlapped micro-op programs at the sizes the bench runs -- 32768^3 on 2 / 4 /
8 ranks, the pzgemm multiply (16384^3 complex on 8), both transports, beta == 0 and beta != 0 -- and of the small forced programs the
lock-step interpreter (tests/test_overlap_cpu.py) executes. The lock-step interpreter proves the DATA FLOW of a program run in program
order; this proves that the executor's two streams and the ring mate's pushes cannot reorder it: every conflicting pair of accesses is
ordered by stream order and wait edges, nothing touches a landing zone while the mate may be writing it, and the caller's stream ends
after every communication op (csrc/multiply_exec.cu::plan_run_overlapped waits for the last one only)."""
import pytest

import overlap_hazards as H
from cosma_b200.distributed import MultiplyPlan
from test_overlap_cpu import FORCED, IDS

BENCH = [(32768, 32768, 32768, 2, "", "d"), (32768, 32768, 32768, 4, "", "d"), (32768, 32768, 32768, 8, "", "d"),
         (16384, 16384, 16384, 8, "", "z"), (16384, 16384, 16384, 8, "", "s"), (16384, 16384, 16384, 4, "", "c"),
         (32768, 32768, 32768, 2, "pn2", "d"), (32768, 32768, 32768, 2, "pm2", "d")]


def _check(m, n, k, P, steps, dtype, copy_engine):
    seen = 0
    for rank in range(P):
        pl = MultiplyPlan(None, m, n, k, steps, dtype, rank=rank, nranks=P, allocate=False)
        ov = pl.overlap()
        if not ov["enabled"]:
            pl.destroy()
            continue
        micro, sched = ov["ops"], pl.ops()
        seen += 1
        if copy_engine:
            assert not any(o["kind"] == "gemm" and o["narrow"] for o in micro)
            # the executor copies the own pieces of ALL overlapped allgathers at the start of the communication stream: they must be the first ops on it
            comm = [o["kind"] for o in micro if o["stream"] == 1]
            assert comm == sorted(comm, key=lambda kd: kd != "allgather"), comm
        for beta_zero in (True, False):
            found = H.hazards(micro, sched, beta_zero, copy_engine)
            assert not found, (m, n, k, P, steps, dtype, "rank %d" % rank, "beta == 0" if beta_zero else "beta != 0", found)
        assert H.final_order(micro) == [], (rank, H.final_order(micro))
        pl.destroy()
    return seen


@pytest.mark.parametrize("case", BENCH, ids=lambda c: "%dx%dx%d_P%d_%s_%s" % c)
@pytest.mark.parametrize("copy_engine", [False, True], ids=["nccl", "copy_engine"])
def test_bench_size_programs_have_no_hazard(lib, monkeypatch, case, copy_engine):
    for v in ("COSMA_OVERLAP_COMM_AND_COMP", "COSMA_B200_OVERLAP_GRANULE", "COSMA_B200_OVERLAP_SMS", "COSMA_B200_OVERLAP_GBPS", "COSMA_B200_OVERLAP_ZERO_SM"):
        monkeypatch.delenv(v, raising=False)
    if case[4]:
        monkeypatch.setenv("COSMA_OVERLAP_COMM_AND_COMP", "FORCE")
    if copy_engine:
        monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")
    m, n, k, P, steps, dtype = case
    seen = _check(m, n, k, P, steps, dtype, copy_engine)
    # every rank lowers or none does (the planner may find no gain, e.g. single precision beside NCCL kernels); the FP64 configs always do
    assert seen in ((P,) if dtype in "dz" else (0, P)), seen


@pytest.mark.parametrize("case", FORCED, ids=IDS)
@pytest.mark.parametrize("copy_engine", [False, True], ids=["nccl", "copy_engine"])
def test_forced_programs_have_no_hazard(lib, monkeypatch, case, copy_engine):
    monkeypatch.setenv("COSMA_OVERLAP_COMM_AND_COMP", "FORCE")
    monkeypatch.setenv("COSMA_B200_OVERLAP_GRANULE", "8")
    if copy_engine:
        monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")
    m, n, k, P, steps = case
    for dtype in ("d", "z"):
        assert _check(m, n, k, P, steps, dtype, copy_engine) == P


def test_the_checker_sees_a_missing_wait(lib, monkeypatch):
    """Remove one wait edge / move a landing zone: the analysis must object (it is only worth something if it can fail)."""
    monkeypatch.delenv("COSMA_OVERLAP_COMM_AND_COMP", raising=False)
    monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")
    pl = MultiplyPlan(None, 32768, 32768, 32768, "", "d", rank=3, nranks=8, allocate=False)
    micro, sched = pl.overlap()["ops"], pl.ops()
    assert H.hazards(micro, sched, True, True) == []
    # a GEMM panel that no longer waits for the allgathers
    broken = [dict(o) for o in micro]
    victim = next(i for i, o in enumerate(broken) if o["kind"] == "gemm" and o["wait"])
    broken[victim]["wait"] = []
    assert H.hazards(broken, sched, True, True)
    # the accumulate that adds the received half no longer waits for the exchange
    broken = [dict(o) for o in micro]
    ex = next(i for i, o in enumerate(broken) if o["kind"] == "exchange")
    for o in broken:
        o["wait"] = [w for w in o["wait"] if w != ex]
    assert H.hazards(broken, sched, False, True)
    # the exchange no longer waits for the panel it sends
    broken = [dict(o) for o in micro]
    broken[ex]["wait"] = []
    assert H.hazards(broken, sched, True, True)
    pl.destroy()


def test_random_forced_programs(lib, monkeypatch):
    """Random shapes and ring-of-two strategies (seeded): wherever every rank lowers, the programs of both transports are hazard-free and
    the lock-step interpretation gives the dense product exactly, with poisoned workspaces."""
    import numpy as np
    from schedule_sim import simulate
    monkeypatch.setenv("COSMA_OVERLAP_COMM_AND_COMP", "FORCE")
    rng = np.random.default_rng(20261018)
    lowered = 0
    for _ in range(40):
        P = int(rng.choice([2, 4, 8]))
        n_steps = {2: 1, 4: 2, 8: 3}[P]
        steps = ",".join("p%s2" % rng.choice(list("mnk")) for _ in range(n_steps))
        g = int(rng.choice([8, 16]))
        m, n, k = (int(rng.integers(2, 9)) * 4 * g for _ in range(3))
        dtype = str(rng.choice(["d", "z"]))
        beta = float(rng.choice([0.0, 1.0, -0.5]))
        monkeypatch.setenv("COSMA_B200_OVERLAP_GRANULE", str(g))
        for copy_engine in (False, True):
            if copy_engine:
                monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")
            else:
                monkeypatch.delenv("COSMA_B200_OVERLAP_ZERO_SM", raising=False)
            seen = _check(m, n, k, P, steps, dtype, copy_engine)
            assert seen in (0, P), (m, n, k, P, steps, seen)
            if seen:
                got, want, _ = simulate(m, n, k, P, steps, alpha=2.0, beta=beta, dtype=dtype, overlapped=True, poison=True)
                assert np.array_equal(got, want), (m, n, k, P, steps, dtype, beta, copy_engine)
                lowered += 1
    assert lowered >= 40, lowered
