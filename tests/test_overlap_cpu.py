"""Communication / computation overlap (include/cosma/overlap.hpp, csrc/host/overlap.cpp) without a GPU: the micro-op programs the
executor would issue on its two streams are interpreted for ALL ranks in lock-step (tests/schedule_sim.py::run_overlapped) and
the gathered C must equal the dense product EXACTLY on integer-valued inputs -- with every communication workspace poisoned with NaN
first (nothing is read before it has arrived) and, for beta == 0, with C itself holding NaN (C is never read then).

Reference behaviour being re-told: overlap_m_split / overlap_n_split / overlap_k_split (src/cosma/one_sided_communicator.cpp:417-656,
776-1016), switched by COSMA_OVERLAP_COMM_AND_COMP (environment_variables.hpp) and gated by strategy.cpp:851-901."""
import numpy as np
import pytest

from schedule_sim import simulate
from cosma_b200.distributed import MultiplyPlan

FORCED = [
    (64, 64, 64, 2, "pk2"), (64, 64, 64, 2, "pm2"), (64, 64, 64, 2, "pn2"), (128, 128, 128, 4, "pn2,pk2"), (128, 128, 128, 4, "pm2,pk2"),
    (128, 128, 128, 8, "pm2,pn2,pk2"), (128, 128, 128, 8, "pn2,pm2,pk2"), (128, 128, 128, 8, "pk2,pm2,pn2"), (128, 128, 128, 4, "pm2,pn2"),
    (128, 96, 64, 4, "pk2,pm2"), (96, 128, 160, 8, "pn2,pm2,pk2"), (128, 128, 128, 4, "pk2,pk2"), (128, 128, 128, 4, "pm2,pm2"),
    (64, 192, 128, 8, "pn2,pn2,pk2"), (256, 64, 64, 4, "pm2,pk2"), (48, 80, 112, 8, "pm2,pn2,pk2"),
]
IDS = lambda c: "%dx%dx%d_P%d_%s" % c


@pytest.fixture
def forced(monkeypatch):
    monkeypatch.setenv("COSMA_OVERLAP_COMM_AND_COMP", "FORCE")
    monkeypatch.setenv("COSMA_B200_OVERLAP_GRANULE", "8")


@pytest.mark.parametrize("case", FORCED, ids=IDS)
@pytest.mark.parametrize("dtype", ["d", "z"])
@pytest.mark.parametrize("zero_sm", [False, True], ids=["nccl", "copy_engine"])
def test_overlapped_programs_in_lock_step(lib, forced, monkeypatch, case, dtype, zero_sm):
    """Both shapes of the program: panels beside NCCL kernels (narrow launches) and panels for the copy-engine transport (no narrow
    launch; what cosma_b200_plan_bind_arenas switches to)."""
    if zero_sm:
        monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")
    m, n, k, P, steps = case
    for beta in (0.0, 1.0, 2.0):
        got, want, P_used = simulate(m, n, k, P, steps, alpha=2.0, beta=beta, dtype=dtype, overlapped=True, poison=True)
        assert P_used == P and np.array_equal(got, want), (case, beta)


@pytest.mark.parametrize("case", [(64, 64, 64, 2, "pk2"), (128, 128, 128, 8, "pm2,pn2,pk2"), (128, 128, 128, 4, "pk2,pk2")], ids=IDS)
def test_c_is_never_read_when_beta_is_zero(lib, forced, case):
    """ScaLAPACK's rule (reference utils/pxgemm_utils.hpp:603-637): with beta == 0 the exchange lands in C itself and the
    'beta * C' step of the micro-op program is skipped at run time."""
    m, n, k, P, steps = case
    rng = np.random.default_rng(5)
    A, B = rng.integers(0, 10, size=(m, k)).astype(float), rng.integers(0, 10, size=(k, n)).astype(float)
    got, _, _ = simulate(m, n, k, P, steps, alpha=1.0, beta=0.0, inputs=(A, B, np.full((m, n), np.nan)), overlapped=True, poison=True)
    assert np.array_equal(got, A @ B)


def _plan(m, n, k, P, steps="", rank=0, dtype="d"):
    return MultiplyPlan(None, m, n, k, steps, dtype, rank=rank, nranks=P, allocate=False)


def test_copy_engine_program_for_the_headline_config(lib, monkeypatch):
    """32768^3 on 8 ranks planned for the zero-SM transport: no narrow launch, and every GEMM panel but the last is a whole number of
    waves of the 148-CTA grid (128 tile rows: widths that are multiples of 37 tile columns)."""
    monkeypatch.delenv("COSMA_OVERLAP_COMM_AND_COMP", raising=False)
    monkeypatch.delenv("COSMA_B200_OVERLAP_GRANULE", raising=False)
    monkeypatch.setenv("COSMA_B200_OVERLAP_ZERO_SM", "ON")
    for rank in (0, 3, 5, 6):
        pl = _plan(32768, 32768, 32768, 8, rank=rank)
        ov = pl.overlap()
        assert ov["enabled"], ov["why"]
        gemms = [o for o in ov["ops"] if o["kind"] == "gemm"]
        assert not any(o["narrow"] for o in gemms)
        waste = 0.0
        for o in gemms:
            tiles = (o["m"] // 128) * (o["n"] // 128)
            waves = -(-tiles // 148)
            waste += (waves * 148 - tiles) / 148.0 * (o["k"] / 16384.0)   # in full-depth tile rounds
        assert waste < 1.0, (rank, waste, [(o["n"], o["k"]) for o in gemms])  # the undivided GEMM alone wastes 0.3 of a round
        pl.destroy()


def test_baseline_configs_are_lowered_by_default(lib, monkeypatch):
    """BASELINE configs[2] (32768^3 at 2 / 4 / 8 ranks): the default (no environment) overlaps every ring-of-two collective; the first
    panel needs nothing from the network and is narrow (the NCCL kernels run beside it), the peer's half of C is complete before the
    exchange starts, and the panels still cover the GEMM exactly once."""
    for v in ("COSMA_OVERLAP_COMM_AND_COMP", "COSMA_B200_OVERLAP_GRANULE", "COSMA_B200_OVERLAP_SMS", "COSMA_B200_OVERLAP_GBPS", "COSMA_B200_OVERLAP_ZERO_SM"):
        monkeypatch.delenv(v, raising=False)
    for P, steps, n_ag in ((8, "pm2,pn2,pk2", 2), (4, "pn2,pk2", 1), (2, "pk2", 0)):
        for rank in range(P):
            pl = _plan(32768, 32768, 32768, P, rank=rank)
            ov = pl.overlap()
            assert pl.strategy == steps and ov["enabled"], ov["why"]
            ops = ov["ops"]
            assert [o["kind"] for o in ops].count("allgather") == n_ag and all(o["stream"] == 1 for o in ops if o["kind"] in ("allgather", "exchange"))
            gemms = [o for o in ops if o["kind"] == "gemm"]
            g = next(o for o in pl.ops() if o["kind"] == "gemm")
            assert sum(o["n"] * o["k"] for o in gemms) == g["n"] * g["k"] and all(o["m"] == g["m"] for o in gemms)
            if n_ag:
                assert gemms[0]["narrow"] == 1 and gemms[0]["wait"] == [] and gemms[0]["k"] == g["k"] // (2 if P >= 4 else 1)
                assert all(o["wait"] for o in gemms[1:])
            ex = next(i for i, o in enumerate(ops) if o["kind"] == "exchange")
            assert ops[ex]["wait"] == [ex - 1] and ops[ex - 1]["kind"] == "gemm" and ops[ex + 1]["kind"] == "gemm" and ops[ex + 1]["narrow"] == 1
            # whole waves: all launches together waste less than 1.5 full-depth tile rounds (the undivided GEMM alone wastes 0.2 - 0.3)
            waste = 0.0
            for o in gemms:
                ctas = 140 if o["narrow"] else 148
                tiles = (o["m"] // 128) * (o["n"] // 128)
                waves = -(-tiles // ctas)
                waste += (waves * ctas - tiles) / float(ctas) * (o["k"] / float(g["k"]))
            assert waste < 1.5, (P, rank, waste)
            serial, overlapped, comm = ov["est_ms"]
            assert overlapped < serial and comm > 0
            pl.destroy()


def test_what_is_not_lowered(lib, monkeypatch):
    monkeypatch.delenv("COSMA_OVERLAP_COMM_AND_COMP", raising=False)
    monkeypatch.delenv("COSMA_B200_OVERLAP_GRANULE", raising=False)
    for (m, n, k, P, steps, reason) in ((8192, 8192, 1048576, 8, "", "ring-of-two"),             # pk8: one ring of eight
                                        (16384, 16384, 16384, 1, "", "multi-rank"),                # no communication at all
                                        (512, 512, 512, 4, "sm2,pn2,pk2", "more than one"),       # sequential steps: several GEMMs
                                        (300, 301, 302, 6, "pk3,pm2", "ring-of-two")):
        pl = _plan(m, n, k, P, steps)
        ov = pl.overlap()
        assert not ov["enabled"] and reason in ov["why"], ov
        pl.destroy()
    monkeypatch.setenv("COSMA_OVERLAP_COMM_AND_COMP", "OFF")
    pl = _plan(32768, 32768, 32768, 8)
    assert not pl.overlap()["enabled"] and "switched off" in pl.overlap()["why"]
    pl.destroy()


def test_irregular_rings_fall_back_on_every_rank(lib, forced):
    """Odd sizes: some rings carry unequal pieces. The verdict is taken over ALL ranks' schedules, so ring mates never disagree about
    the protocol -- either every rank lowers or none does."""
    for (m, n, k, P, steps) in ((97, 101, 103, 4, "pn2,pk2"), (130, 126, 66, 8, "pm2,pn2,pk2"), (64, 64, 64, 4, "pn2,pk2")):
        verdicts = []
        for rank in range(P):
            pl = _plan(m, n, k, P, steps, rank=rank)
            verdicts.append(pl.overlap()["enabled"])
            pl.destroy()
        assert len(set(verdicts)) == 1, (m, n, k, P, steps, verdicts)
        got, want, _ = simulate(m, n, k, P, steps, alpha=1.0, beta=1.0, overlapped=verdicts[0], poison=True)
        assert np.array_equal(got, want)
