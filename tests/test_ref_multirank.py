"""Distributed parity against the UNMODIFIED reference running on several ranks (oracle/_ref/ref_driver under
oracle/minirun.py: minimpi processes + OpenBLAS): for the reference's own 40 distributed cases
(tests/multiply.cpp:142-321), its mixed sequential/parallel case (tests/scalar_matmul.cpp) and BASELINE configs[0]
(2000^3 on 2 ranks), every rank's raw local C buffer of cosma::multiply must equal BIT FOR BIT the buffer our compiled
schedule produces for that rank (interpreted on the CPU by tests/schedule_sim.py; the GPU tests run the same comparison
on the device). Inputs are integer valued, so every partial sum is exact and the comparison is order independent.

Also writes / checks the committed fixtures under tests/golden/ref_multirank_*.npz (generated here by
tests/golden/make_multirank_golden.py from the same driver) so the pin travels to boxes without /root/reference."""
import os

import numpy as np
import pytest

from cases import REFERENCE_MULTIPLY_CASES, SCALAR_MATMUL_CASE
from schedule_sim import simulate

IDS = lambda c: "%dx%dx%d_P%d_%s" % (c[0], c[1], c[2], c[3], c[4] or "auto")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs(m, n, k, dtype, seed):
    rng = np.random.default_rng(seed)
    def one(r, c):
        v = rng.integers(-4, 6, size=(r, c)).astype(np.float64)
        if dtype in "cz":
            v = v + 1j * rng.integers(-4, 6, size=(r, c))
        return v
    return one(m, k), one(k, n), one(m, n)


def _compare(ref_locals, our_locals):
    assert len(ref_locals) == len(our_locals)
    for r, (a, b) in enumerate(zip(ref_locals, our_locals)):
        assert (a is None) == (b is None), "rank %d: idle on one side only" % r
        if a is not None:
            assert a.dtype == b.dtype and a.shape == b.shape, "rank %d: local size %s vs %s" % (r, a.shape, b.shape)
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "rank %d: local C differs from the reference" % r


@pytest.fixture(scope="module")
def refd(ref):
    if not ref.have_ref_driver():
        pytest.skip("oracle/_ref/ref_driver not built")
    return ref


@pytest.mark.parametrize("case", REFERENCE_MULTIPLY_CASES, ids=IDS)
def test_reference_cases_per_rank_bit_exact(lib, refd, case):
    m, n, k, P, steps = case
    A, B, C = _inputs(m, n, k, "d", seed=m * 7 + n * 5 + k * 3 + P)
    ref_locals, _ = refd.ref_multiply_ranks("d", m, n, k, P, steps, 1.0, 1.0, A, B, C)
    ours = []
    got, want, _ = simulate(m, n, k, P, steps, alpha=1.0, beta=1.0, inputs=(A, B, C), local_c=ours)
    assert np.array_equal(got, want)
    _compare(ref_locals, ours)


@pytest.mark.parametrize("dtype", ["s", "d", "c", "z"])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (2.0, -1.0)])
def test_scalar_matmul_all_types(lib, refd, dtype, alpha, beta):
    m, n, k, P, steps = SCALAR_MATMUL_CASE
    A, B, C = _inputs(m, n, k, dtype, seed=11)
    ref_locals, _ = refd.ref_multiply_ranks(dtype, m, n, k, P, steps, alpha, beta, A, B, C)
    ours = []
    sim_dtype = "z" if dtype in "cz" else "d"
    simulate(m, n, k, P, steps, alpha=alpha, beta=beta, dtype=sim_dtype, inputs=(A, B, C), local_c=ours)
    ours = [None if x is None else x.astype(refd.NPDT[dtype]) for x in ours]  # integers: exact in every type
    _compare(ref_locals, ours)


def test_baseline_config0_2000_cubed_on_2_ranks(lib, refd):
    """BASELINE.json configs[0]: square dgemm m=n=k=2000, 2 ranks (reference correctness run)."""
    m = n = k = 2000
    A, B, C = _inputs(m, n, k, "d", seed=2000)
    ref_locals, _ = refd.ref_multiply_ranks("d", m, n, k, 2, "", 1.0, 0.0, A, B, C, threads=2)
    ours = []
    simulate(m, n, k, 2, "", alpha=1.0, beta=0.0, inputs=(A, B, C), local_c=ours)
    _compare(ref_locals, ours)


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.startswith("ref_multirank_")) if os.path.isdir(GOLDEN) else [])
def test_committed_reference_fixtures(lib, name):
    """Fixtures produced by the reference itself (tests/golden/make_multirank_golden.py); no reference needed to check them."""
    z = np.load(os.path.join(GOLDEN, name))
    m, n, k, P = (int(z[x]) for x in "mnkP")
    steps, dtype = str(z["steps"]), str(z["dtype"])
    alpha, beta = complex(z["alpha"]), complex(z["beta"])
    if dtype in "sd":
        alpha, beta = alpha.real, beta.real
    ours = []
    simulate(m, n, k, P, steps, alpha=alpha, beta=beta, dtype="z" if dtype in "cz" else "d", inputs=(z["A"], z["B"], z["C"]), local_c=ours)
    for r in range(P):
        key = "local_c_%d" % r
        if key in z.files:
            assert np.array_equal(z[key], ours[r].astype(z[key].dtype)), "rank %d" % r
        else:
            assert ours[r] is None


def test_randomised_schedules_against_the_live_reference(lib, refd):
    """A seeded slice of tests/fuzz/fuzz_schedule_vs_reference.py (755 random problems offline: random shapes, 1-8 ranks, automatic or random
    explicit strategies mixing sequential and parallel steps, three (alpha, beta) pairs): wherever the reference's answer is right, every
    rank's local C of our compiled schedule is bit-identical to it; the tool found NO case where ours is wrong, and 21 exotic explicit
    strategies (e.g. sm2,pm2,pm2 or pk2,sn2,pn2) where the REFERENCE's result differs from the dense product while ours
    equals it."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz", "fuzz_schedule_vs_reference.py"), "7", "25"], capture_output=True, text=True,
                         timeout=900, cwd=root)
    tail = [ln for ln in out.stdout.splitlines() if ln.startswith("cases run")]
    assert out.returncode == 0 and tail and tail[0].endswith("OUR mismatches 0"), out.stdout[-3000:]
    assert int(tail[0].split()[2].rstrip(",")) >= 12
