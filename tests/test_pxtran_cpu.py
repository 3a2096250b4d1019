"""p?tran / p?tranu / p?tranc and p?gemr2d on the CPU: the transform plans our entry points build for
block-cyclic -> block-cyclic moves (cosma_b200_scalapack_layout + cosma_b200_transform_plan_create, the two calls
xptransform in csrc/layout_multiply.cu makes) are interpreted for ALL ranks in lock-step with the oracle's
copy_and_transform and every rank's local array must equal BIT FOR BIT what the UNMODIFIED reference wrappers
(costa::pxtran_op, costa::pxgemr2d on minimpi ranks with the miniblacs grid) leave in it -- including padding rows
(lld > local rows) and everything outside sub(C). The device execution of the same plans is tests/test_costa_gpu.py."""
import numpy as np
import pytest

import costa_sim as sim
from cosma_b200 import costa

EB = {"s": 4, "d": 8, "c": 8, "z": 16}

TRAN_CASES = [
    # m, n (of sub(C)), blocks a, blocks c, (ia, ja), (ic, jc), extra rows/cols around the sub-matrices, alpha, beta
    dict(m=40, n=56, ba=(8, 8), bc=(8, 8), sa=(1, 1), sc=(1, 1), extra=0, alpha=1.0, beta=0.0),
    dict(m=37, n=53, ba=(5, 7), bc=(4, 9), sa=(3, 2), sc=(2, 6), extra=11, alpha=2.0, beta=-1.0),
    dict(m=64, n=16, ba=(16, 4), bc=(8, 32), sa=(1, 5), sc=(7, 1), extra=3, alpha=1.0, beta=1.0),
]


@pytest.fixture(scope="module")
def refd(ref):
    if not ref.have_ref_driver():
        pytest.skip("oracle/_ref/ref_driver not built")
    return ref


def _simulate(oracle, dtype, op, m, n, alpha, beta, bcA, bcC, a_loc, c_loc, ia, ja, ic, jc, P):
    """What xptransform does, on host memory: layouts of sub(A), sub(C) for every rank -> plans -> lock-step run."""
    a_subm, a_subn = (m, n) if op == "N" else (n, m)
    plans = []
    keep = []
    for r in range(P):
        lays = []
        for bc, loc, i0, j0, sm, sn in ((bcA, a_loc[r], ia, ja, a_subm, a_subn), (bcC, c_loc[r], ic, jc, m, n)):
            in_grid = r < bc.nprow * bc.npcol
            lld = bc.local_shape(r)[0] if in_grid else 1
            lays.append(costa.block_cyclic_layout(bc.M, bc.N, bc.mb, bc.nb, i0, j0, sm, sn, bc.nprow, bc.npcol, bc.order, bc.rsrc, bc.csrc,
                                                  loc.ctypes.data, lld, "C", r if in_grid else -1, EB[dtype]))
        keep.append(lays)
        tp = costa.TransformPlan(None, dtype, [(lays[0], lays[1], op, alpha, beta)], rank=r, nranks=P)
        plans.append(tp.export())
        tp.destroy()
    sim.simulate(oracle, dtype, plans, [(alpha, beta)])


@pytest.mark.parametrize("nprow,npcol,order", [(1, 1, "R"), (2, 2, "R"), (2, 3, "C"), (4, 2, "R")])
@pytest.mark.parametrize("dtype,op", [("d", "T"), ("s", "T"), ("z", "T"), ("z", "C"), ("c", "C")])
def test_pxtran_matches_reference_per_rank(lib, oracle, refd, nprow, npcol, order, dtype, op):
    P = nprow * npcol
    for case in TRAN_CASES:
        m, n, extra = case["m"], case["n"], case["extra"]
        (ia, ja), (ic, jc) = case["sa"], case["sc"]
        alpha, beta = case["alpha"], case["beta"]
        if dtype in "cz":
            alpha = alpha * (1 - 0.5j)
        rng = np.random.default_rng(m * 13 + n)
        GA = sim.random_values(rng, (n + ia - 1 + extra, m + ja - 1 + extra), dtype)
        GC = sim.random_values(rng, (m + ic - 1 + extra, n + jc - 1 + extra), dtype)
        rsrc, csrc = (1 % nprow, 1 % npcol) if extra else (0, 0)
        bcA = sim.BlockCyclic(GA.shape[0], GA.shape[1], case["ba"][0], case["ba"][1], nprow, npcol, order, rsrc, csrc, lld_pad=2)
        bcC = sim.BlockCyclic(GC.shape[0], GC.shape[1], case["bc"][0], case["bc"][1], nprow, npcol, order, 0, 0, lld_pad=1)
        a_loc = [bcA.scatter(GA, r, fill=77) for r in range(P)]
        c_loc = [bcC.scatter(GC, r, fill=55) for r in range(P)]
        want = refd.ref_pxtran_ranks(dtype, order, nprow, npcol, op, m, n, alpha, a_loc, ia, ja, [bcA.desc(r) for r in range(P)], beta,
                                     [x.copy() for x in c_loc], ic, jc, [bcC.desc(r) for r in range(P)])
        _simulate(oracle, dtype, op, m, n, alpha, beta, bcA, bcC, a_loc, c_loc, ia, ja, ic, jc, P)
        for r in range(P):
            assert np.array_equal(c_loc[r].view(np.uint8), want[r].view(np.uint8)), (case, r)
        # and the reference agrees with the dense definition
        got = GC.copy()
        for r in range(P):
            bcC.gather_into(got, want[r], r)
        dense = GC.copy()
        sub = sim.apply_op(GA[ia - 1:ia - 1 + n, ja - 1:ja - 1 + m], op)
        dense[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = alpha * sub + beta * GC[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n]
        assert np.array_equal(got, dense.astype(got.dtype))


@pytest.mark.parametrize("nprow,npcol,order,orderc", [(1, 1, "R", "R"), (2, 2, "R", "C"), (2, 4, "R", "R"), (3, 2, "C", "R")])
@pytest.mark.parametrize("dtype", ["d", "z", "s"])
def test_pxgemr2d_matches_reference_per_rank(lib, oracle, refd, nprow, npcol, order, orderc, dtype):
    P = nprow * npcol
    for case in TRAN_CASES:
        m, n, extra = case["m"], case["n"], case["extra"]
        (ia, ja), (ic, jc) = case["sa"], case["sc"]
        rng = np.random.default_rng(m * 17 + n)
        GA = sim.random_values(rng, (m + ia - 1 + extra, n + ja - 1 + extra), dtype)
        GC = sim.random_values(rng, (m + ic - 1 + extra, n + jc - 1 + extra), dtype)
        bcA = sim.BlockCyclic(GA.shape[0], GA.shape[1], case["ba"][0], case["ba"][1], nprow, npcol, order, 0, 0, lld_pad=2)
        bcC = sim.BlockCyclic(GC.shape[0], GC.shape[1], case["bc"][0], case["bc"][1], nprow, npcol, orderc, 0, 0, lld_pad=0)
        a_loc = [bcA.scatter(GA, r, fill=77) for r in range(P)]
        c_loc = [bcC.scatter(GC, r, fill=55) for r in range(P)]
        want = refd.ref_pxgemr2d_ranks(dtype, order, nprow, npcol, m, n, a_loc, ia, ja, [bcA.desc(r) for r in range(P)], [x.copy() for x in c_loc], ic,
                                       jc, [bcC.desc(r) for r in range(P)], orderc=orderc)
        _simulate(oracle, dtype, "N", m, n, 1.0, 0.0, bcA, bcC, a_loc, c_loc, ia, ja, ic, jc, P)
        for r in range(P):
            assert np.array_equal(c_loc[r].view(np.uint8), want[r].view(np.uint8)), (case, r)
