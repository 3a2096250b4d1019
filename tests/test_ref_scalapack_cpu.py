"""The UNMODIFIED reference ScaLAPACK wrappers -- cosma::pxgemm, costa::pxgemr2d, costa::pxtran_op -- running on several
minimpi ranks with the miniblacs grid stand-in (oracle/ref_driver.cpp), checked against dense numpy on integer-valued
block-cyclic matrices. This pins the stand-ins: with them in place the reference itself is the oracle of our p?gemm /
p?gemr2d / p?tran entry points (tests/test_costa_gpu.py, tests/test_pxtran_cpu.py) -- ScaLAPACK proper is absent from
the image, so without this the p?gemm boundary would be 'parity unpinned' (SURVEY.md 8c)."""
import numpy as np
import pytest

import costa_sim as sim

PX = [
    dict(m=96, n=80, k=64, ta="N", tb="N", alpha=1.0, beta=0.0, blk=((8, 8), (8, 8), (8, 8))),
    dict(m=96, n=80, k=64, ta="T", tb="N", alpha=2.0, beta=1.0, blk=((16, 8), (8, 4), (32, 8))),
    dict(m=50, n=60, k=70, ta="N", tb="T", alpha=1.0, beta=-1.0, blk=((7, 5), (3, 9), (4, 4)), sub=((3, 2), (5, 4), (2, 6)), extra=12),
    dict(m=33, n=17, k=129, ta="C", tb="C", alpha=1.0, beta=0.0, blk=((5, 5), (6, 6), (7, 7)), sub=((1, 4), (2, 1), (3, 3)), extra=9),
]


@pytest.fixture(scope="module")
def refd(ref):
    if not ref.have_ref_driver():
        pytest.skip("oracle/_ref/ref_driver not built")
    return ref


def px_problem(case, dtype, nprow, npcol, order, seed=None):
    """Global matrices, block-cyclic descriptions and the dense expectation of one p?gemm case."""
    m, n, k, ta, tb = case["m"], case["n"], case["k"], case["ta"], case["tb"]
    alpha, beta = case["alpha"], case["beta"]
    if dtype in "zc" and alpha != 0.0:
        alpha = alpha * (1 + 0.5j)
    extra = case.get("extra", 0)
    subs = case.get("sub", ((1, 1), (1, 1), (1, 1)))
    (ia, ja), (ib, jb), (ic, jc) = subs
    am, an = (m, k) if ta == "N" else (k, m)
    bm, bn = (k, n) if tb == "N" else (n, k)
    shapes = [(max(am, 1) + ia - 1 + extra, max(an, 1) + ja - 1 + extra), (max(bm, 1) + ib - 1 + extra, max(bn, 1) + jb - 1 + extra),
              (max(m, 1) + ic - 1 + extra, max(n, 1) + jc - 1 + extra)]
    rng = np.random.default_rng(m * 31 + n * 17 + k if seed is None else seed)
    G = [sim.random_values(rng, s, dtype) for s in shapes]
    rsrc, csrc = (1 % nprow, 1 % npcol) if extra else (0, 0)
    bc = [sim.BlockCyclic(s[0], s[1], blk[0], blk[1], nprow, npcol, order, rsrc, csrc, lld_pad=1) for s, blk in zip(shapes, case["blk"])]
    want = G[2].copy()
    wide = np.complex128 if dtype in "zc" else np.float64
    As = sim.apply_op(G[0][ia - 1:ia - 1 + am, ja - 1:ja - 1 + an], ta).astype(wide)
    Bs = sim.apply_op(G[1][ib - 1:ib - 1 + bm, jb - 1:jb - 1 + bn], tb).astype(wide)
    want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = (alpha * (As @ Bs) + beta * G[2][ic - 1:ic - 1 + m, jc - 1:jc - 1 + n].astype(wide)).astype(want.dtype)
    return G, bc, subs, alpha, beta, want


@pytest.mark.parametrize("nprow,npcol,order", [(1, 1, "R"), (2, 1, "R"), (2, 2, "C"), (2, 4, "R")])
@pytest.mark.parametrize("dtype", ["d", "z"])
def test_reference_pxgemm_on_ranks(lib, refd, nprow, npcol, order, dtype):
    P = nprow * npcol
    for case in PX:
        G, bc, ((ia, ja), (ib, jb), (ic, jc)), alpha, beta, want = px_problem(case, dtype, nprow, npcol, order)
        locs = [[bc[x].scatter(G[x], r) for r in range(P)] for x in range(3)]
        descs = [[bc[x].desc(r) for r in range(P)] for x in range(3)]
        outs, _ = refd.ref_pxgemm_ranks(dtype, order, nprow, npcol, case["ta"], case["tb"], case["m"], case["n"], case["k"], alpha, locs[0], ia, ja,
                                        descs[0], locs[1], ib, jb, descs[1], beta, locs[2], ic, jc, descs[2])
        got = np.zeros_like(G[2])
        for r in range(P):
            bc[2].gather_into(got, outs[r], r)
        assert np.array_equal(got, want), case


def _pdgemm_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pdgemm_cases.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("chunk", range(5))
def test_reference_pdgemm_parameter_sets(lib, refd, chunk):
    """The 50 parameter sets of the reference's tests/pdgemm.cpp (tests/golden/pdgemm_cases.json, extracted by make_pdgemm_cases.py),
    run through the UNMODIFIED reference cosma::pxgemm on the grid each set names (1-16 minimpi ranks) and compared with the dense
    definition on integer-valued matrices. The reference checks them against a vendor ScaLAPACK, absent here: this is what pins the
    reference-with-stand-ins as the oracle of our p?gemm on exactly these cases (tests/cpp/test_pxgemm.cpp runs the same sets through
    our C++ API; sub-matrix offsets, rectangular grids, k = 43417, m|n|k = 0, alpha = 0 and beta = 0 included)."""
    cases = _pdgemm_cases()
    assert len(cases) == 50
    for idx in range(chunk * 10, chunk * 10 + 10):
        c = cases[idx]
        nprow, npcol, order, P = c["p_rows"], c["p_cols"], c["order"], c["p_rows"] * c["p_cols"]
        rng = np.random.default_rng(1000 + idx)
        shapes = [(c["ma"], c["na"]), (c["mb"], c["nb"]), (c["mc"], c["nc"])]
        blks = [(c["bma"], c["bna"]), (c["bmb"], c["bnb"]), (c["bmc"], c["bnc"])]
        srcs = [(c["src_ma"], c["src_na"]), (c["src_mb"], c["src_nb"]), (c["src_mc"], c["src_nc"])]
        G = [sim.random_values(rng, s, "d") for s in shapes]
        bc = [sim.BlockCyclic(s[0], s[1], b[0], b[1], nprow, npcol, order, r[0], r[1]) for s, b, r in zip(shapes, blks, srcs)]
        m, n, k, ta, tb, alpha, beta = c["m"], c["n"], c["k"], c["ta"], c["tb"], c["alpha"], c["beta"]
        (ia, ja), (ib, jb), (ic, jc) = (c["ia"], c["ja"]), (c["ib"], c["jb"]), (c["ic"], c["jc"])
        am, an = (m, k) if ta == "N" else (k, m)
        bm, bn = (k, n) if tb == "N" else (n, k)
        want = G[2].copy()
        if m and n:
            As = sim.apply_op(G[0][ia - 1:ia - 1 + am, ja - 1:ja - 1 + an], ta)
            Bs = sim.apply_op(G[1][ib - 1:ib - 1 + bm, jb - 1:jb - 1 + bn], tb)
            sub = G[2][ic - 1:ic - 1 + m, jc - 1:jc - 1 + n]
            want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = (alpha * (As @ Bs) if k and alpha != 0 else 0) + (beta * sub if beta != 0 else 0)
        locs = [[bc[x].scatter(G[x], r) for r in range(P)] for x in range(3)]
        descs = [[bc[x].desc(r) for r in range(P)] for x in range(3)]
        outs, _ = refd.ref_pxgemm_ranks("d", order, nprow, npcol, ta, tb, m, n, k, alpha, locs[0], ia, ja, descs[0], locs[1], ib, jb, descs[1], beta,
                                        locs[2], ic, jc, descs[2])
        got = np.zeros_like(G[2])
        for r in range(P):
            bc[2].gather_into(got, outs[r], r)
        assert np.allclose(got, want, rtol=1e-14, atol=0), (idx, c)
