"""CPU tests of the COSTA relayout path (no GPU): the oracle against the reference's golden vectors and against the
unmodified reference (oracle/_ref); our host planning layer (block-cyclic layouts, grid overlay, piece lists, per-peer
byte counts) against the reference and against dense global oracles, for every rank, by lock-step simulation.

Reference tests mirrored: libs/COSTA/tests/unit/test_utils.cpp (known-answer arrays); tests/pdgemm.cpp descriptor
cases (sub-matrices, unequal blocks, row/col-major grids, rsrc/csrc) for the layout formulas."""
import ctypes
import json
import os

import numpy as np
import pytest

import costa_sim as sim
from cosma_b200 import costa

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "costa_copy_and_transform.json")) as f:
        return json.load(f)["cases"]


def _golden_io(case):
    if "in" in case:
        return np.array(case["in"], dtype=np.int32)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(100)
    n = case["n_cols"] * case["src_ld"]
    v = np.array([(i + libc.rand()) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32).astype(np.int32)
    assert v[:16].tolist() == case["in_head"], "C library rand() differs from the one the golden file was made with"
    return v


def _check_golden(case, src, out):
    nr, nc, ls, ld = case["n_rows"], case["n_cols"], case["src_ld"], case["dst_ld"]
    if "expected" in case:
        exp = np.array(case["expected"], dtype=np.int32)
        for i in range(nr):
            for j in range(nc):
                idx = i * ld + j if case["dst_ordering"] == "R" else j * ld + i
                assert out[idx] == exp[idx], (case["name"], i, j)
    else:  # predicate of the reference test: out[i*dst_ld + j] == in[j*src_ld + i]
        O = out[:nr * ld].reshape(nr, ld)[:, :nc]
        I = src[:nc * ls].reshape(nc, ls)[:, :nr].T
        assert np.array_equal(O, I)


@pytest.mark.parametrize("case", _golden(), ids=lambda c: c["name"])
def test_oracle_matches_reference_golden_vectors(oracle, case):
    src = _golden_io(case)
    out = np.zeros(max(case["n_rows"], case["n_cols"]) * case["dst_ld"], dtype=np.int32)
    oracle.copy_and_transform(case["n_rows"], case["n_cols"], src, case["src_ld"], case["src_ordering"], out, case["dst_ld"],
                              case["dst_ordering"], case["transpose"], case["conjugate"], case["alpha"], case["beta"])
    _check_golden(case, src, out)


@pytest.mark.parametrize("case", _golden(), ids=lambda c: c["name"])
def test_reference_build_matches_its_golden_vectors(ref, case):
    src = _golden_io(case)
    out = np.zeros(max(case["n_rows"], case["n_cols"]) * case["dst_ld"], dtype=np.int32)
    ref.ref_copy_and_transform("i", case["n_rows"], case["n_cols"], src, case["src_ld"], case["src_ordering"], out, case["dst_ld"],
                               case["dst_ordering"], case["transpose"], case["conjugate"], case["alpha"], case["beta"])
    _check_golden(case, src, out)


@pytest.mark.parametrize("dtype", ["d", "z", "s", "c"])
@pytest.mark.parametrize("src_ord,dst_ord,transpose", [(a, b, t) for a in "CR" for b in "CR" for t in (0, 1)])
def test_oracle_equals_reference_copy_and_transform(oracle, ref, dtype, src_ord, dst_ord, transpose):
    """Oracle restatement vs the unmodified reference kernel, all 8 ordering/transpose combinations, with and without
    conjugation, alpha/beta identity and general. beta != 0 everywhere (the vendored COSTA reads dest when alpha != 1
    even for beta == 0; we do not -- DESIGN.md deviations)."""
    rng = np.random.default_rng(hash((dtype, src_ord, dst_ord, transpose)) % 2**32)
    nr, nc = 37, 53
    dr, dc = (nc, nr) if transpose else (nr, nc)
    sld = (nr if src_ord == "C" else nc) + 3
    dld = (dr if dst_ord == "C" else dc) + 5
    for conj in (0, 1):
        for alpha, beta in ((1.0, 0.0), (2.0, -0.5), (0.75 - 1.5j, 0.25 + 2j) if dtype in "cz" else (-3.0, 1.0)):
            src = sim.random_values(rng, sld * max(nr, nc), dtype, ints=False)
            d0 = sim.random_values(rng, dld * max(dr, dc), dtype, ints=False)
            a, b = d0.copy(), d0.copy()
            oracle.copy_and_transform(nr, nc, src, sld, src_ord, a, dld, dst_ord, transpose, conj, alpha, beta)
            ref.ref_copy_and_transform(dtype, nr, nc, src, sld, src_ord, b, dld, dst_ord, transpose, conj, alpha, beta)
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (conj, alpha, beta)


DESCS = [
    # lld, M, N, ia, ja, sub_m, sub_n, mb, nb, nprow, npcol, order, rsrc, csrc
    (40, 100, 90, 1, 1, 100, 90, 8, 8, 2, 3, "R", 0, 0),
    (64, 100, 90, 1, 1, 100, 90, 32, 16, 2, 2, "C", 0, 0),
    (50, 100, 90, 5, 7, 60, 50, 8, 4, 3, 2, "R", 0, 0),        # sub-matrix not aligned to blocks
    (50, 100, 90, 9, 17, 40, 30, 8, 8, 2, 4, "C", 1, 2),       # rsrc/csrc != 0
    (128, 256, 256, 1, 1, 256, 256, 256, 256, 2, 4, "R", 0, 0),  # one block
    (30, 57, 43, 2, 3, 55, 40, 7, 5, 2, 2, "R", 1, 1),
    (17, 16, 16, 1, 1, 16, 16, 1, 1, 4, 2, "C", 0, 0),         # 1x1 blocks
]


@pytest.mark.parametrize("desc", DESCS)
@pytest.mark.parametrize("data_ordering", ["C", "R"])
def test_scalapack_layout_matches_reference(ref, desc, data_ordering):
    lld, M, N, ia, ja, sm, sn, mb, nb, pr, pc, order, rsrc, csrc = desc
    for rank in range(pr * pc):
        ours = costa.scalapack_grid(lld, M, N, ia, ja, sm, sn, mb, nb, pr, pc, order, rsrc, csrc, data_ordering, rank)
        theirs = ref.ref_scalapack_layout(lld, M, N, ia, ja, sm, sn, mb, nb, pr, pc, order, rsrc, csrc, data_ordering, rank)
        assert ours[0].tolist() == theirs[0].tolist() and ours[1].tolist() == theirs[1].tolist()
        assert np.array_equal(ours[2], theirs[2])
        assert sorted(ours[3]) == sorted(theirs[3])


def test_numroc():
    # ScaLAPACK NUMROC known answers: the pieces of a dimension sum to the dimension for every (nb, source, nprocs)
    for n, nb, src, p in ((100, 8, 0, 3), (57, 7, 1, 2), (16, 1, 3, 4), (5, 8, 0, 4), (0, 4, 0, 2)):
        assert sum(costa.numroc(n, nb, i, src, p) for i in range(p)) == n
    assert [costa.numroc(100, 8, i, 0, 3) for i in range(3)] == [36, 32, 32]


def _random_dist(rng, m, n, P, dtype, ordering, max_parts=5, pad=None):
    rs = sim.random_split(rng, m, int(rng.integers(1, max_parts + 1)))
    cs = sim.random_split(rng, n, int(rng.integers(1, max_parts + 1)))
    owners = rng.integers(0, P, size=(len(rs) - 1, len(cs) - 1))
    return sim.DistMatrix(rs, cs, owners, P, dtype, ordering, pad=int(rng.integers(0, 4)) if pad is None else pad)


@pytest.mark.parametrize("dtype", ["d", "z"])
@pytest.mark.parametrize("op", ["N", "T", "C"])
@pytest.mark.parametrize("orderings", ["CC", "RC", "CR", "RR"])
def test_single_rank_plan_matches_reference_transform(oracle, ref, dtype, op, orderings):
    """P = 1: our plan (grid overlay -> pieces) interpreted with the oracle kernel vs the unmodified costa::transform on
    the same random block layouts. alpha = 1, beta = 0 -> bit-exact."""
    rng = np.random.default_rng(hash((dtype, op, orderings)) % 2**32)
    for trial in range(4):
        m, n = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        sm, sn = (m, n) if op == "N" else (n, m)
        F = _random_dist(rng, sm, sn, 1, dtype, orderings[0])
        T1 = _random_dist(rng, m, n, 1, dtype, orderings[1])
        T2 = sim.DistMatrix(T1.rowsplit, T1.colsplit, T1.owners, 1, dtype, orderings[1], pad=T1.pad)
        G = sim.random_values(rng, (sm, sn), dtype)
        F.scatter(G)
        for T in (T1, T2):
            T.fill_padding(77)
        tp = costa.TransformPlan(None, dtype, [(F.layout(0), T1.layout(0), op, 1.0, 0.0)], rank=0, nranks=1)
        sim.simulate(oracle, dtype, [tp.export()], [(1.0, 0.0)])
        tp.destroy()
        ref.ref_transform_p1(dtype, F.ref_tuple(), T2.ref_tuple(), op, 1.0, 0.0)
        assert np.array_equal(T1.gather(), sim.apply_op(G, op))
        for key in T1.store:
            assert np.array_equal(T1.store[key][0].view(np.uint8), T2.store[key][0].view(np.uint8)), key


@pytest.mark.parametrize("P", [2, 4, 7])
@pytest.mark.parametrize("dtype,op", [("d", "N"), ("d", "T"), ("z", "C"), ("s", "T"), ("c", "C")])
def test_multi_rank_plans_against_dense_oracle(oracle, P, dtype, op):
    """Every rank's plan, executed in lock-step (pack -> exchange -> unpack), reproduces
    target = beta*target + alpha*op(source) on random block layouts with random owners; send and receive byte counts
    of every rank pair agree; two transforms batched in one exchange."""
    rng = np.random.default_rng(1000 * P + ord(op))
    for trial in range(3):
        specs, mats = [], []
        for t in range(2):
            m, n = int(rng.integers(1, 90)), int(rng.integers(1, 90))
            sm, sn = (m, n) if op == "N" else (n, m)
            F = _random_dist(rng, sm, sn, P, dtype, "CR"[int(rng.integers(0, 2))])
            T = _random_dist(rng, m, n, P, dtype, "CR"[int(rng.integers(0, 2))])
            G, H = sim.random_values(rng, (sm, sn), dtype), sim.random_values(rng, (m, n), dtype)
            F.scatter(G)
            T.fill_padding(55)
            T.scatter(H)
            alpha, beta = ((1.0, 0.0), (2.0, 3.0))[t] if trial else ((1.0, 0.0), (1.0, 0.0))[t]
            specs.append((alpha, beta))
            mats.append((F, T, alpha * sim.apply_op(G, op) + beta * H))
        plans = []
        for r in range(P):
            tp = costa.TransformPlan(None, dtype, [(F.layout(r), T.layout(r), op, a, b) for (F, T, _), (a, b) in zip(mats, specs)],
                                     rank=r, nranks=P)
            plans.append(tp.export())
            tp.destroy()
        sim.simulate(oracle, dtype, plans, specs)
        for F, T, want in mats:
            assert np.array_equal(T.gather(), want.astype(T.gather().dtype))


def test_block_cyclic_to_cosma_layout_cfg5_shape(oracle):
    """The pzgemm configuration in miniature: A stored k x m block-cyclic on a 2 x 4 row-major grid, conjugate-transposed
    into COSMA's native layout for P = 8 (strategy pm2,pn2,pk2 -> Mapper grid), all 8 ranks simulated."""
    from cosma_b200 import planning
    P, m, k, nb = 8, 96, 80, 8
    rng = np.random.default_rng(5)
    lay = planning.mapper_layout("A", m, 64, k, P, "pm2,pn2,pk2")  # per rank: list of (r0, r1, c0, c1) inclusive
    rows = sorted({b[0] for blocks in lay for b in blocks} | {m})
    cols = sorted({b[2] for blocks in lay for b in blocks} | {k})
    owners = np.zeros((len(rows) - 1, len(cols) - 1), dtype=np.int32)
    for r, blocks in enumerate(lay):
        for (r0, r1, c0, c1) in blocks:
            owners[rows.index(r0), cols.index(c0)] = r
    T = sim.DistMatrix(rows, cols, owners, P, "z", "C")
    G = sim.random_values(rng, (k, m), "z")
    # block-cyclic source: one local array per rank
    locs, lays = [], []
    for r in range(P):
        pr, pc = r // 4, r % 4
        lr, lc = costa.numroc(k, nb, pr, 0, 2), costa.numroc(m, nb, pc, 0, 4)
        lld = lr + 2
        loc = np.full(lld * max(lc, 1), 99, dtype=np.complex128)
        locs.append((loc, lld))
        lays.append(costa.block_cyclic_layout(k, m, nb, nb, 1, 1, k, m, 2, 4, "R", 0, 0, loc.ctypes.data, lld, "C", r, 16))
    for r in range(P):
        loc, lld = locs[r]
        for (bi, bj, addr, ld) in lays[r].blocks:
            off = (addr - loc.ctypes.data) // 16
            r0, r1, c0, c1 = lays[r].rowsplit[bi], lays[r].rowsplit[bi + 1], lays[r].colsplit[bj], lays[r].colsplit[bj + 1]
            for j in range(c1 - c0):
                loc[off + j * lld: off + j * lld + (r1 - r0)] = G[r0:r1, c0 + j]
    plans = []
    for r in range(P):
        tp = costa.TransformPlan(None, "z", [(lays[r], T.layout(r), "C", 1.0, 0.0)], rank=r, nranks=P)
        plans.append(tp.export())
        tp.destroy()
    sim.simulate(oracle, "z", plans, [(1.0, 0.0)])
    assert np.array_equal(T.gather(), G.conj().T)


def test_plan_errors(lib):
    F = sim.DistMatrix([0, 4], [0, 6], [[0]], 1, "d")
    T = sim.DistMatrix([0, 5], [0, 6], [[0]], 1, "d")
    with pytest.raises(Exception, match="target is"):
        costa.TransformPlan(None, "d", [(F.layout(0), T.layout(0), "N", 1.0, 0.0)], rank=0, nranks=1)
    T2 = sim.DistMatrix([0, 4], [0, 6], [[3]], 4, "d")
    with pytest.raises(Exception, match="owner"):
        costa.TransformPlan(None, "d", [(F.layout(0), T2.layout(0), "N", 1.0, 0.0)], rank=0, nranks=2)
