"""Our p?gemm pipeline executed END TO END ON THE CPU, all ranks in lock-step, on the 50 parameter sets of the reference's
tests/pdgemm.cpp (tests/golden/pdgemm_cases.json) -- with the very plans the GPU executes:

    block-cyclic sub(A), sub(B)  --costa transform plan (op = transa / transb, alpha 1, beta 0)-->  COSMA's native layout
    compiled multiply schedule (allgather / GEMM / reduce ops on the arenas, alpha 1, beta 0)
    native C  --costa transform plan (alpha, beta)-->  block-cyclic sub(C)

(the three phases of cosma_b200_p?gemm, csrc/layout_multiply.cu, reference cosma_pxgemm.cpp:16-388). The transform plans come from
cosma_b200_transform_plan_create (planning only, one per rank), the schedules from cosma_b200_plan_create; tests/costa_sim.py and
tests/schedule_sim.py interpret them with numpy. Results are compared with the dense definition on integer-valued matrices
(exact) -- the same expectation the unmodified reference meets on these sets in tests/test_ref_scalapack_cpu.py."""
import json
import os

import numpy as np
import pytest

import costa_sim as sim
import schedule_sim
from cosma_b200 import costa
from cosma_b200.distributed import MultiplyPlan

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    with open(os.path.join(HERE, "golden", "pdgemm_cases.json")) as f:
        return json.load(f)["cases"]


def _native_layouts(plans, label, arenas, x, shape, P):
    """COSMA's native layout of one matrix as per-rank costa layouts whose blocks point into the ranks' arenas."""
    P_used = plans[0].P_used
    per_rank = [plans[0].local_blocks(label, r) for r in range(P)]
    rs = sorted({b[0] for bl in per_rank for b in bl} | {shape[0]})
    cs = sorted({b[2] for bl in per_rank for b in bl} | {shape[1]})
    owners = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
    for r, bl in enumerate(per_rank):
        for (r0, r1, c0, c1) in bl:
            owners[rs.index(r0), cs.index(c0)] = r
    layouts = []
    for r in range(P):
        blocks, pos = [], 0
        for (r0, r1, c0, c1) in (per_rank[r] if r < P_used else []):
            nr, nc = r1 - r0 + 1, c1 - c0 + 1
            blocks.append((rs.index(r0), cs.index(c0), arenas[r][x].ctypes.data + pos * 8, nr))
            pos += nr * nc
        layouts.append(costa.custom_layout(rs, cs, owners, blocks, "C"))
    return layouts


def run_pdgemm_on_cpu(oracle, c, seed, steps="", stats=None):
    nprow, npcol, order, P = c["p_rows"], c["p_cols"], c["order"], c["p_rows"] * c["p_cols"]
    m, n, k, ta, tb, alpha, beta = c["m"], c["n"], c["k"], c["ta"], c["tb"], c["alpha"], c["beta"]
    subs = ((c["ia"], c["ja"]), (c["ib"], c["jb"]), (c["ic"], c["jc"]))
    shapes = [(c["ma"], c["na"]), (c["mb"], c["nb"]), (c["mc"], c["nc"])]
    blks = [(c["bma"], c["bna"]), (c["bmb"], c["bnb"]), (c["bmc"], c["bnc"])]
    srcs = [(c["src_ma"], c["src_na"]), (c["src_mb"], c["src_nb"]), (c["src_mc"], c["src_nc"])]
    am, an = (m, k) if ta == "N" else (k, m)
    bm, bn = (k, n) if tb == "N" else (n, k)
    subdims = [(am, an), (bm, bn), (m, n)]
    rng = np.random.default_rng(seed)
    G = [sim.random_values(rng, s, "d") for s in shapes]
    want = G[2].copy()
    (ia, ja), (ib, jb), (ic, jc) = subs
    As = sim.apply_op(G[0][ia - 1:ia - 1 + am, ja - 1:ja - 1 + an], ta)
    Bs = sim.apply_op(G[1][ib - 1:ib - 1 + bm, jb - 1:jb - 1 + bn], tb)
    want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = alpha * (As @ Bs) + (beta * G[2][ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] if beta != 0 else 0)
    # the caller's block-cyclic local arrays (C poisoned with NaN inside sub(C) when beta == 0: it must not be read)
    bc = [sim.BlockCyclic(s[0], s[1], b[0], b[1], nprow, npcol, order, r[0], r[1], lld_pad=1) for s, b, r in zip(shapes, blks, srcs)]
    Cin = G[2].copy()
    if beta == 0:
        Cin[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = np.nan
    locs = [[bc[x].scatter((G[0], G[1], Cin)[x], r) for r in range(P)] for x in range(3)]
    user = [[costa.block_cyclic_layout(shapes[x][0], shapes[x][1], blks[x][0], blks[x][1], subs[x][0], subs[x][1], subdims[x][0], subdims[x][1],
                                       nprow, npcol, order, srcs[x][0], srcs[x][1], locs[x][r].ctypes.data, bc[x].local_shape(r)[0], "C", r, 8)
             for r in range(P)] for x in range(3)]
    # phase 0: the multiply plans (automatic strategy, as p?gemm uses) and their arenas
    plans = [MultiplyPlan(None, m, n, k, steps, "d", rank=r, nranks=P, allocate=False) for r in range(P)]
    arenas = [[np.zeros(max(pl.arena_elements[x], 1), dtype=np.float64) for x in range(3)] for pl in plans]
    native = [_native_layouts(plans, "ABC"[x], arenas, x, ((m, k), (k, n), (m, n))[x], P) for x in range(3)]
    # phase 1: relayout op(sub(A)), op(sub(B)) into the native layout -- one exchange, two transforms
    tin = []
    for r in range(P):
        tp = costa.TransformPlan(None, "d", [(user[0][r], native[0][r], ta, 1.0, 0.0), (user[1][r], native[1][r], tb, 1.0, 0.0)], rank=r, nranks=P)
        tin.append(tp.export()); tp.destroy()
        if stats is not None:  # how much of A alone would leave this rank
            ta_only = costa.TransformPlan(None, "d", [(user[0][r], native[0][r], ta, 1.0, 0.0)], rank=r, nranks=P)
            stats["a_remote"] = stats.get("a_remote", 0) + ta_only.stats()["remote_elements"]
            stats["a_local"] = stats.get("a_local", 0) + ta_only.stats()["local_elements"]
            ta_only.destroy()
    sim.simulate(oracle, "d", tin, [(1.0, 0.0), (1.0, 0.0)])
    # phase 2: the compiled schedule, alpha = 1, beta = 0
    schedule_sim.run_schedules(plans, arenas, 1.0, 0.0)
    # phase 3: native C -> sub(C) with the caller's alpha and beta
    tout = []
    for r in range(P):
        tp = costa.TransformPlan(None, "d", [(native[2][r], user[2][r], "N", alpha, beta)], rank=r, nranks=P)
        tout.append(tp.export()); tp.destroy()
    sim.simulate(oracle, "d", tout, [(alpha, beta)])
    strategy = plans[0].strategy
    for pl in plans:
        pl.destroy()
    got = np.zeros_like(G[2])
    for r in range(P):
        bc[2].gather_into(got, locs[2][r], r)
    return got, want, strategy


@pytest.mark.parametrize("chunk", range(5))
def test_pdgemm_parameter_sets_in_lock_step(lib, oracle, chunk):
    cases = _cases()
    ran = 0
    for idx in range(chunk * 10, chunk * 10 + 10):
        c = cases[idx]
        if c["m"] == 0 or c["n"] == 0 or c["k"] == 0 or c["alpha"] == 0:
            continue  # no product: p?gemm only scales sub(C) (corner cases of the BLAS standard; covered by the GPU tests)
        got, want, strategy = run_pdgemm_on_cpu(oracle, c, 2000 + idx)
        assert np.allclose(got, want, rtol=1e-14, atol=0), (idx, strategy, c)
        assert not np.isnan(got).any()
        ran += 1
    assert ran >= 2


def test_most_parameter_sets_have_a_product():
    cases = _cases()
    assert len(cases) == 50
    assert sum(1 for c in cases if c["m"] and c["n"] and c["k"] and c["alpha"] != 0) >= 40


@pytest.mark.parametrize("order,prefix", [("R", "sm4,sk6,pm2,pk2"), ("C", "sm4,sk6,pk2,pm2")])
def test_adapted_strategy_leaves_the_largest_operand_in_place(lib, oracle, order, prefix):
    """COSMA_ADAPT_STRATEGY in miniature (the size threshold of the reference, 1e7 elements per rank, is out of reach of a numpy
    simulation, so the prefix is written out by the same rule: sequential steps over the block-cycle repetitions, then the process
    grid): with the adapted strategy COSMA's native layout of A IS the caller's block-cyclic layout -- its relayout moves nothing
    between ranks -- and the product is still exact; with the automatic strategy most of A travels."""
    c = dict(ma=64, na=96, mb=96, nb=32, mc=64, nc=32, bma=8, bna=8, bmb=8, bnb=8, bmc=8, bnc=8, ia=1, ja=1, ib=1, jb=1, ic=1, jc=1, m=64, n=32, k=96,
             ta="N", tb="N", alpha=1.0, beta=1.0, p_rows=2, p_cols=2, order=order, src_ma=0, src_na=0, src_mb=0, src_nb=0, src_mc=0, src_nc=0)
    adapted, automatic = {}, {}
    got, want, strategy = run_pdgemm_on_cpu(oracle, c, 7, steps=prefix, stats=adapted)
    assert strategy.startswith(prefix) and np.array_equal(got, want)
    assert adapted["a_remote"] == 0 and adapted["a_local"] == 64 * 96
    got, want, strategy = run_pdgemm_on_cpu(oracle, c, 7, stats=automatic)
    assert np.array_equal(got, want) and automatic["a_remote"] > 0


def _random_layout(rng, rows, cols, P, dtype):
    rs = sim.random_split(rng, rows, int(rng.integers(1, 5)))
    cs = sim.random_split(rng, cols, int(rng.integers(1, 5)))
    owners = rng.integers(0, P, size=(len(rs) - 1, len(cs) - 1))
    return sim.DistMatrix(rs, cs, owners, P, dtype, "C", pad=int(rng.integers(0, 3)))


def _relabelling(lib, P, user, native_grids, ta, tb):
    """The permutation csrc/layout_multiply.cu derives (reference multiply.cpp:136-152): volumes of op(A), op(B) into COSMA's grids
    and of C out of it, then the matching. user / native_grids: (rowsplit, colsplit, owners) per matrix."""
    import ctypes

    def ptr(a, t=ctypes.c_int):
        return a.ctypes.data_as(ctypes.POINTER(t))

    def volume(ga, gb, trans):
        out = np.zeros(P * P, dtype=np.int64)
        ga = [np.ascontiguousarray(x, dtype=np.int32).reshape(-1) for x in ga]
        gb = [np.ascontiguousarray(x, dtype=np.int32).reshape(-1) for x in gb]
        assert lib.cosma_b200_comm_volume(len(ga[0]) - 1, len(ga[1]) - 1, ptr(ga[0]), ptr(ga[1]), ptr(ga[2]), len(gb[0]) - 1, len(gb[1]) - 1,
                                          ptr(gb[0]), ptr(gb[1]), ptr(gb[2]), ctypes.c_char(trans.encode()), P, ptr(out, ctypes.c_longlong)) == 0
        return out.reshape(P, P)
    total = volume(user[0], native_grids[0], ta) + volume(user[1], native_grids[1], tb) + volume(native_grids[2], user[2], "N")
    perm = np.zeros(P, dtype=np.int32)
    flag = ctypes.c_int(0)
    assert lib.cosma_b200_optimal_reordering(P, ptr(np.ascontiguousarray(total), ctypes.c_longlong), ptr(perm), ctypes.byref(flag)) == 0
    assert np.array_equal(perm[perm], np.arange(P))  # a matching: always an involution
    return perm if flag.value else np.arange(P, dtype=np.int32)


@pytest.mark.parametrize("relabel", [False, True], ids=["labels_kept", "relabelled"])
@pytest.mark.parametrize("P", [3, 4, 6, 8])
@pytest.mark.parametrize("dtype,ta,tb", [("d", "N", "N"), ("d", "T", "N"), ("z", "C", "N"), ("z", "N", "C"), ("d", "T", "T")])
def test_multiply_using_layout_in_lock_step(lib, oracle, P, dtype, ta, tb, relabel):
    """cosma::multiply_using_layout (reference multiply.cpp:78-213) END TO END on the CPU with the plans the GPU executes: random block
    layouts with random owners for A, B and C (padded leading dimensions), op(A), op(B) relayouted into COSMA's native layout, the
    compiled schedule, the result relayouted into C's layout with (alpha, beta); exact on integer-valued matrices, real and complex."""
    rng = np.random.default_rng(100 * P + ord(ta) + 3 * ord(tb))
    npdt = np.float64 if dtype == "d" else np.complex128
    for trial in range(2):
        m, n, k = (int(rng.integers(220, 300)), int(rng.integers(210, 280)), int(rng.integers(200, 320)))
        alpha, beta = ((1.0, 0.0), (2.0, -1.0))[trial]
        A = sim.random_values(rng, (m, k) if ta == "N" else (k, m), dtype)
        B = sim.random_values(rng, (k, n) if tb == "N" else (n, k), dtype)
        C = sim.random_values(rng, (m, n), dtype)
        dA, dB, dC = (_random_layout(rng, X.shape[0], X.shape[1], P, dtype) for X in (A, B, C))
        dA.scatter(A); dB.scatter(B)
        dC.fill_padding(7)
        dC.scatter(C if beta != 0.0 else np.full_like(C, np.nan))
        probe = MultiplyPlan(None, m, n, k, "", dtype, rank=0, nranks=P, allocate=False)
        grids = []
        for x, shape in enumerate(((m, k), (k, n), (m, n))):
            per_rank = [probe.local_blocks("ABC"[x], r) for r in range(P)]
            rs = sorted({b[0] for bl in per_rank for b in bl} | {shape[0]})
            cs = sorted({b[2] for bl in per_rank for b in bl} | {shape[1]})
            owners = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
            for r, bl in enumerate(per_rank):
                for (r0, r1, c0, c1) in bl:
                    owners[rs.index(r0), cs.index(c0)] = r
            grids.append((rs, cs, owners, per_rank))
        # relabelling as in csrc/layout_multiply.cu: physical rank r plays COSMA rank perm[r]; COSMA rank q is played by perm[q]
        perm = np.arange(P, dtype=np.int32)
        if relabel:
            perm = _relabelling(lib, P, [(d.rowsplit, d.colsplit, d.owners) for d in (dA, dB, dC)], [g[:3] for g in grids], ta, tb)
        probe.destroy()
        plans = [MultiplyPlan(None, m, n, k, "", dtype, rank=int(perm[r]), nranks=P, allocate=False) for r in range(P)]
        arenas = [[np.zeros(max(pl.arena_elements[x], 1), dtype=npdt) for x in range(3)] for pl in plans]
        eb = 8 if dtype == "d" else 16
        native = []
        for x, (rs, cs, owners, per_rank) in enumerate(grids):
            lays = []
            for r in range(P):
                blocks, pos = [], 0
                for (r0, r1, c0, c1) in per_rank[perm[r]]:
                    nr, nc = r1 - r0 + 1, c1 - c0 + 1
                    blocks.append((rs.index(r0), cs.index(c0), arenas[r][x].ctypes.data + pos * eb, nr))
                    pos += nr * nc
                lays.append(costa.custom_layout(rs, cs, perm[owners], blocks, "C"))
            native.append(lays)
        tin = []
        for r in range(P):
            tp = costa.TransformPlan(None, dtype, [(dA.layout(r), native[0][r], ta, 1.0, 0.0), (dB.layout(r), native[1][r], tb, 1.0, 0.0)], rank=r, nranks=P)
            tin.append(tp.export()); tp.destroy()
        sim.simulate(oracle, dtype, tin, [(1.0, 0.0), (1.0, 0.0)])
        inv = np.argsort(perm)  # the schedules are indexed by COSMA rank
        schedule_sim.run_schedules([plans[inv[q]] for q in range(P)], [arenas[inv[q]] for q in range(P)], 1.0, 0.0)
        tout = []
        for r in range(P):
            tp = costa.TransformPlan(None, dtype, [(native[2][r], dC.layout(r), "N", alpha, beta)], rank=r, nranks=P)
            tout.append(tp.export()); tp.destroy()
        sim.simulate(oracle, dtype, tout, [(alpha, beta)])
        for pl in plans:
            pl.destroy()
        want = alpha * (sim.apply_op(A, ta) @ sim.apply_op(B, tb)) + (beta * C if beta != 0.0 else 0)
        assert np.array_equal(dC.gather(), want.astype(npdt)), (m, n, k, trial)
