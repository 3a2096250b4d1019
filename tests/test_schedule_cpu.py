"""The compiled multiply schedule (cosma::Schedule via the C ABI) on CPU: every rank's op list is interpreted in
lockstep with numpy (tests/schedule_sim.py) and the gathered C must equal the dense product EXACTLY on
integer-valued inputs -- for the reference's 40 distributed test cases (tests/multiply.cpp:142-321, alpha=beta=1 as in
utils/cosma_utils.hpp:43-44), its mixed sequential/parallel all-types case (tests/scalar_matmul.cpp) and small
versions of the BASELINE strategies."""
import numpy as np
import pytest

from cases import REFERENCE_MULTIPLY_CASES, SCALAR_MATMUL_CASE
from schedule_sim import simulate

IDS = lambda c: "%dx%dx%d_P%d_%s" % (c[0], c[1], c[2], c[3], c[4] or "auto")


@pytest.mark.parametrize("case", REFERENCE_MULTIPLY_CASES, ids=IDS)
def test_reference_multiply_cases(lib, case):
    m, n, k, P, steps = case
    got, want, _ = simulate(m, n, k, P, steps, alpha=1.0, beta=1.0)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", ["d", "z"])
@pytest.mark.parametrize("beta", [0.0, 1.0, 2.0])
def test_scalar_matmul_strategy(lib, dtype, beta):
    m, n, k, P, steps = SCALAR_MATMUL_CASE
    got, want, _ = simulate(m, n, k, P, steps, alpha=1.0, beta=beta, dtype=dtype)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", [
    (512, 512, 512, 8, "pm2,pn2,pk2"), (512, 512, 512, 4, "pn2,pk2"), (512, 512, 512, 2, "pk2"), (256, 256, 4096, 8, "pk8"),
    (384, 256, 640, 8, "sk2,sm2,pk2,pm4"), (300, 301, 302, 6, "pk3,pm2"), (97, 101, 103, 4, "pk2,pk2"),
    (64, 64, 640, 4, "pk2,sk5,pk2"), (250, 130, 70, 8, "pn2,sm3,pm2,sk2,pk2"),
], ids=IDS)
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_baseline_shaped_strategies(lib, case, beta):
    m, n, k, P, steps = case
    got, want, _ = simulate(m, n, k, P, steps, alpha=2.0, beta=beta)
    assert np.array_equal(got, want)


def test_idle_ranks_when_problem_is_small(lib):
    # automatic strategy drops ranks when a local dimension would fall under COSMA_MIN_LOCAL_DIMENSION
    got, want, P_used = simulate(300, 300, 300, 8, "", alpha=1.0, beta=0.0)
    assert P_used < 8
    assert np.array_equal(got, want)


def test_arena_is_bounded(lib):
    """Stack allocation of communication buffers: the arena never exceeds initial + sum over nesting depth."""
    from cosma_b200.distributed import MultiplyPlan
    pl = MultiplyPlan(None, 32768, 32768, 32768, "", "d", rank=3, nranks=8, allocate=False)
    assert pl.strategy == "pm2,pn2,pk2"
    local = 16384 * 8192
    assert pl.initial_elements == [local, local, local]
    # A expanded once (pn2), B once (pm2): arena = local + 2*local; C expanded once (pk2) plus the staging slice the
    # reduce-scatter lands in when beta != 0: local + 2*local + local
    assert pl.arena_elements == [3 * local, 3 * local, 4 * local]
    kinds = [(op["kind"], op.get("matrix"), op.get("ring")) for op in pl.ops()]
    assert kinds == [("allgather", 1, [3, 7]), ("allgather", 0, [1, 3]), ("gemm", None, None), ("reduce", 2, [2, 3])]
    pl.destroy()


def test_statistics_tool(lib, capsys):
    """cosma_b200.statistics (the reference's miniapp/cosma_statistics.cpp in spirit): BASELINE configs[2] at 8 ranks -- two allgathers of
    2.1 GB, one 16384^3 GEMM, one reduce-scatter; per rank 3/2 x 2.1 GB on the wire and 10.7 GB of arenas."""
    from cosma_b200 import statistics
    d = statistics.describe(32768, 32768, 32768, 8)
    assert d["strategy"] == "pm2,pn2,pk2" and d["P_used"] == 8 and len(d["lines"]) == 4
    assert d["wire_bytes"] == 3 * 16384 * 8192 * 8 and abs(d["flops"] - 2.0 * 16384 ** 3) < 1
    assert d["arena_bytes"] == (3 * 16384 * 8192 + 3 * 16384 * 16384 - 0) * 8 or d["arena_bytes"] > 9e9
    assert statistics.main(["-m", "8192", "-n", "8192", "-k", "1048576", "-P", "8"]) == 0
    out = capsys.readouterr().out
    assert "strategy  : pk8" in out and "reduce     C  ring of 8" in out


def test_statistics_pxgemm_relayout(lib, capsys):
    """--pxgemm: relayout volumes of BASELINE configs[4] (pzgemm 16384^3, 2 x 4 grid, 256^2 blocks, A conjugate-transposed). The adapted
    strategy leaves A where it is; a COSMA layout handed in under reversed rank labels is found again by the relabelling."""
    from cosma_b200 import planning, statistics
    r = statistics.pxgemm_relayout(16384, 16384, 16384, 8, 2, 4, "R", "CN", ((256, 256),) * 3)
    assert r["strategy"] == "pm2,pn2,pk2" and r["elements"] == 3 * 16384 ** 2
    assert [x[0] for x in r["matrices"]] == ["A", "B", "C"] and all(x[2] + x[3] == 16384 ** 2 for x in r["matrices"])
    assert not r["reordered"] and r["stay_relabelled"] == r["stay"]            # all rank pairs exchange the same volume
    ad = statistics.pxgemm_relayout(16384, 16384, 16384, 8, 2, 4, "R", "CN", ((256, 256),) * 3, "sk32,sm16,pk2,pm4")
    assert ad["matrices"][0][3] == 0 and ad["stay"] > r["stay"]                # A does not move
    # native layout against itself with the rank labels reversed: nothing is in place, relabelling recovers everything
    _, nat = statistics.layouts(4096, 4096, 4096, 8)
    rs, cs, ow = nat["A"]
    rev = (rs, cs, [[7 - o for o in row] for row in ow])
    vol = planning.comm_volume(rev, (rs, cs, ow), "N", 8)
    perm, flag = planning.optimal_reordering(vol)
    assert flag and sum(vol[u][u] for u in range(8)) == 0 and perm == [7 - u for u in range(8)]
    assert statistics.main(["-m", "2048", "-n", "2048", "-k", "2048", "-P", "4", "--pxgemm", "--block_a", "64,64", "--transpose", "TN"]) == 0
    out = capsys.readouterr().out
    assert "p?gemm    : grid 2 x 2 (R), op = TN" in out and "relayout of C" in out and "in place:" in out


def test_over_divided_dimension_is_refused(lib):
    """m = 2 cut into 3 sequential parts: the reference's Strategy accepts it and its multiply then returns a wrong product
    (Interval::subinterval returns the whole interval, interval.cpp:84-98). The plan refuses with a clear message instead."""
    from cosma_b200.distributed import MultiplyPlan
    with pytest.raises(Exception, match="divides dimension m = 2 into 3 parts"):
        MultiplyPlan(None, 2, 14, 55, "sm3", "d", rank=0, nranks=1, allocate=False)
    pl = MultiplyPlan(None, 3, 14, 55, "sm3", "d", rank=0, nranks=1, allocate=False)  # exactly one row per part is fine
    pl.destroy()


@pytest.mark.parametrize("P,steps,m,n,k,c", [(8, "pm2,pn2,pk2", 64, 96, 80, 4), (4, "pn2,pk2", 48, 64, 40, 2), (2, "pk2", 40, 36, 64, 3),
                                             (8, "pk8", 32, 48, 128, 3), (8, "pm2,pn2,pk2", 60, 72, 56, 3), (4, "pm2,pk2", 36, 40, 44, 5),
                                             (2, "pm2", 40, 36, 64, 3), (6, "pm3,pn2", 30, 24, 50, 2), (4, "pm2,pn2", 28, 32, 20, 4)])
def test_column_panels_of_the_local_buffers_are_independent_subproblems(lib, P, steps, m, n, k, c):
    """The plan for end-to-end pipelining at N > 1: cut every rank's LOCAL C into c column chunks (contiguous in the
    column-major local buffer) and its local B into the matching columns -- for every C column range [c0, c1) that lies inside the rank's
    B columns (one per member of the k-ring that shares this B), the j-th c-th of it; each piece is contiguous in local B. (In general,
    e.g. pm steps where B is divided more finely than C, both sides are cut along the common refinement of the B and C column ranges:
    cosma_b200_plan_host_panel returns piece lists for both.) Chunk j of all ranks is then exactly the native layout of the smaller problem (m, n / c, k) under the same strategy (a product over a subset of the
    columns), so c sub-plans, with A uploaded and gathered once, give the full product bit for bit while chunk j + 1 travels up and
    chunk j - 1 travels down under the GEMM of chunk j. (Taking contiguous c-ths of local B instead is WRONG whenever C is divided
    more finely than B, e.g. pn2,pk2: the values land on other ranks.)"""
    import numpy as np
    from cosma_b200.distributed import MultiplyPlan, fill_local_from_global
    from schedule_sim import run_schedules
    rng = np.random.default_rng(P * 1000 + n)
    Ag, Bg = rng.integers(-5, 6, size=(m, k)).astype(np.float64), rng.integers(-5, 6, size=(k, n)).astype(np.float64)
    Cg = rng.integers(-5, 6, size=(m, n)).astype(np.float64)
    alpha, beta = 2.0, -1.0
    full = [MultiplyPlan(None, m, n, k, steps, "d", rank=r, nranks=P, allocate=False) for r in range(P)]
    sub = [MultiplyPlan(None, m, n // c, k, steps, "d", rank=r, nranks=P, allocate=False) for r in range(P)]
    assert full[0].strategy == sub[0].strategy == steps
    arenas = []
    for pl in full:
        bufs = [np.zeros(max(pl.arena_elements[x], 1)) for x in range(3)]
        for x, (label, G) in enumerate((("A", Ag), ("B", Bg), ("C", Cg))):
            fill_local_from_global(pl, label, bufs[x], G)
        arenas.append(bufs)
    locals_in = [[a[x][:pl.initial_elements[x]].copy() for x in range(3)] for a, pl in zip(arenas, full)]
    run_schedules(full, arenas, alpha, beta)
    want = [a[2][:pl.initial_elements[2]].copy() for a, pl in zip(arenas, full)]
    blocks = {x: [full[r].local_blocks(x) for r in range(P)] for x in "BC"}
    for r in range(P):
        assert sub[r].initial_elements[0] == full[r].initial_elements[0]                      # A is shared by all chunks
        assert sub[r].initial_elements[1] * c == full[r].initial_elements[1] and sub[r].initial_elements[2] * c == full[r].initial_elements[2]
        assert len(blocks["B"][r]) == 1 and len(blocks["C"][r]) == 1                          # one column-major block each
    from panel_cut import host_panel
    got = [np.empty_like(w) for w in want]
    c_ranges = sorted({(b[0][2], b[0][3] + 1) for b in blocks["C"]})
    for j in range(c):
        sa, cuts = [], []
        for r in range(P):
            bufs = [np.zeros(max(sub[r].arena_elements[x], 1)) for x in range(3)]
            bufs[0][:sub[r].initial_elements[0]] = locals_in[r][0]
            # the pieces as the library plans them (cosma_b200_plan_host_panel)
            ok, planned, cplanned = host_panel(lib, full[r].handle, c, j)
            assert ok
            if all(bc[0][3] - bc[0][2] <= bb[0][3] - bb[0][2] for bb, bc in zip(blocks["B"], blocks["C"])) and not any(
                    cr[0] < blocks["B"][q][0][2] < cr[1] for q in range(P) for cr in c_ranges):
                # C ranges nest inside B ranges: the rule stated above, derived here from the block tables
                (r0, r1, b0, b1) = blocks["B"][r][0]
                rows, pos, expect = r1 - r0 + 1, 0, []
                for (c0, c1) in c_ranges:
                    if c0 < b0 or c1 > b1 + 1:
                        continue
                    w = (c1 - c0) // c
                    expect.append(((c0 - b0 + j * w) * rows, w * rows, pos))
                    pos += w * rows
                nc = sub[r].initial_elements[2]
                assert planned == expect and cplanned == [(j * nc, nc, 0)]
            assert sum(p[1] for p in planned) == sub[r].initial_elements[1] and sum(p[1] for p in cplanned) == sub[r].initial_elements[2]
            for (src, ln, dst) in planned:
                bufs[1][dst:dst + ln] = locals_in[r][1][src:src + ln]
            for (src, ln, dst) in cplanned:
                bufs[2][dst:dst + ln] = locals_in[r][2][src:src + ln]
            cuts.append(cplanned)
            sa.append(bufs)
        run_schedules(sub, sa, alpha, beta)
        for r in range(P):
            for (src, ln, dst) in cuts[r]:
                got[r][src:src + ln] = sa[r][2][dst:dst + ln]
    for r in range(P):
        assert np.array_equal(got[r], want[r]), r
    # layouts that cannot be cut: a width that c does not divide
    assert not host_panel(lib, full[0].handle, 7 if n % 7 else 11, 0)[0]
    for pl in full + sub:
        pl.destroy()


@pytest.mark.parametrize("seed", range(13))
def test_column_panels_randomised(lib, seed):
    """Random shapes, rank counts, strategies (automatic and explicit, parallel and sequential steps) and panel counts: whenever
    cosma_b200_plan_host_panel calls a layout eligible on EVERY rank and the sub-problem keeps the strategy, the panels planned by the
    library reproduce the full product bit for bit in lock-step. (Idle ranks, several GEMMs, several blocks per rank, indivisible widths
    must come back as not eligible -- never as a wrong product.)"""
    import numpy as np
    from cosma_b200.distributed import MultiplyPlan, fill_local_from_global
    from panel_cut import host_panel
    from schedule_sim import run_schedules
    rng = np.random.default_rng(7000 + seed)
    checked = 0
    for _ in range(10):
        P = int(rng.choice([2, 3, 4, 6, 8, 12]))
        c = int(rng.choice([2, 3, 4]))
        m, k = int(rng.integers(8, 60)), int(rng.integers(8, 60))
        n = c * P * int(rng.integers(1, 5))
        explicit = {2: ["pk2", "pn2", "pm2"], 3: ["pn3", "pk3"], 4: ["pn2,pk2", "pm2,pn2", "pk2,pn2", "pn4"], 6: ["pn2,pk3", "pm3,pn2"],
                    8: ["pm2,pn2,pk2", "pn2,pn2,pk2", "pk2,pm2,pn2", "sm2,pn2,pk4"], 12: ["pm2,pn2,pk3", "pn3,pk4"]}[P]
        steps = str(rng.choice(explicit + [""]))
        try:
            full = [MultiplyPlan(None, m, n, k, steps, "d", rank=r, nranks=P, allocate=False) for r in range(P)]
        except Exception:
            continue
        real_steps = full[0].strategy
        verdicts = [host_panel(lib, full[r].handle, c, 0)[0] for r in range(P)]
        used = full[0].P_used
        one_gemm = all(sum(op["kind"] == "gemm" for op in full[r].ops()) == 1 for r in range(used))
        try:
            sub = [MultiplyPlan(None, m, n // c, k, real_steps, "d", rank=r, nranks=P, allocate=False) for r in range(P)] if real_steps else None
        except Exception:
            sub = None
        if used == P and all(verdicts) and one_gemm and sub is not None and sub[0].strategy == real_steps and \
                all(sub[r].initial_elements[1] * c == full[r].initial_elements[1] and sub[r].initial_elements[2] * c == full[r].initial_elements[2]
                    and sub[r].initial_elements[0] == full[r].initial_elements[0] for r in range(P)):
            Ag, Bg = rng.integers(-5, 6, size=(m, k)).astype(np.float64), rng.integers(-5, 6, size=(k, n)).astype(np.float64)
            Cg = rng.integers(-5, 6, size=(m, n)).astype(np.float64)
            arenas = []
            for pl in full:
                bufs = [np.zeros(max(pl.arena_elements[x], 1)) for x in range(3)]
                for x, (label, G) in enumerate((("A", Ag), ("B", Bg), ("C", Cg))):
                    fill_local_from_global(pl, label, bufs[x], G)
                arenas.append(bufs)
            locals_in = [[a[x][:pl.initial_elements[x]].copy() for x in range(3)] for a, pl in zip(arenas, full)]
            run_schedules(full, arenas, 1.0, 1.0)
            for j in range(c):
                sa, cuts = [], []
                for r in range(P):
                    bufs = [np.zeros(max(sub[r].arena_elements[x], 1)) for x in range(3)]
                    bufs[0][:sub[r].initial_elements[0]] = locals_in[r][0]
                    _, bp, cp = host_panel(lib, full[r].handle, c, j)
                    for (src, ln, dst) in bp:
                        bufs[1][dst:dst + ln] = locals_in[r][1][src:src + ln]
                    for (src, ln, dst) in cp:
                        bufs[2][dst:dst + ln] = locals_in[r][2][src:src + ln]
                    cuts.append(cp)
                    sa.append(bufs)
                run_schedules(sub, sa, 1.0, 1.0)
                for r in range(P):
                    for (src, ln, dst) in cuts[r]:
                        assert np.array_equal(sa[r][2][dst:dst + ln], arenas[r][2][src:src + ln]), (P, real_steps, m, n, k, c, j, r)
            checked += 1
        for pl in full + (sub or []):
            pl.destroy()
    assert checked >= 1


def test_host_panel_pipeline_applies_at_the_bench_shapes(lib):
    """32768^3 on 4 and 8 ranks: everything multiply_host_panels (csrc/multiply_exec.cu) asks for before it pipelines the host-pointer
    multiply as c = 4 column panels holds on every rank -- the layout can be cut, the panel problem (m, n / 4, k) keeps the strategy,
    local A is shared, local B / C are a quarter each, the panel's A arena fits the plan's. On 2 ranks (pk2: nothing is gathered) the plain
    path is taken, which streams A and B under the kernel."""
    from cosma_b200.distributed import MultiplyPlan
    from panel_cut import host_panel
    N, c = 32768, 4
    for P, gathered in ((2, set()), (4, {0}), (8, {0, 1})):
        full = [MultiplyPlan(None, N, N, N, "", "d", rank=r, nranks=P, allocate=False) for r in range(P)]
        steps = full[0].strategy
        sub = [MultiplyPlan(None, N, N // c, N, steps, "d", rank=r, nranks=P, allocate=False) for r in range(P)]
        assert {o["matrix"] for o in full[0].ops() if o["kind"] == "allgather"} == gathered
        for r in range(P):
            assert sum(o["kind"] == "gemm" for o in full[r].ops()) == 1 and sub[r].strategy == steps
            assert all(host_panel(lib, full[r].handle, c, j)[0] for j in range(c))
            assert sub[r].initial_elements[0] == full[r].initial_elements[0] and sub[r].arena_elements[0] <= full[r].arena_elements[0]
            assert sub[r].initial_elements[1] * c == full[r].initial_elements[1] and sub[r].initial_elements[2] * c == full[r].initial_elements[2]
            pieces = [host_panel(lib, full[r].handle, c, j) for j in range(c)]
            assert all(sum(p[1] for p in bp) == sub[r].initial_elements[1] and sum(p[1] for p in cp) == sub[r].initial_elements[2] for _, bp, cp in pieces)
        for pl in full + sub:
            pl.destroy()
