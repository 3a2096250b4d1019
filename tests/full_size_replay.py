"""TEST INFRASTRUCTURE: the programs the library emits for FULL-SIZE problems (32768^3 on 2 / 4 / 8 ranks ...), replayed numerically.

The overlap planner's decisions depend on the real sizes (panel widths in whole waves of 128-wide tiles on 148 CTAs), so a small problem
never yields the bench's program. But neither the schedule compiler nor the planner ever cuts the m dimension inside a rank's local
matrix: every offset, count, leading dimension and piece size in the A and C arenas is a whole number of columns of m_loc rows. Dividing
all of them by f = m / m_small therefore gives an isomorphic program on matrices with m_small rows -- same panels, same columns, same k
blocks, same exchanges, same waits -- while B (k x n) keeps its full size. tests/schedule_sim.py then executes all ranks in lock-step
with poisoned workspaces and the gathered C must equal A.B exactly (integer-valued operands in float32: every partial sum stays below
2^24)."""
import numpy as np

from cosma_b200.distributed import MultiplyPlan
import schedule_sim as sim


class ScaledPlan:
    """A MultiplyPlan whose A / C arenas are measured in units of f elements (see the module docstring). Quacks like the plan for
    schedule_sim.run_overlapped / run_schedules and fill_local_from_global / gather_local_to_global."""

    def __init__(self, plan, f):
        self.plan, self.f = plan, f
        self.P_used, self.rank, self.idle, self.strategy = plan.P_used, plan.rank, plan.idle, plan.strategy

    def _div(self, v, what):
        assert v % self.f == 0, "%s = %d is not a whole number of %d-row columns" % (what, v, self.f)
        return v // self.f

    @property
    def arena_elements(self):
        a = self.plan.arena_elements
        return [self._div(a[0], "A arena"), a[1], self._div(a[2], "C arena")]

    @property
    def initial_elements(self):
        a = self.plan.initial_elements
        return [self._div(a[0], "local A"), a[1], self._div(a[2], "local C")]

    def local_blocks(self, label, rank=None):
        blocks = self.plan.local_blocks(label, rank)
        if label == "A":  # rows are m
            return [(self._div(r0, "row"), self._div(r1 + 1, "row") - 1, c0, c1) for (r0, r1, c0, c1) in blocks]
        if label == "C":
            return [(self._div(r0, "row"), self._div(r1 + 1, "row") - 1, c0, c1) for (r0, r1, c0, c1) in blocks]
        return blocks

    def ops(self):
        out = []
        for o in self.plan.ops():
            o = dict(o)
            if o["kind"] == "gemm":
                for key in ("a_off", "c_off", "m"):
                    o[key] = self._div(o[key], key)
            elif o["matrix"] != 1:
                for key in ("src_off", "dst_off") + (("tmp_off",) if "tmp_off" in o else ()):
                    o[key] = self._div(o[key], key)
                o["piece"] = [[self._div(v, "piece") for v in row] for row in o["piece"]]
            out.append(o)
        return out

    def overlap(self):
        ov = dict(self.plan.overlap())
        ops = []
        for o in ov["ops"]:
            o = dict(o)
            if o["kind"] == "gemm":
                keys = ("a_off", "c_off", "lda", "ldc", "m")
            elif o["kind"] == "exchange":
                keys = ("send_off", "recv_off", "recv_off_zero", "count")
            elif o["kind"] == "accumulate":
                keys = ("dst_off", "add_off", "count")
            else:
                keys = ()
            for key in keys:
                o[key] = self._div(o[key], key)
            ops.append(o)
        ov["ops"] = ops
        return ov

    def destroy(self):
        self.plan.destroy()


def _fast_sub(buf, off, rows, cols, ld):
    """View (no index arrays) of a column-major rows x cols sub-matrix."""
    return np.lib.stride_tricks.as_strided(buf[off:], shape=(rows, cols), strides=(buf.itemsize, ld * buf.itemsize), writeable=False)


def _micro_gemm(o, A, B, C, alpha, user_beta):
    a = _fast_sub(A, o["a_off"], o["m"], o["k"], o["lda"])
    b = _fast_sub(B, o["b_off"], o["k"], o["n"], o["ldb"])
    beta = sim._beta(o["beta"], user_beta)
    res = alpha * (a @ b)
    cv = np.lib.stride_tricks.as_strided(C[o["c_off"]:], shape=(o["m"], o["n"]), strides=(C.itemsize, o["ldc"] * C.itemsize))
    if beta != 0:
        res = res + beta * cv
    cv[...] = res


def random_b(k, n, seed=0):
    """Integer-valued k x n operand in COLUMN-major storage (the local buffers are column-major: block copies stay contiguous)."""
    rng = np.random.default_rng(1000 + seed)
    return rng.integers(0, 10, size=(n, k), dtype=np.int8).astype(np.float32).T


def replay(m, n, k, P, steps="", dtype="d", m_small=16, alpha=1.0, beta=0.0, overlapped=True, Bg=None, seed=0):
    """-> (got, want, info): the gathered m_small x n result of the full-size plans' programs, the dense product, and what ran."""
    assert m % m_small == 0
    f = m // m_small
    rng = np.random.default_rng(seed)
    plans = [ScaledPlan(MultiplyPlan(None, m, n, k, steps, dtype, rank=r, nranks=P, allocate=False), f) for r in range(P)]
    P_used = plans[0].P_used
    Ag = rng.integers(0, 10, size=(m_small, k)).astype(np.float32)
    if Bg is None:
        Bg = random_b(k, n, seed)
    Cg = rng.integers(0, 4, size=(m_small, n)).astype(np.float32)
    arenas = []
    for r, pl in enumerate(plans):
        bufs = [np.full(max(pl.arena_elements[x], 1), np.nan, dtype=np.float32) for x in range(3)]
        if r < P_used:
            for x, (label, full) in enumerate((("A", Ag), ("B", Bg), ("C", Cg))):
                got = sim.fill_local_from_global(pl, label, bufs[x], full)
                assert got == pl.initial_elements[x]
            if beta == 0.0:
                bufs[2][:pl.initial_elements[2]] = np.nan  # C must never be read
        arenas.append(bufs)
    info = {"strategy": plans[0].strategy, "overlap": plans[0].plan.overlap()["why"], "f": f}
    saved = sim.micro_gemm_cpu
    sim.micro_gemm_cpu = _micro_gemm
    try:
        if overlapped:
            assert sim.run_overlapped(plans, arenas, alpha, beta), "the plans are not overlapped: " + info["overlap"]
            info["gemm_panels"] = [sum(1 for o in pl.overlap()["ops"] if o["kind"] == "gemm") for pl in plans[:P_used]]
        else:
            sim.run_schedules(plans, arenas, alpha, beta)
    finally:
        sim.micro_gemm_cpu = saved
    got = np.zeros((m_small, n), dtype=np.float32)
    for r in range(P_used):
        sim.gather_local_to_global(plans[r], "C", arenas[r][2], got)
    want = alpha * (Ag @ Bg)  # float32 is exact here: integer partial sums below 2^24 whatever the order
    assert float(np.abs(want).max()) * max(1.0, abs(alpha)) < 2 ** 24
    if beta != 0.0:
        want = want + np.float32(beta) * Cg
    for pl in plans:
        pl.destroy()
    return got, want.astype(np.float32), info
