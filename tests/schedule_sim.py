"""TEST INFRASTRUCTURE: interprets compiled schedules (cosma_b200_plan_export) on the CPU with numpy.

`simulate` runs ALL ranks of a job inside one process in lockstep (a rank blocks at a collective until every member
of its ring has reached the same op), so the compiled data placement, ring membership, offsets and beta handling
can be checked against a dense product without any GPU. The arithmetic itself is numpy's; GEMM parity is the
business of the GPU tests."""
import numpy as np

from cosma_b200.distributed import MultiplyPlan, fill_local_from_global, gather_local_to_global


def _beta(op_beta, user_beta):
    return {0: 0.0, 1: 1.0, 2: user_beta}[op_beta]


def gemm_cpu(op, A, B, C, alpha, user_beta):
    m, n, k = op["m"], op["n"], op["k"]
    a = A[op["a_off"]:op["a_off"] + m * k].reshape(k, m).T
    b = B[op["b_off"]:op["b_off"] + k * n].reshape(n, k).T
    cview = C[op["c_off"]:op["c_off"] + m * n]
    beta = _beta(op["beta"], user_beta)
    res = alpha * (a @ b)
    if beta != 0:
        res = res + beta * cview.reshape(n, m).T
    cview[:] = res.T.reshape(-1)


def allgather_pieces(op, member_src):
    """member_src[g] = that member's contiguous piece buffer. Returns the expanded (bucket-major) buffer."""
    nb = len(op["piece"][0])
    out, pos = [], [0] * len(op["ring"])
    for b in range(nb):
        for g in range(len(op["ring"])):
            cnt = op["piece"][g][b]
            out.append(member_src[g][pos[g]:pos[g] + cnt])
            pos[g] += cnt
    return np.concatenate(out) if out else np.zeros(0)


def reduce_slices(op, member_src, g):
    """Sum over members of the slice of the expanded partial result that belongs to member g (bucket order)."""
    nb = len(op["piece"][0])
    outs, off = [], 0
    for b in range(nb):
        for gg in range(len(op["ring"])):
            cnt = op["piece"][gg][b]
            if gg == g:
                outs.append(sum(src[off:off + cnt] for src in member_src))
            off += cnt
    return np.concatenate(outs) if outs else np.zeros(0)


def simulate(m, n, k, P, steps, alpha=1.0, beta=0.0, dtype="d", seed=0, ints=True, inputs=None, local_c=None, overlapped=False, poison=False):
    """Returns (C_got, C_want, P_used) with C as dense m x n arrays. inputs = (A, B, C) overrides the random operands;
    local_c (a list) receives every rank's raw local C buffer (matrix_pointer() contents; None for idle ranks).
    overlapped: run the plans' overlapped micro-op programs (they must have one: AssertionError otherwise);
    poison: fill the communication workspace of every arena with NaN first (nothing may be read before it is written)."""
    rng = np.random.default_rng(seed)
    npdt = np.float64 if dtype == "d" else np.complex128
    def rnd(r, c):
        v = rng.integers(0, 10, size=(r, c)).astype(np.float64) if ints else rng.random((r, c))
        if dtype == "z":
            v = v + 1j * (rng.integers(0, 10, size=(r, c)) if ints else rng.random((r, c)))
        return v.astype(npdt)
    Ag, Bg, Cg = rnd(m, k), rnd(k, n), rnd(m, n)
    if inputs is not None:
        Ag, Bg, Cg = (np.asarray(x, dtype=npdt) for x in inputs)
    plans = [MultiplyPlan(None, m, n, k, steps, dtype, rank=r, nranks=P, allocate=False) for r in range(P)]
    P_used = plans[0].P_used
    arenas = []
    for r, pl in enumerate(plans):
        bufs = [np.full(max(pl.arena_elements[x], 1), np.nan if poison else 0.0, dtype=npdt) for x in range(3)]
        if r < P_used:
            for x, (label, full) in enumerate((("A", Ag), ("B", Bg), ("C", Cg))):
                got = fill_local_from_global(pl, label, bufs[x], full)
                assert got == pl.initial_elements[x]
        arenas.append(bufs)
    if overlapped:
        assert run_overlapped(plans, arenas, alpha, beta), "the plans are not overlapped: %s" % plans[0].overlap()["why"]
    else:
        run_schedules(plans, arenas, alpha, beta)
    got = np.zeros((m, n), dtype=npdt)
    for r in range(P_used):
        gather_local_to_global(plans[r], "C", arenas[r][2], got)
    want = alpha * (Ag @ Bg) + beta * Cg
    if local_c is not None:
        local_c.extend(arenas[r][2][:plans[r].initial_elements[2]].copy() if r < P_used else None for r in range(P))
    for pl in plans:
        pl.destroy()
    return got, want, P_used


def run_schedules(plans, arenas, alpha, beta):
    """Executes the compiled schedules of all ranks in lock-step on the given arenas ([rank][matrix] numpy buffers whose first
    initial_elements hold the rank's local matrix)."""
    P = len(plans)
    P_used = plans[0].P_used
    progs = [pl.ops() if r < P_used else [] for r, pl in enumerate(plans)]
    pc = [0] * P
    while any(pc[r] < len(progs[r]) for r in range(P)):
        progressed = False
        for r in range(P):
            while pc[r] < len(progs[r]) and progs[r][pc[r]]["kind"] == "gemm":
                gemm_cpu(progs[r][pc[r]], *arenas[r], alpha, beta)
                pc[r] += 1
                progressed = True
        for r in range(P):
            if pc[r] >= len(progs[r]):
                continue
            op = progs[r][pc[r]]
            if op["kind"] == "gemm":
                continue  # unblocked by a collective earlier in this sweep; runs in the next one
            ring = op["ring"]
            if r != ring[0]:
                continue
            # all members must be waiting at an op of the same kind/step with the same ring
            if not all(pc[q] < len(progs[q]) and progs[q][pc[q]]["kind"] == op["kind"] and progs[q][pc[q]]["step"] == op["step"]
                       and progs[q][pc[q]]["ring"] == ring for q in ring):
                continue
            x = op["matrix"]
            mops = [progs[q][pc[q]] for q in ring]
            for g, q in enumerate(ring):
                assert mops[g]["my_pos"] == g and mops[g]["piece"] == op["piece"]
            if op["kind"] == "allgather":
                srcs = [arenas[q][x][mops[g]["src_off"]:mops[g]["src_off"] + sum(op["piece"][g])].copy() for g, q in enumerate(ring)]
                exp = allgather_pieces(op, srcs)
                for g, q in enumerate(ring):
                    arenas[q][x][mops[g]["dst_off"]:mops[g]["dst_off"] + len(exp)] = exp
            else:
                total = sum(sum(p) for p in op["piece"])
                srcs = [arenas[q][x][mops[g]["src_off"]:mops[g]["src_off"] + total].copy() for g, q in enumerate(ring)]
                for g, q in enumerate(ring):
                    mine = reduce_slices(op, srcs, g)
                    b = _beta(mops[g]["beta"], beta)
                    dst = arenas[q][x][mops[g]["dst_off"]:mops[g]["dst_off"] + len(mine)]
                    dst[:] = mine if b == 0 else b * dst + mine
            for q in ring:
                pc[q] += 1
            progressed = True
        assert progressed, "schedule deadlock: %s" % [(r, pc[r], len(progs[r])) for r in range(P)]


# ---- overlapped schedules (include/cosma/overlap.hpp): the micro-op programs of all ranks in lock-step --------------------------

def _sub(buf, off, rows, cols, ld):
    """Copy of the column-major rows x cols sub-matrix at element offset `off` with leading dimension ld."""
    idx = off + np.arange(rows)[:, None] + ld * np.arange(cols)[None, :]
    return buf[idx]


def _store(buf, off, rows, cols, ld, value):
    idx = off + np.arange(rows)[:, None] + ld * np.arange(cols)[None, :]
    buf[idx] = value


def micro_gemm_cpu(o, A, B, C, alpha, user_beta):
    a = _sub(A, o["a_off"], o["m"], o["k"], o["lda"])
    b = _sub(B, o["b_off"], o["k"], o["n"], o["ldb"])
    beta = _beta(o["beta"], user_beta)
    res = alpha * (a @ b)
    if beta != 0:
        res = res + beta * _sub(C, o["c_off"], o["m"], o["n"], o["ldc"])
    _store(C, o["c_off"], o["m"], o["n"], o["ldc"], res)


def run_overlapped(plans, arenas, alpha, beta):
    """Executes the OVERLAPPED programs (MultiplyPlan.overlap()) of all ranks in lock-step, each rank's micro-ops in program order (a
    valid serialisation of its two streams: every wait points backwards). Returns False if the plans are not overlapped."""
    P = len(plans)
    P_used = plans[0].P_used
    progs, sched = [], []
    for r, pl in enumerate(plans):
        ov = pl.overlap() if r < P_used else {"enabled": True, "ops": []}
        if not ov["enabled"]:
            return False
        for i, o in enumerate(ov["ops"]):
            assert all(0 <= w < i for w in o["wait"]), "a micro-op waits for a later one"
        progs.append(ov["ops"])
        sched.append(pl.ops() if r < P_used else [])
    pc = [0] * P

    def coll(r):
        """The collective rank r is waiting at: (kind, ring ranks, schedule op or micro-op)."""
        o = progs[r][pc[r]]
        if o["kind"] in ("allgather", "serial"):
            sop = sched[r][o["op"]]
            return sop["kind"], sop["ring"], sop
        if o["kind"] == "exchange":
            ring = next(s["ring"] for s in sched[r] if s["kind"] != "gemm" and s["ring_index"] == o["ring_index"])
            return "exchange", ring, o
        return None

    while any(pc[r] < len(progs[r]) for r in range(P)):
        progressed = False
        for r in range(P):
            while pc[r] < len(progs[r]) and progs[r][pc[r]]["kind"] in ("gemm", "accumulate"):
                o = progs[r][pc[r]]
                if o["kind"] == "gemm":
                    micro_gemm_cpu(o, *arenas[r], alpha, beta)
                else:
                    C = arenas[r][2]
                    b = _beta(o["beta"], beta)
                    if not (o["beta_term"] and b == 0):  # beta == 0 at run time: the exchange landed in C itself
                        C[o["dst_off"]:o["dst_off"] + o["count"]] = b * C[o["dst_off"]:o["dst_off"] + o["count"]] + C[o["add_off"]:o["add_off"] + o["count"]]
                pc[r] += 1
                progressed = True
        for r in range(P):
            if pc[r] >= len(progs[r]) or progs[r][pc[r]]["kind"] in ("gemm", "accumulate"):
                continue
            kind, ring, op = coll(r)
            if r != ring[0]:
                continue
            mates = []
            for q in ring:
                if pc[q] >= len(progs[q]) or progs[q][pc[q]]["kind"] in ("gemm", "accumulate"):
                    break
                kq, rq, oq = coll(q)
                if kq != kind or rq != ring:
                    break
                mates.append(oq)
            if len(mates) != len(ring):
                continue
            if kind == "exchange":
                assert len(ring) == 2
                sent = [arenas[q][2][mates[g]["send_off"]:mates[g]["send_off"] + mates[g]["count"]].copy() for g, q in enumerate(ring)]
                for g, q in enumerate(ring):
                    assert ring[mates[g]["peer"]] == ring[1 - g] and mates[g]["count"] == mates[1 - g]["count"]
                    at = mates[g]["recv_off_zero"] if _beta(mates[g]["beta"], beta) == 0 else mates[g]["recv_off"]
                    arenas[q][2][at:at + mates[g]["count"]] = sent[1 - g]
            else:
                x = op["matrix"]
                for g, q in enumerate(ring):
                    assert mates[g]["my_pos"] == g and mates[g]["piece"] == op["piece"] and mates[g]["step"] == op["step"]
                if kind == "allgather":
                    srcs = [arenas[q][x][mates[g]["src_off"]:mates[g]["src_off"] + sum(op["piece"][g])].copy() for g, q in enumerate(ring)]
                    exp = allgather_pieces(op, srcs)
                    for g, q in enumerate(ring):
                        arenas[q][x][mates[g]["dst_off"]:mates[g]["dst_off"] + len(exp)] = exp
                else:
                    total = sum(sum(p) for p in op["piece"])
                    srcs = [arenas[q][x][mates[g]["src_off"]:mates[g]["src_off"] + total].copy() for g, q in enumerate(ring)]
                    for g, q in enumerate(ring):
                        mine = reduce_slices(op, srcs, g)
                        b = _beta(mates[g]["beta"], beta)
                        dst = arenas[q][x][mates[g]["dst_off"]:mates[g]["dst_off"] + len(mine)]
                        dst[:] = mine if b == 0 else b * dst + mine
            for q in ring:
                pc[q] += 1
            progressed = True
        assert progressed, "overlapped schedule deadlock: %s" % [(r, pc[r], len(progs[r])) for r in range(P)]
    return True
