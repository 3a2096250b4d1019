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


def simulate(m, n, k, P, steps, alpha=1.0, beta=0.0, dtype="d", seed=0, ints=True, inputs=None, local_c=None):
    """Returns (C_got, C_want, P_used) with C as dense m x n arrays. inputs = (A, B, C) overrides the random operands;
    local_c (a list) receives every rank's raw local C buffer (matrix_pointer() contents; None for idle ranks)."""
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
        bufs = [np.zeros(max(pl.arena_elements[x], 1), dtype=npdt) for x in range(3)]
        if r < P_used:
            for x, (label, full) in enumerate((("A", Ag), ("B", Bg), ("C", Cg))):
                got = fill_local_from_global(pl, label, bufs[x], full)
                assert got == pl.initial_elements[x]
        arenas.append(bufs)
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
