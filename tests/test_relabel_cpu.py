"""Rank relabelling in multiply_using_layout (reference multiply.cpp:136-152), END TO END on the CPU in lock-step:
the caller's matrices are in COSMA's own layout but with the ranks numbered backwards. The volume graph + matching (the C ABI
functions the device path calls) must return that permutation; with physical rank r playing COSMA rank perm[r] -- native grids
relabelled, plans built for the relabelled rank, exactly what csrc/layout_multiply.cu does under COSMA_B200_REORDER_RANKS=ON --
the relayouts move NOTHING between ranks and the product is still right. Without relabelling the same problem sends almost
everything over the wire."""
import ctypes

import numpy as np
import pytest

import costa_sim as sim
import schedule_sim
from cosma_b200 import costa
from cosma_b200.distributed import MultiplyPlan


def _ptr(a, t=ctypes.c_int):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _volume(lib, ga, gb, P):
    out = np.zeros(P * P, dtype=np.int64)
    assert lib.cosma_b200_comm_volume(len(ga[0]) - 1, len(ga[1]) - 1, _ptr(ga[0]), _ptr(ga[1]), _ptr(ga[2]), len(gb[0]) - 1, len(gb[1]) - 1,
                                      _ptr(gb[0]), _ptr(gb[1]), _ptr(gb[2]), ctypes.c_char(b"N"), P, _ptr(out, ctypes.c_longlong)) == 0
    return out.reshape(P, P)


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("relabel", [True, False])
def test_reversed_cosma_layout(lib, oracle, P, relabel):
    m = n = k = 640
    alpha, beta = 2.0, -1.0
    rng = np.random.default_rng(P)
    probe = MultiplyPlan(None, m, n, k, "", "d", rank=0, nranks=P, allocate=False)
    assert probe.P_used == P
    dims = {"A": (m, k), "B": (k, n), "C": (m, n)}
    sigma = np.array([P - 1 - r for r in range(P)], dtype=np.int32)
    grids, user, G = {}, {}, {}
    for label in "ABC":
        per_rank = [probe.local_blocks(label, r) for r in range(P)]
        rs = np.array(sorted({b[0] for bl in per_rank for b in bl} | {dims[label][0]}), dtype=np.int32)
        cs = np.array(sorted({b[2] for bl in per_rank for b in bl} | {dims[label][1]}), dtype=np.int32)
        native = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
        for r, bl in enumerate(per_rank):
            for (r0, r1, c0, c1) in bl:
                native[list(rs).index(r0), list(cs).index(c0)] = r
        grids[label] = (rs, cs, native, per_rank)
        user[label] = sim.DistMatrix(rs, cs, sigma[native], P, "d", "C")
        G[label] = sim.random_values(rng, dims[label], "d")
        user[label].scatter(G[label])
    # the decision the device path takes: volumes of A and B into the native grids, of C out of it; then the matching
    total = np.zeros((P, P), dtype=np.int64)
    for label in "ABC":
        rs, cs, native, _ = grids[label]
        u, v = (rs, cs, np.ascontiguousarray(sigma[native].reshape(-1))), (rs, cs, np.ascontiguousarray(native.reshape(-1)))
        total += _volume(lib, u, v, P) if label != "C" else _volume(lib, v, u, P)
    perm = np.zeros(P, dtype=np.int32)
    flag = ctypes.c_int(0)
    assert lib.cosma_b200_optimal_reordering(P, _ptr(np.ascontiguousarray(total), ctypes.c_longlong), _ptr(perm), ctypes.byref(flag)) == 0
    assert flag.value == 1 and np.array_equal(perm, sigma)
    if not relabel:
        perm = np.arange(P, dtype=np.int32)
    # physical rank r plays COSMA rank perm[r]
    plans = [MultiplyPlan(None, m, n, k, "", "d", rank=int(perm[r]), nranks=P, allocate=False) for r in range(P)]
    arenas = [[np.zeros(max(pl.arena_elements[x], 1), dtype=np.float64) for x in range(3)] for pl in plans]
    native_l = {}
    for x, label in enumerate("ABC"):
        rs, cs, native, per_rank = grids[label]
        layouts = []
        for r in range(P):
            blocks, pos = [], 0
            for (r0, r1, c0, c1) in per_rank[perm[r]]:
                nr, nc = r1 - r0 + 1, c1 - c0 + 1
                blocks.append((list(rs).index(r0), list(cs).index(c0), arenas[r][x].ctypes.data + pos * 8, nr))
                pos += nr * nc
            layouts.append(costa.custom_layout(rs, cs, perm[native], blocks, "C"))  # COSMA rank q is physical rank perm[q]
        native_l[label] = layouts
    tin, remote_in = [], 0
    for r in range(P):
        tp = costa.TransformPlan(None, "d", [(user["A"].layout(r), native_l["A"][r], "N", 1.0, 0.0), (user["B"].layout(r), native_l["B"][r], "N", 1.0, 0.0)],
                                 rank=r, nranks=P)
        remote_in += tp.stats()["remote_elements"]
        tin.append(tp.export()); tp.destroy()
    sim.simulate(oracle, "d", tin, [(1.0, 0.0), (1.0, 0.0)])
    inv = np.argsort(perm)
    schedule_sim.run_schedules([plans[inv[q]] for q in range(P)], [arenas[inv[q]] for q in range(P)], 1.0, 0.0)
    tout, remote_out = [], 0
    for r in range(P):
        tp = costa.TransformPlan(None, "d", [(native_l["C"][r], user["C"].layout(r), "N", alpha, beta)], rank=r, nranks=P)
        remote_out += tp.stats()["remote_elements"]
        tout.append(tp.export()); tp.destroy()
    sim.simulate(oracle, "d", tout, [(alpha, beta)])
    for pl in plans + [probe]:
        pl.destroy()
    assert np.array_equal(user["C"].gather(), alpha * (G["A"] @ G["B"]) + beta * G["C"])
    if relabel:
        assert remote_in == 0 and remote_out == 0
    else:
        assert remote_in > (m * k + k * n) // 2 and remote_out > (m * n) // 2
