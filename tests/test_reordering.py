"""costa::communication_volume / costa::optimal_reordering (reference libs/COSTA/src/costa/grid2grid/transform.cpp:9-44,
ranks_reordering.cpp:4-61) against the unmodified reference built into oracle/_ref: the volume graph is an exact integer match on
random grids with random owners (plain and transposed); the relabelling equals the reference's wherever the greedy matching has no
ties (the reference breaks ties by hash-map iteration order, we by rank pair), and is always an involution that never lowers the
volume kept in place."""
import ctypes

import numpy as np
import pytest

from cosma_b200 import _lib


def _grid(rng, rows, cols, n_ranks, max_blocks=7):
    def split(n):
        k = int(rng.integers(1, min(max_blocks, n) + 1))
        cuts = sorted(rng.choice(np.arange(1, n), size=k - 1, replace=False).tolist()) if k > 1 else []
        return np.array([0] + cuts + [n], dtype=np.int32)
    rs, cs = split(rows), split(cols)
    owners = rng.integers(0, n_ranks, size=(len(rs) - 1) * (len(cs) - 1)).astype(np.int32)
    return rs, cs, owners


def _ptr(a, t=ctypes.c_int):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _volume(fn, ga, gb, trans, P):
    out = np.zeros(P * P, dtype=np.int64)
    rc = fn(len(ga[0]) - 1, len(ga[1]) - 1, _ptr(ga[0]), _ptr(ga[1]), _ptr(ga[2]), len(gb[0]) - 1, len(gb[1]) - 1, _ptr(gb[0]), _ptr(gb[1]),
            _ptr(gb[2]), ctypes.c_char(trans.encode()), P, _ptr(out, ctypes.c_longlong))
    assert rc == 0
    return out.reshape(P, P)


def _reorder(fn, vol):
    P = vol.shape[0]
    perm = np.zeros(P, dtype=np.int32)
    flag = ctypes.c_int(0)
    v = np.ascontiguousarray(vol, dtype=np.int64)
    assert fn(P, _ptr(v, ctypes.c_longlong), _ptr(perm), ctypes.byref(flag)) == 0
    return perm, bool(flag.value)


def _kept_in_place(vol, perm):
    """elements that stay local after relabelling: rank r holds what was sent between r and perm[r] (or kept on r)"""
    P = vol.shape[0]
    return sum(int(vol[min(r, perm[r]), max(r, perm[r])]) for r in range(P) if perm[r] >= r)


@pytest.mark.parametrize("seed", range(12))
def test_comm_volume_matches_reference(lib, ref, seed):
    rng = np.random.default_rng(seed)
    R = ref.ref()
    P = int(rng.integers(1, 9))
    rows, cols = int(rng.integers(2, 60)), int(rng.integers(2, 60))
    trans = "NTC"[seed % 3]
    ga = _grid(rng, cols if trans != "N" else rows, rows if trans != "N" else cols, P)
    gb = _grid(rng, rows, cols, P)
    ours = _volume(lib.cosma_b200_comm_volume, ga, gb, trans, P)
    theirs = _volume(R.ref_comm_volume, ga, gb, trans, P)
    assert np.array_equal(ours, theirs)
    assert ours.sum() == rows * cols  # every element is counted exactly once
    assert np.array_equal(ours, np.triu(ours))


@pytest.mark.parametrize("seed", range(20))
def test_optimal_reordering(lib, ref, seed):
    rng = np.random.default_rng(100 + seed)
    R = ref.ref()
    P = int(rng.integers(1, 13))
    # distinct weights -> distinct gains almost surely; keep only instances without tied positive gains
    vol = np.triu(rng.permutation(np.arange(1, P * P + 1) * 17 % 1009 + 1).reshape(P, P)).astype(np.int64)
    if seed % 3 == 0:  # a layout that is a pure relabelling: all volume sits on the pairs (r, sigma(r))
        sigma = rng.permutation(P)
        vol[:] = 0
        for r in range(P):
            vol[min(r, sigma[r]), max(r, sigma[r])] += 1000 + 37 * r
    # every rank keeps something (as in any real relayout): the reference looks the self edges up with operator[] WHILE it
    # iterates over its hash map (ranks_reordering.cpp:27-28), which inserts -- and may rehash -- when one is missing
    for r in range(P):
        vol[r, r] = max(int(vol[r, r]), 1 + (r * 7) % 5)
    ours, ours_flag = _reorder(lib.cosma_b200_optimal_reordering, vol)
    theirs, theirs_flag = _reorder(R.ref_optimal_reordering, vol)
    # always: an involution that does not lose local volume
    assert np.array_equal(ours[ours], np.arange(P))
    assert _kept_in_place(vol, ours) >= int(np.trace(vol))
    assert ours_flag == bool((ours != np.arange(P)).any())
    gains = []
    for u in range(P):
        for v in range(u + 1, P):
            g = int(vol[u, v]) - int(vol[u, u]) - int(vol[v, v])
            if g > 0:
                gains.append(g)
    if len(gains) == len(set(gains)) and 1 not in gains:  # tie-free (gain 1 would tie with the "stay" edges)
        assert np.array_equal(ours, theirs), (vol, ours, theirs)
        assert ours_flag == theirs_flag
    else:
        assert _kept_in_place(vol, ours) > 0 or P == 1 or vol.sum() == 0


def test_relabelling_recovers_a_permuted_block_cyclic_layout(lib):
    """Two block-cyclic layouts that differ only by how the ranks are numbered (row- vs column-major grid): after relabelling
    nothing has to travel."""
    P, nprow, npcol, M, N, mb, nb = 6, 2, 3, 48, 60, 4, 5

    def grid(order):
        rs = np.arange(0, M + 1, mb, dtype=np.int32)
        cs = np.arange(0, N + 1, nb, dtype=np.int32)
        own = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
        for i in range(own.shape[0]):
            for j in range(own.shape[1]):
                pr, pc = i % nprow, j % npcol
                own[i, j] = pr * npcol + pc if order == "R" else pc * nprow + pr
        return rs, cs, own.reshape(-1)

    vol = _volume(lib.cosma_b200_comm_volume, grid("R"), grid("C"), "N", P)
    perm, flag = _reorder(lib.cosma_b200_optimal_reordering, vol)
    assert np.array_equal(perm[perm], np.arange(P))
    # row-major label r = pr*npcol + pc maps to column-major pc*nprow + pr; the matching pairs up exactly those labels
    kept = _kept_in_place(vol, perm)
    direct = int(np.trace(vol))
    assert kept >= direct and flag


@pytest.mark.parametrize("world", [2, 4, 8, 6])
def test_relabelling_finds_cosma_layout_with_reversed_ranks(lib, world):
    """The case multiply_using_layout meets when the caller's layout IS COSMA's native layout up to a renumbering of the ranks
    (reference multiply.cpp:136-152): the volume graph summed over A, B, C puts everything on the pairs (r, sigma(r)), and the
    matching returns sigma -- after which every element is already in place."""
    from cosma_b200 import planning
    m = n = k = 640  # every split keeps the local dimensions above COSMA_MIN_LOCAL_DIMENSION (200)
    steps, P_used, _ = planning.strategy(m, n, k, world)
    assert P_used == world
    sigma = [world - 1 - r for r in range(world)]
    total = np.zeros((world, world), dtype=np.int64)
    for label, (rows, cols) in (("A", (m, k)), ("B", (k, n)), ("C", (m, n))):
        per_rank = planning.mapper_layout(label, m, n, k, world, steps)
        rs = np.array(sorted({b[0] for bl in per_rank for b in bl} | {rows}), dtype=np.int32)
        cs = np.array(sorted({b[2] for bl in per_rank for b in bl} | {cols}), dtype=np.int32)
        native = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
        for r, bl in enumerate(per_rank):
            for (r0, r1, c0, c1) in bl:
                for bi in range(len(rs) - 1):
                    for bj in range(len(cs) - 1):
                        if r0 <= rs[bi] <= r1 and c0 <= cs[bj] <= c1:
                            native[bi, bj] = r
        user = np.array([sigma[r] for r in native.reshape(-1)], dtype=np.int32)
        a, b = (rs, cs, user), (rs, cs, native.reshape(-1).copy())
        total += _volume(lib.cosma_b200_comm_volume, a, b, "N", world) if label != "C" else _volume(lib.cosma_b200_comm_volume, b, a, "N", world)
    assert int(np.trace(total)) == 0 or world % 2 == 1
    perm, flag = _reorder(lib.cosma_b200_optimal_reordering, total)
    assert flag and list(perm) == sigma
    assert _kept_in_place(total, perm) == 3 * m * n


def test_baseline_pzgemm_configuration_needs_no_relabelling(lib):
    """BASELINE configs[4] (16384^2, 256 x 256 blocks on a 2 x 4 row-major grid, A conjugate-transposed) against COSMA's layout for
    pm2,pn2,pk2: block-cyclic spreads every COSMA block evenly over all ranks, so every pair of ranks exchanges the same volume and the
    matching keeps the identity."""
    from cosma_b200 import planning
    P, m, n, k = 8, 16384, 16384, 16384
    steps = planning.strategy(m, n, k, P)[0]
    total = np.zeros((P, P), dtype=np.int64)
    for label, (rows, cols), tr in (("A", (m, k), "C"), ("B", (k, n), "N"), ("C", (m, n), "N")):
        per = planning.mapper_layout(label, m, n, k, P, steps)
        rs = np.array(sorted({b[0] for bl in per for b in bl} | {rows}), dtype=np.int32)
        cs = np.array(sorted({b[2] for bl in per for b in bl} | {cols}), dtype=np.int32)
        native = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
        for r, bl in enumerate(per):
            for (r0, r1, c0, c1) in bl:
                native[list(rs).index(r0), list(cs).index(c0)] = r
        ur, uc = (cols, rows) if tr != "N" else (rows, cols)
        brs, bcs = np.arange(0, ur + 1, 256, dtype=np.int32), np.arange(0, uc + 1, 256, dtype=np.int32)
        own = np.array([[(i % 2) * 4 + (j % 4) for j in range(len(bcs) - 1)] for i in range(len(brs) - 1)], dtype=np.int32).reshape(-1)
        u, v = (brs, bcs, own), (rs, cs, np.ascontiguousarray(native.reshape(-1)))
        total += _volume(lib.cosma_b200_comm_volume, u, v, tr, P) if label != "C" else _volume(lib.cosma_b200_comm_volume, v, u, "N", P)
    assert int(total.sum()) == 3 * m * n
    off = total[np.triu_indices(P, 1)]
    assert (off == off[0]).all() and (np.diag(total) == total[0, 0]).all()
    perm, flag = _reorder(lib.cosma_b200_optimal_reordering, total)
    assert not flag and list(perm) == list(range(P))
