"""cosma::adapt_strategy_to_block_cyclic_grid (reference src/cosma/cosma_pxgemm.cpp:517-650) against the unmodified reference: the
strategy prefix that reproduces the ScaLAPACK grid of the largest operand. Exact string parity on the BASELINE pzgemm configuration,
on a sweep of shapes / grids / block sizes / transposes / numberings around the reference's conditions (big enough, whole matrix,
perfectly tiled), and on inputs that must yield no prefix; the completed strategy (prefix + automatic rest) matches too."""
import ctypes
import itertools

import numpy as np
import pytest

from cosma_b200 import planning


def _desc(rows, cols, mb, nb, lld=1):
    return np.array([1, 0, rows, cols, mb, nb, 0, 0, max(lld, 1)], dtype=np.int32)


def _call(fn, m, n, k, P, da, ia, ja, db, ib, jb, dc, ic, jc, ta, tb, pr, pc, order):
    out = ctypes.create_string_buffer(512)
    pi = ctypes.POINTER(ctypes.c_int)
    rc = fn(m, n, k, P, da.ctypes.data_as(pi), ia, ja, db.ctypes.data_as(pi), ib, jb, dc.ctypes.data_as(pi), ic, jc, ctypes.c_char(ta.encode()),
            ctypes.c_char(tb.encode()), pr, pc, ctypes.c_char(order.encode()), out, 512)
    assert rc == 0
    return out.value.decode()


def _both(lib, R, m, n, k, pr, pc, order, ta, tb, blocks, sub=((1, 1), (1, 1), (1, 1)), extra=0):
    am, an = (m, k) if ta == "N" else (k, m)
    bm, bn = (k, n) if tb == "N" else (n, k)
    da = _desc(am + extra, an + extra, *blocks[0])
    db = _desc(bm + extra, bn + extra, *blocks[1])
    dc = _desc(m + extra, n + extra, *blocks[2])
    args = (m, n, k, pr * pc, da, sub[0][0], sub[0][1], db, sub[1][0], sub[1][1], dc, sub[2][0], sub[2][1], ta, tb, pr, pc, order)
    return _call(lib.cosma_b200_adapt_strategy, *args), _call(R.ref_adapt_strategy, *args)


def test_baseline_pzgemm_configuration(lib, ref):
    """BASELINE configs[4]: 16384^3, 256 x 256 blocks, 2 x 4 row-major grid, A conjugate-transposed: the reference turns it into
    32 x 16 sequential repetitions of a 2 x 4 parallel grid over A (= 512 local GEMMs per rank)."""
    R = ref.ref()
    ours, theirs = _both(lib, R, 16384, 16384, 16384, 2, 4, "R", "C", "N", ((256, 256),) * 3)
    assert ours == theirs == "sk32,sm16,pk2,pm4"
    steps, P_used, _ = planning.strategy(16384, 16384, 16384, 8, 0, ours)
    assert P_used == 8 and steps.startswith("sk32,sm16,pk2,pm4")


@pytest.mark.parametrize("order", ["R", "C"])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("C", "C")])
def test_sweep_matches_reference(lib, ref, order, ta, tb):
    R = ref.ref()
    seen = set()
    shapes = [(16384, 16384, 16384), (32768, 8192, 8192), (8192, 32768, 8192), (8192, 8192, 65536), (20000, 20000, 20000), (4096, 4096, 4096),
              (12288, 24576, 12288)]
    for (m, n, k), (pr, pc), blk in itertools.product(shapes, [(2, 4), (4, 2), (1, 8), (8, 1), (2, 2), (3, 2), (1, 1)], [256, 512, 1000, 96]):
        ours, theirs = _both(lib, R, m, n, k, pr, pc, order, ta, tb, ((blk, blk), (blk, 2 * blk), (2 * blk, blk)))
        assert ours == theirs, (m, n, k, pr, pc, blk, ours, theirs)
        seen.add(ours != "")
        if ours:
            # the completed strategy agrees with the reference's completion of the same prefix
            P = pr * pc
            try:
                mine = planning.strategy(m, n, k, P, 0, ours)[0]
            except Exception:
                continue  # a prefix the Strategy rejects (the reference throws as well): nothing to compare
            out = ctypes.create_string_buffer(4096)
            Pout, mem = ctypes.c_int(0), ctypes.c_longlong(0)
            rc = R.ref_strategy(m, n, k, P, ctypes.c_longlong(0), ours.encode(), out, 4096, ctypes.byref(Pout), ctypes.byref(mem))
            if rc >= 0:
                assert out.value.decode() == mine
    assert seen == {True, False}


def test_no_prefix_when_conditions_fail(lib, ref):
    R = ref.ref()
    big = ((256, 256),) * 3
    # sub-matrix, not the whole matrix
    assert _both(lib, R, 16384, 16384, 16384, 2, 4, "R", "N", "N", big, sub=((2, 1), (1, 1), (1, 1)), extra=1) == ("", "")
    # too small per rank (<= 1e7 elements)
    assert _both(lib, R, 4096, 4096, 4096, 2, 4, "R", "N", "N", big) == ("", "")
    # blocks do not divide the matrix / grid does not divide the block grid
    assert _both(lib, R, 16384, 16384, 16384, 2, 4, "R", "N", "N", ((250, 256),) * 3) == ("", "")
    assert _both(lib, R, 16384, 16384, 16384, 3, 2, "R", "N", "N", big) == ("", "")
