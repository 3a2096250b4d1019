"""GPU parity of the COSTA relayout path through the C ABI: the batched copy/transpose kernel (R3/R4) against the
oracle's copy_and_transform (bit-exact, also for general alpha/beta: the kernel uses unfused multiplies and adds in the
reference's order), costa::transform on the device, multiply_using_layout and p?gemm against dense oracles
(integer-valued inputs -> exact), single GPU and 2/4/8 GPUs.

Mirrors the reference's tests/pdgemm.cpp (descriptor cases incl. sub-matrices, transposes, NaN-filled C with beta = 0),
tests/multiply_using_layout.cpp and libs/COSTA/tests/unit/test_utils.cpp."""
import os
import time
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import costa_sim as sim  # noqa: E402
from cosma_b200 import costa  # noqa: E402

TDT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dtype", ["s", "d", "c", "z"])
@pytest.mark.parametrize("n_rows,n_cols", [(37, 53), (1, 1), (32, 32), (33, 31), (200, 257), (3, 500)])
def test_relayout_kernel_bit_exact(oracle, dtype, n_rows, n_cols):
    rng = np.random.default_rng(n_rows * 1000 + n_cols)
    pieces, checks = [], []
    for src_ord in "CR":
        for dst_ord in "CR":
            for transpose in (0, 1):
                for conj in (0, 1):
                    for alpha, beta in ((1.0, 0.0), (2.0, -0.5), ((0.75 - 1.5j, 0.25 + 2j) if dtype in "cz" else (-3.0, 1.0)), (3.0, 0.0)):
                        dr, dc = (n_cols, n_rows) if transpose else (n_rows, n_cols)
                        sld = (n_rows if src_ord == "C" else n_cols) + int(rng.integers(0, 4))
                        dld = (dr if dst_ord == "C" else dc) + int(rng.integers(0, 4))
                        src = sim.random_values(rng, sld * max(n_rows, n_cols), dtype, ints=False)
                        d0 = sim.random_values(rng, dld * max(dr, dc), dtype, ints=False)
                        want = oracle.copy_and_transform(n_rows, n_cols, src, sld, src_ord, d0.copy(), dld, dst_ord, transpose, conj, alpha, beta)
                        ds, dd = _dev(src), _dev(d0)
                        pieces.append({"src": ds.data_ptr(), "dst": dd.data_ptr(), "src_ld": sld, "dst_ld": dld, "n_rows": n_rows, "n_cols": n_cols,
                                       "src_ordering": src_ord, "dst_ordering": dst_ord, "transpose": transpose, "conjugate": conj,
                                       "alpha": alpha, "beta": beta})
                        checks.append((ds, dd, want, (src_ord, dst_ord, transpose, conj, alpha, beta)))
    costa.relayout_batch(dtype, pieces)  # ONE launch for all 128 pieces
    torch.cuda.synchronize()
    for ds, dd, want, what in checks:
        got = dd.cpu().numpy()
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), what


def test_relayout_beta_zero_never_reads_dest(oracle):
    rng = np.random.default_rng(3)
    src = sim.random_values(rng, 64 * 48, "z", ints=False)
    dst = torch.full((64 * 48,), float("nan"), dtype=torch.complex128, device="cuda")
    ds = _dev(src)
    costa.relayout_batch("z", [{"src": ds.data_ptr(), "dst": dst.data_ptr(), "n_rows": 64, "n_cols": 48, "transpose": 1, "conjugate": 1,
                                "alpha": 2.0 + 1j, "beta": 0.0}])
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    assert not np.isnan(got).any()
    want = oracle.copy_and_transform(64, 48, src, 64, "C", np.zeros(64 * 48, dtype=np.complex128), 48, "C", 1, 1, 2.0 + 1j, 0.0)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


class DeviceDist:
    """sim.DistMatrix mirrored into device memory (one allocation per block) for single-process tests."""

    def __init__(self, host):
        self.host = host
        self.dev = {key: _dev(arr) for key, (arr, ld) in host.store.items()}

    def layout(self, rank):
        h = self.host
        blocks = [(bi, bj, self.dev[(bi, bj)].data_ptr(), ld) for (bi, bj), (arr, ld) in sorted(h.store.items()) if h.owners[bi, bj] == rank]
        return costa.custom_layout(h.rowsplit, h.colsplit, h.owners, blocks, h.ordering)

    def download(self):
        for key, t in self.dev.items():
            self.host.store[key][0][...] = t.cpu().numpy()
        return self.host.gather()


def _rand_dist(rng, m, n, P, dtype, ordering=None):
    rs = sim.random_split(rng, m, int(rng.integers(1, 5)))
    cs = sim.random_split(rng, n, int(rng.integers(1, 5)))
    owners = rng.integers(0, P, size=(len(rs) - 1, len(cs) - 1))
    return sim.DistMatrix(rs, cs, owners, P, dtype, ordering or "CR"[int(rng.integers(0, 2))], pad=int(rng.integers(0, 3)))


@pytest.mark.parametrize("dtype,op", [("d", "N"), ("d", "T"), ("z", "C"), ("s", "T"), ("c", "C")])
def test_single_gpu_transform(lib, dtype, op):
    rng = np.random.default_rng(ord(op) + ord(dtype))
    for trial in range(3):
        m, n = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        sm, sn = (m, n) if op == "N" else (n, m)
        F, T = _rand_dist(rng, sm, sn, 1, dtype), _rand_dist(rng, m, n, 1, dtype)
        G, H = sim.random_values(rng, (sm, sn), dtype), sim.random_values(rng, (m, n), dtype)
        F.scatter(G); T.fill_padding(55); T.scatter(H)
        dF, dT = DeviceDist(F), DeviceDist(T)
        alpha, beta = ((1.0, 0.0), (2.0, 3.0), (1.0, 1.0))[trial]
        tp = costa.TransformPlan(None, dtype, [(dF.layout(0), dT.layout(0), op, alpha, beta)], rank=0, nranks=1)
        tp.run(); tp.run() if beta == 0.0 else None  # plans are reusable
        torch.cuda.synchronize()
        assert 1 <= tp.stats()["launches"] <= 4  # one launch per non-empty piece class (copy|transpose x 16-byte|element requests)
        tp.destroy()
        want = alpha * sim.apply_op(G, op) + beta * H
        assert np.array_equal(dT.download(), want.astype(sim.NP[dtype]))
        for key, (arr, ld) in T.store.items():  # padding untouched
            r, c = T.rowsplit[key[0] + 1] - T.rowsplit[key[0]], T.colsplit[key[1] + 1] - T.colsplit[key[1]]
            inner, outer = (r, c) if T.ordering == "C" else (c, r)
            pad = arr[:ld * outer].reshape(outer, ld)[:, inner:]
            assert (pad == 55).all()


@pytest.mark.parametrize("dtype", ["d", "z", "s", "c"])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "C"), ("C", "T")])
def test_single_gpu_multiply_using_layout(lib, dtype, ta, tb):
    """tests/multiply_using_layout.cpp on one rank, plus transposes: custom block layouts in, dense oracle out."""
    from cosma_b200.distributed import init_comm
    comm = init_comm()
    rng = np.random.default_rng(ord(ta) * 7 + ord(tb))
    for (m, n, k, alpha, beta) in ((100, 80, 60, 1.0, 1.0), (257, 130, 95, 2.0, 0.0), (64, 64, 64, 1.0, -1.0)):
        if dtype in "zc":
            alpha = alpha * (1 - 0.5j)
        A = sim.random_values(rng, (m, k) if ta == "N" else (k, m), dtype)
        B = sim.random_values(rng, (k, n) if tb == "N" else (n, k), dtype)
        C = sim.random_values(rng, (m, n), dtype)
        dA, dB, dC = (_rand_dist(rng, X.shape[0], X.shape[1], 1, dtype, "C") for X in (A, B, C))
        dA.scatter(A); dB.scatter(B)
        dC.fill_padding(7)
        dC.scatter(C if beta != 0.0 else np.full_like(C, np.nan))
        gA, gB, gC = DeviceDist(dA), DeviceDist(dB), DeviceDist(dC)
        costa.multiply_using_layout(comm, dtype, ta, tb, alpha, gA.layout(0), gB.layout(0), beta, gC.layout(0))
        torch.cuda.synchronize()
        wide = np.complex128 if dtype in "zc" else np.float64
        want = alpha * (sim.apply_op(A, ta).astype(wide) @ sim.apply_op(B, tb).astype(wide)) + (beta * C.astype(wide) if beta != 0.0 else 0)
        assert np.array_equal(gC.download(), want.astype(C.dtype))
    comm.destroy()


PX_CASES = [
    # M_a, N_a, M_b, N_b, M_c, N_c, mb/nb a, b, c, (ia, ja), (ib, jb), (ic, jc), m, n, k, ta, tb, alpha, beta
    dict(m=96, n=80, k=64, ta="N", tb="N", alpha=1.0, beta=0.0, blk=((8, 8), (8, 8), (8, 8))),
    dict(m=96, n=80, k=64, ta="T", tb="N", alpha=2.0, beta=1.0, blk=((16, 8), (8, 4), (32, 8))),
    dict(m=50, n=60, k=70, ta="N", tb="T", alpha=1.0, beta=-1.0, blk=((7, 5), (3, 9), (4, 4)), sub=((3, 2), (5, 4), (2, 6)), extra=12),
    dict(m=33, n=17, k=129, ta="C", tb="C", alpha=1.0, beta=0.0, blk=((5, 5), (6, 6), (7, 7)), sub=((1, 4), (2, 1), (3, 3)), extra=9),
    dict(m=64, n=64, k=0, ta="N", tb="N", alpha=1.0, beta=3.0, blk=((8, 8), (8, 8), (8, 8))),      # k = 0: C *= beta
    dict(m=64, n=64, k=32, ta="N", tb="N", alpha=0.0, beta=0.0, blk=((8, 8), (8, 8), (8, 8))),     # alpha = 0, beta = 0: C = 0
    dict(m=0, n=64, k=32, ta="N", tb="N", alpha=1.0, beta=0.0, blk=((8, 8), (8, 8), (8, 8))),      # m = 0: no-op
]


def _pxgemm_case(comm, grid, case, dtype, nprow, npcol, order, host_pointers, gather):
    """Runs one p?gemm case on the calling rank; `gather(loc)` returns the list of all ranks' local C arrays (numpy)."""
    rank, P = comm.rank, comm.size
    m, n, k, ta, tb = case["m"], case["n"], case["k"], case["ta"], case["tb"]
    alpha, beta = case["alpha"], case["beta"]
    if dtype in "zc" and alpha not in (0.0,):
        alpha = alpha * (1 + 0.5j)
    extra = case.get("extra", 0)
    (ia, ja), (ib, jb), (ic, jc) = case.get("sub", ((1, 1), (1, 1), (1, 1)))
    am, an = (m, k) if ta == "N" else (k, m)
    bm, bn = (k, n) if tb == "N" else (n, k)
    shapes = [(max(am, 1) + ia - 1 + extra, max(an, 1) + ja - 1 + extra), (max(bm, 1) + ib - 1 + extra, max(bn, 1) + jb - 1 + extra),
              (max(m, 1) + ic - 1 + extra, max(n, 1) + jc - 1 + extra)]
    rng = np.random.default_rng(m * 31 + n * 17 + k)  # same on every rank
    G = [sim.random_values(rng, s, dtype) for s in shapes]
    rsrc, csrc = (1 % nprow, 1 % npcol) if extra else (0, 0)
    bc = [sim.BlockCyclic(s[0], s[1], blk[0], blk[1], nprow, npcol, order, rsrc, csrc, lld_pad=1) for s, blk in zip(shapes, case["blk"])]
    Cin = G[2].copy()
    if beta == 0.0:
        Cin[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = np.nan  # C must not be read when beta == 0 (utils/pxgemm_utils.hpp:603-637)
    locs = [bc[0].scatter(G[0], rank), bc[1].scatter(G[1], rank), bc[2].scatter(Cin, rank)]
    if host_pointers:
        bufs = [torch.from_numpy(l).pin_memory() for l in locs]
    else:
        bufs = [_dev(l) for l in locs]
    costa.pxgemm(grid, dtype, ta, tb, m, n, k, alpha, bufs[0].data_ptr(), ia, ja, bc[0].desc(rank), bufs[1].data_ptr(), ib, jb, bc[1].desc(rank),
                 beta, bufs[2].data_ptr(), ic, jc, bc[2].desc(rank))
    torch.cuda.synchronize()
    all_c = gather(bufs[2].cpu().numpy())
    if rank != 0:
        return True
    got = np.zeros_like(G[2])
    for r in range(P):
        bc[2].gather_into(got, all_c[r], r)
    want = G[2].copy()
    if m and n:
        wide = np.complex128 if dtype in "zc" else np.float64
        As = sim.apply_op(G[0][ia - 1:ia - 1 + am, ja - 1:ja - 1 + an], ta).astype(wide)
        Bs = sim.apply_op(G[1][ib - 1:ib - 1 + bm, jb - 1:jb - 1 + bn], tb).astype(wide)
        prod = alpha * (As @ Bs) if k and alpha != 0 else 0
        want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = np.asarray(prod + (beta * G[2][ic - 1:ic - 1 + m, jc - 1:jc - 1 + n].astype(wide) if beta != 0.0 else 0)).astype(want.dtype)
    return bool(np.array_equal(got, want))


@pytest.mark.parametrize("dtype", ["d", "z", "s", "c"])
@pytest.mark.parametrize("host_pointers", [False, True])
def test_single_gpu_pxgemm(lib, dtype, host_pointers):
    from cosma_b200.distributed import init_comm
    comm = init_comm()
    grid = costa.Grid(comm, "R", 1, 1)
    for case in PX_CASES:
        assert _pxgemm_case(comm, grid, case, dtype, 1, 1, "R", host_pointers, lambda loc: [loc]), case
    grid.destroy(); comm.destroy()


def _reference_sets(max_ranks):
    """The reference's own p?gemm parameter sets (tests/pdgemm.cpp via tests/golden/pdgemm_cases.json) that fit max_ranks ranks."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pdgemm_cases.json")) as f:
        return [c for c in json.load(f)["cases"] if c["p_rows"] * c["p_cols"] <= max_ranks]


def _reference_set_case(comm, c, dtype, host_pointers, gather):
    """One parameter set of the reference's tests/pdgemm.cpp through the C ABI, every argument as the set names it (global sizes, block
    sizes, sub-matrix origins, rsrc/csrc per matrix, lld = the set's own -- mostly exactly the local row count, i.e. ODD pitches).
    Integer-valued operands: the dense product is exact in every type; sub(C) holds NaN when beta == 0 (pxgemm_utils.hpp:603-637)."""
    rank = comm.rank
    nprow, npcol, order = c["p_rows"], c["p_cols"], c["order"]
    P = nprow * npcol
    m, n, k, ta, tb, alpha, beta = c["m"], c["n"], c["k"], c["ta"], c["tb"], c["alpha"], c["beta"]
    (ia, ja), (ib, jb), (ic, jc) = (c["ia"], c["ja"]), (c["ib"], c["jb"]), (c["ic"], c["jc"])
    shapes = [(c["ma"], c["na"]), (c["mb"], c["nb"]), (c["mc"], c["nc"])]
    blks = [(c["bma"], c["bna"]), (c["bmb"], c["bnb"]), (c["bmc"], c["bnc"])]
    srcs = [(c["src_ma"], c["src_na"]), (c["src_mb"], c["src_nb"]), (c["src_mc"], c["src_nc"])]
    llds = [c["lld_a"], c["lld_b"], c["lld_c"]]
    am, an = (m, k) if ta == "N" else (k, m)
    bm, bn = (k, n) if tb == "N" else (n, k)
    rng = np.random.default_rng(m * 31 + n * 17 + k)  # same on every rank
    G = [sim.random_values(rng, s, dtype) for s in shapes]
    grid = costa.Grid(comm, order, nprow, npcol)
    bc = [sim.BlockCyclic(s[0], s[1], blk[0], blk[1], nprow, npcol, order, src[0], src[1], lld=(lld if lld > 0 else None))
          for s, blk, src, lld in zip(shapes, blks, srcs, llds)]
    Cin = G[2].copy()
    if beta == 0.0:
        Cin[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = np.nan
    locs = [bc[0].scatter(G[0], rank), bc[1].scatter(G[1], rank), bc[2].scatter(Cin, rank)]
    bufs = [torch.from_numpy(l).pin_memory() for l in locs] if host_pointers else [_dev(l) for l in locs]
    costa.pxgemm(grid, dtype, ta, tb, m, n, k, alpha, bufs[0].data_ptr(), ia, ja, bc[0].desc(rank), bufs[1].data_ptr(), ib, jb, bc[1].desc(rank),
                 beta, bufs[2].data_ptr(), ic, jc, bc[2].desc(rank))
    torch.cuda.synchronize()
    all_c = gather(bufs[2].cpu().numpy())
    grid.destroy()
    if rank != 0:
        return True
    got = np.zeros_like(G[2])
    for r in range(P):
        bc[2].gather_into(got, all_c[r], r)
    wide = np.complex128 if dtype in "zc" else np.float64
    As = sim.apply_op(G[0][ia - 1:ia - 1 + am, ja - 1:ja - 1 + an], ta).astype(wide)
    Bs = sim.apply_op(G[1][ib - 1:ib - 1 + bm, jb - 1:jb - 1 + bn], tb).astype(wide)
    want = G[2].copy()
    want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = np.asarray(alpha * (As @ Bs) + (beta * G[2][ic - 1:ic - 1 + m, jc - 1:jc - 1 + n].astype(wide) if beta != 0.0 else 0)).astype(want.dtype)
    if float(2 * alpha).is_integer() and float(2 * beta).is_integer():
        return bool(np.array_equal(got, want))  # integers and halves: every operation is exact in every type
    # scalars like 1.2 (the reference uses them) are rounded to the matrix type by the library but not by the expectation formed in
    # double: element-wise relative tolerance of a few units in the last place of the type
    tol = 4e-7 if dtype in "sc" else 1e-15
    return bool(np.isfinite(got).all() and np.all(np.abs(got - want) <= tol * np.maximum(np.abs(want), 1.0)))


@pytest.mark.parametrize("dtype", ["d", "z", "s", "c"])
@pytest.mark.parametrize("host_pointers", [False, True])
def test_single_gpu_pxgemm_reference_sets(lib, dtype, host_pointers):
    """The seven P = 1 sets of the reference's tests/pdgemm.cpp below the C++ layer: odd sizes (83 / 77 / 13 / 11 / 7 / 3) with lld == local
    rows -- the operands TMA cannot address -- and 128 x 128 x 1280 'T','N'."""
    from cosma_b200.distributed import init_comm
    comm = init_comm()
    sets = _reference_sets(1)
    assert len(sets) == 7
    for c in sets:
        assert _reference_set_case(comm, c, dtype, host_pointers, lambda loc: [loc]), c
    comm.destroy()


TRAN_CASES = [
    dict(m=40, n=56, ba=(8, 8), bc=(8, 8), sa=(1, 1), sc=(1, 1), extra=0, alpha=1.0, beta=0.0),
    dict(m=37, n=53, ba=(5, 7), bc=(4, 9), sa=(3, 2), sc=(2, 6), extra=11, alpha=2.0, beta=-1.0),
    dict(m=64, n=16, ba=(16, 4), bc=(8, 32), sa=(1, 5), sc=(7, 1), extra=3, alpha=1.0, beta=1.0),
    dict(m=300, n=200, ba=(32, 32), bc=(64, 16), sa=(1, 1), sc=(1, 1), extra=0, alpha=1.0, beta=0.0),
]


def _pxtran_case(comm, grid, grid2, case, dtype, op, nprow, npcol, order, order2, host_pointers, gather):
    """One p?tran (op 'T'|'C') or p?gemr2d (op 'N', A on `grid`, C on `grid2`) case; every rank's local C is compared bit for
    bit with the dense definition scattered back (and, on rank 0, with the live reference wrappers when oracle/_ref exists)."""
    rank, P = comm.rank, comm.size
    m, n, extra = case["m"], case["n"], case["extra"]
    (ia, ja), (ic, jc) = case["sa"], case["sc"]
    alpha, beta = (case["alpha"], case["beta"]) if op != "N" else (1.0, 0.0)
    if dtype in "cz" and op != "N":
        alpha = alpha * (1 - 0.5j)
    am, an = (m, n) if op == "N" else (n, m)
    rng = np.random.default_rng(m * 13 + n)
    GA = sim.random_values(rng, (am + ia - 1 + extra, an + ja - 1 + extra), dtype)
    GC = sim.random_values(rng, (m + ic - 1 + extra, n + jc - 1 + extra), dtype)
    bcA = sim.BlockCyclic(GA.shape[0], GA.shape[1], case["ba"][0], case["ba"][1], nprow, npcol, order, 0, 0, lld_pad=2)
    bcC = sim.BlockCyclic(GC.shape[0], GC.shape[1], case["bc"][0], case["bc"][1], nprow, npcol, order2, 0, 0, lld_pad=1)
    Cin = GC.copy()
    if beta == 0.0:
        Cin[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = np.nan  # must not be read
    a_loc, c_loc = bcA.scatter(GA, rank, fill=77), bcC.scatter(Cin, rank, fill=55)
    if host_pointers:
        bufs = [torch.from_numpy(a_loc).pin_memory(), torch.from_numpy(c_loc).pin_memory()]
    else:
        bufs = [_dev(a_loc), _dev(c_loc)]
    if op == "N":
        costa.pxgemr2d(grid, grid2, dtype, m, n, bufs[0].data_ptr(), ia, ja, bcA.desc(rank), bufs[1].data_ptr(), ic, jc, bcC.desc(rank))
    else:
        costa.pxtran(grid, dtype, op, m, n, alpha, bufs[0].data_ptr(), ia, ja, bcA.desc(rank), beta, bufs[1].data_ptr(), ic, jc, bcC.desc(rank))
    torch.cuda.synchronize()
    dense = GC.copy()
    dense[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = alpha * sim.apply_op(GA[ia - 1:ia - 1 + am, ja - 1:ja - 1 + an], op) + \
        (beta * GC[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] if beta != 0.0 else 0)
    want = bcC.scatter(dense.astype(GC.dtype), rank, fill=55)
    return bool(np.array_equal(bufs[1].cpu().numpy().view(np.uint8), want.view(np.uint8)))


@pytest.mark.parametrize("dtype,op", [("d", "T"), ("s", "T"), ("z", "T"), ("z", "C"), ("c", "C"), ("d", "N"), ("z", "N"), ("s", "N"), ("c", "N")])
@pytest.mark.parametrize("host_pointers", [False, True])
def test_single_gpu_pxtran_pxgemr2d(lib, dtype, op, host_pointers):
    from cosma_b200.distributed import init_comm
    comm = init_comm()
    grid = costa.Grid(comm, "R", 1, 1)
    for case in TRAN_CASES:
        assert _pxtran_case(comm, grid, grid, case, dtype, op, 1, 1, "R", "R", host_pointers, None), case
    grid.destroy(); comm.destroy()


# ---- multi-GPU ------------------------------------------------------------------------------------------------------

def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, nprow, npcol, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cosma_b200.distributed import init_comm
    comm = init_comm()

    def gather(loc):
        n = torch.tensor([loc.view(np.uint8).size], device="cuda")
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n)
        mx = max(int(s.item()) for s in sizes)
        mine = torch.zeros(max(mx, 1), dtype=torch.uint8, device="cuda")
        mine[:loc.view(np.uint8).size] = torch.from_numpy(loc.view(np.uint8).copy()).cuda()
        allb = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        return [allb[r][:int(sizes[r].item())].cpu().numpy().view(loc.dtype) for r in range(world)]

    ok = []
    # (1) costa::transform between random layouts with random owners, two transforms per exchange
    for dtype, op in (("d", "T"), ("z", "C"), ("d", "N")):
        rng = np.random.default_rng(77 + ord(op))  # same on every rank
        mats, specs = [], []
        for t in range(2):
            m, n = int(rng.integers(40, 400)), int(rng.integers(40, 400))
            sm, sn = (m, n) if op == "N" else (n, m)
            F, T = _rand_dist(rng, sm, sn, world, dtype), _rand_dist(rng, m, n, world, dtype)
            Gm, Hm = sim.random_values(rng, (sm, sn), dtype), sim.random_values(rng, (m, n), dtype)
            F.scatter(Gm); T.scatter(Hm)
            specs.append(((1.0, 0.0), (2.0, 1.0))[t])
            mats.append((DeviceDist(F), DeviceDist(T), specs[t][0] * sim.apply_op(Gm, op) + specs[t][1] * Hm))
        tp = costa.TransformPlan(comm, dtype, [(F.layout(rank), T.layout(rank), op, a, b) for (F, T, _), (a, b) in zip(mats, specs)])
        tp.run()
        torch.cuda.synchronize()
        tp.destroy()
        for F, T, want in mats:
            # every rank checks the blocks it owns
            got = T.download()
            h = T.host
            for bi in range(len(h.rowsplit) - 1):
                for bj in range(len(h.colsplit) - 1):
                    if h.owners[bi, bj] == rank:
                        sl = (slice(h.rowsplit[bi], h.rowsplit[bi + 1]), slice(h.colsplit[bj], h.colsplit[bj + 1]))
                        ok.append(bool(np.array_equal(got[sl], want[sl].astype(got.dtype))))
    # (2) p?gemm on a block-cyclic grid
    for order in ("R", "C"):
        grid = costa.Grid(comm, order, nprow, npcol)
        for dtype in ("d", "z", "s", "c"):
            for host in (False, True):
                for case in PX_CASES:
                    ok.append(_pxgemm_case(comm, grid, case, dtype, nprow, npcol, order, host, gather))
        grid.destroy()
    # (2b) the reference's own parameter sets (tests/pdgemm.cpp) whose grid fits this job, each on the grid it names (the first
    # p_rows x p_cols ranks; the others call with empty local arrays, as ScaLAPACK requires of every process of the context's parent)
    for c in _reference_sets(world):
        for dtype, host in (("d", False), ("c", True)):
            ok.append(_reference_set_case(comm, c, dtype, host, gather))
    # (3) p?tran / p?tranu / p?tranc and p?gemr2d (between a row-major and a column-major numbering of the same grid)
    gr, gc2 = costa.Grid(comm, "R", nprow, npcol), costa.Grid(comm, "C", nprow, npcol)
    for dtype, op in (("d", "T"), ("z", "C"), ("c", "T"), ("s", "N"), ("z", "N")):
        for host in (False, True):
            for case in TRAN_CASES:
                ok.append(_pxtran_case(comm, gr, gc2 if op == "N" else gr, case, dtype, op, nprow, npcol, "R", "C" if op == "N" else "R", host, gather))
    gr.destroy(); gc2.destroy()
    # (4) multiply_using_layout across ranks: (a) random layouts with random owners; (b) COSMA's own layout with the rank
    # labels reversed -- the relabelling must find the permutation, so that no element leaves its GPU
    from cosma_b200 import planning

    def own_blocks_equal(D, dev, want):
        got = dev.download()
        h = D
        res = []
        for bi in range(len(h.rowsplit) - 1):
            for bj in range(len(h.colsplit) - 1):
                if h.owners[bi, bj] == rank:
                    sl = (slice(h.rowsplit[bi], h.rowsplit[bi + 1]), slice(h.colsplit[bj], h.colsplit[bj + 1]))
                    res.append(bool(np.array_equal(got[sl], want[sl].astype(got.dtype))))
        return all(res)

    for trial, (dtype, ta, tb, m, n, k, alpha, beta) in enumerate((("d", "N", "N", 96, 80, 64, 1.0, 0.0), ("z", "C", "N", 70, 90, 50, 2.0, 1.0),
                                                                   ("s", "N", "T", 128, 64, 96, 1.0, -1.0), ("c", "T", "C", 60, 60, 60, 1.0, 0.0))):
        rng = np.random.default_rng(900 + trial)  # same on every rank
        A = sim.random_values(rng, (m, k) if ta == "N" else (k, m), dtype)
        B = sim.random_values(rng, (k, n) if tb == "N" else (n, k), dtype)
        C = sim.random_values(rng, (m, n), dtype)
        dA, dB, dC = (_rand_dist(rng, X.shape[0], X.shape[1], world, dtype, "C") for X in (A, B, C))
        dA.scatter(A); dB.scatter(B); dC.scatter(C if beta != 0.0 else np.full_like(C, np.nan))
        gA, gB, gC = DeviceDist(dA), DeviceDist(dB), DeviceDist(dC)
        costa.multiply_using_layout(comm, dtype, ta, tb, alpha, gA.layout(rank), gB.layout(rank), beta, gC.layout(rank))
        torch.cuda.synchronize()
        wide = np.complex128 if dtype in "zc" else np.float64
        want = alpha * (sim.apply_op(A, ta).astype(wide) @ sim.apply_op(B, tb).astype(wide)) + (beta * C.astype(wide) if beta != 0.0 else 0)
        ok.append(own_blocks_equal(dC, gC, want))
    m = n = k = 640  # every split keeps the local dimensions above COSMA_MIN_LOCAL_DIMENSION (200)
    steps, P_used, _ = planning.strategy(m, n, k, world)
    sigma = [world - 1 - r for r in range(world)]
    mats = {}
    rng = np.random.default_rng(4242)
    for label, (rows, cols) in (("A", (m, k)), ("B", (k, n)), ("C", (m, n))):
        per_rank = planning.mapper_layout(label, m, n, k, world, steps)
        rs = sorted({b[0] for bl in per_rank for b in bl} | {rows})
        cs = sorted({b[2] for bl in per_rank for b in bl} | {cols})
        owners = np.zeros((len(rs) - 1, len(cs) - 1), dtype=np.int32)
        for r, bl in enumerate(per_rank):
            for (r0, r1, c0, c1) in bl:
                for bi in range(len(rs) - 1):
                    for bj in range(len(cs) - 1):
                        if r0 <= rs[bi] <= r1 and c0 <= cs[bj] <= c1:
                            owners[bi, bj] = sigma[r]
        D = sim.DistMatrix(rs, cs, owners, world, "d", "C")
        G = sim.random_values(rng, (rows, cols), "d")
        D.scatter(G if label != "C" else np.full_like(G, np.nan))
        mats[label] = (D, DeviceDist(D), G)
    costa.multiply_using_layout(comm, "d", "N", "N", 1.0, mats["A"][1].layout(rank), mats["B"][1].layout(rank), 0.0, mats["C"][1].layout(rank))
    torch.cuda.synchronize()
    st = costa.last_layout_multiply_stats(comm)
    ok.append(own_blocks_equal(mats["C"][0], mats["C"][1], mats["A"][2] @ mats["B"][2]))
    if P_used == world and os.environ.get("COSMA_B200_REORDER_RANKS", "ON").upper() == "ON":  # the default, as in the reference
        ok.append(st["in_remote_elements"] == 0 and st["out_remote_elements"] == 0 and st["in_local_elements"] > 0)
    if not all(ok):
        print("rank %d: failed checks (index of %d): %s" % (rank, len(ok), [i for i, v in enumerate(ok) if not v]), flush=True)
    t = torch.tensor([1 if all(ok) else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(bool(t.item()))
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


def _run_world(world, nprow, npcol):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nprow, npcol, q)) for r in range(world)]
    for p in procs:
        p.start()
    # one wall-clock limit for the whole world: a rank stuck in a collective must cost minutes, not the GPU call
    deadline = time.time() + 300
    for p in procs:
        p.join(max(1.0, deadline - time.time()))
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.terminate()
    assert not hung, "ranks still running after the time limit"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=10)


def test_two_gpus(lib):
    _run_world(2, 2, 1)


def test_four_gpus(lib):
    _run_world(4, 2, 2)


def test_eight_gpus(lib):
    _run_world(8, 2, 4)
