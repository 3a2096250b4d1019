"""Test helpers for the COSTA relayout path: distributed matrices over host (numpy) storage, and a CPU lock-step
interpreter of exported transform plans (pack -> exchange -> local/unpack) that uses the ORACLE's copy_and_transform on
raw addresses. Lets the planning layer of every rank be checked without a GPU or a multi-process launch."""
import numpy as np

from cosma_b200 import costa

NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128, "i": np.int32}


class DistMatrix:
    """A global matrix cut by (rowsplit, colsplit) with owners[i][j]; every rank stores each of its blocks in its own
    numpy allocation with leading dimension = tight + pad, ordering 'C' or 'R'."""

    def __init__(self, rowsplit, colsplit, owners, n_ranks, dtype, ordering="C", pad=0):
        self.rowsplit, self.colsplit = list(rowsplit), list(colsplit)
        self.owners = np.asarray(owners, dtype=np.int32).reshape(len(self.rowsplit) - 1, len(self.colsplit) - 1)
        self.n_ranks, self.dtype, self.ordering, self.pad = n_ranks, dtype, ordering, pad
        self.store = {}  # (bi, bj) -> (array, ld)
        for bi in range(len(self.rowsplit) - 1):
            for bj in range(len(self.colsplit) - 1):
                r, c = self.rowsplit[bi + 1] - self.rowsplit[bi], self.colsplit[bj + 1] - self.colsplit[bj]
                ld = (r if ordering == "C" else c) + pad
                n = ld * (c if ordering == "C" else r)
                self.store[(bi, bj)] = (np.zeros(max(n, 1), dtype=NP[dtype]), max(ld, 1))

    @property
    def shape(self):
        return self.rowsplit[-1], self.colsplit[-1]

    def _view(self, bi, bj):
        arr, ld = self.store[(bi, bj)]
        r, c = self.rowsplit[bi + 1] - self.rowsplit[bi], self.colsplit[bj + 1] - self.colsplit[bj]
        if self.ordering == "C":
            return arr[:ld * c].reshape(c, ld).T[:r, :]
        return arr[:ld * r].reshape(r, ld)[:, :c]

    def scatter(self, G):
        for (bi, bj) in self.store:
            self._view(bi, bj)[...] = G[self.rowsplit[bi]:self.rowsplit[bi + 1], self.colsplit[bj]:self.colsplit[bj + 1]]

    def gather(self):
        G = np.zeros(self.shape, dtype=NP[self.dtype])
        for (bi, bj) in self.store:
            G[self.rowsplit[bi]:self.rowsplit[bi + 1], self.colsplit[bj]:self.colsplit[bj + 1]] = self._view(bi, bj)
        return G

    def fill_padding(self, value):
        """Poison everything (call before scatter) so that stray writes into padding are detected."""
        for arr, _ in self.store.values():
            arr[...] = value

    def layout(self, rank):
        blocks = [(bi, bj, arr.ctypes.data, ld) for (bi, bj), (arr, ld) in sorted(self.store.items()) if self.owners[bi, bj] == rank]
        return costa.custom_layout(self.rowsplit, self.colsplit, self.owners, blocks, self.ordering)

    def ref_tuple(self, rank=0):
        blocks = [(bi, bj, arr.ctypes.data, ld) for (bi, bj), (arr, ld) in sorted(self.store.items()) if self.owners[bi, bj] == rank]
        return (self.rowsplit, self.colsplit, self.owners, blocks, self.ordering)


def random_split(rng, total, parts):
    if parts == 1 or total == 0:
        return [0, total] if parts == 1 else [0] + sorted(rng.integers(0, total + 1, size=parts - 1).tolist()) + [total]
    cuts = sorted(rng.choice(np.arange(1, total), size=min(parts - 1, total - 1), replace=False).tolist())
    return [0] + cuts + [total]


def random_values(rng, shape, dtype, ints=True):
    if ints:
        # non-zero integers: exact in every dtype, and no signed zeros (the reference's conjugating copy path evaluates
        # 0*dest + 1*conj(x) even for alpha = 1, beta = 0, which turns -0.0 into +0.0; memory_utils.hpp:28-41)
        def nz(size):
            return rng.integers(1, 10, size=size) * rng.choice([-1, 1], size=size)
        v = nz(shape).astype(np.float64)
        if dtype in ("c", "z"):
            v = v + 1j * nz(shape)
    else:
        v = rng.standard_normal(shape)
        if dtype in ("c", "z"):
            v = v + 1j * rng.standard_normal(shape)
    return v.astype(NP[dtype])


def apply_op(G, op):
    return G if op == "N" else (G.T if op == "T" else G.conj().T)


def run_pieces(oracle, dtype, pieces, specs, src_base=0, dst_base=0):
    for p in pieces:
        alpha, beta = (1.0, 0.0) if p["transform"] < 0 else specs[p["transform"]]
        oracle.copy_and_transform_raw(dtype, p["n_rows"], p["n_cols"], p["src"] + src_base, p["src_ld"], p["src_ordering"], p["dst"] + dst_base,
                                      p["dst_ld"], p["dst_ordering"], p["transpose"], p["conjugate"], alpha, beta)


def simulate(oracle, dtype, plans, specs):
    """plans: exported plan per rank; specs: [(alpha, beta)] per transform. Executes all ranks in lock-step on the CPU."""
    P = len(plans)
    send = [np.zeros(max(pl["total_send"], 1), dtype=np.uint8) for pl in plans]
    recv = [np.zeros(max(pl["total_recv"], 1), dtype=np.uint8) for pl in plans]
    for r, pl in enumerate(plans):  # stage 1: pack + local
        run_pieces(oracle, dtype, pl["pack"], specs, dst_base=send[r].ctypes.data)
        run_pieces(oracle, dtype, pl["local"], specs)
    for r, pl in enumerate(plans):  # the all-to-all-v
        for p in range(P):
            n = pl["send_bytes"][p]
            assert n == plans[p]["recv_bytes"][r], "send/recv byte counts disagree between ranks %d and %d" % (r, p)
            if n:
                recv[p][plans[p]["recv_off"][r]:plans[p]["recv_off"][r] + n] = send[r][pl["send_off"][p]:pl["send_off"][p] + n]
    for r, pl in enumerate(plans):  # stage 2: unpack
        run_pieces(oracle, dtype, pl["unpack"], specs, src_base=recv[r].ctypes.data)


class BlockCyclic:
    """A ScaLAPACK 2D block-cyclic matrix: per-rank local arrays (column-major, leading dimension lld) and the
    global <-> local maps, built from the layout formulas that test_costa_cpu.py pins against the reference."""

    def __init__(self, M, N, mb, nb, nprow, npcol, order="R", rsrc=0, csrc=0, lld_pad=0, lld=None):
        self.M, self.N, self.mb, self.nb = M, N, mb, nb
        self.nprow, self.npcol, self.order, self.rsrc, self.csrc = nprow, npcol, order, rsrc, csrc
        self.lld_pad, self.lld = lld_pad, lld  # lld: the caller's leading dimension (at least the local row count), else rows + pad

    def coords(self, rank):
        if self.order == "C":
            return rank % self.nprow, rank // self.nprow
        return rank // self.npcol, rank % self.npcol

    def local_shape(self, rank):
        pr, pc = self.coords(rank)
        lr = costa.numroc(self.M, self.mb, pr, self.rsrc, self.nprow)
        lc = costa.numroc(self.N, self.nb, pc, self.csrc, self.npcol)
        if self.lld is not None:
            return max(self.lld, lr, 1), lc
        return max(lr, 1) + self.lld_pad, lc  # (lld, local columns)

    def desc(self, rank):
        return costa.descinit(self.M, self.N, self.mb, self.nb, self.rsrc, self.csrc, self.local_shape(rank)[0])

    def _blocks(self, rank):
        lld = self.local_shape(rank)[0]
        rs, cs, _, blocks = costa.scalapack_grid(lld, self.M, self.N, 1, 1, self.M, self.N, self.mb, self.nb, self.nprow, self.npcol, self.order,
                                                 self.rsrc, self.csrc, "C", rank)
        return lld, rs, cs, blocks

    def scatter(self, G, rank, fill=0):
        """-> the rank's local array (1-D, lld * local columns) holding its part of G."""
        lld, lc = self.local_shape(rank)
        loc = np.full(max(lld * lc, 1), fill, dtype=G.dtype)
        if rank >= self.nprow * self.npcol:
            return loc
        lld, rs, cs, blocks = self._blocks(rank)
        L = loc[:lld * lc].reshape(lc, lld).T if lc else None
        for (bi, bj, off) in blocks:
            r0, c0 = off % lld, off // lld
            L[r0:r0 + rs[bi + 1] - rs[bi], c0:c0 + cs[bj + 1] - cs[bj]] = G[rs[bi]:rs[bi + 1], cs[bj]:cs[bj + 1]]
        return loc

    def gather_into(self, G, loc, rank):
        if rank >= self.nprow * self.npcol:
            return
        lld, lc = self.local_shape(rank)
        lld, rs, cs, blocks = self._blocks(rank)
        L = loc[:lld * lc].reshape(lc, lld).T if lc else None
        for (bi, bj, off) in blocks:
            r0, c0 = off % lld, off // lld
            G[rs[bi]:rs[bi + 1], cs[bj]:cs[bj + 1]] = L[r0:r0 + rs[bi + 1] - rs[bi], c0:c0 + cs[bj + 1] - cs[bj]]
