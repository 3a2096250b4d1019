"""Host-side mirror of the COSTA layout / transform interface and of the layout-based multiply entry points.

  lay = block_cyclic_layout(m, n, mb, nb, 1, 1, m, n, nprow, npcol, 'R', 0, 0, ptr, lld, 'C', rank)   # costa::block_cyclic_layout
  lay = custom_layout(rowsplit, colsplit, owners, [(row, col, ptr, ld), ...])                        # costa::custom_layout
  tp = TransformPlan(comm, 'z', [(src, dst, 'C', 1.0, 0.0)])                                         # costa::transformer::schedule
  tp.run()                                                                                            # costa::transformer::transform
  multiply_using_layout(comm, 'd', 'N', 'T', alpha, A, B, beta, C)                                    # cinterface.hpp:42-76
  pxgemm(grid, 'z', 'C', 'N', m, n, k, alpha, a, 1, 1, desca, b, 1, 1, descb, beta, c, 1, 1, descc)   # pxgemm.h:6-107

Mirrors reference libs/COSTA/src/costa/layout.hpp:14-86, grid2grid/transformer.hpp:8-63, src/cosma/cinterface.hpp,
src/cosma/cosma_pxgemm.hpp:15-33. Pointers are raw addresses (ints): device addresses for execution, anything for
planning-only use."""
import ctypes

import numpy as np

from . import _lib
from .gemm import _stream_ptr

vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
pi = ctypes.POINTER(ctypes.c_int)
pd = ctypes.POINTER(ctypes.c_double)


class CBlock(ctypes.Structure):
    _fields_ = [("data", vp), ("ld", ci), ("row", ci), ("col", ci)]


class CLayout(ctypes.Structure):
    _fields_ = [("rowblocks", ci), ("colblocks", ci), ("rowsplit", pi), ("colsplit", pi), ("owners", pi),
                ("nlocalblocks", ci), ("localblocks", ctypes.POINTER(CBlock))]


class CPiece(ctypes.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("src_ld", i64), ("dst_ld", i64), ("n_rows", ci), ("n_cols", ci),
                ("src_ordering", ctypes.c_char), ("dst_ordering", ctypes.c_char), ("transpose", ctypes.c_char),
                ("conjugate", ctypes.c_char), ("alpha", ctypes.c_double * 2), ("beta", ctypes.c_double * 2)]


def _declare(lib):
    if getattr(lib, "_costa_declared", False):
        return
    cp = ctypes.c_char_p
    lib.cosma_b200_relayout_batch.argtypes = [vp, ctypes.c_char, ci, ctypes.POINTER(CPiece)]
    lib.cosma_b200_transform_plan_create.argtypes = [vp, ci, ci, ctypes.c_char, ci, ctypes.POINTER(CLayout), ctypes.POINTER(CLayout),
                                                     cp, cp, cp, pd, pd, ctypes.POINTER(vp)]
    lib.cosma_b200_transform_run.argtypes = [vp, vp]
    lib.cosma_b200_transform_plan_destroy.argtypes = [vp]
    lib.cosma_b200_transform_plan_export.argtypes = [vp, ctypes.POINTER(i64), i64, ctypes.POINTER(i64)]
    lib.cosma_b200_transform_plan_stats.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), pi]
    lib.cosma_b200_scalapack_layout.argtypes = [ci] * 11 + [ctypes.c_char, ci, ci, ctypes.c_char, ci, pi, pi, pi, pi, pi, pi, pi, pi,
                                                ctypes.POINTER(i64)]
    lib.cosma_b200_numroc.argtypes = [ci] * 5
    for name in ("cosma_b200_%smultiply_using_layout" % t for t in "sdcz"):
        getattr(lib, name).argtypes = [vp, cp, cp, pd, ctypes.POINTER(CLayout), ctypes.POINTER(CLayout), pd, ctypes.POINTER(CLayout), vp]
    lib.cosma_b200_grid_create.argtypes = [vp, ctypes.c_char, ci, ci, ctypes.POINTER(vp)]
    lib.cosma_b200_grid_destroy.argtypes = [vp]
    lib.cosma_b200_grid_info.argtypes = [vp, pi, pi, pi, pi]
    for name in ("cosma_b200_p%sgemm" % t for t in "sdcz"):
        getattr(lib, name).argtypes = [vp, ctypes.c_char, ctypes.c_char, ci, ci, ci, pd, vp, ci, ci, pi, vp, ci, ci, pi, pd, vp, ci, ci, pi, vp]
    lib.cosma_b200_pxtran.argtypes = [vp, ctypes.c_char, ctypes.c_char, ci, ci, pd, vp, ci, ci, pi, pd, vp, ci, ci, pi, vp]
    lib.cosma_b200_pxgemr2d.argtypes = [vp, vp, ctypes.c_char, ci, ci, vp, ci, ci, pi, vp, ci, ci, pi, vp]
    lib.cosma_b200_last_layout_multiply_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i64), cp, ci, pi]
    lib._costa_declared = True


def lib():
    L = _lib.load()
    _declare(L)
    return L


ELEM_BYTES = {"s": 4, "d": 8, "c": 8, "z": 16}
NP_DTYPE = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


class Layout:
    """costa::grid_layout in the shape of the C interface: split points, owners (row-major), this rank's blocks."""

    def __init__(self, rowsplit, colsplit, owners, blocks, ordering="C"):
        self.rowsplit = np.ascontiguousarray(rowsplit, dtype=np.int32)
        self.colsplit = np.ascontiguousarray(colsplit, dtype=np.int32)
        self.owners = np.ascontiguousarray(owners, dtype=np.int32).reshape(len(self.rowsplit) - 1, len(self.colsplit) - 1)
        self.blocks = [(int(r), int(c), int(p), int(ld)) for (r, c, p, ld) in blocks]  # (block row, block col, address, ld)
        self.ordering = ordering
        self._cblocks = (CBlock * max(len(self.blocks), 1))()
        for i, (r, c, p, ld) in enumerate(self.blocks):
            self._cblocks[i] = CBlock(p, ld, r, c)

    @property
    def shape(self):
        return int(self.rowsplit[-1]), int(self.colsplit[-1])

    def c_struct(self):
        return CLayout(len(self.rowsplit) - 1, len(self.colsplit) - 1, self.rowsplit.ctypes.data_as(pi), self.colsplit.ctypes.data_as(pi),
                       self.owners.ctypes.data_as(pi), len(self.blocks), self._cblocks)

    def block_shape(self, r, c):
        return int(self.rowsplit[r + 1] - self.rowsplit[r]), int(self.colsplit[c + 1] - self.colsplit[c])


def custom_layout(rowsplit, colsplit, owners, blocks, ordering="C"):
    return Layout(rowsplit, colsplit, owners, blocks, ordering)


def numroc(n, nb, iproc, isrcproc, nprocs):
    return lib().cosma_b200_numroc(n, nb, iproc, isrcproc, nprocs)


def scalapack_grid(lld, mat_rows, mat_cols, ia, ja, sub_m, sub_n, mb, nb, nprow, npcol, grid_order, rsrc, csrc, data_ordering, rank):
    """cosma_b200_scalapack_layout -> (rowsplit, colsplit, owners, [(block row, block col, element offset)])."""
    L = lib()
    nr, nc, nl = ci(), ci(), ci()
    args = [lld, mat_rows, mat_cols, ia, ja, sub_m, sub_n, mb, nb, nprow, npcol, grid_order.encode(), rsrc, csrc, data_ordering.encode(), rank]
    _lib.check(L.cosma_b200_scalapack_layout(*args, ctypes.byref(nr), ctypes.byref(nc), None, None, None, ctypes.byref(nl), None, None, None),
               "cosma_b200_scalapack_layout")
    rs = np.zeros(nr.value + 1, dtype=np.int32); cs = np.zeros(nc.value + 1, dtype=np.int32)
    ow = np.zeros(max(nr.value * nc.value, 1), dtype=np.int32)
    lr = np.zeros(max(nl.value, 1), dtype=np.int32); lc = np.zeros(max(nl.value, 1), dtype=np.int32)
    lo = np.zeros(max(nl.value, 1), dtype=np.int64)
    _lib.check(L.cosma_b200_scalapack_layout(*args, ctypes.byref(nr), ctypes.byref(nc), rs.ctypes.data_as(pi), cs.ctypes.data_as(pi),
                                             ow.ctypes.data_as(pi), ctypes.byref(nl), lr.ctypes.data_as(pi), lc.ctypes.data_as(pi),
                                             lo.ctypes.data_as(ctypes.POINTER(i64))), "cosma_b200_scalapack_layout")
    blocks = [(int(lr[i]), int(lc[i]), int(lo[i])) for i in range(nl.value)]
    return rs, cs, ow[:nr.value * nc.value].reshape(nr.value, nc.value), blocks


def block_cyclic_layout(m, n, block_m, block_n, i, j, sub_m, sub_n, p_m, p_n, order, rsrc, csrc, ptr, lld, ordering, rank, elem_bytes):
    """costa::block_cyclic_layout<T> (reference libs/COSTA/src/costa/layout.hpp:50-86); ptr = address of the local array."""
    rs, cs, ow, blocks = scalapack_grid(lld, m, n, i, j, sub_m, sub_n, block_m, block_n, p_m, p_n, order, rsrc, csrc, ordering, rank)
    return Layout(rs, cs, ow, [(r, c, ptr + off * elem_bytes, lld) for (r, c, off) in blocks], ordering)


def _scalars(values, dtype, n):
    out = (ctypes.c_double * (2 * n))()
    for i, v in enumerate(values):
        v = complex(v)
        out[2 * i], out[2 * i + 1] = v.real, v.imag
    return out


class TransformPlan:
    """costa::transformer<T>: a batch of (from, to, op, alpha, beta) executed as one exchange."""

    def __init__(self, comm, dtype, transforms, rank=None, nranks=None):
        """transforms: list of (from Layout, to Layout, op 'N'|'T'|'C', alpha, beta). comm: distributed.Comm or None
        (plan for (rank, nranks) without a communicator)."""
        self.lib = lib()
        self.dtype = dtype
        n = len(transforms)
        self._keep = transforms
        F = (CLayout * max(n, 1))(*[t[0].c_struct() for t in transforms])
        T = (CLayout * max(n, 1))(*[t[1].c_struct() for t in transforms])
        of = "".join(t[0].ordering for t in transforms).encode()
        ot = "".join(t[1].ordering for t in transforms).encode()
        ops = "".join(t[2] for t in transforms).encode()
        al = _scalars([t[3] for t in transforms], dtype, n)
        be = _scalars([t[4] for t in transforms], dtype, n)
        h = vp()
        ch = comm.handle if comm is not None else None
        r = comm.rank if comm is not None else (rank or 0)
        s = comm.size if comm is not None else (nranks or 1)
        _lib.check(self.lib.cosma_b200_transform_plan_create(ch, r, s, dtype.encode(), n, F, T, of, ot, ops, al, be, ctypes.byref(h)),
                   "cosma_b200_transform_plan_create")
        self.handle = h

    def export(self):
        n = i64()
        self.lib.cosma_b200_transform_plan_export(self.handle, None, 0, ctypes.byref(n))
        buf = (i64 * n.value)()
        self.lib.cosma_b200_transform_plan_export(self.handle, buf, n.value, ctypes.byref(n))
        return parse_transform_plan(list(buf))

    def run(self, stream=None):
        _lib.check(self.lib.cosma_b200_transform_run(self.handle, _stream_ptr(stream)), "cosma_b200_transform_run")

    def stats(self):
        a, b, l = i64(), i64(), ci()
        self.lib.cosma_b200_transform_plan_stats(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(l))
        return {"local_elements": a.value, "remote_elements": b.value, "launches": l.value}

    def destroy(self):
        if self.handle:
            self.lib.cosma_b200_transform_plan_destroy(self.handle)
            self.handle = None


def parse_transform_plan(flat):
    """Decodes cosma_b200_transform_plan_export (format: csrc/transform_exec.cu)."""
    P, eb, ts, tr, npack, nloc, nunp = flat[:7]
    pos = 7
    arrs = []
    for _ in range(4):
        arrs.append(flat[pos:pos + P]); pos += P
    pieces = []
    names = ("kind", "src", "dst", "src_ld", "dst_ld", "n_rows", "n_cols", "src_ordering", "dst_ordering", "transpose", "conjugate",
             "transform", "peer")
    for _ in range(npack + nloc + nunp):
        rec = dict(zip(names, flat[pos:pos + 13])); pos += 13
        rec["src_ordering"] = chr(rec["src_ordering"]); rec["dst_ordering"] = chr(rec["dst_ordering"])
        pieces.append(rec)
    return {"n_ranks": P, "elem_bytes": eb, "total_send": ts, "total_recv": tr, "send_off": arrs[0], "send_bytes": arrs[1],
            "recv_off": arrs[2], "recv_bytes": arrs[3], "pack": pieces[:npack], "local": pieces[npack:npack + nloc],
            "unpack": pieces[npack + nloc:]}


def relayout_batch(dtype, pieces, stream=None):
    """pieces: list of dicts with the fields of cosma_b200_piece (copy_and_transform argument meaning)."""
    L = lib()
    arr = (CPiece * max(len(pieces), 1))()
    for i, p in enumerate(pieces):
        a, b = complex(p.get("alpha", 1.0)), complex(p.get("beta", 0.0))
        arr[i] = CPiece(p["src"], p["dst"], p.get("src_ld", 0), p.get("dst_ld", 0), p["n_rows"], p["n_cols"], p.get("src_ordering", "C").encode(),
                        p.get("dst_ordering", "C").encode(), bytes([1 if p.get("transpose") else 0]), bytes([1 if p.get("conjugate") else 0]),
                        (ctypes.c_double * 2)(a.real, a.imag), (ctypes.c_double * 2)(b.real, b.imag))
    _lib.check(L.cosma_b200_relayout_batch(_stream_ptr(stream), dtype.encode(), len(pieces), arr), "cosma_b200_relayout_batch")


def multiply_using_layout(comm, dtype, transa, transb, alpha, A, B, beta, C, stream=None):
    """{d,z}multiply_using_layout (reference src/cosma/cinterface.hpp:42-76). A, B, C: Layout with device blocks."""
    L = lib()
    fn = getattr(L, "cosma_b200_%smultiply_using_layout" % dtype)
    al, be = _scalars([alpha], dtype, 1), _scalars([beta], dtype, 1)
    a, b, c = A.c_struct(), B.c_struct(), C.c_struct()
    _lib.check(fn(comm.handle, transa.encode(), transb.encode(), al, ctypes.byref(a), ctypes.byref(b), be, ctypes.byref(c), _stream_ptr(stream)),
               "cosma_b200_%smultiply_using_layout" % dtype)


class Grid:
    """The process grid a BLACS context would describe (Cblacs_gridinit + Cblacs_gridinfo)."""

    def __init__(self, comm, order, nprow, npcol):
        self.lib = lib()
        h = vp()
        _lib.check(self.lib.cosma_b200_grid_create(comm.handle, order.encode(), nprow, npcol, ctypes.byref(h)), "cosma_b200_grid_create")
        self.handle, self.order, self.nprow, self.npcol = h, order, nprow, npcol
        r, c = ci(), ci()
        self.lib.cosma_b200_grid_info(h, None, None, ctypes.byref(r), ctypes.byref(c))
        self.myrow, self.mycol = r.value, c.value

    def destroy(self):
        if self.handle:
            self.lib.cosma_b200_grid_destroy(self.handle)
            self.handle = None


def descinit(m, n, mb, nb, rsrc, csrc, lld, ctxt=0):
    """ScaLAPACK DESCINIT: the 9-int array descriptor (reference src/cosma/scalapack.hpp:11-47)."""
    return np.array([1, ctxt, m, n, mb, nb, rsrc, csrc, lld], dtype=np.int32)


def pxgemm(grid, dtype, transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc, stream=None):
    """p{d,z}gemm (reference src/cosma/pxgemm.h:6-107). a, b, c: addresses of the rank's local arrays (device or host)."""
    L = lib()
    fn = getattr(L, "cosma_b200_p%sgemm" % dtype)
    al, be = _scalars([alpha], dtype, 1), _scalars([beta], dtype, 1)
    da, db, dc = (np.ascontiguousarray(d, dtype=np.int32) for d in (desca, descb, descc))
    _lib.check(fn(grid.handle, transa.encode(), transb.encode(), m, n, k, al, vp(a), ia, ja, da.ctypes.data_as(pi), vp(b), ib, jb,
                  db.ctypes.data_as(pi), be, vp(c), ic, jc, dc.ctypes.data_as(pi), _stream_ptr(stream)), "cosma_b200_p%sgemm" % dtype)


def last_layout_multiply_stats(comm):
    """Phase timings (ms) and relayout volumes of the last multiply_using_layout / pxgemm on comm (synchronise first)."""
    L = lib()
    ms = (ctypes.c_float * 3)()
    el = (i64 * 4)()
    buf = ctypes.create_string_buffer(256)
    n = ci()
    _lib.check(L.cosma_b200_last_layout_multiply_stats(comm.handle, ms, el, buf, 256, ctypes.byref(n)), "cosma_b200_last_layout_multiply_stats")
    return {"ms_relayout_in": ms[0], "ms_multiply": ms[1], "ms_relayout_out": ms[2], "in_local_elements": el[0], "in_remote_elements": el[1],
            "out_local_elements": el[2], "out_remote_elements": el[3], "strategy": buf.value.decode(), "launches": n.value}


def pxtran(grid, dtype, op, m, n, alpha, a, ia, ja, desca, beta, c, ic, jc, descc, stream=None):
    """p?tran / p?tranu (op 'T') / p?tranc (op 'C') (reference libs/COSTA/src/costa/pxtran/pxtran.h:7-20,
    pxtran_op/costa_pxtran_op.cpp:14-172): sub(C) (m x n) = beta*sub(C) + alpha*op(sub(A)) (sub(A) is n x m).
    a, c: addresses of the rank's local arrays (device or host)."""
    L = lib()
    al, be = _scalars([alpha], dtype, 1), _scalars([beta], dtype, 1)
    da, dc = (np.ascontiguousarray(d, dtype=np.int32) for d in (desca, descc))
    _lib.check(L.cosma_b200_pxtran(grid.handle, dtype.encode(), op.encode(), m, n, al, vp(a), ia, ja, da.ctypes.data_as(pi), be, vp(c), ic, jc,
                                   dc.ctypes.data_as(pi), _stream_ptr(stream)), "cosma_b200_pxtran")


def pxgemr2d(grid_a, grid_c, dtype, m, n, a, ia, ja, desca, c, ic, jc, descc, stream=None):
    """p?gemr2d (reference libs/COSTA/src/costa/pxgemr2d/pxgemr2d.h:7-41, costa_pxgemr2d.cpp:14-168): sub(C) = sub(A)
    between two block-cyclic distributions (possibly on different process grids of one communicator)."""
    L = lib()
    da, dc = (np.ascontiguousarray(d, dtype=np.int32) for d in (desca, descc))
    _lib.check(L.cosma_b200_pxgemr2d(grid_a.handle, grid_c.handle, dtype.encode(), m, n, vp(a), ia, ja, da.ctypes.data_as(pi), vp(c), ic, jc,
                                     dc.ctypes.data_as(pi), _stream_ptr(stream)), "cosma_b200_pxgemr2d")
