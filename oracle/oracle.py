"""TEST INFRASTRUCTURE ONLY -- ctypes access to liboracle.so (our C restatement) and _ref/libcosma_ref.so
(the unmodified reference built by oracle/Makefile). numpy in, numpy out; column-major everywhere."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcosma_ref.so")
REF_MINIAPP = os.path.join(HERE, "_ref", "cosma_miniapp_ref")
_oracle = None
_ref = None
i64 = ctypes.c_int64


def build(ref=True):
    """make liboracle.so (+ _ref when /root/reference is present). Building the checker is not using it."""
    subprocess.check_call(["make", "-C", HERE, "-j8", "all" if ref else "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _oracle = ctypes.CDLL(ORACLE_SO)
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libcosma_ref.so not built (needs /root/reference; run make -C oracle ref)")
        _ref = ctypes.CDLL(REF_SO)
    return _ref


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def gemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    """Naive triple-loop oracle (restating local_multiply_cpu). A, B, C: 1-D numpy arrays (col-major storage);
    dtype float64/float32/complex128/complex64. C is updated in place and returned."""
    L = lib()
    dt = C.dtype
    if dt == np.float64:
        L.oracle_dgemm(ctypes.c_char(ta.encode()), ctypes.c_char(tb.encode()), i64(m), i64(n), i64(k), ctypes.c_double(alpha), _p(A), i64(lda), _p(B),
                       i64(ldb), ctypes.c_double(beta), _p(C), i64(ldc))
    elif dt == np.float32:
        L.oracle_sgemm(ctypes.c_char(ta.encode()), ctypes.c_char(tb.encode()), i64(m), i64(n), i64(k), ctypes.c_float(alpha), _p(A), i64(lda), _p(B),
                       i64(ldb), ctypes.c_float(beta), _p(C), i64(ldc))
    elif dt == np.complex128:
        al = np.array([complex(alpha)], dtype=np.complex128)
        be = np.array([complex(beta)], dtype=np.complex128)
        L.oracle_zgemm(ctypes.c_char(ta.encode()), ctypes.c_char(tb.encode()), i64(m), i64(n), i64(k), _p(al), _p(A), i64(lda), _p(B), i64(ldb), _p(be),
                       _p(C), i64(ldc))
    elif dt == np.complex64:
        al = np.array([complex(alpha)], dtype=np.complex64)
        be = np.array([complex(beta)], dtype=np.complex64)
        L.oracle_cgemm(ctypes.c_char(ta.encode()), ctypes.c_char(tb.encode()), i64(m), i64(n), i64(k), _p(al), _p(A), i64(lda), _p(B), i64(ldb), _p(be),
                       _p(C), i64(ldc))
    else:
        raise TypeError(dt)
    return C


def copy_and_transform(n_rows, n_cols, src, src_ld, src_ord, dst, dst_ld, dst_ord, transpose, conjugate, alpha, beta):
    L = lib()
    dt = dst.dtype
    args = [i64(n_rows), i64(n_cols), _p(src), i64(src_ld), ctypes.c_char(src_ord.encode()), _p(dst), i64(dst_ld),
            ctypes.c_char(dst_ord.encode()), ctypes.c_int(int(transpose)), ctypes.c_int(int(conjugate))]
    if dt == np.float64:
        L.oracle_copy_and_transform_d(*args, ctypes.c_double(alpha), ctypes.c_double(beta))
    elif dt == np.float32:
        L.oracle_copy_and_transform_s(*args, ctypes.c_float(alpha), ctypes.c_float(beta))
    elif dt == np.int32:
        L.oracle_copy_and_transform_i(*args, ctypes.c_int(int(alpha)), ctypes.c_int(int(beta)))
    elif dt == np.complex128:
        al = np.array([complex(alpha)], dtype=np.complex128); be = np.array([complex(beta)], dtype=np.complex128)
        L.oracle_copy_and_transform_z(*args, _p(al), _p(be))
    elif dt == np.complex64:
        al = np.array([complex(alpha)], dtype=np.complex64); be = np.array([complex(beta)], dtype=np.complex64)
        L.oracle_copy_and_transform_c(*args, _p(al), _p(be))
    else:
        raise TypeError(dt)
    return dst


# ---- the real reference (oracle/_ref) --------------------------------------------------------------

def ref_dgemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, threads=None):
    """Reference base-case GEMM: cosma::gemm -> cblas_dgemm (src/cosma/blas.cpp:24-49), always 'N','N'."""
    R = ref()
    if threads:
        R.ref_set_blas_threads(ctypes.c_int(threads))
    R.ref_dgemm(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_double(alpha), _p(A), ctypes.c_int(lda), _p(B),
                ctypes.c_int(ldb), ctypes.c_double(beta), _p(C), ctypes.c_int(ldc))
    return C


def ref_zgemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, threads=None):
    R = ref()
    if threads:
        R.ref_set_blas_threads(ctypes.c_int(threads))
    al = np.array([complex(alpha)], dtype=np.complex128); be = np.array([complex(beta)], dtype=np.complex128)
    R.ref_zgemm(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), _p(al), _p(A), ctypes.c_int(lda), _p(B),
                ctypes.c_int(ldb), _p(be), _p(C), ctypes.c_int(ldc))
    return C


def ref_strategy(m, n, k, P, mem_limit=0, prefix=""):
    """Reference Strategy(m,n,k,P[,prefix steps][,mem_limit]) -> (steps string, P actually used, memory_used)."""
    R = ref()
    out = ctypes.create_string_buffer(4096)
    P_out = ctypes.c_int(0)
    mem = ctypes.c_longlong(0)
    rc = R.ref_strategy(ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(P), ctypes.c_longlong(mem_limit),
                        prefix.encode(), out, ctypes.c_int(4096), ctypes.byref(P_out), ctypes.byref(mem))
    if rc < 0:
        raise RuntimeError("reference Strategy threw")
    return out.value.decode(), P_out.value, mem.value


def ref_mapper_layout(label, m, n, k, P, steps):
    """Reference Mapper(label, Strategy(m,n,k,P,steps)).complete_layout() -> list over ranks of lists of
    (row_first,row_last,col_first,col_last) inclusive intervals."""
    R = ref()
    counts = (ctypes.c_int * max(P, 1))()
    cap = 4 * 65536
    out = (ctypes.c_int * cap)()
    total = R.ref_mapper_layout(ctypes.c_char(label.encode()), ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k),
                                ctypes.c_int(P), steps.encode(), counts, out, ctypes.c_int(cap))
    if total < 0:
        raise RuntimeError("reference Mapper threw (%d)" % total)
    res, pos = [], 0
    for r in range(P):
        blocks = []
        for _ in range(counts[r]):
            blocks.append(tuple(out[4 * pos:4 * pos + 4]))
            pos += 1
        res.append(blocks)
    return res


# ---- COSTA relayout ------------------------------------------------------------------------------

_CAT = {"d": ("oracle_copy_and_transform_d", ctypes.c_double), "s": ("oracle_copy_and_transform_s", ctypes.c_float),
        "i": ("oracle_copy_and_transform_i", ctypes.c_int), "z": ("oracle_copy_and_transform_z", None),
        "c": ("oracle_copy_and_transform_c", None)}


def copy_and_transform_raw(dtype, n_rows, n_cols, src_addr, src_ld, src_ord, dst_addr, dst_ld, dst_ord, transpose, conjugate, alpha, beta):
    """oracle/relayout_oracle.c on raw addresses (so tests can interpret an exported transform plan piece by piece)."""
    L = lib()
    name, ctype = _CAT[dtype]
    args = [i64(n_rows), i64(n_cols), ctypes.c_void_p(src_addr), i64(src_ld), ctypes.c_char(src_ord.encode()), ctypes.c_void_p(dst_addr),
            i64(dst_ld), ctypes.c_char(dst_ord.encode()), ctypes.c_int(int(transpose)), ctypes.c_int(int(conjugate))]
    if ctype is not None:
        a, b = complex(alpha).real, complex(beta).real
        getattr(L, name)(*args, ctype(int(a)) if dtype == "i" else ctype(a), ctype(int(b)) if dtype == "i" else ctype(b))
    else:
        npdt = np.complex128 if dtype == "z" else np.complex64
        al = np.array([complex(alpha)], dtype=npdt); be = np.array([complex(beta)], dtype=npdt)
        getattr(L, name)(*args, _p(al), _p(be))


def ref_copy_and_transform(dtype, n_rows, n_cols, src, src_ld, src_ord, dst, dst_ld, dst_ord, transpose, conjugate, alpha, beta):
    """The reference's costa::memory::copy_and_transform (memory_utils.hpp:287-346) via oracle/_ref. dtype in i,s,d,c,z."""
    R = ref()
    al = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    be = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
    rc = R.ref_copy_and_transform(ctypes.c_char(dtype.encode()), ctypes.c_int(n_rows), ctypes.c_int(n_cols), _p(src), ctypes.c_int(src_ld),
                                  ctypes.c_int(1 if src_ord == "C" else 0), _p(dst), ctypes.c_int(dst_ld), ctypes.c_int(1 if dst_ord == "C" else 0),
                                  ctypes.c_int(int(transpose)), ctypes.c_int(int(conjugate)), al, be)
    if rc != 0:
        raise RuntimeError("reference copy_and_transform failed (%d)" % rc)
    return dst


def ref_scalapack_layout(lld, mat_rows, mat_cols, ia, ja, sub_m, sub_n, mb, nb, nprow, npcol, grid_order, rsrc, csrc, data_ordering, rank):
    """The reference's costa::get_scalapack_layout (scalapack_layout.cpp:178-285) ->
    (rowsplit, colsplit, owners[rows][cols], [(block row, block col, element offset)])."""
    R = ref()
    ci = ctypes.c_int
    nr, nc, nl = ci(), ci(), ci()
    args = [ci(lld), ci(mat_rows), ci(mat_cols), ci(ia), ci(ja), ci(sub_m), ci(sub_n), ci(mb), ci(nb), ci(nprow), ci(npcol),
            ctypes.c_char(grid_order.encode()), ci(rsrc), ci(csrc), ctypes.c_char(data_ordering.encode()), ci(rank)]
    rc = R.ref_scalapack_layout(*args, ctypes.byref(nr), ctypes.byref(nc), None, None, None, ctypes.byref(nl), None, None, None)
    if rc != 0:
        raise RuntimeError("reference get_scalapack_layout threw")
    rs = np.zeros(nr.value + 1, dtype=np.int32); cs = np.zeros(nc.value + 1, dtype=np.int32)
    ow = np.zeros(max(nr.value * nc.value, 1), dtype=np.int32)
    lr = np.zeros(max(nl.value, 1), dtype=np.int32); lc = np.zeros(max(nl.value, 1), dtype=np.int32)
    lo = np.zeros(max(nl.value, 1), dtype=np.int64)
    R.ref_scalapack_layout(*args, ctypes.byref(nr), ctypes.byref(nc), _p(rs), _p(cs), _p(ow), ctypes.byref(nl), _p(lr), _p(lc), _p(lo))
    return rs, cs, ow[:nr.value * nc.value].reshape(nr.value, nc.value), [(int(lr[i]), int(lc[i]), int(lo[i])) for i in range(nl.value)]


class _RefBlock(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("ld", ctypes.c_int), ("row", ctypes.c_int), ("col", ctypes.c_int)]


class _RefLayout(ctypes.Structure):
    _fields_ = [("rowblocks", ctypes.c_int), ("colblocks", ctypes.c_int), ("rowsplit", ctypes.c_void_p), ("colsplit", ctypes.c_void_p),
                ("owners", ctypes.c_void_p), ("nlocalblocks", ctypes.c_int), ("localblocks", ctypes.POINTER(_RefBlock))]


def ref_transform_p1(dtype, from_layout, to_layout, trans, alpha, beta):
    """The reference's costa::transform on one rank. Layout arguments: (rowsplit, colsplit, owners, [(row, col, addr, ld)], ordering)."""
    R = ref()
    keep = []

    def mk(l):
        rs, cs, ow, blocks, _ = l
        rs = np.ascontiguousarray(rs, dtype=np.int32); cs = np.ascontiguousarray(cs, dtype=np.int32)
        ow = np.ascontiguousarray(ow, dtype=np.int32)
        arr = (_RefBlock * max(len(blocks), 1))()
        for i, (r, c, addr, ld) in enumerate(blocks):
            arr[i] = _RefBlock(addr, ld, r, c)
        keep.extend([rs, cs, ow, arr])
        return _RefLayout(len(rs) - 1, len(cs) - 1, rs.ctypes.data, cs.ctypes.data, ow.ctypes.data, len(blocks), arr)

    F, T = mk(from_layout), mk(to_layout)
    al = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    be = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
    rc = R.ref_transform_p1(ctypes.c_char(dtype.encode()), ctypes.byref(F), ctypes.c_char(from_layout[4].encode()), ctypes.byref(T),
                            ctypes.c_char(to_layout[4].encode()), ctypes.c_char(trans.encode()), al, be)
    if rc != 0:
        raise RuntimeError("reference costa::transform failed (%d)" % rc)


# ---- the unmodified reference on SEVERAL ranks (oracle/minirun.py + stubs/minimpi.cpp + oracle/ref_driver.cpp) --------
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")
NPDT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def have_ref_driver():
    return os.path.exists(REF_DRIVER) and have_ref()


def _scalar_arg(v):
    v = complex(v)
    return "%r,%r" % (v.real, v.imag)


def _run_driver(nranks, args, threads=1, timeout=600):
    import sys
    sys.path.insert(0, HERE)
    from minirun import launch
    argv = [REF_DRIVER] + ["%s=%s" % kv for kv in args.items()]
    code, outs = launch(nranks, argv, threads=threads, stdout=subprocess.PIPE, timeout=timeout)
    if code != 0:
        raise RuntimeError("reference driver failed with exit code %s: %s" % (code, " ".join(argv)))
    text = outs[0].decode()
    return [float(line.split()[1]) for line in text.splitlines() if line.startswith("REF_TIME_MS")]


def ref_multiply_ranks(dtype, m, n, k, P, steps, alpha, beta, A, B, C, threads=1, reps=1, tmpdir=None):
    """cosma::multiply of the unmodified reference on P minimpi ranks. A (m x k), B (k x n), C (m x n): dense numpy
    arrays. Returns (list of per-rank raw local C buffers -- ranks idle under the strategy give None --, times in ms)."""
    import tempfile
    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        dt = NPDT[dtype]
        for name, M in (("A", A), ("B", B), ("C", C)):
            np.asfortranarray(M.astype(dt)).T.tofile(os.path.join(d, name + ".bin"))  # column-major bytes
        times = _run_driver(P, dict(op="multiply", dtype=dtype, m=m, n=n, k=k, steps=steps or "auto", alpha=_scalar_arg(alpha),
                                    beta=_scalar_arg(beta), dir=d, reps=reps), threads=threads)
        outs = []
        for r in range(P):
            f = os.path.join(d, "Cout.%d.bin" % r)
            outs.append(np.fromfile(f, dtype=dt) if os.path.exists(f) else None)
    return outs, times


def _desc_arg(desc):
    # 9-int ScaLAPACK descriptor -> "M,N,MB,NB,RSRC,CSRC,LLD"
    return ",".join(str(int(desc[i])) for i in (2, 3, 4, 5, 6, 7, 8))


def _llds(descs):
    return ",".join(str(int(x[8])) for x in descs)


def ref_pxgemm_ranks(dtype, order, nprow, npcol, ta, tb, m, n, k, alpha, a_loc, ia, ja, desca, b_loc, ib, jb, descb, beta, c_loc, ic, jc,
                     descc, threads=1, reps=1):
    """cosma::pxgemm<T> of the unmodified reference on nprow*npcol minimpi ranks (miniblacs grid). a_loc/b_loc/c_loc:
    per-rank local arrays (1-D numpy, LLD x local columns); desc*: per-rank descriptors (LLD may differ). Returns
    (per-rank local C arrays after the call, times in ms)."""
    import tempfile
    P = nprow * npcol
    dt = NPDT[dtype]
    with tempfile.TemporaryDirectory() as d:
        for r in range(P):
            for name, loc in (("a", a_loc), ("b", b_loc), ("c", c_loc)):
                np.ascontiguousarray(loc[r], dtype=dt).tofile(os.path.join(d, "%s.%d.bin" % (name, r)))
        times = _run_driver(P, dict(op="pxgemm", dtype=dtype, ta=ta, tb=tb, m=m, n=n, k=k, alpha=_scalar_arg(alpha), beta=_scalar_arg(beta),
                                    ia=ia, ja=ja, ib=ib, jb=jb, ic=ic, jc=jc, order=order, nprow=nprow, npcol=npcol,
                                    desca=_desc_arg(desca[0]), descb=_desc_arg(descb[0]), descc=_desc_arg(descc[0]), llda=_llds(desca),
                                    lldb=_llds(descb), lldc=_llds(descc), dir=d, reps=reps),
                            threads=threads)
        outs = [np.fromfile(os.path.join(d, "cout.%d.bin" % r), dtype=dt) for r in range(P)]
    return outs, times


def ref_pxgemr2d_ranks(dtype, order, nprow, npcol, m, n, a_loc, ia, ja, desca, c_loc, ic, jc, descc, orderc=None):
    """costa::pxgemr2d<T> of the unmodified reference. Returns the per-rank local C arrays."""
    import tempfile
    P = nprow * npcol
    dt = NPDT[dtype]
    with tempfile.TemporaryDirectory() as d:
        for r in range(P):
            for name, loc in (("a", a_loc), ("c", c_loc)):
                np.ascontiguousarray(loc[r], dtype=dt).tofile(os.path.join(d, "%s.%d.bin" % (name, r)))
        _run_driver(P, dict(op="pxgemr2d", dtype=dtype, m=m, n=n, ia=ia, ja=ja, ic=ic, jc=jc, order=order, orderc=orderc or order, nprow=nprow,
                            npcol=npcol, desca=_desc_arg(desca[0]), descc=_desc_arg(descc[0]), llda=_llds(desca), lldc=_llds(descc), dir=d))
        return [np.fromfile(os.path.join(d, "cout.%d.bin" % r), dtype=dt) for r in range(P)]


def ref_pxtran_ranks(dtype, order, nprow, npcol, trans, m, n, alpha, a_loc, ia, ja, desca, beta, c_loc, ic, jc, descc):
    """costa::pxtran_op<T> of the unmodified reference: sub(C) (m x n) = beta*sub(C) + alpha*op(sub(A)) (n x m)."""
    import tempfile
    P = nprow * npcol
    dt = NPDT[dtype]
    with tempfile.TemporaryDirectory() as d:
        for r in range(P):
            for name, loc in (("a", a_loc), ("c", c_loc)):
                np.ascontiguousarray(loc[r], dtype=dt).tofile(os.path.join(d, "%s.%d.bin" % (name, r)))
        _run_driver(P, dict(op="pxtran", dtype=dtype, trans=trans, m=m, n=n, alpha=_scalar_arg(alpha), beta=_scalar_arg(beta), ia=ia, ja=ja, ic=ic,
                            jc=jc, order=order, nprow=nprow, npcol=npcol, desca=_desc_arg(desca[0]), descc=_desc_arg(descc[0]), llda=_llds(desca),
                            lldc=_llds(descc), dir=d))
        return [np.fromfile(os.path.join(d, "cout.%d.bin" % r), dtype=dt) for r in range(P)]
