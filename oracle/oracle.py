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
