"""Local GEMM through the C ABI (cosma_b200_dgemm / cosma_b200_zgemm), column-major, device pointers.

Mirrors the call the reference makes at the base case: local_multiply(ctx, A, B, C, m, n, k, alpha, beta)
(reference src/cosma/local_multiply.hpp:7-16) -> gemm('N','N', m, n, k, alpha, A, lda=m, B, ldb=k, beta, C, ldc=m)."""
import ctypes

from . import _lib


def _stream_ptr(stream):
    if stream is None:
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if isinstance(stream, int):
        return ctypes.c_void_p(stream)
    return ctypes.c_void_p(stream.cuda_stream)


def gemm_raw(dtype, transa, transb, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, stream=None):
    """dtype: 'd' | 'z' (FP64 DMMA kernels) or 's' | 'c' (3xTF32 tcgen05 kernels). a_ptr/b_ptr/c_ptr: device addresses
    (int). alpha/beta: python float or complex."""
    lib = _lib.load()
    ctype = ctypes.c_double if dtype in "dz" else ctypes.c_float
    if dtype in "ds":
        al = (ctype * 1)(float(alpha))
        be = (ctype * 1)(float(beta))
    elif dtype in "zc":
        al = (ctype * 2)(complex(alpha).real, complex(alpha).imag)
        be = (ctype * 2)(complex(beta).real, complex(beta).imag)
    else:
        raise ValueError(dtype)
    fn = getattr(lib, "cosma_b200_%sgemm" % dtype)
    if dtype in "sc" and not getattr(fn, "_declared", False):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_char, ctypes.c_char, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_float),
                       ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int64]
        fn.restype = ctypes.c_int
        fn._declared = True
    st = fn(_stream_ptr(stream), transa.encode(), transb.encode(), m, n, k, al, ctypes.c_void_p(a_ptr), lda,
            ctypes.c_void_p(b_ptr), ldb, be, ctypes.c_void_p(c_ptr), ldc)
    _lib.check(st, "cosma_b200_%sgemm" % dtype)


def local_multiply(A, B, C, m, n, k, alpha, beta, stream=None):
    """A, B, C: 1-D torch CUDA tensors holding column-major m x k, k x n, m x n (lda=m, ldb=k, ldc=m)."""
    import torch
    dt = {torch.float64: "d", torch.complex128: "z", torch.float32: "s", torch.complex64: "c"}[C.dtype]
    gemm_raw(dt, "N", "N", m, n, k, alpha, A.data_ptr(), max(m, 1), B.data_ptr(), max(k, 1), beta, C.data_ptr(),
             max(m, 1), stream)


def gemm_host(dtype, m, n, k, alpha, hA, lda, hB, ldb, beta, hC, ldc, stream=None):
    """cosma_b200_{d,z}gemm_host: 'N','N' GEMM on HOST (pinned) torch tensors, PCIe pipelined under the kernel -- the
    calling convention of the reference's GPU base case (gpu::gemm with host pointers, local_multiply.cpp:219-269)."""
    lib = _lib.load()
    fn = getattr(lib, "cosma_b200_%sgemm_host" % dtype)
    fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 2 + [ctypes.c_int64] + \
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    fn.restype = ctypes.c_int
    if dtype == "d":
        al = (ctypes.c_double * 1)(float(alpha)); be = (ctypes.c_double * 1)(float(beta))
    else:
        al = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        be = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
    st = fn(_stream_ptr(stream), m, n, k, al, hA.data_ptr(), lda, hB.data_ptr(), ldb, be, hC.data_ptr(), ldc)
    _lib.check(st, "cosma_b200_%sgemm_host" % dtype)
    return lib.cosma_b200_last_launch_count()
