"""ctypes binding of libcosma_b200.so (the C ABI in include/cosma_b200.h).

There is no CPU fallback: if the library is missing this raises. `load(build=True)` compiles it first
(nvcc cross-compiles without a GPU)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcosma_b200.so")
_lib = None

c_i64 = ctypes.c_int64
c_dp = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p


class CosmaB200Error(RuntimeError):
    pass


STATUS = {0: "OK", 1: "INVALID_ARG", 2: "CUDA_ERROR", 3: "NCCL_ERROR", 4: "OUT_OF_MEMORY", 5: "NOT_SUPPORTED",
          6: "INTERNAL_ERROR"}


def _declare(lib):
    lib.cosma_b200_version.restype = ctypes.c_char_p
    lib.cosma_b200_last_error.restype = ctypes.c_char_p
    gemm_args = [c_vp, ctypes.c_char, ctypes.c_char, c_i64, c_i64, c_i64, c_dp, c_vp, c_i64, c_vp, c_i64, c_dp, c_vp,
                 c_i64]
    for name in ("cosma_b200_dgemm", "cosma_b200_zgemm"):
        if hasattr(lib, name):
            getattr(lib, name).argtypes = gemm_args
            getattr(lib, name).restype = ctypes.c_int
    lib.cosma_b200_last_gemm_path.restype = ctypes.c_int


def load(build=False):
    global _lib
    if _lib is not None:
        return _lib
    if build or os.environ.get("COSMA_B200_BUILD") == "1":
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise CosmaB200Error(
            "libcosma_b200.so not found at %s -- run `python -m cosma_b200.build` (there is no CPU fallback)" % LIB_PATH)
    _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    _declare(_lib)
    return _lib


def check(status, what=""):
    if status != 0:
        lib = load()
        msg = lib.cosma_b200_last_error().decode()
        raise CosmaB200Error("%s failed: %s %s" % (what, STATUS.get(status, status), msg))
