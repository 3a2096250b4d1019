"""GPU smoke/timing of the DGEMM kernel against torch.matmul(float64) (cuBLAS). Development tool."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cosma_b200 import gemm, _lib

dev = torch.device("cuda:0")
lib = _lib.load()

def colmajor(m, n, ld=None, seed=0, ints=False):
    ld = ld or m
    g = torch.Generator(device=dev); g.manual_seed(seed)
    buf = torch.zeros(n * ld, device=dev, dtype=torch.float64)
    v = buf.view(n, ld)
    if ints:
        v[:, :m] = torch.randint(0, 10, (n, m), device=dev, generator=g).double()
    else:
        v[:, :m] = torch.rand(n, m, device=dev, generator=g, dtype=torch.float64) * 10
    return buf

def as_mat(buf, m, n, ld):  # logical m x n
    return buf.view(n, ld)[:, :m].t()

def check(ta, tb, m, n, k, alpha=1.0, beta=0.0, pad=0, ints=False):
    am, ak = (k, m) if ta != 'N' else (m, k)
    bk, bn = (n, k) if tb != 'N' else (k, n)
    lda, ldb, ldc = max(1, am + pad), max(1, bk + pad), max(1, m + pad)
    A = colmajor(am, ak, lda, 1, ints); B = colmajor(bk, bn, ldb, 2, ints); C = colmajor(m, n, ldc, 3, ints)
    if beta == 0.0:
        C.fill_(float('nan'))
    C0 = C.clone()
    opA = as_mat(A, am, ak, lda); opA = opA.t() if ta != 'N' else opA
    opB = as_mat(B, bk, bn, ldb); opB = opB.t() if tb != 'N' else opB
    ref = alpha * (opA @ opB)
    if beta != 0.0:
        ref = ref + beta * as_mat(C0, m, n, ldc)
    gemm.gemm_raw('d', ta, tb, m, n, k, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta, C.data_ptr(), ldc)
    torch.cuda.synchronize()
    out = as_mat(C, m, n, ldc)
    err = ((out - ref).norm() / ref.norm()).item() if ref.numel() else 0.0
    exact = bool((out == ref).all().item()) if ints else None
    path = lib.cosma_b200_last_gemm_path()
    ok = err < 1e-13 and not torch.isnan(out).any().item()
    print(json.dumps({"ta": ta, "tb": tb, "m": m, "n": n, "k": k, "alpha": alpha, "beta": beta, "pad": pad, "path": path,
                      "rel_err": err, "exact": exact, "ok": ok}), flush=True)
    return ok

def bench(n, reps=5):
    A = colmajor(n, n, seed=1); B = colmajor(n, n, seed=2); C = torch.empty(n * n, device=dev, dtype=torch.float64)
    f = lambda: gemm.gemm_raw('d', 'N', 'N', n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    a2 = A.view(n, n); b2 = B.view(n, n); c2 = torch.empty(n, n, device=dev, dtype=torch.float64)
    torch.matmul(b2, a2, out=c2); torch.cuda.synchronize()   # row-major view: C^T = B^T A^T
    cb = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(b2, a2, out=c2); e1.record(); torch.cuda.synchronize()
        cb = min(cb, e0.elapsed_time(e1))
    err = ((C.view(n, n) - c2).norm() / c2.norm()).item()
    print(json.dumps({"bench_n": n, "ms": best, "tflops": 2 * n**3 / best * 1e-9, "cublas_ms": cb,
                      "cublas_tflops": 2 * n**3 / cb * 1e-9, "rel_err_vs_cublas": err}), flush=True)

if __name__ == "__main__":
    allok = True
    for (ta, tb) in [('N', 'N'), ('N', 'T'), ('T', 'N'), ('T', 'T')]:
        allok &= check(ta, tb, 256, 256, 64, ints=True)
        allok &= check(ta, tb, 300, 200, 100, alpha=1.0, beta=1.0)
        allok &= check(ta, tb, 130, 70, 18, alpha=2.5, beta=-0.5, pad=2)
        allok &= check(ta, tb, 1000, 1000, 1000)
    allok &= check('N', 'N', 2000, 2000, 1000, ints=True)
    allok &= check('N', 'N', 127, 129, 33, pad=1)      # generic (odd lda) path
    allok &= check('T', 'T', 65, 63, 17, pad=0, alpha=1.0, beta=1.0)
    allok &= check('N', 'N', 5, 3, 0, beta=2.0)
    allok &= check('N', 'N', 8, 1, 4)
    print(json.dumps({"all_ok": bool(allok)}), flush=True)
    if "--bench" in sys.argv:
        for n in (4096, 8192, 16384):
            bench(n)
