"""GPU parity of the single-precision local GEMM (K3 SGEMM, K4 CGEMM: 3xTF32 on tcgen05/TMEM) through the C ABI.

Tolerance (north_star): normwise ||C - C_ref||_F / ||C_ref||_F <= 1e-6 against the FP32 reference result; here C_ref is
formed in FP64 from the same FP32 inputs (stricter than comparing with an FP32 BLAS, whose own error is ~1e-7).
Integer-valued inputs (Tiled-MM convention, libs/Tiled-MM/tests/test-multiply.cpp:60-68) are exact in TF32, so those
cases must be BIT-exact against the oracle's naive loop (reference src/cosma/local_multiply.cpp:277-297). The
reference's own element-wise criterion for float (utils/cosma_utils.hpp:366-377, 1e-5) is checked too."""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-6
# COSMA_B200_TF32_KERNEL = v1 | v2 (csrc/gemm_tf32x3_sm100.cu): v1 has no tensor path for CGEMM with a transposed / conjugated operand
V1_ONLY = os.environ.get("COSMA_B200_TF32_KERNEL", "").lower().startswith("v1")


def _run(oracle, dtype, ta, tb, m, n, k, alpha, beta, pad=0, ints=False, seed=0, expect_path=None):
    from cosma_b200 import gemm, _lib
    rng = np.random.default_rng(seed)
    cplx = dtype == "c"
    npdt = np.complex64 if cplx else np.float32
    ar, ac = (k, m) if ta != "N" else (m, k)
    br, bc = (n, k) if tb != "N" else (k, n)
    lda, ldb, ldc = max(1, ar + pad), max(1, br + pad), max(1, m + pad)
    if expect_path is None:  # TMA needs 16-byte row pitches: ld % 4 == 0 (float), ld % 2 == 0 (complex float)
        q = 2 if cplx else 4
        expect_path = 1 if (lda % q == 0 and ldb % q == 0 and (not cplx or not V1_ONLY or (ta == "N" and tb == "N"))) else 2

    def fill(cols, ld):
        cnt = max(1, ld * cols)
        if ints:
            v = rng.integers(0, 10, size=cnt).astype(np.float64)
            if cplx:
                v = v + 1j * rng.integers(0, 10, size=cnt)
        else:
            v = rng.random(cnt) * 10  # U[0,10) like the reference miniapp (miniapp/cosma_miniapp.cpp:21-25)
            if cplx:
                v = v + 1j * rng.random(cnt) * 10
        return v.astype(npdt)

    A, B, C = fill(ac, lda), fill(bc, ldb), fill(n, ldc)
    if beta == 0:
        C[:] = np.nan  # beta == 0 must not read C
    dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, C))
    gemm.gemm_raw(dtype, ta, tb, m, n, k, alpha, dA.data_ptr(), lda, dB.data_ptr(), ldb, beta, dC.data_ptr(), ldc)
    torch.cuda.synchronize()
    if m and n and k and alpha != 0:
        assert _lib.load().cosma_b200_last_gemm_path() == expect_path
    got = dC.cpu().numpy()

    def view(X, rows, cols, ld):
        return X[:ld * cols].reshape(cols, ld).T[:rows, :]

    def op(X, t):
        return X if t == "N" else (X.T if t == "T" else X.conj().T)
    Am, Bm = op(view(A, ar, ac, lda), ta).astype(np.complex128 if cplx else np.float64), op(view(B, br, bc, ldb), tb).astype(np.complex128 if cplx else np.float64)
    Cm = view(C, m, n, ldc).astype(Am.dtype)
    want = alpha * (Am @ Bm) + (beta * Cm if beta != 0 else 0)
    G = view(got, m, n, ldc)
    if pad and m and n:  # padding rows untouched
        assert np.array_equal(got[:ldc * n].reshape(n, ldc)[:, m:].view(np.uint8), C[:ldc * n].reshape(n, ldc)[:, m:].view(np.uint8))
    if ints:
        assert np.array_equal(G, want.astype(npdt))
        return 0.0
    if G.size == 0:
        return 0.0
    assert np.isfinite(G).all()
    err = np.linalg.norm(G - want) / max(np.linalg.norm(want), 1e-30)
    assert err <= TOL, err
    rel = np.abs(G - want) / np.maximum(np.abs(want), 1e-30)
    assert rel.max() < 1e-5  # the reference's element-wise float criterion
    return err


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (128, 128, 64), (256, 128, 96), (300, 260, 200), (1000, 1000, 1000), (64, 64, 4096),
                                   (2048, 2048, 2048), (129, 127, 33), (4, 4, 4)])
def test_sgemm_nn(oracle, m, n, k):
    _run(oracle, "s", "N", "N", m, n, k, 1.0, 0.0, ints=True)
    _run(oracle, "s", "N", "N", m, n, k, 1.0, 0.0)
    _run(oracle, "s", "N", "N", m, n, k, 0.5, 2.0, seed=1)


@pytest.mark.parametrize("ta,tb", [("N", "T"), ("T", "N"), ("T", "T")])
def test_sgemm_transposes(oracle, ta, tb):
    for (m, n, k) in ((256, 384, 128), (300, 260, 200)):
        _run(oracle, "s", ta, tb, m, n, k, 1.0, 1.0, ints=True)
        _run(oracle, "s", ta, tb, m, n, k, -1.5, 0.0)


@pytest.mark.parametrize("m,n,k", [(64, 128, 16), (128, 128, 64), (300, 260, 200), (1000, 1000, 500), (1024, 1024, 1024), (65, 63, 17)])
def test_cgemm_nn(oracle, m, n, k):
    _run(oracle, "c", "N", "N", m, n, k, 1.0, 0.0, ints=True)
    _run(oracle, "c", "N", "N", m, n, k, 1.0 - 0.5j, 0.0)
    _run(oracle, "c", "N", "N", m, n, k, 0.5 + 1j, 2.0 - 1j, seed=2)


@pytest.mark.parametrize("ta,tb", [("N", "T"), ("N", "C"), ("T", "N"), ("C", "N"), ("T", "T"), ("C", "C"), ("T", "C"), ("C", "T")])
def test_cgemm_transposes(oracle, ta, tb):
    """op(A), op(B) in {T, C}: the split stages of the second-generation kernel build the real embedding of op(A) and the real view of
    op(B) from either storage order (conjugation = a sign flip); the first generation sends these to the generic kernel."""
    for (m, n, k) in ((128, 128, 64), (300, 260, 200), (65, 63, 17)):
        _run(oracle, "c", ta, tb, m, n, k, 1.0, 0.0, ints=True)
        _run(oracle, "c", ta, tb, m, n, k, 0.5 - 1j, 1.0 + 0.5j, seed=3)


def test_generic_paths(oracle):
    # unaligned leading dimensions -> generic kernel
    _run(oracle, "s", "N", "N", 100, 90, 80, 1.0, 1.0, pad=1, expect_path=2)
    _run(oracle, "c", "C", "N", 64, 64, 64, 1.0, 0.0, pad=1, expect_path=2)
    _run(oracle, "c", "N", "T", 60, 50, 40, 2.0, 1.0, pad=1, expect_path=2)


def test_padded_ld_and_degenerate(oracle):
    _run(oracle, "s", "N", "N", 256, 256, 256, 1.0, 1.0, pad=4)       # ld multiple of 4 keeps the TMA path
    _run(oracle, "c", "N", "N", 128, 128, 128, 1.0, 1.0, pad=2)
    _run(oracle, "s", "N", "N", 100, 100, 0, 1.0, 2.0, ints=True)     # k = 0: C *= beta
    _run(oracle, "s", "N", "N", 0, 10, 10, 1.0, 0.0)
    _run(oracle, "s", "N", "N", 64, 64, 64, 0.0, 0.0, ints=True)      # alpha = 0, beta = 0: zeros, NaN not propagated


@pytest.mark.skipif("COSMA_B200_TF32_KERNEL" in os.environ, reason="already a run with an explicit kernel generation")
@pytest.mark.parametrize("generation", ["v1", "v2"])
def test_both_kernel_generations(generation):
    """The whole module again with the kernel generation forced (whichever is the default, the other one stays covered)."""
    env = dict(os.environ, COSMA_B200_TF32_KERNEL=generation)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"], capture_output=True, text=True,
                         timeout=900, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
