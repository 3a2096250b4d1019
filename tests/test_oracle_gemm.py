"""The GEMM oracle (oracle/gemm_oracle.c, restating local_multiply_cpu, reference
src/cosma/local_multiply.cpp:277-297) pinned against the real reference (oracle/_ref: cosma::gemm ->
OpenBLAS cblas_?gemm, src/cosma/blas.cpp:24-130) and against exact integer arithmetic."""
import numpy as np
import pytest


def _ints(rng, n):
    # Tiled-MM's convention (libs/Tiled-MM/tests/test-multiply.cpp:60-68): integers 0..9 -> FP64 GEMM is exact
    return rng.integers(0, 10, size=n).astype(np.float64)


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (8, 4, 2), (33, 17, 9), (100, 100, 100), (257, 129, 65)])
def test_oracle_exact_integers_vs_int64(oracle, m, n, k):
    rng = np.random.default_rng(42)
    A, B, C = _ints(rng, m * k), _ints(rng, k * n), _ints(rng, m * n)
    want = (A.reshape(k, m).T.astype(np.int64) @ B.reshape(n, k).T.astype(np.int64)) + C.reshape(n, m).T.astype(np.int64)
    got = oracle.gemm("N", "N", m, n, k, 1.0, A, m, B, k, 1.0, C.copy(), m).reshape(n, m).T
    assert np.array_equal(got, want.astype(np.float64))


@pytest.mark.parametrize("m,n,k,alpha,beta", [(64, 48, 32, 1.0, 0.0), (100, 100, 100, 1.0, 1.0), (123, 77, 211, 2.5, -0.5)])
def test_oracle_vs_reference_dgemm(ref, m, n, k, alpha, beta):
    rng = np.random.default_rng(1)
    A, B, C = rng.random(m * k) * 10, rng.random(k * n) * 10, rng.random(m * n)
    got = ref.gemm("N", "N", m, n, k, alpha, A, m, B, k, beta, C.copy(), m)
    want = ref.ref_dgemm(m, n, k, alpha, A, m, B, k, beta, C.copy(), m)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 1e-14
    # the reference's own element-wise criterion (utils/cosma_utils.hpp:366-377): rel err < 1e-8
    assert np.all(np.abs(got - want) <= 1e-8 * np.maximum(np.abs(want), 1e-300))


def test_oracle_vs_reference_zgemm(ref):
    m, n, k = 37, 29, 53
    rng = np.random.default_rng(2)
    A = (rng.random(m * k) + 1j * rng.random(m * k)).astype(np.complex128)
    B = (rng.random(k * n) + 1j * rng.random(k * n)).astype(np.complex128)
    C = (rng.random(m * n) + 1j * rng.random(m * n)).astype(np.complex128)
    al, be = 0.7 - 0.2j, 0.3 + 0.4j
    got = ref.gemm("N", "N", m, n, k, al, A, m, B, k, be, C.copy(), m)
    want = ref.ref_zgemm(m, n, k, al, A, m, B, k, be, C.copy(), m)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-14


@pytest.mark.parametrize("ta,tb", [("N", "T"), ("T", "N"), ("T", "T"), ("C", "N"), ("N", "C"), ("C", "C")])
def test_oracle_transposes_against_numpy(oracle, ta, tb):
    m, n, k = 13, 11, 7
    rng = np.random.default_rng(3)
    shp_a = (k, m) if ta != "N" else (m, k)
    shp_b = (n, k) if tb != "N" else (k, n)
    A = rng.random(shp_a) + 1j * rng.random(shp_a)
    B = rng.random(shp_b) + 1j * rng.random(shp_b)
    op = lambda X, t: X if t == "N" else (X.T if t == "T" else X.conj().T)
    want = op(A, ta) @ op(B, tb)
    C = np.full(m * n, np.nan + 0j, dtype=np.complex128)   # beta == 0 must not read C
    got = oracle.gemm(ta, tb, m, n, k, 1.0, np.asfortranarray(A).ravel(order="F"), shp_a[0],
                      np.asfortranarray(B).ravel(order="F"), shp_b[0], 0.0, C, m).reshape(n, m).T
    assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_oracle_float32(oracle):
    m, n, k = 31, 19, 23
    rng = np.random.default_rng(4)
    A, B = rng.random(m * k).astype(np.float32), rng.random(k * n).astype(np.float32)
    C = np.zeros(m * n, dtype=np.float32)
    got = oracle.gemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m).reshape(n, m).T
    want = A.reshape(k, m).T.astype(np.float64) @ B.reshape(n, k).T.astype(np.float64)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-6
