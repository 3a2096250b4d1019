/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's base-case multiply.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * (cosma_b200/) never does.
 *
 * Restates local_multiply_cpu (reference src/cosma/local_multiply.cpp:277-297), the naive triple loop
 * the reference's own tests use as THEIR oracle (utils/cosma_utils.hpp:226-283):
 *     for mi, ni:  C(mi,ni) *= beta;  for ki: C(mi,ni) += alpha * A(mi,ki) * B(ki,ni)
 * column-major, lda = m, ldb = k, ldc = m in the reference; leading dimensions are explicit here.
 * Extension (ours): op(A)/op(B) flags and beta == 0 not reading C, to check the BLAS-style entry points.
 * Pinned by: tests/test_oracle.py against oracle/_ref (reference blas.cpp -> OpenBLAS) and integer-exact
 * inputs (Tiled-MM's convention, libs/Tiled-MM/tests/test-multiply.cpp:60-68).
 * OpenMP over columns only (each C element is still accumulated sequentially in k order). */
#include <complex.h>
#include <stdint.h>

static inline int64_t idx(int64_t i, int64_t j, int64_t ld) { return i + j * ld; }

void oracle_dgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    const int tA = !(ta == 'N' || ta == 'n'), tB = !(tb == 'N' || tb == 'n');
#pragma omp parallel for schedule(static)
    for (int64_t ni = 0; ni < n; ++ni) {
        for (int64_t mi = 0; mi < m; ++mi) {
            double c = (beta == 0.0) ? 0.0 : C[idx(mi, ni, ldc)] * beta;
            for (int64_t ki = 0; ki < k; ++ki) {
                const double a = tA ? A[idx(ki, mi, lda)] : A[idx(mi, ki, lda)];
                const double b = tB ? B[idx(ni, ki, ldb)] : B[idx(ki, ni, ldb)];
                c += alpha * a * b;
            }
            C[idx(mi, ni, ldc)] = c;
        }
    }
}

void oracle_sgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc) {
    const int tA = !(ta == 'N' || ta == 'n'), tB = !(tb == 'N' || tb == 'n');
#pragma omp parallel for schedule(static)
    for (int64_t ni = 0; ni < n; ++ni) {
        for (int64_t mi = 0; mi < m; ++mi) {
            float c = (beta == 0.0f) ? 0.0f : C[idx(mi, ni, ldc)] * beta;
            for (int64_t ki = 0; ki < k; ++ki) {
                const float a = tA ? A[idx(ki, mi, lda)] : A[idx(mi, ki, lda)];
                const float b = tB ? B[idx(ni, ki, ldb)] : B[idx(ki, ni, ldb)];
                c += alpha * a * b;
            }
            C[idx(mi, ni, ldc)] = c;
        }
    }
}

static inline double complex opz(char t, double complex v) { return (t == 'C' || t == 'c') ? conj(v) : v; }

/* complex128: interleaved (re, im) doubles; alpha/beta point at 2 doubles */
void oracle_zgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, const double* alpha_, const double* A_, int64_t lda,
                  const double* B_, int64_t ldb, const double* beta_, double* C_, int64_t ldc) {
    const double complex alpha = alpha_[0] + alpha_[1] * I, beta = beta_[0] + beta_[1] * I;
    const double complex* A = (const double complex*)A_;
    const double complex* B = (const double complex*)B_;
    double complex* C = (double complex*)C_;
    const int tA = !(ta == 'N' || ta == 'n'), tB = !(tb == 'N' || tb == 'n');
#pragma omp parallel for schedule(static)
    for (int64_t ni = 0; ni < n; ++ni) {
        for (int64_t mi = 0; mi < m; ++mi) {
            double complex c = (beta == 0.0) ? 0.0 : C[idx(mi, ni, ldc)] * beta;
            for (int64_t ki = 0; ki < k; ++ki) {
                const double complex a = opz(ta, tA ? A[idx(ki, mi, lda)] : A[idx(mi, ki, lda)]);
                const double complex b = opz(tb, tB ? B[idx(ni, ki, ldb)] : B[idx(ki, ni, ldb)]);
                c += alpha * a * b;
            }
            C[idx(mi, ni, ldc)] = c;
        }
    }
}

static inline float complex opc(char t, float complex v) { return (t == 'C' || t == 'c') ? conjf(v) : v; }

void oracle_cgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, const float* alpha_, const float* A_, int64_t lda,
                  const float* B_, int64_t ldb, const float* beta_, float* C_, int64_t ldc) {
    const float complex alpha = alpha_[0] + alpha_[1] * I, beta = beta_[0] + beta_[1] * I;
    const float complex* A = (const float complex*)A_;
    const float complex* B = (const float complex*)B_;
    float complex* C = (float complex*)C_;
    const int tA = !(ta == 'N' || ta == 'n'), tB = !(tb == 'N' || tb == 'n');
#pragma omp parallel for schedule(static)
    for (int64_t ni = 0; ni < n; ++ni) {
        for (int64_t mi = 0; mi < m; ++mi) {
            float complex c = (beta == 0.0f) ? 0.0f : C[idx(mi, ni, ldc)] * beta;
            for (int64_t ki = 0; ki < k; ++ki) {
                const float complex a = opc(ta, tA ? A[idx(ki, mi, lda)] : A[idx(mi, ki, lda)]);
                const float complex b = opc(tb, tB ? B[idx(ni, ki, ldb)] : B[idx(ki, ni, ldb)]);
                c += alpha * a * b;
            }
            C[idx(mi, ni, ldc)] = c;
        }
    }
}
