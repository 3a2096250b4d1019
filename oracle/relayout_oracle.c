/* TEST INFRASTRUCTURE ONLY -- CPU restatement of COSTA's block copy/transform kernel.
 * Restates costa::memory::copy_and_transform (reference libs/COSTA/src/costa/grid2grid/memory_utils.hpp:287-346)
 * and its helpers copy2D (:47-85) / transpose (:88-250):
 *     dest = beta * dest + alpha * op(src),  op in {identity, transpose, conjugate-transpose},
 * for a block of n_rows x n_cols SOURCE elements, with independent storage orders ('C' col-major / 'R' row-major)
 * and leading dimensions for source and destination. With alpha = 1, beta = 0 this is pure data movement and the
 * comparison is bit-exact (conjugation only flips the sign bit of the imaginary part).
 * Pinned by the reference's golden vectors libs/COSTA/tests/unit/test_utils.cpp (tests/golden/costa_copy_and_transform.json). */
#include <stdint.h>

/* element (i,j) of a block stored with leading dimension ld in ordering ord */
static inline int64_t at(int64_t i, int64_t j, int64_t ld, char ord) { return ord == 'R' ? i * ld + j : i + j * ld; }

#define DEFINE_REAL(NAME, T)                                                                                       \
    void NAME(int64_t n_rows, int64_t n_cols, const T* src, int64_t src_ld, char src_ord, T* dst, int64_t dst_ld,  \
              char dst_ord, int transpose, int conjugate, T alpha, T beta) {                                      \
        (void)conjugate;                                                                                           \
        for (int64_t j = 0; j < n_cols; ++j)                                                                       \
            for (int64_t i = 0; i < n_rows; ++i) {                                                                 \
                const T s = src[at(i, j, src_ld, src_ord)];                                                        \
                T* d = transpose ? &dst[at(j, i, dst_ld, dst_ord)] : &dst[at(i, j, dst_ld, dst_ord)];              \
                *d = (beta == (T)0) ? alpha * s : beta * *d + alpha * s;                                           \
            }                                                                                                      \
    }
DEFINE_REAL(oracle_copy_and_transform_d, double)
DEFINE_REAL(oracle_copy_and_transform_s, float)
DEFINE_REAL(oracle_copy_and_transform_i, int)

/* complex: interleaved (re, im); alpha/beta point at 2 values */
#define DEFINE_CPLX(NAME, T)                                                                                       \
    void NAME(int64_t n_rows, int64_t n_cols, const T* src, int64_t src_ld, char src_ord, T* dst, int64_t dst_ld,  \
              char dst_ord, int transpose, int conjugate, const T* alpha, const T* beta) {                        \
        const int beta_zero = beta[0] == (T)0 && beta[1] == (T)0;                                                  \
        for (int64_t j = 0; j < n_cols; ++j)                                                                       \
            for (int64_t i = 0; i < n_rows; ++i) {                                                                 \
                const T* s = &src[2 * at(i, j, src_ld, src_ord)];                                                  \
                const T sr = s[0], si = conjugate ? -s[1] : s[1];                                                  \
                T* d = transpose ? &dst[2 * at(j, i, dst_ld, dst_ord)] : &dst[2 * at(i, j, dst_ld, dst_ord)];      \
                T vr = alpha[0] * sr - alpha[1] * si, vi = alpha[0] * si + alpha[1] * sr;                          \
                if (alpha[0] == (T)1 && alpha[1] == (T)0) { vr = sr; vi = si; }                                    \
                if (!beta_zero) {                                                                                  \
                    const T dr = d[0], di = d[1];                                                                  \
                    vr += beta[0] * dr - beta[1] * di;                                                             \
                    vi += beta[0] * di + beta[1] * dr;                                                             \
                }                                                                                                  \
                d[0] = vr;                                                                                         \
                d[1] = vi;                                                                                         \
            }                                                                                                      \
    }
DEFINE_CPLX(oracle_copy_and_transform_z, double)
DEFINE_CPLX(oracle_copy_and_transform_c, float)
