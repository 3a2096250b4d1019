/* TEST INFRASTRUCTURE ONLY -- declarations of the four cblas_?gemm entry points the reference calls
 * (src/cosma/blas.cpp:35-129); the image has OpenBLAS only as wheel-private shared objects without headers. */
#ifndef ORACLE_CBLAS_H
#define ORACLE_CBLAS_H
/* The fastest BLAS in the image is scipy's bundled OpenBLAS 0.3.30 (LP64), whose symbols carry a scipy_ prefix. */
#ifdef ORACLE_SCIPY_OPENBLAS
#define cblas_sgemm scipy_cblas_sgemm
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_cgemm scipy_cblas_cgemm
#define cblas_zgemm scipy_cblas_zgemm
#define openblas_set_num_threads scipy_openblas_set_num_threads
#define openblas_get_num_threads scipy_openblas_get_num_threads
#endif
#ifdef __cplusplus
extern "C" {
#endif
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, float, const float*, int, const float*, int, float, float*, int);
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, double, const double*, int, const double*, int, double, double*, int);
void cblas_cgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void*, const void*, int, const void*, int, const void*, void*, int);
void cblas_zgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void*, const void*, int, const void*, int, const void*, void*, int);
void openblas_set_num_threads(int);
int openblas_get_num_threads(void);
#ifdef __cplusplus
}
#endif
#endif
