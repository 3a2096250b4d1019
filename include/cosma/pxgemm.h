// ScaLAPACK p?gemm symbols, for interposition (reference src/cosma/pxgemm.h:6-107): all-pointer Fortran ABI, lower / upper case, with and
// without the trailing underscore. Complex scalars and arrays are interleaved (re, im) float / double.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSMA_B200_PXGEMM_ABI(NAME, T)                                                                                   \
    void NAME(const char* trans_a, const char* trans_b, const int* m, const int* n, const int* k, const T* alpha,       \
              const T* a, const int* ia, const int* ja, const int* desca, const T* b, const int* ib, const int* jb,     \
              const int* descb, const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSMA_B200_PXGEMM_ABI(psgemm, float); COSMA_B200_PXGEMM_ABI(psgemm_, float); COSMA_B200_PXGEMM_ABI(PSGEMM, float); COSMA_B200_PXGEMM_ABI(PSGEMM_, float);
COSMA_B200_PXGEMM_ABI(pdgemm, double); COSMA_B200_PXGEMM_ABI(pdgemm_, double); COSMA_B200_PXGEMM_ABI(PDGEMM, double); COSMA_B200_PXGEMM_ABI(PDGEMM_, double);
COSMA_B200_PXGEMM_ABI(pcgemm, float); COSMA_B200_PXGEMM_ABI(pcgemm_, float); COSMA_B200_PXGEMM_ABI(PCGEMM, float); COSMA_B200_PXGEMM_ABI(PCGEMM_, float);
COSMA_B200_PXGEMM_ABI(pzgemm, double); COSMA_B200_PXGEMM_ABI(pzgemm_, double); COSMA_B200_PXGEMM_ABI(PZGEMM, double); COSMA_B200_PXGEMM_ABI(PZGEMM_, double);
#undef COSMA_B200_PXGEMM_ABI
#ifdef __cplusplus
}
#endif
