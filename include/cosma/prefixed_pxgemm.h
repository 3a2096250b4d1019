// The same entry points under a cosma_ prefix, for explicit calls (reference src/cosma/prefixed_pxgemm.h): all-pointer Fortran ABI, lower / upper case, with and
// without the trailing underscore. Complex scalars and arrays are interleaved (re, im) float / double.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSMA_B200_PXGEMM_ABI(NAME, T)                                                                                   \
    void NAME(const char* trans_a, const char* trans_b, const int* m, const int* n, const int* k, const T* alpha,       \
              const T* a, const int* ia, const int* ja, const int* desca, const T* b, const int* ib, const int* jb,     \
              const int* descb, const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSMA_B200_PXGEMM_ABI(cosma_psgemm, float); COSMA_B200_PXGEMM_ABI(cosma_psgemm_, float); COSMA_B200_PXGEMM_ABI(COSMA_PSGEMM, float); COSMA_B200_PXGEMM_ABI(COSMA_PSGEMM_, float);
COSMA_B200_PXGEMM_ABI(cosma_pdgemm, double); COSMA_B200_PXGEMM_ABI(cosma_pdgemm_, double); COSMA_B200_PXGEMM_ABI(COSMA_PDGEMM, double); COSMA_B200_PXGEMM_ABI(COSMA_PDGEMM_, double);
COSMA_B200_PXGEMM_ABI(cosma_pcgemm, float); COSMA_B200_PXGEMM_ABI(cosma_pcgemm_, float); COSMA_B200_PXGEMM_ABI(COSMA_PCGEMM, float); COSMA_B200_PXGEMM_ABI(COSMA_PCGEMM_, float);
COSMA_B200_PXGEMM_ABI(cosma_pzgemm, double); COSMA_B200_PXGEMM_ABI(cosma_pzgemm_, double); COSMA_B200_PXGEMM_ABI(COSMA_PZGEMM, double); COSMA_B200_PXGEMM_ABI(COSMA_PZGEMM_, double);
#undef COSMA_B200_PXGEMM_ABI
#ifdef __cplusplus
}
#endif
