// p{c,z}tranu: plain transpose of a complex matrix under a costa_ prefix (reference libs/COSTA/src/costa/pxtranu/prefixed_pxtranu.h): all-pointer Fortran ABI in lower / upper case, with and without
// the trailing underscore; sub(C) (m x n) = beta * sub(C) + alpha * op(sub(A)) with sub(A) n x m.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_TRAN_ABI(NAME, T)                                                                                    \
    void NAME(const int* m, const int* n, const T* alpha, const T* a, const int* ia, const int* ja, const int* desca,      \
              const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSTA_B200_TRAN_ABI(costa_pctranu, float); COSTA_B200_TRAN_ABI(costa_pctranu_, float); COSTA_B200_TRAN_ABI(COSTA_PCTRANU, float); COSTA_B200_TRAN_ABI(COSTA_PCTRANU_, float);
COSTA_B200_TRAN_ABI(costa_pztranu, double); COSTA_B200_TRAN_ABI(costa_pztranu_, double); COSTA_B200_TRAN_ABI(COSTA_PZTRANU, double); COSTA_B200_TRAN_ABI(COSTA_PZTRANU_, double);
#undef COSTA_B200_TRAN_ABI
#ifdef __cplusplus
}
#endif
