// p{c,z}tranu: plain transpose of a complex matrix (reference libs/COSTA/src/costa/pxtranu/pxtranu.h:7-20): all-pointer Fortran ABI in lower / upper case, with and without
// the trailing underscore; sub(C) (m x n) = beta * sub(C) + alpha * op(sub(A)) with sub(A) n x m.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_TRAN_ABI(NAME, T)                                                                                    \
    void NAME(const int* m, const int* n, const T* alpha, const T* a, const int* ia, const int* ja, const int* desca,      \
              const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSTA_B200_TRAN_ABI(pctranu, float); COSTA_B200_TRAN_ABI(pctranu_, float); COSTA_B200_TRAN_ABI(PCTRANU, float); COSTA_B200_TRAN_ABI(PCTRANU_, float);
COSTA_B200_TRAN_ABI(pztranu, double); COSTA_B200_TRAN_ABI(pztranu_, double); COSTA_B200_TRAN_ABI(PZTRANU, double); COSTA_B200_TRAN_ABI(PZTRANU_, double);
#undef COSTA_B200_TRAN_ABI
#ifdef __cplusplus
}
#endif
