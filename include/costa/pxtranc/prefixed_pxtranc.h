// p{c,z}tranc: conjugate transpose of a complex matrix under a costa_ prefix (reference libs/COSTA/src/costa/pxtranc/prefixed_pxtranc.h): all-pointer Fortran ABI in lower / upper case, with and without
// the trailing underscore; sub(C) (m x n) = beta * sub(C) + alpha * op(sub(A)) with sub(A) n x m.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_TRAN_ABI(NAME, T)                                                                                    \
    void NAME(const int* m, const int* n, const T* alpha, const T* a, const int* ia, const int* ja, const int* desca,      \
              const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSTA_B200_TRAN_ABI(costa_pctranc, float); COSTA_B200_TRAN_ABI(costa_pctranc_, float); COSTA_B200_TRAN_ABI(COSTA_PCTRANC, float); COSTA_B200_TRAN_ABI(COSTA_PCTRANC_, float);
COSTA_B200_TRAN_ABI(costa_pztranc, double); COSTA_B200_TRAN_ABI(costa_pztranc_, double); COSTA_B200_TRAN_ABI(COSTA_PZTRANC, double); COSTA_B200_TRAN_ABI(COSTA_PZTRANC_, double);
#undef COSTA_B200_TRAN_ABI
#ifdef __cplusplus
}
#endif
