// p{s,d}tran: transpose of a real matrix (reference libs/COSTA/src/costa/pxtran/pxtran.h:7-20): all-pointer Fortran ABI in lower / upper case, with and without
// the trailing underscore; sub(C) (m x n) = beta * sub(C) + alpha * op(sub(A)) with sub(A) n x m.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_TRAN_ABI(NAME, T)                                                                                    \
    void NAME(const int* m, const int* n, const T* alpha, const T* a, const int* ia, const int* ja, const int* desca,      \
              const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSTA_B200_TRAN_ABI(pstran, float); COSTA_B200_TRAN_ABI(pstran_, float); COSTA_B200_TRAN_ABI(PSTRAN, float); COSTA_B200_TRAN_ABI(PSTRAN_, float);
COSTA_B200_TRAN_ABI(pdtran, double); COSTA_B200_TRAN_ABI(pdtran_, double); COSTA_B200_TRAN_ABI(PDTRAN, double); COSTA_B200_TRAN_ABI(PDTRAN_, double);
#undef COSTA_B200_TRAN_ABI
#ifdef __cplusplus
}
#endif
