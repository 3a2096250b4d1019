// p{s,d}tran: transpose of a real matrix under a costa_ prefix (reference libs/COSTA/src/costa/pxtran/prefixed_pxtran.h): all-pointer Fortran ABI in lower / upper case, with and without
// the trailing underscore; sub(C) (m x n) = beta * sub(C) + alpha * op(sub(A)) with sub(A) n x m.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_TRAN_ABI(NAME, T)                                                                                    \
    void NAME(const int* m, const int* n, const T* alpha, const T* a, const int* ia, const int* ja, const int* desca,      \
              const T* beta, T* c, const int* ic, const int* jc, const int* descc)
COSTA_B200_TRAN_ABI(costa_pstran, float); COSTA_B200_TRAN_ABI(costa_pstran_, float); COSTA_B200_TRAN_ABI(COSTA_PSTRAN, float); COSTA_B200_TRAN_ABI(COSTA_PSTRAN_, float);
COSTA_B200_TRAN_ABI(costa_pdtran, double); COSTA_B200_TRAN_ABI(costa_pdtran_, double); COSTA_B200_TRAN_ABI(COSTA_PDTRAN, double); COSTA_B200_TRAN_ABI(COSTA_PDTRAN_, double);
#undef COSTA_B200_TRAN_ABI
#ifdef __cplusplus
}
#endif
