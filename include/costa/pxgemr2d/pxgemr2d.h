// p?gemr2d ScaLAPACK symbols (reference libs/COSTA/src/costa/pxgemr2d/pxgemr2d.h:7-41): all-pointer Fortran ABI in lower / upper case, with
// and without the trailing underscore; complex arrays are interleaved (re, im).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_GEMR2D_ABI(NAME, T)                                                                                 \
    void NAME(const int* m, const int* n, const T* a, const int* ia, const int* ja, const int* desca, T* b, const int* ib,  \
              const int* jb, const int* descb, const int* ictxt)
COSTA_B200_GEMR2D_ABI(psgemr2d, float); COSTA_B200_GEMR2D_ABI(psgemr2d_, float); COSTA_B200_GEMR2D_ABI(PSGEMR2D, float); COSTA_B200_GEMR2D_ABI(PSGEMR2D_, float);
COSTA_B200_GEMR2D_ABI(pdgemr2d, double); COSTA_B200_GEMR2D_ABI(pdgemr2d_, double); COSTA_B200_GEMR2D_ABI(PDGEMR2D, double); COSTA_B200_GEMR2D_ABI(PDGEMR2D_, double);
COSTA_B200_GEMR2D_ABI(pcgemr2d, float); COSTA_B200_GEMR2D_ABI(pcgemr2d_, float); COSTA_B200_GEMR2D_ABI(PCGEMR2D, float); COSTA_B200_GEMR2D_ABI(PCGEMR2D_, float);
COSTA_B200_GEMR2D_ABI(pzgemr2d, double); COSTA_B200_GEMR2D_ABI(pzgemr2d_, double); COSTA_B200_GEMR2D_ABI(PZGEMR2D, double); COSTA_B200_GEMR2D_ABI(PZGEMR2D_, double);
#undef COSTA_B200_GEMR2D_ABI
#ifdef __cplusplus
}
#endif
