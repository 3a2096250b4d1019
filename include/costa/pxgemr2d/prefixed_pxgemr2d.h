// p?gemr2d ScaLAPACK symbols under a costa_ prefix (reference libs/COSTA/src/costa/pxgemr2d/prefixed_pxgemr2d.h): all-pointer Fortran ABI in lower / upper case, with
// and without the trailing underscore; complex arrays are interleaved (re, im).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define COSTA_B200_GEMR2D_ABI(NAME, T)                                                                                 \
    void NAME(const int* m, const int* n, const T* a, const int* ia, const int* ja, const int* desca, T* b, const int* ib,  \
              const int* jb, const int* descb, const int* ictxt)
COSTA_B200_GEMR2D_ABI(costa_psgemr2d, float); COSTA_B200_GEMR2D_ABI(costa_psgemr2d_, float); COSTA_B200_GEMR2D_ABI(COSTA_PSGEMR2D, float); COSTA_B200_GEMR2D_ABI(COSTA_PSGEMR2D_, float);
COSTA_B200_GEMR2D_ABI(costa_pdgemr2d, double); COSTA_B200_GEMR2D_ABI(costa_pdgemr2d_, double); COSTA_B200_GEMR2D_ABI(COSTA_PDGEMR2D, double); COSTA_B200_GEMR2D_ABI(COSTA_PDGEMR2D_, double);
COSTA_B200_GEMR2D_ABI(costa_pcgemr2d, float); COSTA_B200_GEMR2D_ABI(costa_pcgemr2d_, float); COSTA_B200_GEMR2D_ABI(COSTA_PCGEMR2D, float); COSTA_B200_GEMR2D_ABI(COSTA_PCGEMR2D_, float);
COSTA_B200_GEMR2D_ABI(costa_pzgemr2d, double); COSTA_B200_GEMR2D_ABI(costa_pzgemr2d_, double); COSTA_B200_GEMR2D_ABI(COSTA_PZGEMR2D, double); COSTA_B200_GEMR2D_ABI(COSTA_PZGEMR2D_, double);
#undef COSTA_B200_GEMR2D_ABI
#ifdef __cplusplus
}
#endif
