// libcosma_pxgemm.so: the ScaLAPACK names themselves, so that linking (or LD_PRELOADing) this library in front of ScaLAPACK
// routes every p?gemm of an unmodified application here (reference src/cosma/pxgemm.cpp, CMakeLists.txt:71-87).
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <dlfcn.h>

#include <cosma/cosma_pxgemm.hpp>
#include <cosma/pxgemm.h>
#define COSMA_B200_DELEGATE 1
#define COSMA_B200_SYM(x) x
#define COSMA_B200_SYM_UP(x) x
#include "pxgemm_symbols.inc"
