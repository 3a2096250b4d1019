// libcosma_prefixed_pxgemm.so: cosma_p?gemm / COSMA_P?GEMM, for callers that want both ScaLAPACK's and COSMA's p?gemm in
// one executable (reference src/cosma/prefixed_pxgemm.cpp).
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <dlfcn.h>

#include <cosma/cosma_pxgemm.hpp>
#include <cosma/prefixed_pxgemm.h>
#define COSMA_B200_DELEGATE 0
#define COSMA_B200_SYM(x) cosma_##x
#define COSMA_B200_SYM_UP(x) COSMA_##x
#include "pxgemm_symbols.inc"
