// libcosta_prefixed_scalapack.so: the same entry points as costa_p?gemr2d, costa_p?tran ... (reference costa_prefixed_scalapack).
#include <costa/pxgemr2d/costa_pxgemr2d.hpp>
#include <costa/pxgemr2d/prefixed_pxgemr2d.h>
#include <costa/pxtran/prefixed_pxtran.h>
#include <costa/pxtran_op/costa_pxtran_op.hpp>
#include <costa/pxtranc/prefixed_pxtranc.h>
#include <costa/pxtranu/prefixed_pxtranu.h>
#define COSTA_B200_SYM(x) costa_##x
#define COSTA_B200_SYM_UP(x) COSTA_##x
#include "costa_symbols.inc"
