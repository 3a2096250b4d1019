// libcosta_scalapack.so: p?gemr2d and p?tran / p?tranu / p?tranc under their ScaLAPACK names, for interposition in front of
// ScaLAPACK (reference libs/COSTA/src/costa/CMakeLists.txt: costa_scalapack).
#include <costa/pxgemr2d/costa_pxgemr2d.hpp>
#include <costa/pxgemr2d/pxgemr2d.h>
#include <costa/pxtran/pxtran.h>
#include <costa/pxtran_op/costa_pxtran_op.hpp>
#include <costa/pxtranc/pxtranc.h>
#include <costa/pxtranu/pxtranu.h>
#define COSTA_B200_SYM(x) x
#define COSTA_B200_SYM_UP(x) x
#include "costa_symbols.inc"
