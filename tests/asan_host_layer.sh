#!/bin/bash
# AddressSanitizer + UBSan over the whole C++ host layer on the CPU: the api/ and host/ sources, the CPU stand-in for the C ABI
# (tests/cpp/mock_b200.cpp) and a test program are linked into ONE instrumented executable per program and run on several ranks.
# Usage: tests/asan_host_layer.sh [outdir]   (takes a few minutes; last run: clean at 1-8 ranks)
set -e
R=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/cosma_b200_asan}
mkdir -p "$OUT"
gcc -O1 -g -fsanitize=address,undefined -fopenmp -c "$R/oracle/gemm_oracle.c" -o "$OUT/gemm_oracle.o"
A=$R/cosma_b200/csrc/api; H=$R/cosma_b200/csrc/host
CORE="$A/process_group.cpp $A/runtime.cpp $A/context.cpp $A/matrix.cpp $A/multiply.cpp $A/costa_api.cpp $A/cinterface.cpp $H/strategy.cpp $H/mapper.cpp
      $H/interval.cpp $H/math_utils.cpp $H/environment_variables.cpp $H/costa_layout.cpp $H/costa_reorder.cpp $R/tests/cpp/mock_b200.cpp"
PX="$A/scalapack.cpp $A/cosma_pxgemm.cpp $A/costa_pxtransform.cpp $A/blacs_lite.cpp $A/pxgemm.cpp $A/prefixed_pxgemm.cpp $A/costa_scalapack.cpp
    $A/costa_prefixed_scalapack.cpp"
FL="-O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -I $R/include -I $A"
g++ $FL "$R/tests/cpp/test_multiply.cpp" $CORE "$OUT/gemm_oracle.o" -o "$OUT/test_multiply"
g++ $FL "$R/tests/cpp/test_multiply_using_layout.cpp" $CORE "$OUT/gemm_oracle.o" -o "$OUT/test_multiply_using_layout"
g++ $FL "$R/tests/cpp/test_pxgemm.cpp" $CORE $PX "$OUT/gemm_oracle.o" -o "$OUT/test_pxgemm"
g++ $FL "$R/tests/cpp/test_pxtran.cpp" $CORE $PX -o "$OUT/test_pxtran"
g++ $FL "$R/tests/cpp/test_costa_examples.cpp" $CORE -o "$OUT/test_costa_examples"
g++ $FL "$R/miniapp/pxgemr2d_miniapp.cpp" $CORE $PX -o "$OUT/pxgemr2d_miniapp"
g++ $FL "$R/miniapp/pxtran_miniapp.cpp" $CORE $PX -o "$OUT/pxtran_miniapp"
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 COSMA_B200_PDGEMM_CASES=$R/tests/golden/pdgemm_cases.txt
cd "$R"
rc=0
for spec in "test_multiply 1" "test_multiply 8" "test_multiply_using_layout 4" "test_pxgemm 4" "test_pxtran 6" "test_costa_examples 4"; do
    set -- $spec
    echo "== $1 on $2 rank(s)"
    python -m cosma_b200.launch -np "$2" --timeout 850 "$OUT/$1" 2>&1 | grep -E "checks passed|ERROR|runtime error|SUMMARY|terminate" || rc=1
done
# the COSTA miniapps with --test (they print "Result is CORRECT!" instead of a check count)
for spec in "pxgemr2d_miniapp 6 -m 301 -n 203 --block_a 32,16 --block_c 7,50 -p 2,3 -q 3,2 -t zdouble" \
            "pxgemr2d_miniapp 6 -m 128 -n 256 --block_a 32,32 --block_c 32,32 -p 2,2 -q 1,6 -t float" \
            "pxtran_miniapp 6 -m 301 -n 203 --block_a 32,16 --block_c 7,50 -p 2,3 -t zdouble --op C --alpha 2 --beta -1"; do
    set -- $spec
    app=$1; np=$2; shift 2
    echo "== $app on $np rank(s)"
    python -m cosma_b200.launch -np "$np" --timeout 300 "$OUT/$app" "$@" --test 2>&1 | grep -E "Result is|ERROR|runtime error|SUMMARY|terminate" || rc=1
done
exit $rc
