#!/bin/bash
# AddressSanitizer + UBSan over the plan-time host code of libcosma_b200.so (Strategy, Mapper, schedule compiler, overlap planner) with
# random problems (tests/cpp/fuzz_planner.cpp), and over the COSTA message-list planner with random layouts (tests/cpp/fuzz_transform_planner.cpp). Usage: tests/asan_planner.sh [outdir] [seeds...]   (about a minute per seed)
set -e
R=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/cosma_b200_asan_planner}
shift || true
SEEDS=${@:-1 2 3}
mkdir -p "$OUT"
H=$R/cosma_b200/csrc/host
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -I "$R/include" "$R/tests/cpp/fuzz_planner.cpp" \
    "$H/strategy.cpp" "$H/mapper.cpp" "$H/interval.cpp" "$H/math_utils.cpp" "$H/environment_variables.cpp" "$H/schedule.cpp" "$H/overlap.cpp" "$H/auto_strategy.cpp" \
    -o "$OUT/fuzz_planner"
export ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -I "$R/include" "$R/tests/cpp/fuzz_transform_planner.cpp" \
    "$H/costa_transform.cpp" "$H/costa_layout.cpp" "$H/costa_reorder.cpp" -o "$OUT/fuzz_transform_planner"
for s in $SEEDS; do "$OUT/fuzz_planner" "$s" 400; "$OUT/fuzz_transform_planner" "$s" 600; done
