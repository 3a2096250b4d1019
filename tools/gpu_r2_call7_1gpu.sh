#!/bin/bash
# Round 2, the last 50 GPU-seconds (1 GPU): what changed on the GPU side after call 5 -- smoke() with its SGEMM / relayout legs, the
# restructured bench.py (line assembled stage by stage) on the short cfg2 workload, the single-GPU multiply schedules.
#   gpurun --timeout 45 -- 'bash tools/gpu_r2_call7_1gpu.sh'
mkdir -p gpurun_out
( timeout 12 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r2h_smoke.txt; cat gpurun_out/r2h_smoke.txt
timeout 20 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_cfg2_n1.json 2> gpurun_out/r2h_bench_cfg2_n1.err; tail -c 600 gpurun_out/r2h_bench_cfg2_n1.json; tail -3 gpurun_out/r2h_bench_cfg2_n1.err
timeout 25 python -m pytest tests/test_multiply_gpu.py -m gpu -x -q -k single_gpu > gpurun_out/r2h_pytest_multiply_n1.txt 2>&1; tail -2 gpurun_out/r2h_pytest_multiply_n1.txt
