#!/bin/bash
# Round 2, last call (1 GPU, the remaining minutes): the whole single-GPU suite at HEAD, the N = 1 bench line in its new form (32768^3),
# its ncu launch list, the repack probe, smoke().
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu_n1.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu_n1.txt; tail -4 gpurun_out/r2e_pytest_gpu_n1.txt
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; tail -c 1500 gpurun_out/r2e_bench_n1.json; tail -5 gpurun_out/r2e_bench_n1.err
timeout 30 python tools/gemm_time.py --dtype d --m 8191 --n 8192 --k 8191 --reps 3 > gpurun_out/r2e_repack_auto.json 2>&1; cut -c1-300 gpurun_out/r2e_repack_auto.json
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-parity > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/r2e_launches_bench_n1.csv | cut -c1-200
