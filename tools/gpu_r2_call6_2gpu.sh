#!/bin/bash
# Round 2, the last GPU seconds (2 GPUs): the N = 2 bench line at HEAD (overlap on copy engines with the own pieces placed first, exact
# parity), then as much of the 2-GPU multiply suite (host panels by default) as the remaining budget allows.
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; tail -c 1200 gpurun_out/r2f_bench_n2.json; tail -3 gpurun_out/r2f_bench_n2.err
timeout 60 python -m pytest tests/test_multiply_gpu.py -m gpu -q -x -k two_gpus > gpurun_out/r2f_pytest_multiply_n2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest_multiply_n2.txt; tail -3 gpurun_out/r2f_pytest_multiply_n2.txt
