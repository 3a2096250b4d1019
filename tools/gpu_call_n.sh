#!/bin/bash
# usage: gpu_call_n.sh N   -- multi-GPU parity tests and the multi-GPU workloads at N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multiply_gpu.py tests/test_costa_gpu.py -x -q -m gpu > gpurun_out/pytest_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_n$N.log
tail -6 gpurun_out/pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/bench_pxgemm.py --steps 5 --warmup 3 > gpurun_out/bench_pzgemm_n$N.json 2> gpurun_out/bench_pzgemm_n$N.err; tail -c 1800 gpurun_out/bench_pzgemm_n$N.json; tail -3 gpurun_out/bench_pzgemm_n$N.err
timeout 600 $TR tools/bench_pxgemm.py --steps 5 --warmup 3 --dtype d --transa T > gpurun_out/bench_pdgemm_n$N.json 2> gpurun_out/bench_pdgemm_n$N.err; tail -c 1500 gpurun_out/bench_pdgemm_n$N.json
if [ "$N" = "8" ]; then
  timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 1500 gpurun_out/bench_n8.json
  timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 --mnk 8192,8192,1048576 --no-e2e > gpurun_out/bench_largek_n8.json 2> gpurun_out/bench_largek_n8.err; tail -c 1500 gpurun_out/bench_largek_n8.json; tail -3 gpurun_out/bench_largek_n8.err
fi
