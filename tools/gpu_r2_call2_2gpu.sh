#!/bin/bash
# Round 2, second call (2 GPUs, <= 6 min of box time = 12 GPU-minutes): the multi-rank paths written after the round-1 GPU budget ran
# out (DESIGN 9 items 1-3 and the copy-engine probe of item 5). EVERY step carries its own short wall-clock limit and the host layer's
# receive timeout, so that a protocol error costs seconds, not the call.
#   gpurun --gpus 2 --timeout 400 -- 'bash tools/gpu_r2_call2_2gpu.sh'
mkdir -p gpurun_out
export COSMA_B200_PG_RECV_TIMEOUT=40
nvidia-smi -L > gpurun_out/gpus.txt
# the established 2-GPU suite first (its "host" cases -- streamed un-gathered operands -- have not run on 2 GPUs since they were written)
timeout 90 python -m pytest tests/test_multiply_gpu.py -m gpu -q -k two_gpus > gpurun_out/r2_pytest_two_gpus.txt 2>&1; tail -2 gpurun_out/r2_pytest_two_gpus.txt
COSMA_B200_CPP_MULTIRANK=1 COSMA_B200_TRACE=ON timeout 120 python -m pytest tests/test_z_cpp_api.py -m gpu -q -k "2-" > gpurun_out/r2_pytest_cpp_n2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_cpp_n2.txt; grep -v "^\[cosma rank" gpurun_out/r2_pytest_cpp_n2.txt | tail -8
COSMA_B200_REORDER_RANKS=ON timeout 90 python -m pytest tests/test_costa_gpu.py -m gpu -q -k two_gpus > gpurun_out/r2_pytest_relabel_n2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_relabel_n2.txt; tail -3 gpurun_out/r2_pytest_relabel_n2.txt
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 \
    > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 900 gpurun_out/r2_bench_n2.json
COSMA_B200_TEST_HOST_PANELS=1 timeout 200 python -m pytest tests/test_zz_optin_gpu.py -m gpu -q -k "host_panels and 2-" > gpurun_out/r2_pytest_host_panels_n2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_host_panels_n2.txt; tail -3 gpurun_out/r2_pytest_host_panels_n2.txt
timeout 60 python tools/ce_overlap_probe.py > gpurun_out/r2_ce_overlap_probe.json 2>&1; tail -3 gpurun_out/r2_ce_overlap_probe.json
for app in pxgemr2d_miniapp pxtran_miniapp; do
  timeout 60 python -m cosma_b200.launch -np 2 tests/cpp/bin/$app -m 16384 -n 16384 --block_a 256,256 --block_c 128,512 -p 1,2 -t zdouble -r 4 >> gpurun_out/r2_costa_miniapps_n2.txt 2>&1
done
tail -8 gpurun_out/r2_costa_miniapps_n2.txt
