#!/bin/bash
# Round 2, third call (8 GPUs, <= 5 min of box time = 40 GPU-minutes; only after call 2 was green): the multi-GPU suites and the N = 8
# bench line with and without the pipelined host panels. Every step under its own wall-clock limit.
#   gpurun --gpus 8 --timeout 330 -- 'bash tools/gpu_r2_call3_8gpu.sh'
mkdir -p gpurun_out
export COSMA_B200_PG_RECV_TIMEOUT=40
run_bench() {  # $1 = tag, rest = environment assignments
  tag=$1; shift
  env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 3 --warmup 3 \
      > gpurun_out/r2_bench_n8_$tag.json 2> gpurun_out/r2_bench_n8_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2_bench_n8_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "match", d["e2e"].get("matches_device_path"), d["config"].get("host_affinity_rank0"), d.get("collectives"))
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
# how much of the host link each rank keeps when eight share the host, without and with NUMA placement
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 tools/pcie_probe.py > gpurun_out/r2_pcie_probe_n8.jsonl 2>&1
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29583 tools/pcie_probe.py --bind > gpurun_out/r2_pcie_probe_n8_bound.jsonl 2>&1
grep -h h2d_contig gpurun_out/r2_pcie_probe_n8.jsonl gpurun_out/r2_pcie_probe_n8_bound.jsonl | cut -c1-200 | head -4
run_bench plain COSMA_B200_HOST_PANELS=0
run_bench panels4 COSMA_B200_HOST_PANELS=4
run_bench panels8 COSMA_B200_HOST_PANELS=8
COSMA_B200_TEST_HOST_PANELS=1 timeout 100 python -m pytest tests/test_zz_optin_gpu.py -m gpu -q -k "host_panels" > gpurun_out/r2_pytest_host_panels_n8.txt 2>&1; tail -2 gpurun_out/r2_pytest_host_panels_n8.txt
