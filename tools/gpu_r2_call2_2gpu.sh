#!/bin/bash
# Round 2, second call (2 GPUs): multi-rank parity at HEAD (multiply incl. the overlapped schedules, COSTA / p?gemm incl. the reference's own
# parameter sets, live reference, the C++ programs), rank relabelling and COSMA_ADAPT_STRATEGY switched on, host panels, and the N = 2 bench
# line with the overlap on / off and 4 / 8 / 16 reserved SMs. EVERY step carries its own wall-clock limit.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2_call2_2gpu.sh'
mkdir -p gpurun_out
export COSMA_B200_PG_RECV_TIMEOUT=40
nvidia-smi -L > gpurun_out/gpus.txt
t() { # $1 = log name, $2 = limit, rest = command
  log=gpurun_out/$1; lim=$2; shift 2
  timeout $lim "$@" > $log 2>&1; echo "rc=$?" >> $log; echo "== $log: $(tail -3 $log | tr '\n' ' ' | cut -c1-300)"
}
t r2_pytest_multiply_n2.txt 200 python -m pytest tests/test_multiply_gpu.py -m gpu -q -x -k two_gpus
t r2_pytest_costa_n2.txt 240 python -m pytest tests/test_costa_gpu.py -m gpu -q -x -k two_gpus
COSMA_B200_REORDER_RANKS=ON t r2_pytest_costa_relabel_n2.txt 240 python -m pytest tests/test_costa_gpu.py -m gpu -q -x -k two_gpus
COSMA_ADAPT_STRATEGY=ON t r2_pytest_costa_adapt_n2.txt 240 python -m pytest tests/test_costa_gpu.py -m gpu -q -x -k two_gpus
t r2_pytest_reflive_n2.txt 200 python -m pytest tests/test_ref_live_gpu.py -m gpu -q -x -k two_gpus
COSMA_B200_CPP_MULTIRANK=1 t r2_pytest_cpp_n2.txt 300 python -m pytest tests/test_z_cpp_api.py -m gpu -q -k "2-"
COSMA_B200_TEST_HOST_PANELS=1 t r2_pytest_host_panels_n2.txt 200 python -m pytest tests/test_zz_optin_gpu.py -m gpu -q -k "host_panels and 2-"
bench() { # $1 = tag, rest: env assignments
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 $EXTRA \
      > gpurun_out/r2_bench_n2_$tag.json 2> gpurun_out/r2_bench_n2_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_%s.json" % sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"].get("value", 0), 1), "parity", d.get("parity") and d["parity"].get("ok"), d.get("collectives"))
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
EXTRA=""
bench auto X=1
EXTRA="--no-e2e --no-parity --no-cpu-baseline"
bench serial COSMA_OVERLAP_COMM_AND_COMP=OFF
bench sms4 COSMA_B200_OVERLAP_SMS=4
bench sms16 COSMA_B200_OVERLAP_SMS=16
EXTRA="--no-e2e --no-cpu-baseline --strategy pn2"
bench pn2_auto X=1
EXTRA="--no-e2e --no-parity --no-cpu-baseline --strategy pn2"
bench pn2_serial COSMA_OVERLAP_COMM_AND_COMP=OFF
ls -la gpurun_out | tail -20
