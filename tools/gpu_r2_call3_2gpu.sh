#!/bin/bash
# Round 2, third call (2 GPUs): the second-generation 3xTF32 kernel (A through TMEM) on GPU 0, then the copy-engine peer transport of the
# overlapped schedules (tests with both transports, bench lines). EVERY step carries its own wall-clock limit.
#   gpurun --gpus 2 --timeout 1100 -- 'bash tools/gpu_r2_call3_2gpu.sh'
mkdir -p gpurun_out
export COSMA_B200_PG_RECV_TIMEOUT=40
t() { log=gpurun_out/$1; lim=$2; shift 2; timeout $lim "$@" > $log 2>&1; echo "rc=$?" >> $log; echo "== $log: $(tail -4 $log | tr '\n' ' ' | cut -c1-400)"; }
# 1. K3/K4 second generation
COSMA_B200_TF32_KERNEL=v2 t r2c_pytest_f32_v2.txt 400 python -m pytest tests/test_gemm_f32_gpu.py -m gpu -q -x
for gen in v1 v2; do
  for dt in s c; do COSMA_B200_TF32_KERNEL=$gen timeout 60 python tools/gemm_time.py --dtype $dt --m 8192 --n 8192 --k 8192 --reps 3 >> gpurun_out/r2c_gemm_f32_$gen.jsonl 2>&1; done
  COSMA_B200_TF32_KERNEL=$gen timeout 60 python tools/gemm_time.py --dtype s --m 16384 --n 16384 --k 16384 --reps 3 >> gpurun_out/r2c_gemm_f32_$gen.jsonl 2>&1
  COSMA_B200_TF32_KERNEL=$gen timeout 60 python tools/gemm_time.py --dtype s --transa T --m 8192 --n 8192 --k 8192 --reps 3 >> gpurun_out/r2c_gemm_f32_$gen.jsonl 2>&1
done
COSMA_B200_TF32_KERNEL=v2 timeout 60 python tools/gemm_time.py --dtype c --transa C --m 8192 --n 8192 --k 8192 --reps 3 >> gpurun_out/r2c_gemm_f32_v2.jsonl 2>&1
COSMA_B200_TF32_KERNEL=v2 timeout 120 python tools/gemm_f32_probe.py > gpurun_out/r2c_gemm_f32_probe_v2.jsonl 2>&1
grep -h tflops gpurun_out/r2c_gemm_f32_v*.jsonl | cut -c1-160
COSMA_B200_TF32_KERNEL=v2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_v2 -s 2 -c 1 -o gpurun_out/r2c_sgemm8192_v2 \
    python tools/gemm_time.py --dtype s --m 8192 --n 8192 --k 8192 --reps 1 > gpurun_out/ncu_sgemm_v2.log 2>&1
# 2. copy-engine peer transport
t r2c_pytest_multiply_n2.txt 200 python -m pytest tests/test_multiply_gpu.py -m gpu -q -x -k two_gpus
COSMA_B200_PEER_COPY=OFF t r2c_pytest_multiply_nccl_n2.txt 200 python -m pytest tests/test_multiply_gpu.py -m gpu -q -x -k two_gpus
t r2c_pytest_costa_n2.txt 240 python -m pytest tests/test_costa_gpu.py -m gpu -q -x -k two_gpus
COSMA_B200_CPP_MULTIRANK=1 t r2c_pytest_cpp_n2.txt 300 python -m pytest tests/test_z_cpp_api.py -m gpu -q -k "2-"
bench() { tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 $EXTRA \
      > gpurun_out/r2c_bench_n2_$tag.json 2> gpurun_out/r2c_bench_n2_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/r2c_bench_n2_%s.json" % sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"].get("value", 0), 1), "parity", d.get("parity") and d["parity"].get("ok"), d.get("collectives"))
except Exception as e:
    print(sys.argv[1], "no line:", e); print(open("gpurun_out/r2c_bench_n2_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
EXTRA="--no-cpu-baseline"
bench ce X=1
EXTRA="--no-e2e --no-parity --no-cpu-baseline"
bench nccl COSMA_B200_PEER_COPY=OFF
EXTRA="--no-e2e --no-cpu-baseline --strategy pn2"
bench pn2_ce X=1
EXTRA="--no-e2e --no-cpu-baseline --strategy pm2"
bench pm2_ce X=1
ls -la gpurun_out | tail -12
