// FP64 pipe probe for B200 (sm_100a): how fast do DMMA.8x8x4 and DFMA issue?
// Output: one JSON object per line. Used once per pod to fix the FP64 roofline
// denominator (MEASURED_PEAKS.json has no FP64 figure). Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-12 * threadIdx.x;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 1 << 20));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  const int iters = 20000;
  int warps_list[] = {1, 2, 4, 8, 12, 16, 32};
  for (int wi = 0; wi < 7; wi++) {
    int warps = warps_list[wi];
    {
      float ms = time_ms([&] { dmma_loop<16><<<sms, warps * 32>>>(out, iters); });
      double flops = 2.0 * 256 * 16.0 * iters * warps * sms;
      printf("{\"probe\": \"dmma884\", \"acc\": 16, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma_loop<4><<<sms, warps * 32>>>(out, iters); });
      double flops = 2.0 * 256 * 4.0 * iters * warps * sms;
      printf("{\"probe\": \"dmma884\", \"acc\": 4, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dfma_loop<16><<<sms, warps * 32>>>(out, iters); });
      double flops = 2.0 * 32 * 16.0 * iters * warps * sms;
      printf("{\"probe\": \"dfma\", \"acc\": 16, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9);
    }
  }
  // a longer sustained DMMA run (~2 s) to see the power-capped figure
  {
    float ms = time_ms([&] { dmma_loop<16><<<sms * 2, 512>>>(out, 2000000); }, 1);
    double flops = 2.0 * 256 * 16.0 * 2000000.0 * 16 * sms * 2;
    printf("{\"probe\": \"dmma884_sustained\", \"ms\": %.2f, \"tflops\": %.3f}\n", ms, flops / ms * 1e-9);
  }
  return 0;
}
