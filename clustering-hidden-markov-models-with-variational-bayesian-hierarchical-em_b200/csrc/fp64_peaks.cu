// fp64_peaks.cu -- measures the fp64 roofline denominators on the box itself.
//
// MEASURED_PEAKS.json (driver-written) holds HBM and bf16 numbers only; the
// E-step is fp64-pipe bound, so the denominator has to
// be measured here: DFMA chains, DMMA (mma.sync m8n8k4 f64) chains, both
// interleaved, and the libdevice exp/log call rates.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
// Run  : ./fp64_peaks [iters]          (prints one JSON object)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// NF DFMAs per DMMA interleaved: do the two share one pipe?
template <int NACC, int NF>
__global__ void k_mixed(double *out, int iters, double a, double b) {
  double c0[NACC], c1[NACC], f[NACC * NF];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < NACC * NF; i++) f[i] = i + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      dmma(c0[i], c1[i], a, b);
#pragma unroll
      for (int j = 0; j < NF; j++) f[i * NF + j] = fma(f[i * NF + j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NACC * NF; i++) s += f[i];
  if (s == 12345.678) out[0] = s;
}

template <int WHICH>
__global__ void k_func(double *out, int iters, double a) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = -1.0 - 1e-3 * threadIdx.x - i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (WHICH == 0) x[i] = exp(x[i]) - 1.5 + a;        // keeps args in a sane negative range
      else if (WHICH == 1) x[i] = log(x[i] * x[i] + 1.5) + a;
      else x[i] = 1.0 / (x[i] - 2.0) + a;
    }
  }
  double s = x[0] + x[1] + x[2] + x[3];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char **argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, 64));
  const int threads = 512, blocks = sms * 4;          // 64 warps / SM
  const double nthr = (double)threads * blocks;

  double t;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"iters\": %d", p.name, sms, iters);
  t = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 0.999999, 1e-7); }, 5);
  printf(", \"dfma_tflops\": %.3f", nthr * iters * 8 * 2 / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dmma<4><<<blocks, threads>>>(out, iters, 0.999999, 1e-7); }, 5);
  // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
  printf(", \"dmma_tflops\": %.3f", (nthr / 32) * iters * 4 * 512.0 / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 0.999999, 1e-7); }, 5);
  printf(", \"dmma8_tflops\": %.3f", (nthr / 32) * iters * 8 * 512.0 / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_mixed<4, 8><<<blocks, threads>>>(out, iters, 0.999999, 1e-7); }, 5);
  printf(", \"mixed_1dmma_8dfma_tflops\": %.3f", (nthr / 32) * iters * 4 * (512.0 + 8 * 64.0) / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_mixed<4, 2><<<blocks, threads>>>(out, iters, 0.999999, 1e-7); }, 5);
  printf(", \"mixed_1dmma_2dfma_tflops\": %.3f", (nthr / 32) * iters * 4 * (512.0 + 2 * 64.0) / (t * 1e-3) / 1e12);
  int fi = iters / 8;
  t = time_ms([&] { k_func<0><<<blocks, threads>>>(out, fi, 1e-9); }, 5);
  printf(", \"exp_gcalls\": %.2f", nthr * fi * 4 / (t * 1e-3) / 1e9);
  t = time_ms([&] { k_func<1><<<blocks, threads>>>(out, fi, 1e-9); }, 5);
  printf(", \"log_gcalls\": %.2f", nthr * fi * 4 / (t * 1e-3) / 1e9);
  t = time_ms([&] { k_func<2><<<blocks, threads>>>(out, fi, 1e-9); }, 5);
  printf(", \"div_gcalls\": %.2f", nthr * fi * 4 / (t * 1e-3) / 1e9);
  // sustained: 2 s of back-to-back DFMA, then one more measurement
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    float ms = 0;
    do {
      for (int r = 0; r < 10; r++) k_dfma<8><<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
      n += 10;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    } while (ms < 3000.f);
    printf(", \"dfma_tflops_sustained\": %.3f", nthr * iters * 8 * 2 * n / (ms * 1e-3) / 1e12);
  }
  printf("}\n");
  return 0;
}
