// Accuracy of the kernels' fast reciprocal (MUFU.RCP64H seed + one cubic correction; vbhem_lattice.cu, vbhmm_fb.cu)
// against the IEEE division, over 2^26 inputs spread over [2^-900, 2^10].  Prints the worst relative error of the seed
// and of the corrected value.   nvcc -arch=sm_100a -O3 -o _build/rcp_check csrc/rcp_check.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double seed(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double rcp_pos(double x) {
  const double r = seed(x), e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
__global__ void k(unsigned long long n, double *out) {
  double w0 = 0.0, w1 = 0.0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long h = i * 0x9E3779B97F4A7C15ull;                  // mantissa bits
    const int e = -900 + (int)(i % 911);                                      // exponent -900 .. 10
    const double x = ldexp(1.0 + (double)(h >> 12) * 0x1p-52, e);
    const double t = 1.0 / x;
    w0 = fmax(w0, fabs(seed(x) - t) / t);
    w1 = fmax(w1, fabs(rcp_pos(x) - t) / t);
  }
  for (int o = 16; o; o >>= 1) { w0 = fmax(w0, __shfl_xor_sync(~0u, w0, o)); w1 = fmax(w1, __shfl_xor_sync(~0u, w1, o)); }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long *)out, (unsigned long long)__double_as_longlong(w0));
    atomicMax((unsigned long long *)out + 1, (unsigned long long)__double_as_longlong(w1));
  }
}
int main() {
  double *d, h[2] = {0, 0};
  cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  k<<<592, 256>>>(1ull << 26, d);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("{\"inputs\": %llu, \"seed_max_rel\": %.3e, \"rcp_pos_max_rel\": %.3e, \"ulp\": %.3e}\n", 1ull << 26, h[0], h[1], 0x1p-53);
  return cudaGetLastError() != cudaSuccess;
}
