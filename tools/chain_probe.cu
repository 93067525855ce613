// Probe: cost per element of a sequential (order-dependent) f64 sum in one thread, as latest_kernel runs them.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/chain_probe tools/chain_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

template <int MODE>
__global__ void probe(double *out, long long *clk, int n, double a) {
  extern __shared__ double buf[];
  for (int i = threadIdx.x; i < n + 64; i += blockDim.x) buf[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 5 && lane < (MODE == 4 ? 32 : 1)) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(buf);
    const long long t0 = clock64();
    double acc = 0.0, acc2 = 0.0;
    if (MODE == 0) {  // registers only
      for (int p = 0; p < n; p += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc = __dadd_rn(acc, a);
      }
    } else if (MODE == 1 || MODE == 4) {  // loads one step ahead
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = lds_f64(base + 8 * q);
#pragma unroll 1
      for (int p = 0; p < n; p += 8) {
        double w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = lds_f64(base + 8 * (p + 8 + q));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc = __dadd_rn(acc, v[q]);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = w[q];
      }
    } else if (MODE == 2) {  // two independent chains
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = lds_f64(base + 8 * q);
#pragma unroll 1
      for (int p = 0; p < n; p += 8) {
        double w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = lds_f64(base + 8 * (p + 8 + q));
#pragma unroll
        for (int q = 0; q < 8; ++q) acc = __dadd_rn(acc, v[q]), acc2 = __dadd_rn(acc2, __dmul_rn(v[q], a));
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = w[q];
      }
    } else if (MODE == 3) {  // 16 per step
      double v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = lds_f64(base + 8 * q);
#pragma unroll 1
      for (int p = 0; p < n; p += 16) {
        double w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = lds_f64(base + 8 * (p + 16 + q));
#pragma unroll
        for (int q = 0; q < 16; ++q) acc = __dadd_rn(acc, v[q]);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = w[q];
      }
    }
    const long long t1 = clock64();
    if (lane == 0) clk[0] = t1 - t0;
    out[lane] = acc + acc2;
  }
  __syncthreads();
}

int main() {
  double *out;
  long long *clk, h;
  cudaMalloc(&out, 256);
  cudaMalloc(&clk, 8);
  const int n = 8000;
  const char *names[] = {"registers only", "loads one step ahead", "two chains", "16 per step", "32 lanes in lockstep"};
  auto run = [&](auto kern, int mode) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) kern<<<1, 512, (n + 64) * 8>>>(out, clk, n, 1.0000001);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-24s %6.2f clk per element (%lld clk)\n", names[mode], (double)h / n, h);
  };
  run(probe<0>, 0);
  run(probe<1>, 1);
  run(probe<2>, 2);
  run(probe<3>, 3);
  run(probe<4>, 4);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
