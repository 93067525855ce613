// IMMA.16832.S8 issue-rate probe: how many warps per SM sub-partition does the legacy int8 tensor path need,
// with and without operand reuse, to reach its peak?   nvcc -arch=sm_100a -O3 -o imma_probe imma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void imma(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// PATTERN 0: 6 accumulators, one A quad and one B pair reused by all (the old microbench, made unmergeable)
// PATTERN 1: the Gram k-step operand pattern: 2 A quads x 4 B pairs, 6 accumulators, all distinct registers
// PATTERN 2: like 1 plus 12 ALU instructions per step (the LOP3 / SHF traffic of the real loop, no LDS)
template <int PATTERN>
__global__ void k(int *out, unsigned seed) {
  unsigned a0[4], a1[4], b[4][2];
  for (int i = 0; i < 4; ++i) a0[i] = seed * (3 + i) + threadIdx.x, a1[i] = seed * (11 + i) ^ threadIdx.x;
  for (int j = 0; j < 4; ++j) b[j][0] = seed * (17 + j) + threadIdx.x, b[j][1] = seed * (23 + j) - threadIdx.x;
  int c[6][4];
  for (int j = 0; j < 6; ++j) for (int r = 0; r < 4; ++r) c[j][r] = j + r;
  unsigned x = seed, y = threadIdx.x;
#pragma unroll 2
  for (int i = 0; i < ITERS; ++i) {
    if (PATTERN == 0) {
      imma(c[0], a0, b[0]); imma(c[1], a0, b[0]); imma(c[2], a0, b[0]); imma(c[3], a0, b[0]); imma(c[4], a0, b[0]); imma(c[5], a0, b[0]);
    } else {
      if (PATTERN == 2) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          x = __funnelshift_r(x, y, 8) & 0xFF00FFFFu;
          y = (y ^ x) + q;
        }
        a0[1] ^= x & 1, a1[3] ^= y & 1, b[3][0] ^= x & 2, b[3][1] ^= y & 2;
      }
      imma(c[0], a0, b[0]); imma(c[1], a0, b[1]); imma(c[2], a0, b[2]); imma(c[3], a0, b[3]); imma(c[4], a1, b[2]); imma(c[5], a1, b[3]);
    }
  }
  int s = x + y;
  for (int j = 0; j < 6; ++j) for (int r = 0; r < 4; ++r) s += c[j][r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int PATTERN>
static void run(int sms, int warps_per_sm, int *buf, double clk) {
  const int tpb = 32 * warps_per_sm;  // one CTA per SM
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PATTERN><<<sms, tpb>>>(buf, 7u);
  cudaEventRecord(e0);
  k<PATTERN><<<sms, tpb>>>(buf, 7u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double immas = (double)sms * warps_per_sm * ITERS * 6;
  printf("pattern %d warps/SMSP %2d: %.4f IMMA/clk/SMSP (1 per %.1f clk)\n", PATTERN, warps_per_sm / 4,
         immas / (ms * 1e-3) / (sms * 4) / clk, (sms * 4) * clk * (ms * 1e-3) / immas);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  int *buf; cudaMalloc(&buf, 64 << 20);
  for (int w : {4, 8, 16, 24, 32}) { run<0>(p.multiProcessorCount, w, buf, clk); run<1>(p.multiProcessorCount, w, buf, clk); run<2>(p.multiProcessorCount, w, buf, clk); }
  return 0;
}
