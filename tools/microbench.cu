// Instruction-throughput probes used to choose the Gram-kernel design:
// per-SM rates of IMAD, DP4A, legacy IMMA (mma.sync m16n8k32 s8), and the FP64 pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096

__global__ void k_imad(int *out, int a, int b) {
  int x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 1, x5 = 2, x6 = 3, x7 = 4;
  for (int i = 0; i < ITERS; ++i) {
    x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
    x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dp4a(int *out, int a, int b) {
  int x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 1, x5 = 2, x6 = 3, x7 = 4;
  for (int i = 0; i < ITERS; ++i) {
    x0 = __dp4a(a, b, x0); x1 = __dp4a(a, b, x1); x2 = __dp4a(a, b, x2); x3 = __dp4a(a, b, x3);
    x4 = __dp4a(a, b, x4); x5 = __dp4a(a, b, x5); x6 = __dp4a(a, b, x6); x7 = __dp4a(a, b, x7);
    a ^= x0 & 1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void imma(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void k_imma(int *out, unsigned seed) {
  unsigned a[4] = {seed + threadIdx.x, seed * 3, seed * 5, seed * 7};
  unsigned b[2] = {seed * 11, seed * 13 + threadIdx.x};
  int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
  int c4[4] = {0, 0, 0, 0}, c5[4] = {0, 0, 0, 0};
  for (int i = 0; i < ITERS; ++i) {
    imma(c0, a, b); imma(c1, a, b); imma(c2, a, b); imma(c3, a, b); imma(c4, a, b); imma(c5, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[2] + c3[3] + c4[0] + c5[1];
}

template <int OP>
__global__ void k_f64(double *out, double a, double b) {
  double x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 1, x5 = 2, x6 = 3, x7 = 4;
  for (int i = 0; i < ITERS; ++i) {
    if (OP == 0) {
      x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a);
      x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a);
    } else if (OP == 1) {
      x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a);
      x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a);
    } else {
      x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// single dependent chain per thread: latency-bound FP64 (what one flat-block thread looks like)
__global__ void k_f64_chain(double *out, double a) {
  double x = threadIdx.x;
  for (int i = 0; i < ITERS * 8; ++i) x = __dadd_rn(x, a);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk_khz);
  void *buf;
  cudaMalloc(&buf, 64 << 20);
  const int grid = sms * 8, tpb = 256;
  const double nthr = (double)grid * tpb;
  for (int warps_tpb : {256, 1024}) {
    const int g2 = sms * (2048 / warps_tpb);
    float ms = time_ms([&] { k_imad<<<g2, warps_tpb>>>((int *)buf, 3, 5); });
    printf(", \"imad_per_clk_sm_tpb%d\": %.1f", warps_tpb,
           (double)g2 * warps_tpb * ITERS * 8 / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    float ms = time_ms([&] { k_imad<<<grid, tpb>>>((int *)buf, 3, 5); });
    printf(", \"imad_Gops\": %.1f", nthr * ITERS * 8 / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_dp4a<<<grid, tpb>>>((int *)buf, 0x01020304, 0x05060708); });
    printf(", \"dp4a_Ginstr\": %.1f", nthr * ITERS * 8 / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_imma<<<grid, tpb>>>((int *)buf, 7u); });
    const double immas = (double)grid * (tpb / 32) * ITERS * 6;
    printf(", \"imma_m16n8k32_Ginstr\": %.2f, \"imma_TMAC\": %.1f", immas / (ms * 1e-3) / 1e9,
           immas * 4096 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_f64<0><<<grid, tpb>>>((double *)buf, 1.5, 0.5); });
    printf(", \"dadd_Gops\": %.1f", nthr * ITERS * 8 / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_f64<1><<<grid, tpb>>>((double *)buf, 1.0000001, 0.5); });
    printf(", \"dmul_Gops\": %.1f", nthr * ITERS * 8 / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_f64<2><<<grid, tpb>>>((double *)buf, 1.0000001, 0.5); });
    printf(", \"dfma_Gops\": %.1f", nthr * ITERS * 8 / (ms * 1e-3) / 1e9);
    for (int t : {128, 256, 512, 1024}) {
      const int g2 = sms * (t <= 256 ? 2 : 1);
      ms = time_ms([&] { k_f64_chain<<<g2, t>>>((double *)buf, 1.5); });
      printf(", \"dadd_chain_Gops_%dthr_x%d\": %.1f", t, g2 / sms, (double)g2 * t * ITERS * 8 / (ms * 1e-3) / 1e9);
    }
  }
  printf("}\n");
  return 0;
}
