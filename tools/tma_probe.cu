// Stand-alone probe for the TMA staging path: loads a (72 x 35) u16 box at a negative origin from a
// 2-D tensor through a descriptor that lives in GLOBAL memory (as the engine does) and checks the
// zero fill.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool PARAM>
__global__ void probe(const uint8_t *tmaps_g, const __grid_constant__ CUtensorMap pm, int x, int y, uint16_t *out, int mode, int bytes) {
  const void *tmaps = PARAM ? (const void *)&pm : (const void *)tmaps_g;
  __shared__ __align__(128) uint16_t tile[16384];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (mode == 1) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmaps) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(tile)),
        "l"(tmaps), "r"(x), "r"(y), "r"(smem_u32(&bar))
        : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(
          smem_u32(&bar)),
      "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
  const int W = 256, H = 192;
  const int BW = argc > 1 ? atoi(argv[1]) : 72, BH = argc > 2 ? atoi(argv[2]) : 35, X = argc > 3 ? atoi(argv[3]) : 60, Y = argc > 4 ? atoi(argv[4]) : 29, SYNC = argc > 5 ? atoi(argv[5]) : 0;
  std::vector<uint16_t> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (uint16_t)(i % 1021 + 1);
  uint16_t *d, *out;
  cudaMalloc(&d, W * H * 2);
  cudaMalloc(&out, 32768);
  cudaMemcpy(d, h.data(), W * H * 2, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m;
  const cuuint64_t dims[2] = {W, H}, strides[1] = {W * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d entry=%p q=%d sizeof=%zu\n", (int)r, fn, (int)q, sizeof(m));
  uint8_t *dm;
  cudaMalloc(&dm, 128 * 6);
  cudaMemcpy(dm + 128, &m, 128, cudaMemcpyHostToDevice);
  for (int mode = 2; mode >= 0; --mode) {
    if (mode == 2) probe<true><<<1, 128>>>(dm + 128, m, X, Y, out, SYNC, BW*BH*2); else probe<false><<<1, 128>>>(dm + 128, m, X, Y, out, mode, BW*BH*2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d interior: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    if (mode == 2) probe<true><<<1, 128>>>(dm + 128, m, X, Y, out, 0, BW*BH*2); else probe<false><<<1, 128>>>(dm + 128, m, X, Y, out, mode, BW*BH*2);
    e = cudaDeviceSynchronize();
    printf("mode %d negative origin: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> o(BW * BH);
    cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int ty = 0; ty < BH; ++ty)
      for (int tx = 0; tx < BW; ++tx) {
        const int yy = ty + Y, xx = tx + X;
        const uint16_t want = (yy < 0 || xx < 0 || yy >= H || xx >= W) ? 0 : h[yy * W + xx];
        bad += o[ty * BW + tx] != want;
      }
    printf("mode %d mismatches: %d\n", mode, bad);
  }
  return 0;
}
