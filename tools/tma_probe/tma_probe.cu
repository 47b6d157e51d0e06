// Standalone TMA probe: loads one 4-D box with cp.async.bulk.tensor and checks it against a direct read.
// usage: tma_probe W H P B  bw bh bp  cx cy cz cb  [use_fence=1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n, int cx, int cy, int cz, int cb, int use_fence) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* dst = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((n * 4 + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    if (use_fence) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(&tm), "r"(smem_u32(bar)), "r"(cx), "r"(cy), "r"(cz), "r"(cb) : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
      ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 12) { printf("args\n"); return 2; }
  int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]), B = atoi(argv[4]);
  int bw = atoi(argv[5]), bh = atoi(argv[6]), bp = atoi(argv[7]);
  int cx = atoi(argv[8]), cy = atoi(argv[9]), cz = atoi(argv[10]), cb = atoi(argv[11]);
  int use_fence = argc > 12 ? atoi(argv[12]) : 1;
  size_t total = (size_t)W * H * P * B;
  std::vector<float> h(total);
  for (size_t i = 0; i < total; ++i) h[i] = (float)(i % 100003) + 1.0f;
  float *d, *o;
  int n = bw * bh * bp;
  cudaMalloc(&d, total * 4); cudaMalloc(&o, n * 4);
  cudaMemcpy(d, h.data(), total * 4, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  if (!sym) { printf("no entry point\n"); return 3; }
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * P * 4};
  cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bp, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeFn)sym)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 4; }
  size_t smem = ((n * 4 + 127) / 128) * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(tm, o, n, cx, cy, cz, cb, use_fence);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 5; }
  std::vector<float> got(n);
  cudaMemcpy(got.data(), o, n * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int p = 0; p < bp; ++p) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
    int gx = cx + x, gy = cy + y, gp = cz + p;
    float want = 0.0f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H && gp >= 0 && gp < P && cb >= 0 && cb < B)
      want = h[(((size_t)cb * P + gp) * H + gy) * W + gx];
    if (got[(p * bh + y) * bw + x] != want) ++bad;
  }
  printf("ok bad=%d n=%d\n", bad, n);
  return bad ? 1 : 0;
}
