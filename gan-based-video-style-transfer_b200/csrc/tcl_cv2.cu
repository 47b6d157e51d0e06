// NumPy / OpenCV flavour of the path: the dataset generators of the reference restate warp and the forward-backward
// check with cv2.remap and np.gradient / np.linalg.norm on HWC arrays
// (methods/learning-based/dataset-generation/coco-generation.py:66-113, hollywood2-generation.py:63-111,
// sintel-generation.py:89-130).  Same thresholds as fbcCheckTorch, different numerics:
//   * the sample position is exactly (x+u, y+v), converted to fixed point with 5 fractional bits (cvRound(c*32), half to
//     even); the four weights are products of 1/32 fractions (exact in fp32); the value is v00*w0 + v01*w1 + v10*w2 +
//     v11*w3 left to right with every product and sum rounded (no contraction); taps outside the image read 0
//   * np.gradient: halved central differences inside, one-sided differences on the border rows / columns
//   * np.linalg.norm(.)**2.0 = sqrt of the rounded sum of rounded squares, squared again
// Layout is the reference's own: HWC (interleaved) images and flows.  One pixel per lane; the gathers of neighbouring
// lanes land in neighbouring addresses, so the kernels run at L2 / HBM streaming speed (20 B/px for the fused check,
// 8 + 8C B/px for remap).  No tensor cores: nothing here is a contraction.
// Also here: the Sintel occlusion-PNG -> mask conversion of the dataset reader (utils/sintel_dataset.py:64-65), a table lookup.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tcl_b200.h"

namespace tcl {
void set_last_error(const char* msg);   // tcl_kernels.cu
void count_launch();                    // tcl_kernels.cu: diagnostics counter of tclb200_debug_launch_count
}

namespace {

constexpr float kNearBandCv = 1e-6f;

struct Cv2Taps {
  int x0, y0;
  float w0, w1, w2, w3;
  bool p00, p01, p10, p11;   // tap (row, column) inside the image
};

// cv2.remap's fixed-point conversion of one sample position (x + u, y + v)
__device__ __forceinline__ Cv2Taps cv2_taps(float u, float v, int x, int y, int W, int H) {
  // the reference builds the maps as float64 sums cast to fp32 (coco-generation.py:89-92) = the correctly rounded sum
  const float ax = __fadd_rn(u, (float)x), ay = __fadd_rn(v, (float)y);
  const float sxf = rintf(__fmul_rn(ax, 32.0f)), syf = rintf(__fmul_rn(ay, 32.0f));   // cvRound: half to even
  // far outside / non-finite: cvRound saturates and the short coordinates clamp -- no tap lands inside either way
  const bool ok = fabsf(sxf) < 1073741824.0f && fabsf(syf) < 1073741824.0f;
  const int sx = ok ? (int)sxf : 0, sy = ok ? (int)syf : 0;
  Cv2Taps t;
  t.x0 = sx >> 5; t.y0 = sy >> 5;
  const float fx = __fmul_rn((float)(sx & 31), 0.03125f), fy = __fmul_rn((float)(sy & 31), 0.03125f);
  const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
  t.w0 = __fmul_rn(gy, gx); t.w1 = __fmul_rn(gy, fx); t.w2 = __fmul_rn(fy, gx); t.w3 = __fmul_rn(fy, fx);
  const bool xin0 = (unsigned)t.x0 < (unsigned)W, xin1 = (unsigned)(t.x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)t.y0 < (unsigned)H, yin1 = (unsigned)(t.y0 + 1) < (unsigned)H;
  t.p00 = ok && xin0 && yin0; t.p01 = ok && xin1 && yin0; t.p10 = ok && xin0 && yin1; t.p11 = ok && xin1 && yin1;
  return t;
}

__device__ __forceinline__ float cv2_blend(float v00, float v01, float v10, float v11, const Cv2Taps& t) {
  float acc = __fmul_rn(v00, t.w0);
  acc = __fadd_rn(acc, __fmul_rn(v01, t.w1));
  acc = __fadd_rn(acc, __fmul_rn(v10, t.w2));
  acc = __fadd_rn(acc, __fmul_rn(v11, t.w3));
  return acc;
}

// np.linalg.norm([a, b])**2.0 in fp32
__device__ __forceinline__ float np_sqnorm2(float a, float b) {
  const float s = __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
  return __fmul_rn(s, s);
}

// the rounded sum of rounded squares inside it (what the sqrt-free filter compares)
__device__ __forceinline__ float np_sumsq2(float a, float b) { return __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)); }
// sqrt-then-square moves a term by < 3 * 2^-24 relative; the sums and the 0.01 scaling keep a test's two sides within
// 1e-6 relative of their sqrt-free values: 4e-6 is a safe band (same constant as the hot kernel's filter)
constexpr float kCvFilterEps = 4e-6f;

// cv2.remap(src, x+u, y+v, INTER_LINEAR): src (N,H,W,C), flow (N,H,W,2), out (N,H,W,C)
template <int CT>
__global__ void __launch_bounds__(256) cv2_remap_kernel(const float* __restrict__ src, const float2* __restrict__ flow,
                                                        float* __restrict__ out, int H, int W, int C) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), n = blockIdx.z;
  if (x >= W || y >= H) return;
  const int Cc = CT > 0 ? CT : C;
  const size_t px = ((size_t)n * H + y) * W + x;
  const float2 f = __ldcs(flow + px);
  const Cv2Taps t = cv2_taps(f.x, f.y, x, y, W, H);
  const float* img = src + (size_t)n * H * W * Cc;
  const ptrdiff_t o00 = ((ptrdiff_t)t.y0 * W + t.x0) * Cc, row = (ptrdiff_t)W * Cc;
  float* dst = out + px * Cc;
  auto channel = [&](int c) {
    const float v00 = t.p00 ? __ldg(img + o00 + c) : 0.0f, v01 = t.p01 ? __ldg(img + o00 + Cc + c) : 0.0f;
    const float v10 = t.p10 ? __ldg(img + o00 + row + c) : 0.0f, v11 = t.p11 ? __ldg(img + o00 + row + Cc + c) : 0.0f;
    __stcs(dst + c, cv2_blend(v00, v01, v10, v11, t));
  };
  if (CT > 0) {
#pragma unroll
    for (int c = 0; c < CT; ++c) channel(c);
  } else {
    for (int c = 0; c < Cc; ++c) channel(c);
  }
}

// fb_check(warp_flow(ff, bf), bf)  (prewarped: fb_check(ff, bf) with ff already warped by the caller)
// Measured and dropped (round 2): staging the CTA's 34 x 10 `bf` tile in shared memory so that the centre value and the four
// np.gradient neighbours become shared-memory reads -- 111 Gpix/s against 124 with the five L1-served gathers below (the
// barrier and the clamped ring loads cost more than the L1 hits they replace; the four 8-byte `ff` taps are what is left).
__global__ void __launch_bounds__(256) cv2_fb_check_kernel(const float2* __restrict__ ff, const float2* __restrict__ bf,
                                                           float* __restrict__ mask, int H, int W, int flags, int prewarped,
                                                           unsigned long long* near_threshold) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), n = blockIdx.z;
  unsigned near = 0;
  if (x < W && y < H) {
    const size_t img = (size_t)n * H * W;
    const float2* b = bf + img;
    const size_t o = (size_t)y * W + x;
    const float2 c = __ldg(b + o);
    float2 wf;
    if (prewarped) {
      wf = __ldcs(ff + img + o);
    } else {
      const Cv2Taps t = cv2_taps(c.x, c.y, x, y, W, H);
      const float2* s = ff + img;
      const ptrdiff_t o00 = (ptrdiff_t)t.y0 * W + t.x0;
      const float2 z = make_float2(0.0f, 0.0f);
      const float2 v00 = t.p00 ? __ldg(s + o00) : z, v01 = t.p01 ? __ldg(s + o00 + 1) : z;
      const float2 v10 = t.p10 ? __ldg(s + o00 + W) : z, v11 = t.p11 ? __ldg(s + o00 + W + 1) : z;
      wf.x = cv2_blend(v00.x, v01.x, v10.x, v11.x, t);
      wf.y = cv2_blend(v00.y, v01.y, v10.y, v11.y, t);
    }
    // np.linalg.norm(.)**2.0 is sqrt-then-square of a rounded sum of squares: each such term is within a few ulp of the plain
    // sum.  The plain (sqrt-free) evaluation decides a test whenever lhs lies outside rhs * (1 +- kCvFilterEps); only the
    // remaining pixels (a few per million) and calls that want the near-threshold count evaluate the exact sequence.
    const bool exact_all = near_threshold != nullptr;
    const float sb = np_sumsq2(c.x, c.y);
    bool keep = true;
    if (flags & TCLB200_OCC) {
      const float swb = np_sumsq2(__fadd_rn(wf.x, c.x), __fadd_rn(wf.y, c.y)), sw = np_sumsq2(wf.x, wf.y);
      const float r = __fadd_rn(__fmul_rn(0.01f, __fadd_rn(sw, sb)), 0.5f);
      bool occ = swb > r * (1.0f + kCvFilterEps);
      if (exact_all || (!occ && !(swb < r * (1.0f - kCvFilterEps)))) {
        const float norm_wb = np_sqnorm2(__fadd_rn(wf.x, c.x), __fadd_rn(wf.y, c.y));
        const float rhs = __fadd_rn(__fmul_rn(0.01f, __fadd_rn(np_sqnorm2(wf.x, wf.y), np_sqnorm2(c.x, c.y))), 0.5f);
        occ = norm_wb > rhs;
        near += fabsf(__fsub_rn(norm_wb, rhs)) < kNearBandCv;
      }
      if (occ) keep = false;
    }
    if (flags & TCLB200_MOB) {
      // np.gradient: one-sided on the border, halved central difference inside (H, W >= 2 checked by the host)
      const float2 l = __ldg(b + o - (x > 0 ? 1 : 0)), r = __ldg(b + o + (x + 1 < W ? 1 : 0));
      const float2 up = __ldg(b + o - (y > 0 ? (size_t)W : 0)), dn = __ldg(b + o + (y + 1 < H ? (size_t)W : 0));
      const bool xe = x == 0 || x + 1 == W, ye = y == 0 || y + 1 == H;
      float ux = __fsub_rn(r.x, l.x), vx = __fsub_rn(r.y, l.y), uy = __fsub_rn(dn.x, up.x), vy = __fsub_rn(dn.y, up.y);
      if (!xe) { ux = __fmul_rn(ux, 0.5f); vx = __fmul_rn(vx, 0.5f); }
      if (!ye) { uy = __fmul_rn(uy, 0.5f); vy = __fmul_rn(vy, 0.5f); }
      const float lq = __fadd_rn(np_sumsq2(uy, ux), np_sumsq2(vy, vx));
      const float rq = __fadd_rn(__fmul_rn(0.01f, sb), 0.002f);
      bool mob = lq > rq * (1.0f + kCvFilterEps);
      if (exact_all || (!mob && !(lq < rq * (1.0f - kCvFilterEps)))) {
        const float lhs = __fadd_rn(np_sqnorm2(uy, ux), np_sqnorm2(vy, vx));
        const float rhs = __fadd_rn(__fmul_rn(0.01f, np_sqnorm2(c.x, c.y)), 0.002f);
        mob = lhs > rhs;
        near += fabsf(__fsub_rn(lhs, rhs)) < kNearBandCv;
      }
      if (mob) keep = false;
    }
    __stcs(mask + img + o, keep ? 1.0f : 0.0f);
  }
  if (near_threshold) {
    near = __reduce_add_sync(0xffffffffu, near);
    if ((threadIdx.x & 31) == 0 && near) atomicAdd(near_threshold, (unsigned long long)near);
  }
}

// Sintel ground-truth occlusion PNG -> the mask the path consumes (utils/sintel_dataset.py:64-65):
//   mask = io.imread(png) / 255.0 ; mask = 1.0 - mask   (float64) ; later .float()
// 256 possible inputs: the table is computed on the host with exactly that float64 arithmetic, the kernel is a lookup.
__constant__ float c_occ_lut[256];

__global__ void __launch_bounds__(256) occlusion_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n4, size_t n) {
  __shared__ float lut[256];   // (a warp's 128 lookups hit arbitrary entries: shared memory, not the constant cache)
  lut[threadIdx.x] = c_occ_lut[threadIdx.x];
  __syncthreads();
  // four pixels per lane and step (one 32-bit load, one 128-bit store, both fully coalesced), four independent steps in flight
  constexpr int K = 4;
  for (size_t i0 = (size_t)blockIdx.x * (256 * K) + threadIdx.x; i0 < n4; i0 += (size_t)gridDim.x * (256 * K)) {
    uchar4 v[K];
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (i0 + k * 256 < n4) v[k] = __ldcs(reinterpret_cast<const uchar4*>(src) + i0 + k * 256);
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (i0 + k * 256 < n4)
        __stcs(reinterpret_cast<float4*>(dst) + i0 + k * 256, make_float4(lut[v[k].x], lut[v[k].y], lut[v[k].z], lut[v[k].w]));
  }
  if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n - 4 * n4)) {   // the 0..3 trailing pixels (or a tiny unaligned buffer)
    const size_t j = 4 * n4 + threadIdx.x;
    dst[j] = lut[src[j]];
  }
}

int cfail(int code, const char* msg, const char* detail = "") {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s%s", msg, detail);
  tcl::set_last_error(buf);
  return code;
}

int check_dims(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return cfail(TCLB200_ERR_INVALID, "N, H, W must be positive");
  if (H >= 32768 || W >= 32768) return cfail(TCLB200_ERR_UNSUPPORTED, "cv2.remap addresses images below 32768 x 32768");
  if (N > 65535 || (H + 7) / 8 > 65535) return cfail(TCLB200_ERR_UNSUPPORTED, "too many images / rows for one launch");
  return TCLB200_OK;
}

}  // namespace

extern "C" int tclb200_cv2_remap(const float* src, const float* flow, float* out, int N, int H, int W, int C, tclb200_stream_t stream) {
  if (!src || !flow || !out) return cfail(TCLB200_ERR_INVALID, "src, flow and out are required");
  if (C <= 0) return cfail(TCLB200_ERR_INVALID, "C must be positive");
  if (const int rc = check_dims(N, H, W)) return rc;
  if ((reinterpret_cast<uintptr_t>(flow) & 7u) != 0) return cfail(TCLB200_ERR_INVALID, "flow must be 8-byte aligned");
  const dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)N);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float2* f2 = reinterpret_cast<const float2*>(flow);
  if (C == 3) cv2_remap_kernel<3><<<grid, 256, 0, s>>>(src, f2, out, H, W, C);
  else if (C == 2) cv2_remap_kernel<2><<<grid, 256, 0, s>>>(src, f2, out, H, W, C);
  else if (C == 1) cv2_remap_kernel<1><<<grid, 256, 0, s>>>(src, f2, out, H, W, C);
  else cv2_remap_kernel<0><<<grid, 256, 0, s>>>(src, f2, out, H, W, C);
  tcl::count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cfail(TCLB200_ERR_CUDA, "cv2_remap launch: ", cudaGetErrorString(e));
  return TCLB200_OK;
}

extern "C" int tclb200_cv2_fb_check(const float* ff, const float* bf, float* mask, int N, int H, int W, int flags, int prewarped,
                                    unsigned long long* near_threshold, tclb200_stream_t stream) {
  if (!ff || !bf || !mask) return cfail(TCLB200_ERR_INVALID, "ff, bf and mask are required");
  if (!(flags & (TCLB200_OCC | TCLB200_MOB))) return cfail(TCLB200_ERR_INVALID, "flags must request TCLB200_OCC and/or TCLB200_MOB");
  if (const int rc = check_dims(N, H, W)) return rc;
  if ((flags & TCLB200_MOB) && (H < 2 || W < 2)) return cfail(TCLB200_ERR_INVALID, "np.gradient needs at least 2 rows and 2 columns");
  if (((reinterpret_cast<uintptr_t>(ff) | reinterpret_cast<uintptr_t>(bf)) & 7u) != 0) return cfail(TCLB200_ERR_INVALID, "flows must be 8-byte aligned");
  const dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)N);
  cv2_fb_check_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(ff), reinterpret_cast<const float2*>(bf), mask, H, W, flags & (TCLB200_OCC | TCLB200_MOB),
      prewarped, near_threshold);
  tcl::count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cfail(TCLB200_ERR_CUDA, "cv2_fb_check launch: ", cudaGetErrorString(e));
  return TCLB200_OK;
}

extern "C" int tclb200_occlusion_u8_to_mask(const uint8_t* src, float* dst, size_t n, tclb200_stream_t stream) {
  if (!src || !dst) return cfail(TCLB200_ERR_INVALID, "src and dst are required");
  if (n == 0) return cfail(TCLB200_ERR_INVALID, "n must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = (float)(1.0 - (double)v / 255.0);
  // per call and stream-ordered (64 launches' worth of bytes; no process-global "initialised" flag to go stale across devices)
  cudaError_t e = cudaMemcpyToSymbolAsync(c_occ_lut, lut, sizeof(lut), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cfail(TCLB200_ERR_CUDA, "occlusion table upload: ", cudaGetErrorString(e));
  const bool vec = ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  const size_t n4 = vec ? n / 4 : 0;
  if (n - 4 * n4 > 256)   // (the scalar tail covers at most one CTA's worth of pixels)
    return cfail(TCLB200_ERR_UNSUPPORTED, "src must be 4-byte and dst 16-byte aligned (or n <= 256)");
  const size_t want = n4 ? (n4 + 1023) / 1024 : 1, cap = 148 * 8 * 4;   // (a CTA covers 256 x 4 steps x 4 pixels per pass) grid-stride beyond a few waves
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  occlusion_u8_kernel<<<grid, 256, 0, s>>>(src, dst, n4, n);
  tcl::count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return cfail(TCLB200_ERR_CUDA, "occlusion_u8 launch: ", cudaGetErrorString(e));
  return TCLB200_OK;
}
