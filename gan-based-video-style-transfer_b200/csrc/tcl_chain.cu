// Learning-based trainers' warp chains -- two fused helpers on fp32 NCHW tensors:
//
//   tclb200_reconet_loss    ReCoNet's output-level temporal loss (methods/learning-based/fs_reconet.py:63-69):
//                             output_term = styled2 - warp(styled1, flow)
//                             input_term  = luminance(img2 - warp(img1, flow))          (0.2126, 0.7152, 0.0722)
//                             loss        = mean((mask * (output_term - input_term))**2)
//                           with `warp` = fs_lib.warp (bilinear taps times the binarised warp of an all-ones image,
//                           fs_lib.py:5-39).  The two warps share one flow: the sampling position, the four weights and
//                           the validity factor are computed once per pixel and serve all six planes; one pass instead
//                           of 2 warps (2 x 2 grid_sample) + 8 elementwise ops + a reduction.  Also writes the input term
//                           (B,1,H,W), which the backward needs.
//   tclb200_ruder_input     one step of Ruder's recurrent chain (fs_ruder.py:50-75): the network input
//                             cat((img, mask, warp(styled_prev, flow)), 1)                (B,7,H,W)
//                           written in one pass (the warped frame goes straight into channels 4..6, no warp tensor and no
//                           cat pass); optionally also the warped frame on its own (the `loss_warped` of fs_ruder.py:97).
//
// One lane per target pixel, taps straight from global memory (L2 serves the 4-fold reuse), exact rounding sequences of
// tcl_math.cuh; HBM-bound streaming + gather, no tensor cores (nothing here is a contraction).  Algorithmic bytes per
// pixel: reconet 60 read (+4 written), ruder 36 read + 28 written (+12).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tcl_b200.h"
#include "tcl_common.cuh"
#include "tcl_math.cuh"

namespace tcl {
void set_last_error(const char* msg);
void count_launch();

namespace {

struct ChainParams {
  const float* flow;      // (B,2,H,W)
  const float* mask;      // (B,1,H,W) or nullptr (= ones)
  const float* a_prev;    // reconet: styled1      ruder: styled_prev     (B,3,H,W)
  const float* a_cur;     // reconet: styled2      ruder: img
  const float* b_prev;    // reconet: img1
  const float* b_cur;     // reconet: img2
  float* lum_out;         // reconet: input term (B,1,H,W) or nullptr
  float* cat_out;         // ruder: (B,7,H,W)
  float* warp_out;        // ruder: (B,3,H,W) or nullptr
  double* partials;       // reconet: one partial sum per CTA
  Geo geo;
  int B;
};

// fs_lib.warp of one plane given the pixel's taps and validity factor
__device__ __forceinline__ float vwarp(const float* plane, const Taps& t, int W, float valid) {
  return __fmul_rn(sample_global(plane, t, W, kV), valid);
}

__global__ void __launch_bounds__(256) reconet_forward_kernel(const ChainParams p) {
  __shared__ double red[kWarps];
  const Geo& g = p.geo;
  const int W = g.W, H = g.H;
  const size_t plane = (size_t)H * W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int b = blockIdx.z, x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + wrp;
  float err = 0.0f;
  if (x < W && y < H) {
    const size_t o = (size_t)y * W + x;
    const float u = __ldg(p.flow + (size_t)b * 2 * plane + o), v = __ldg(p.flow + ((size_t)b * 2 + 1) * plane + o);
    const Taps t = make_taps(u, v, x, y, g, kV);
    const float valid = binarise_validity(ones_sample(t, kV));   // fs_lib.py:29-37
    const float m = p.mask ? __ldcs(p.mask + (size_t)b * plane + o) : 1.0f;
    float out[3], in[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t base = ((size_t)b * 3 + c) * plane;
      out[c] = __fsub_rn(__ldcs(p.a_cur + base + o), vwarp(p.a_prev + base, t, W, valid));   // styled2 - warp(styled1)
      in[c] = __fsub_rn(__ldcs(p.b_cur + base + o), vwarp(p.b_prev + base, t, W, valid));    // img2 - warp(img1)
    }
    // 0.2126*r + 0.7152*g + 0.0722*b: three rounded products summed left to right (fs_reconet.py:65)
    const float lum = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, in[0]), __fmul_rn(0.7152f, in[1])), __fmul_rn(0.0722f, in[2]));
    if (p.lum_out) __stcs(p.lum_out + (size_t)b * plane + o, lum);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float md = __fmul_rn(m, __fsub_rn(out[c], lum));
      err = __fmaf_rn(md, md, err);
    }
  }
  const double s = block_sum((double)err, red);
  if (threadIdx.x == 0) p.partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

// partial sums of the CTAs in index order (deterministic), times inv_count
__global__ void __launch_bounds__(256) chain_fold_kernel(const double* partials, size_t n, double inv_count, float* loss_out, double* sum_out) {
  __shared__ double red[kWarps];
  double s = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 256) s += partials[i];
  const double S = block_sum(s, red);
  if (threadIdx.x == 0) {
    if (sum_out) *sum_out = S;
    if (loss_out) *loss_out = (float)(S * inv_count);
  }
}

__global__ void __launch_bounds__(256) ruder_input_kernel(const ChainParams p) {
  const Geo& g = p.geo;
  const int W = g.W, H = g.H;
  const size_t plane = (size_t)H * W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int b = blockIdx.z, x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + wrp;
  if (x >= W || y >= H) return;
  const size_t o = (size_t)y * W + x;
  const float u = __ldg(p.flow + (size_t)b * 2 * plane + o), v = __ldg(p.flow + ((size_t)b * 2 + 1) * plane + o);
  const Taps t = make_taps(u, v, x, y, g, kV);
  const float valid = binarise_validity(ones_sample(t, kV));
  float* cat = p.cat_out + (size_t)b * 7 * plane + o;
#pragma unroll
  for (int c = 0; c < 3; ++c) __stcs(cat + (size_t)c * plane, __ldcs(p.a_cur + ((size_t)b * 3 + c) * plane + o));   // img
  __stcs(cat + 3 * plane, p.mask ? __ldcs(p.mask + (size_t)b * plane + o) : 1.0f);                                   // mask
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float w = vwarp(p.a_prev + ((size_t)b * 3 + c) * plane, t, W, valid);                                      // warp(styled_prev, flow)
    __stcs(cat + (size_t)(4 + c) * plane, w);
    if (p.warp_out) __stcs(p.warp_out + ((size_t)b * 3 + c) * plane + o, w);
  }
}

int cfail(int code, const char* msg) {
  set_last_error(msg);
  return code;
}

}  // namespace
}  // namespace tcl

using namespace tcl;

extern "C" size_t tclb200_reconet_scratch_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return sizeof(double) * (size_t)B * ((size_t)(H + 7) / 8) * ((size_t)(W + 31) / 32);
}

extern "C" int tclb200_reconet_loss(const float* flow, const float* mask, const float* styled1, const float* styled2, const float* img1,
                                    const float* img2, float* lum_out, float* loss_out, double* sum_out, void* scratch, size_t scratch_bytes,
                                    int B, int H, int W, tclb200_stream_t stream) {
  if (!flow || !styled1 || !styled2 || !img1 || !img2) return cfail(TCLB200_ERR_INVALID, "flow, styled1, styled2, img1 and img2 are required");
  if (!loss_out && !sum_out) return cfail(TCLB200_ERR_INVALID, "nothing to return: loss_out and sum_out are both NULL");
  if (B <= 0 || H <= 0 || W <= 0) return cfail(TCLB200_ERR_INVALID, "B, H, W must be positive");
  if ((size_t)H * W >= (1u << 30) || B > 65535 || (H + 7) / 8 > 65535) return cfail(TCLB200_ERR_UNSUPPORTED, "shape too large for one launch");
  if (!scratch || scratch_bytes < tclb200_reconet_scratch_bytes(B, H, W)) return cfail(TCLB200_ERR_INVALID, "scratch missing or smaller than tclb200_reconet_scratch_bytes(B,H,W)");
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.flow = flow; p.mask = mask; p.a_prev = styled1; p.a_cur = styled2; p.b_prev = img1; p.b_cur = img2;
  p.lum_out = lum_out; p.partials = reinterpret_cast<double*>(scratch);
  p.geo = make_geo(H, W); p.B = B;
  const dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)B);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  reconet_forward_kernel<<<grid, 256, 0, s>>>(p);
  count_launch();
  chain_fold_kernel<<<1, 256, 0, s>>>(p.partials, (size_t)grid.x * grid.y * grid.z, 1.0 / ((double)B * 3.0 * H * W), loss_out, sum_out);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char msg[256];
    snprintf(msg, sizeof(msg), "reconet loss launch: %s", cudaGetErrorString(e));
    return cfail(TCLB200_ERR_CUDA, msg);
  }
  return TCLB200_OK;
}

extern "C" int tclb200_ruder_input(const float* img, const float* mask, const float* styled_prev, const float* flow, float* cat_out,
                                   float* warp_out, int B, int H, int W, tclb200_stream_t stream) {
  if (!img || !styled_prev || !flow || !cat_out) return cfail(TCLB200_ERR_INVALID, "img, styled_prev, flow and cat_out are required");
  if (B <= 0 || H <= 0 || W <= 0) return cfail(TCLB200_ERR_INVALID, "B, H, W must be positive");
  if ((size_t)H * W >= (1u << 30) || B > 65535 || (H + 7) / 8 > 65535) return cfail(TCLB200_ERR_UNSUPPORTED, "shape too large for one launch");
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.flow = flow; p.mask = mask; p.a_prev = styled_prev; p.a_cur = img; p.cat_out = cat_out; p.warp_out = warp_out;
  p.geo = make_geo(H, W); p.B = B;
  const dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)B);
  ruder_input_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char msg[256];
    snprintf(msg, sizeof(msg), "ruder input launch: %s", cudaGetErrorString(e));
    return cfail(TCLB200_ERR_CUDA, msg);
  }
  return TCLB200_OK;
}
