// Hand-written sm_100a kernels + C ABI for the flow-based temporal-consistency path
// (backward warp, forward-backward occlusion / motion-boundary mask, masked temporal error).
//
// Boundary: include/tcl_b200.h.  Reference semantics: utils/flowtools.py:12-58,
// methods/learning-based/fs_lib.py:5-39, utils/sintel_eval.py:104-110, utils/metrics/eval.py:137-138,
// methods/GAN-based/StarGANv2AdvCon/core/solver.py:427-446 of the upstream repository.
//
// Layout in HBM: planar NCHW, fp32 flows/masks, fp32 or bf16 frames.  A CTA owns a TW x TH tile of
// one frame pair; a warp owns one tile row and each lane PX consecutive pixels, so every streaming
// access (bf, cur, mask, outputs) is one 16-byte vector per lane and 512 contiguous bytes per warp.
// The data-dependent bilinear taps of `ff` and `prev` go through the read-only L1/L2 path.
// The masked error is reduced lane -> warp (shuffles) -> CTA (smem) -> pair -> batch with
// self-resetting tickets, in a fixed order (deterministic), inside the same launch.
// HBM-bound: no tensor cores, nothing here is a contraction.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tcl_b200.h"
#include "tcl_math.cuh"

namespace tcl {

constexpr int kV = V_ATEN_CUDA;  // arithmetic flavour of the product kernels (see tcl_math.cuh)
constexpr int kWarps = 8;        // warps per CTA = tile rows
constexpr int kThreads = 32 * kWarps;
constexpr float kNearBand = 1e-6f;

// ---------------------------------------------------------------------------------------------
// vector I/O: N consecutive elements of T <-> fp32 registers, in the widest aligned chunks
// ---------------------------------------------------------------------------------------------
enum class Ld { Default, Stream };

template <int BYTES> struct Chunk;
template <> struct Chunk<16> { using type = uint4; };
template <> struct Chunk<8> { using type = uint2; };
template <> struct Chunk<4> { using type = uint32_t; };
template <> struct Chunk<2> { using type = uint16_t; };

template <typename T, int N>
struct Raw {
  static constexpr int kBytes = N * (int)sizeof(T);
  static constexpr int kChunk = (kBytes % 16 == 0) ? 16 : (kBytes % 8 == 0) ? 8 : (kBytes % 4 == 0) ? 4 : 2;
  using chunk_t = typename Chunk<kChunk>::type;
  static constexpr int kCount = kBytes / kChunk;
  union {
    chunk_t c[kCount];
    T e[N];
  };
};

template <typename T, int N, Ld MODE = Ld::Default>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&out)[N]) {
  Raw<T, N> r;
  using chunk_t = typename Raw<T, N>::chunk_t;
  const chunk_t* q = reinterpret_cast<const chunk_t*>(p);
#pragma unroll
  for (int i = 0; i < Raw<T, N>::kCount; ++i) r.c[i] = (MODE == Ld::Stream) ? __ldcs(q + i) : __ldg(q + i);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = to_f32(r.e[i]);
}

__device__ __forceinline__ void from_f32(float v, float& o) { o = v; }
__device__ __forceinline__ void from_f32(float v, __nv_bfloat16& o) { o = __float2bfloat16_rn(v); }

template <typename T, int N>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&v)[N]) {
  Raw<T, N> r;
  using chunk_t = typename Raw<T, N>::chunk_t;
#pragma unroll
  for (int i = 0; i < N; ++i) from_f32(v[i], r.e[i]);
  chunk_t* q = reinterpret_cast<chunk_t*>(p);
#pragma unroll
  for (int i = 0; i < Raw<T, N>::kCount; ++i) __stcs(q + i, r.c[i]);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fixed-order CTA sum of per-thread doubles; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* smem /*[kWarps]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kWarps; ++i) s += smem[i];
  }
  return s;
}

struct Scratch {
  double* partials;        // [B * tiles_per_pair]
  unsigned* pair_ticket;   // [B]
  unsigned* batch_ticket;  // [1]
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// fused forward
// ---------------------------------------------------------------------------------------------
struct FwdParams {
  const float* ff;
  const float* bf;
  const float* mask_in;
  const void* prev;
  const void* cur;
  void* warp_out;
  float* mask_out;
  void* blend_out;
  double* pair_sums;
  double* total_sums;
  float* pair_vals;
  float* total_val;
  unsigned long long* near_threshold;
  Scratch scratch;
  Geo geo;
  int B, C;
  int tiles_x, tiles_per_pair;
  int flags, loss, finalize;
  double inv_count;  // 1/(C*H*W)
};

enum : int { MASK_NONE = 0, MASK_GIVEN = 1, MASK_COMPUTED = 2 };

__device__ __forceinline__ float finalise_value(double mean, int finalize) {
  return (float)(finalize == TCLB200_FIN_RMSE ? sqrt(mean) : mean);
}

// FrameT: float / __nv_bfloat16.  PX: pixels per lane (vector width).  MASK: where the mask comes from.
// REDUCE: accumulate the masked error.  CT: compile-time channel count (0 = runtime loop).
template <typename FrameT, int PX, int MASK, bool REDUCE, int CT>
__global__ void __launch_bounds__(kThreads) fused_forward_kernel(const FwdParams p) {
  const Geo& g = p.geo;
  const int W = g.W, H = g.H;
  const size_t plane = (size_t)H * W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const unsigned bid = blockIdx.x;
  const int pair = bid / p.tiles_per_pair;
  const int tile = bid - pair * p.tiles_per_pair;
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int x = (tx * 32 + lane) * PX;
  const int y = ty * kWarps + wrp;
  const bool active = (x < W) && (y < H);  // W % PX == 0 is guaranteed by the launcher
  const int C = CT > 0 ? CT : p.C;

  float err = 0.0f;
  unsigned near = 0;

  if (active) {
    const size_t o = (size_t)y * W + x;
    const float* bu = p.bf + (size_t)pair * 2 * plane;
    const float* bv = bu + plane;
    float u[PX], v[PX], keep[PX];
    load_vec<float, PX>(bu + o, u);
    load_vec<float, PX>(bv + o, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) keep[i] = 1.0f;

    if (MASK == MASK_GIVEN) load_vec<float, PX, Ld::Stream>(p.mask_in + (size_t)pair * plane + o, keep);

    if (MASK == MASK_COMPUTED && (p.flags & TCLB200_MOB)) {
      // zero-padded central differences of the backward flow (flowtools.py:12-16,47-53)
      float uu[PX], ud[PX], vu[PX], vd[PX];
      if (y > 0) { load_vec<float, PX>(bu + o - W, uu); load_vec<float, PX>(bv + o - W, vu); }
      else {
#pragma unroll
        for (int i = 0; i < PX; ++i) uu[i] = vu[i] = 0.0f;
      }
      if (y + 1 < H) { load_vec<float, PX>(bu + o + W, ud); load_vec<float, PX>(bv + o + W, vd); }
      else {
#pragma unroll
        for (int i = 0; i < PX; ++i) ud[i] = vd[i] = 0.0f;
      }
      const float ul_edge = x > 0 ? __ldg(bu + o - 1) : 0.0f, vl_edge = x > 0 ? __ldg(bv + o - 1) : 0.0f;
      const float ur_edge = x + PX < W ? __ldg(bu + o + PX) : 0.0f, vr_edge = x + PX < W ? __ldg(bv + o + PX) : 0.0f;
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const float ul = i > 0 ? u[i > 0 ? i - 1 : 0] : ul_edge, ur = i + 1 < PX ? u[i + 1 < PX ? i + 1 : 0] : ur_edge;
        const float vl = i > 0 ? v[i > 0 ? i - 1 : 0] : vl_edge, vr = i + 1 < PX ? v[i + 1 < PX ? i + 1 : 0] : vr_edge;
        const float nb = sqnorm2(u[i], v[i], kV);
        float margin;
        if (motion_boundary(ul, ur, uu[i], ud[i], vl, vr, vu[i], vd[i], nb, kV, &margin)) keep[i] = 0.0f;
        near += fabsf(margin) < kNearBand;
      }
    }

    Taps t[PX];
    const bool need_taps = (MASK == MASK_COMPUTED && (p.flags & TCLB200_OCC)) || p.prev != nullptr;
    if (need_taps) {
#pragma unroll
      for (int i = 0; i < PX; ++i) t[i] = make_taps(u[i], v[i], x + i, y, g, kV);
    }

    if (MASK == MASK_COMPUTED && (p.flags & TCLB200_OCC)) {
      const float* fu = p.ff + (size_t)pair * 2 * plane;
      const float* fv = fu + plane;
      float wu[PX], wv[PX];
#pragma unroll
      for (int i = 0; i < PX; ++i) { wu[i] = sample_global(fu, t[i], W, kV); wv[i] = sample_global(fv, t[i], W, kV); }
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const float nb = sqnorm2(u[i], v[i], kV);
        float margin;
        if (occluded(wu[i], wv[i], u[i], v[i], nb, kV, &margin)) keep[i] = 0.0f;
        near += fabsf(margin) < kNearBand;
      }
    }

    if (p.mask_out) store_vec<float, PX>(p.mask_out + (size_t)pair * plane + o, keep);

    if (p.prev != nullptr) {
      float valid[PX];
      if (p.flags & TCLB200_VALIDITY) {
#pragma unroll
        for (int i = 0; i < PX; ++i) valid[i] = binarise_validity(ones_sample(t[i], kV));
      }
      const FrameT* prev = reinterpret_cast<const FrameT*>(p.prev) + (size_t)pair * C * plane;
      const FrameT* cur = p.cur ? reinterpret_cast<const FrameT*>(p.cur) + (size_t)pair * C * plane + o : nullptr;
      FrameT* wout = p.warp_out ? reinterpret_cast<FrameT*>(p.warp_out) + (size_t)pair * C * plane + o : nullptr;
      FrameT* bout = p.blend_out ? reinterpret_cast<FrameT*>(p.blend_out) + (size_t)pair * C * plane + o : nullptr;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const FrameT* pl = prev + (size_t)c * plane;
        float w[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) w[i] = sample_global(pl, t[i], W, kV);
        if (p.flags & TCLB200_VALIDITY) {
#pragma unroll
          for (int i = 0; i < PX; ++i) w[i] = __fmul_rn(w[i], valid[i]);
        }
        if (wout) store_vec<FrameT, PX>(wout + (size_t)c * plane, w);
        if (cur) {
          float cv[PX];
          load_vec<FrameT, PX, Ld::Stream>(cur + (size_t)c * plane, cv);
          if (REDUCE) {
#pragma unroll
            for (int i = 0; i < PX; ++i) {
              if (p.loss == TCLB200_L2) {
                const float md = __fmul_rn(keep[i], __fsub_rn(cv[i], w[i]));  // mask*(cur - warp)
                err = __fmaf_rn(md, md, err);
              } else {
                err += __fmul_rn(keep[i], fabsf(__fsub_rn(w[i], cv[i])));    // mask*|warp - cur|
              }
            }
          }
          if (bout) {
            float bl[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i)
              bl[i] = __fadd_rn(__fmul_rn(keep[i], w[i]), __fmul_rn(__fsub_rn(1.0f, keep[i]), cv[i]));
            store_vec<FrameT, PX>(bout + (size_t)c * plane, bl);
          }
        }
      }
    }
  }

  if (p.near_threshold != nullptr) {
    near = warp_sum(near);
    if (lane == 0 && near) atomicAdd(p.near_threshold, (unsigned long long)near);
  }

  if (REDUCE) {
    __shared__ double red[kWarps];
    __shared__ int s_last;
    const double bsum = block_sum((double)err, red);
    const unsigned tpp = p.tiles_per_pair;
    if (threadIdx.x == 0) {
      __stcg(&p.scratch.partials[(size_t)pair * tpp + tile], bsum);
      __threadfence();
      const unsigned tk = atomicAdd(&p.scratch.pair_ticket[pair], 1u);
      s_last = (tk == tpp - 1);
    }
    __syncthreads();
    if (s_last) {
      // last CTA of this pair: fold the pair's tile partials in a fixed order
      __threadfence();
      double s = 0.0;
      const double* pp = p.scratch.partials + (size_t)pair * tpp;
      for (unsigned i = threadIdx.x; i < tpp; i += kThreads) s += __ldcg(pp + i);
      const double S = block_sum(s, red);
      if (threadIdx.x == 0) {
        const float val = finalise_value(S * p.inv_count, p.finalize);
        if (p.pair_sums) p.pair_sums[pair] = S;
        if (p.pair_vals) p.pair_vals[pair] = val;
        // reuse partials[pair*tpp] / [pair*tpp+1] as this pair's (S, val) record for the batch fold
        __stcg(&p.scratch.partials[(size_t)pair * tpp], S);
        p.scratch.pair_ticket[pair] = 0;
        __threadfence();
        const unsigned tk = atomicAdd(p.scratch.batch_ticket, 1u);
        s_last = (tk == (unsigned)p.B - 1) ? 2 : 1;
      }
      __syncthreads();
      if (s_last == 2) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < p.B; i += kThreads) {
          const double Si = __ldcg(p.scratch.partials + (size_t)i * tpp);
          a += Si;
          b += (double)finalise_value(Si * p.inv_count, p.finalize);
        }
        const double A = block_sum(a, red);
        const double Bv = block_sum(b, red);
        if (threadIdx.x == 0) {
          if (p.total_sums) { p.total_sums[0] = A; p.total_sums[1] = Bv; }
          if (p.total_val) *p.total_val = finalise_value(A * p.inv_count / (double)p.B, p.finalize);
          *p.scratch.batch_ticket = 0;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// gradient(x): zero-padded central differences (flowtools.py:12-16)
// ---------------------------------------------------------------------------------------------
template <int PX>
__global__ void __launch_bounds__(kThreads) gradient_kernel(const float* __restrict__ xin, float* __restrict__ out, int B,
                                                            int H, int W, int tiles_x, int tiles_per_img) {
  const size_t plane = (size_t)H * W;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int img = blockIdx.x / tiles_per_img;
  const int tile = blockIdx.x - img * tiles_per_img;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int x = (tx * 32 + lane) * PX, y = ty * kWarps + wrp;
  if (x >= W || y >= H) return;
  const float* src = xin + (size_t)img * plane;
  const size_t o = (size_t)y * W + x;
  float c[PX], up[PX], dn[PX], dx[PX], dy[PX];
  load_vec<float, PX>(src + o, c);
  if (y > 0) load_vec<float, PX>(src + o - W, up);
  else {
#pragma unroll
    for (int i = 0; i < PX; ++i) up[i] = 0.0f;
  }
  if (y + 1 < H) load_vec<float, PX>(src + o + W, dn);
  else {
#pragma unroll
    for (int i = 0; i < PX; ++i) dn[i] = 0.0f;
  }
  const float l_edge = x > 0 ? __ldg(src + o - 1) : 0.0f;
  const float r_edge = x + PX < W ? __ldg(src + o + PX) : 0.0f;
#pragma unroll
  for (int i = 0; i < PX; ++i) {
    const float l = i > 0 ? c[i > 0 ? i - 1 : 0] : l_edge, r = i + 1 < PX ? c[i + 1 < PX ? i + 1 : 0] : r_edge;
    dx[i] = __fmul_rn(__fsub_rn(r, l), 0.5f);
    dy[i] = __fmul_rn(__fsub_rn(dn[i], up[i]), 0.5f);
  }
  store_vec<float, PX>(out + (size_t)img * plane + o, dx);
  store_vec<float, PX>(out + ((size_t)B + img) * plane + o, dy);
}

// ---------------------------------------------------------------------------------------------
// backward kernels.  One lane per target pixel; grad_prev is a bilinear scatter-add (red.global.add.f32).
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  const float* grad_out;   // warp_backward: (B,C,H,W); tcl_backward: unused
  const float* x;          // source frame (prev)
  const float* f;          // flow
  const float* mask;       // tcl_backward only
  const float* cur;        // tcl_backward only
  const float* grad_scale; // tcl_backward only (device scalar)
  float* grad_x;
  float* grad_f;
  float* grad_cur;
  Geo geo;
  int B, C, flags, loss;
};

// FUSED_LOSS: the upstream gradient of warp is derived in-kernel from the masked loss.
template <bool FUSED_LOSS>
__global__ void __launch_bounds__(256) warp_backward_kernel(const BwdParams p) {
  const Geo& g = p.geo;
  const int W = g.W, H = g.H, C = p.C;
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)p.B * plane) return;
  const int b = (int)(idx / plane);
  const size_t o = idx - (size_t)b * plane;
  const int y = (int)(o / W), x = (int)(o - (size_t)y * W);
  const float u = __ldg(p.f + (size_t)b * 2 * plane + o), v = __ldg(p.f + ((size_t)b * 2 + 1) * plane + o);

  // taps, with the un-multiplied fractional parts kept for the coordinate gradient
  const float ix = source_coord(x, u, g.Wf, g.dxf, g.inv_dx, kV);
  const float iy = source_coord(y, v, g.Hf, g.dyf, g.inv_dy, kV);
  const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
  const int x1 = (int)((unsigned)x0 + 1u), y1 = (int)((unsigned)y0 + 1u);
  const float fx1 = (float)x1 - ix, fx0 = ix - (float)x0, fy1 = (float)y1 - iy, fy0 = iy - (float)y0;
  const float nw = fx1 * fy1, ne = fx0 * fy1, sw = fx1 * fy0, se = fx0 * fy0;
  const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)x1 < (unsigned)W;
  const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)y1 < (unsigned)H;
  const bool p00 = xin0 && yin0, p10 = xin1 && yin0, p01 = xin0 && yin1, p11 = xin1 && yin1;
  const long long o00 = (long long)y0 * W + x0;

  float valid = 1.0f;
  if (p.flags & TCLB200_VALIDITY) {
    Taps t; t.p00 = p00; t.p10 = p10; t.p01 = p01; t.p11 = p11; t.nw = nw; t.ne = ne; t.sw = sw; t.se = se; t.o00 = 0;
    valid = binarise_validity(ones_sample(t, kV));
  }
  float m = 1.0f, scale = 1.0f;
  if (FUSED_LOSS) {
    m = p.mask ? __ldg(p.mask + (size_t)b * plane + o) : 1.0f;
    scale = __ldg(p.grad_scale);
  }

  float gix = 0.0f, giy = 0.0f;
  for (int c = 0; c < C; ++c) {
    const size_t base = ((size_t)b * C + c) * plane;
    const float* xp = p.x + base;
    const float v00 = p00 ? __ldg(xp + o00) : 0.0f, v10 = p10 ? __ldg(xp + o00 + 1) : 0.0f;
    const float v01 = p01 ? __ldg(xp + o00 + W) : 0.0f, v11 = p11 ? __ldg(xp + o00 + W + 1) : 0.0f;
    float go;  // d loss / d warp[b,c,y,x]
    if (FUSED_LOSS) {
      float wv = 0.0f;
      if (p00) wv = fmaf(v00, nw, wv);
      if (p10) wv = fmaf(v10, ne, wv);
      if (p01) wv = fmaf(v01, sw, wv);
      if (p11) wv = fmaf(v11, se, wv);
      wv *= valid;
      const float cv = __ldg(p.cur + base + o);
      float gc;  // d loss / d cur
      if (p.loss == TCLB200_L2) {
        gc = 2.0f * scale * m * m * (cv - wv);
      } else {
        const float d = wv - cv;
        gc = -scale * m * (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f));
      }
      if (p.grad_cur) p.grad_cur[base + o] = gc;
      go = -gc;
    } else {
      go = __ldg(p.grad_out + base + o);
    }
    go *= valid;
    if (p.grad_x) {
      float* gp = p.grad_x + base;
      if (p00) atomicAdd(gp + o00, nw * go);
      if (p10) atomicAdd(gp + o00 + 1, ne * go);
      if (p01) atomicAdd(gp + o00 + W, sw * go);
      if (p11) atomicAdd(gp + o00 + W + 1, se * go);
    }
    if (p.grad_f) {
      gix += go * ((v10 - v00) * fy1 + (v11 - v01) * fy0);
      giy += go * ((v01 - v00) * fx1 + (v11 - v10) * fx0);
    }
  }
  if (p.grad_f) {
    // d ix/d gx = W/2 (grid_sampler), d gx/d u = 2/(W-1) (flowtools.py:28)
    p.grad_f[(size_t)b * 2 * plane + o] = 2.0f * (g.Wf * 0.5f * gix) / g.dxf;
    p.grad_f[((size_t)b * 2 + 1) * plane + o] = 2.0f * (g.Hf * 0.5f * giy) / g.dyf;
  }
}

}  // namespace tcl

// =============================================================================================
// C ABI
// =============================================================================================
using namespace tcl;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
#define CUDA_TRY(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) return fail(TCLB200_ERR_CUDA, #expr ": %s", cudaGetErrorString(e__)); \
  } while (0)

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int tclb200_abi_version(void) { return TCLB200_ABI_VERSION; }
extern "C" const char* tclb200_last_error(void) { return g_err; }

static inline int tiles_x_for(int W, int px) { return (W + 32 * px - 1) / (32 * px); }
static inline int tiles_y_for(int H) { return (H + kWarps - 1) / kWarps; }
// scratch is sized for the narrowest tiling (PX = 1) so any vector width fits
extern "C" size_t tclb200_scratch_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const size_t tpp = (size_t)tiles_x_for(W, 1) * tiles_y_for(H);
  return align_up((size_t)B * tpp * sizeof(double), 256) + align_up(((size_t)B + 1) * sizeof(unsigned), 256);
}

template <typename FrameT, int PX, int MASK, bool REDUCE>
static cudaError_t launch_fused(const FwdParams& p, cudaStream_t s) {
  const unsigned grid = (unsigned)((size_t)p.B * p.tiles_per_pair);
  if (p.C == 3) fused_forward_kernel<FrameT, PX, MASK, REDUCE, 3><<<grid, kThreads, 0, s>>>(p);
  else if (p.C == 2) fused_forward_kernel<FrameT, PX, MASK, REDUCE, 2><<<grid, kThreads, 0, s>>>(p);
  else fused_forward_kernel<FrameT, PX, MASK, REDUCE, 0><<<grid, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

template <typename FrameT, int PX>
static cudaError_t dispatch_fused(const FwdParams& p, int mask_kind, bool reduce, cudaStream_t s) {
  if (mask_kind == MASK_COMPUTED) return reduce ? launch_fused<FrameT, PX, MASK_COMPUTED, true>(p, s) : launch_fused<FrameT, PX, MASK_COMPUTED, false>(p, s);
  if (mask_kind == MASK_GIVEN) return reduce ? launch_fused<FrameT, PX, MASK_GIVEN, true>(p, s) : launch_fused<FrameT, PX, MASK_GIVEN, false>(p, s);
  return reduce ? launch_fused<FrameT, PX, MASK_NONE, true>(p, s) : launch_fused<FrameT, PX, MASK_NONE, false>(p, s);
}

static int run_fused(const tclb200_tcl_args* a, cudaStream_t s) {
  if (!a) return fail(TCLB200_ERR_INVALID, "args is NULL");
  if (a->B <= 0 || a->H <= 0 || a->W <= 0) return fail(TCLB200_ERR_INVALID, "B, H, W must be positive");
  if (!a->bf) return fail(TCLB200_ERR_INVALID, "bf (the flow the warp samples with) is required");
  if (a->dtype != TCLB200_F32 && a->dtype != TCLB200_BF16) return fail(TCLB200_ERR_INVALID, "unknown dtype");
  if (a->loss != TCLB200_L2 && a->loss != TCLB200_L1) return fail(TCLB200_ERR_INVALID, "unknown loss");
  if (a->finalize != TCLB200_FIN_MEAN && a->finalize != TCLB200_FIN_RMSE) return fail(TCLB200_ERR_INVALID, "unknown finalize");
  if (a->prev && a->C <= 0) return fail(TCLB200_ERR_INVALID, "C must be positive when frames are given");
  if ((a->cur || a->warp_out || a->blend_out) && !a->prev) return fail(TCLB200_ERR_INVALID, "prev is required with cur / warp_out / blend_out");
  if (a->blend_out && !a->cur) return fail(TCLB200_ERR_INVALID, "blend_out needs cur");
  if ((size_t)a->H * a->W >= (1u << 30)) return fail(TCLB200_ERR_UNSUPPORTED, "H*W must be below 2^30");
  const bool reduce = a->cur && (a->pair_sums || a->total_sums || a->pair_vals || a->total_val);
  const int mask_kind = a->ff ? MASK_COMPUTED : (a->mask_in ? MASK_GIVEN : MASK_NONE);
  if (mask_kind == MASK_COMPUTED && !(a->flags & (TCLB200_OCC | TCLB200_MOB)))
    return fail(TCLB200_ERR_INVALID, "ff given but neither TCLB200_OCC nor TCLB200_MOB requested");
  if (!a->prev && !a->mask_out) return fail(TCLB200_ERR_INVALID, "nothing to compute: no frames and no mask_out");

  // vector width: 16-byte lanes need W % PX == 0 and 16-byte aligned bases
  const int fsz = a->dtype == TCLB200_BF16 ? 2 : 4;
  bool vec4 = (a->W % 4 == 0) && aligned16(a->bf) && aligned16(a->ff) && aligned16(a->mask_in) && aligned16(a->mask_out);
  // frames: PX elements of fsz bytes each -> 8-byte (bf16) or 16-byte (fp32) accesses
  auto frame_ok = [&](const void* q) { return (reinterpret_cast<uintptr_t>(q) & (uintptr_t)(4 * fsz - 1)) == 0; };
  vec4 = vec4 && frame_ok(a->prev) && frame_ok(a->cur) && frame_ok(a->warp_out) && frame_ok(a->blend_out);
  const int px = vec4 ? 4 : 1;

  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.ff = a->ff; p.bf = a->bf; p.mask_in = a->mask_in; p.prev = a->prev; p.cur = a->cur;
  p.warp_out = a->warp_out; p.mask_out = a->mask_out; p.blend_out = a->blend_out;
  p.pair_sums = a->pair_sums; p.total_sums = a->total_sums; p.pair_vals = a->pair_vals; p.total_val = a->total_val;
  p.near_threshold = a->near_threshold;
  p.geo = make_geo(a->H, a->W);
  p.B = a->B; p.C = a->prev ? a->C : 0;
  p.tiles_x = tiles_x_for(a->W, px);
  p.tiles_per_pair = p.tiles_x * tiles_y_for(a->H);
  p.flags = a->flags; p.loss = a->loss; p.finalize = a->finalize;
  p.inv_count = a->prev ? 1.0 / ((double)a->C * a->H * a->W) : 0.0;
  if ((size_t)p.B * p.tiles_per_pair >= 0x7fffffffu) return fail(TCLB200_ERR_UNSUPPORTED, "too many tiles for one launch");
  if (reduce) {
    if (!a->scratch || a->scratch_bytes < tclb200_scratch_bytes(a->B, a->H, a->W))
      return fail(TCLB200_ERR_INVALID, "scratch missing or smaller than tclb200_scratch_bytes(B,H,W)");
    char* base = reinterpret_cast<char*>(a->scratch);
    const size_t tpp1 = (size_t)tiles_x_for(a->W, 1) * tiles_y_for(a->H);
    p.scratch.partials = reinterpret_cast<double*>(base);
    p.scratch.pair_ticket = reinterpret_cast<unsigned*>(base + align_up((size_t)a->B * tpp1 * sizeof(double), 256));
    p.scratch.batch_ticket = p.scratch.pair_ticket + a->B;
  }
  cudaError_t e;
  if (a->dtype == TCLB200_BF16)
    e = px == 4 ? dispatch_fused<__nv_bfloat16, 4>(p, mask_kind, reduce, s) : dispatch_fused<__nv_bfloat16, 1>(p, mask_kind, reduce, s);
  else
    e = px == 4 ? dispatch_fused<float, 4>(p, mask_kind, reduce, s) : dispatch_fused<float, 1>(p, mask_kind, reduce, s);
  if (e != cudaSuccess) return fail(TCLB200_ERR_CUDA, "fused_forward_kernel launch: %s", cudaGetErrorString(e));
  return TCLB200_OK;
}

extern "C" int tclb200_tcl_forward(const tclb200_tcl_args* args, tclb200_stream_t stream) {
  return run_fused(args, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int tclb200_warp(const void* x, const float* f, void* out, int B, int C, int H, int W, int dtype, int flags,
                            tclb200_stream_t stream) {
  if (!x || !f || !out) return fail(TCLB200_ERR_INVALID, "x, f and out are required");
  tclb200_tcl_args a;
  memset(&a, 0, sizeof(a));
  a.bf = f; a.prev = x; a.warp_out = out;
  a.B = B; a.C = C; a.H = H; a.W = W; a.dtype = dtype; a.flags = flags & TCLB200_VALIDITY;
  return run_fused(&a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int tclb200_fbcheck(const float* ff, const float* bf, float* mask_out, int B, int H, int W, int flags,
                               unsigned long long* near_threshold, tclb200_stream_t stream) {
  if (!bf || !mask_out) return fail(TCLB200_ERR_INVALID, "bf and mask_out are required");
  if ((flags & TCLB200_OCC) && !ff) return fail(TCLB200_ERR_INVALID, "ff is required for the occlusion test");
  if (!(flags & (TCLB200_OCC | TCLB200_MOB))) return fail(TCLB200_ERR_INVALID, "flags must request TCLB200_OCC and/or TCLB200_MOB");
  tclb200_tcl_args a;
  memset(&a, 0, sizeof(a));
  a.ff = ff ? ff : bf;  // the motion-boundary-only variant never reads ff
  a.bf = bf; a.mask_out = mask_out; a.near_threshold = near_threshold;
  a.B = B; a.C = 0; a.H = H; a.W = W; a.flags = flags & (TCLB200_OCC | TCLB200_MOB);
  return run_fused(&a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int tclb200_gradient(const float* x, float* out, int B, int H, int W, tclb200_stream_t stream) {
  if (!x || !out) return fail(TCLB200_ERR_INVALID, "x and out are required");
  if (B <= 0 || H <= 0 || W <= 0) return fail(TCLB200_ERR_INVALID, "B, H, W must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool vec4 = (W % 4 == 0) && aligned16(x) && aligned16(out);
  const int px = vec4 ? 4 : 1;
  const int tx = tiles_x_for(W, px), tpi = tx * tiles_y_for(H);
  if ((size_t)B * tpi >= 0x7fffffffu) return fail(TCLB200_ERR_UNSUPPORTED, "too many tiles for one launch");
  const unsigned grid = (unsigned)((size_t)B * tpi);
  if (vec4) gradient_kernel<4><<<grid, kThreads, 0, s>>>(x, out, B, H, W, tx, tpi);
  else gradient_kernel<1><<<grid, kThreads, 0, s>>>(x, out, B, H, W, tx, tpi);
  CUDA_TRY(cudaGetLastError());
  return TCLB200_OK;
}

static int run_backward(const BwdParams& p, bool fused, cudaStream_t s) {
  const size_t n = (size_t)p.B * p.geo.H * p.geo.W;
  if (p.grad_x) CUDA_TRY(cudaMemsetAsync(p.grad_x, 0, n * p.C * sizeof(float), s));
  const size_t blocks = (n + 255) / 256;
  if (blocks >= 0x7fffffffu) return fail(TCLB200_ERR_UNSUPPORTED, "too many pixels for one launch");
  if (fused) warp_backward_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(p);
  else warp_backward_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(p);
  CUDA_TRY(cudaGetLastError());
  return TCLB200_OK;
}

extern "C" int tclb200_warp_backward(const float* grad_out, const float* x, const float* f, float* grad_x, float* grad_f,
                                     int B, int C, int H, int W, int flags, tclb200_stream_t stream) {
  if (!grad_out || !x || !f) return fail(TCLB200_ERR_INVALID, "grad_out, x and f are required");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(TCLB200_ERR_INVALID, "B, C, H, W must be positive");
  if ((size_t)H * W >= (1u << 30)) return fail(TCLB200_ERR_UNSUPPORTED, "H*W must be below 2^30");
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.grad_out = grad_out; p.x = x; p.f = f; p.grad_x = grad_x; p.grad_f = grad_f;
  p.geo = make_geo(H, W); p.B = B; p.C = C; p.flags = flags & TCLB200_VALIDITY;
  return run_backward(p, false, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int tclb200_tcl_backward(const float* bf, const float* mask, const float* prev, const float* cur,
                                    const float* grad_scale, float* grad_prev, float* grad_cur, int B, int C, int H, int W,
                                    int flags, int loss, tclb200_stream_t stream) {
  if (!bf || !prev || !cur || !grad_scale) return fail(TCLB200_ERR_INVALID, "bf, prev, cur and grad_scale are required");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(TCLB200_ERR_INVALID, "B, C, H, W must be positive");
  if (loss != TCLB200_L2 && loss != TCLB200_L1) return fail(TCLB200_ERR_INVALID, "unknown loss");
  if ((size_t)H * W >= (1u << 30)) return fail(TCLB200_ERR_UNSUPPORTED, "H*W must be below 2^30");
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.x = prev; p.f = bf; p.mask = mask; p.cur = cur; p.grad_scale = grad_scale;
  p.grad_x = grad_prev; p.grad_cur = grad_cur;
  p.geo = make_geo(H, W); p.B = B; p.C = C; p.flags = flags & TCLB200_VALIDITY; p.loss = loss;
  return run_backward(p, true, reinterpret_cast<cudaStream_t>(stream));
}
