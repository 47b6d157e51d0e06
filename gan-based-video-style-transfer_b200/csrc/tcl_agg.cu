// Aggregation of the reference's evaluation loop on the device:
//   per-video mean of the per-pair RMSE        StarGANv2AdvCon/core/solver.py:352-354
//   mean over videos / pooled statistics       utils/sintel_eval.py:112-126 (save_dict_as_json)
// Two tiny single-CTA kernels around the one all-reduce of the sharded evaluation (sharding.py):
//   pack    per-pair values + sequence ids -> [sum of values per sequence | pair count per sequence | sum of squared error | element count]
//   unpack  the (all-reduced) packed vector -> [per-sequence mean ... | mean over sequences | mean over pairs | pooled RMSE | pair count]
// They replace ~22 framework launches per step (fills, index_add, clamp, div, ...) with two.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tcl_b200.h"

namespace tcl {
void set_last_error(const char* msg);
void count_launch();
}

namespace {

constexpr int kAggThreads = 256;
constexpr int kPackThreads = 1024;   // pack: few dependent rounds of loads (1041 pairs = 2 rounds)

// Sums of at most a few thousand fp32-valued doubles of similar magnitude are exact in fp64, so the order in which the
// shared-memory atomics land does not show in the result.
__global__ void __launch_bounds__(kPackThreads) pack_sequences_kernel(const float* __restrict__ pair_vals, const double* __restrict__ sum_sq,
                                                                     const long long* __restrict__ seq_of_pair, int n_pairs, int n_seq,
                                                                     double elems_per_pair, double* __restrict__ packed) {
  extern __shared__ double s_acc[];   // [n_seq] sums, [n_seq] counts
  for (int i = threadIdx.x; i < 2 * n_seq; i += kPackThreads) s_acc[i] = 0.0;
  __syncthreads();
  // warp-aggregated: the lanes of a warp that hold pairs of the same sequence (pairs of a clip are neighbours) fold their
  // values in the lowest such lane, which issues ONE shared-memory atomic per sequence and warp (a shared fp64 atomicAdd is a
  // compare-and-swap loop: 32 lanes hammering one address took 23 us for 1041 pairs under ncu, this takes 9, cold)
  const int lane = threadIdx.x & 31;
  for (int base = 0; base < n_pairs; base += kPackThreads) {
    const int i = base + (int)threadIdx.x;
    const bool live = i < n_pairs;
    long long s = live ? seq_of_pair[i] : -1;
    if (s < 0 || s >= n_seq) s = -1;                      // out-of-range ids are ignored
    const double v = (live && s >= 0) ? (double)pair_vals[i] : 0.0;
    const unsigned grp = __match_any_sync(0xffffffffu, s);
    const int leader = __ffs(grp) - 1;
    double sum = 0.0;
    for (unsigned rem = grp; rem; rem &= rem - 1) {       // every lane walks its own group's members (uniform within the group)
      const int src = __ffs(rem) - 1;
      sum += __shfl_sync(grp, v, src);
    }
    if (lane == leader && s >= 0) {
      atomicAdd(&s_acc[s], sum);
      atomicAdd(&s_acc[n_seq + s], (double)__popc(grp));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * n_seq; i += kPackThreads) packed[i] = s_acc[i];
  if (threadIdx.x == 0) {
    packed[2 * n_seq] = (n_pairs > 0 && sum_sq) ? *sum_sq : 0.0;
    packed[2 * n_seq + 1] = (double)n_pairs * elems_per_pair;
  }
}

__device__ __forceinline__ double block_sum_agg(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < kAggThreads / 32; ++i) s += red[i];   // fixed order
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(kAggThreads) unpack_sequences_kernel(const double* __restrict__ packed, int n_seq, double* __restrict__ out) {
  __shared__ double red[kAggThreads / 32];
  double means = 0.0, present = 0.0, sums = 0.0, cnts = 0.0;
  for (int i = threadIdx.x; i < n_seq; i += kAggThreads) {
    const double s = packed[i], c = packed[n_seq + i];
    const double m = s / (c < 1.0 ? 1.0 : c);
    out[i] = m;                                  // per-sequence mean of the per-pair values
    if (c > 0.0) { means += m; present += 1.0; }
    sums += s; cnts += c;
  }
  const double M = block_sum_agg(means, red), P = block_sum_agg(present, red);
  const double S = block_sum_agg(sums, red), C = block_sum_agg(cnts, red);
  if (threadIdx.x == 0) {
    const double elems = packed[2 * n_seq + 1];
    out[n_seq + 0] = M / (P < 1.0 ? 1.0 : P);            // the reference's mean over sequences (utils/sintel_eval.py:116-118)
    out[n_seq + 1] = S / (C < 1.0 ? 1.0 : C);            // mean over all pairs
    out[n_seq + 2] = sqrt(packed[2 * n_seq] / (elems < 1.0 ? 1.0 : elems));   // pooled RMSE
    out[n_seq + 3] = C;                                   // number of pairs
  }
}

int afail(int code, const char* msg) {
  tcl::set_last_error(msg);
  return code;
}

}  // namespace

extern "C" int tclb200_pack_sequence_sums(const float* pair_vals, const double* sum_sq, const long long* seq_of_pair, int n_pairs, int n_seq,
                                          double elems_per_pair, double* packed, tclb200_stream_t stream) {
  if (!packed || n_seq <= 0 || n_pairs < 0) return afail(TCLB200_ERR_INVALID, "packed and a positive n_seq are required");
  if (n_pairs > 0 && (!pair_vals || !seq_of_pair)) return afail(TCLB200_ERR_INVALID, "pair_vals and seq_of_pair are required");
  if (n_seq > 2048) return afail(TCLB200_ERR_UNSUPPORTED, "at most 2048 sequences per call");
  pack_sequences_kernel<<<1, kPackThreads, 2 * (size_t)n_seq * sizeof(double), reinterpret_cast<cudaStream_t>(stream)>>>(
      pair_vals, sum_sq, seq_of_pair, n_pairs, n_seq, elems_per_pair, packed);
  tcl::count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return afail(TCLB200_ERR_CUDA, cudaGetErrorString(e));
  return TCLB200_OK;
}

extern "C" int tclb200_unpack_sequence_means(const double* packed, int n_seq, double* out, tclb200_stream_t stream) {
  if (!packed || !out || n_seq <= 0) return afail(TCLB200_ERR_INVALID, "packed, out and a positive n_seq are required");
  unpack_sequences_kernel<<<1, kAggThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(packed, n_seq, out);
  tcl::count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return afail(TCLB200_ERR_CUDA, cudaGetErrorString(e));
  return TCLB200_OK;
}
