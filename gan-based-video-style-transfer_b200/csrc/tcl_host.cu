// Host-buffer entry of the C ABI: tclb200_tcl_forward_host (include/tcl_b200.h).
//
// The reference's evaluation loops hold their data on the host side of the boundary: every frame of a clip is loaded,
// stylised and compared pair by pair with a `.cpu().numpy()` per pair (utils/sintel_eval.py:206-222,
// StarGANv2AdvCon/core/solver.py:336-347, utils/metrics/eval.py:137-149).  This entry takes the whole job -- flows,
// the frame bank of the clip(s), which frame each pair warps / compares with -- as HOST pointers and runs it as a
// software pipeline:
//
//   copy stream     chunk k: the frames its pairs need that are not on the device yet (each frame crosses PCIe ONCE:
//                   28 B/px per pair instead of 40 for a clip), then the two flows of its pairs into ring slot k % 3
//   caller's stream chunk k: one fused launch (tclb200_tcl_forward in clip mode) on that slot, results into the
//                   device result array; after the last chunk one D2H copy of the per-pair results
//
// PCIe is the bound (a 1024x436 pair is 12.5 MB of flows + frames; its kernel time is 4 us): the ring only has to keep
// the copy engine busy.  No state outlives the call except what the caller owns (workspace, streams are per call).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/tcl_b200.h"

namespace tcl { void set_last_error(const char* msg); }   // tcl_kernels.cu: the message tclb200_last_error() returns

namespace {

constexpr int kRing = 3;
constexpr int kDefaultChunk = 32;

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Layout {
  size_t frames, flows[kRing][2], mask[kRing], prev_idx, cur_idx, vals, sums, scratch, total;
  size_t frame_bytes, flow_bytes, mask_bytes;
  int chunk;
};

Layout make_layout(int P, int F, int C, int H, int W, int dtype, int chunk_pairs, bool with_mask) {
  Layout L;
  memset(&L, 0, sizeof(L));
  L.chunk = chunk_pairs > 0 ? chunk_pairs : kDefaultChunk;
  if (L.chunk > P) L.chunk = P;
  const size_t px = (size_t)H * W;
  L.frame_bytes = px * C * (dtype == TCLB200_BF16 ? 2 : 4);
  L.flow_bytes = px * 2 * sizeof(float);
  L.mask_bytes = px * sizeof(float);
  size_t o = 0;
  L.frames = o; o += up256(L.frame_bytes * (size_t)F);
  for (int r = 0; r < kRing; ++r) {
    for (int j = 0; j < 2; ++j) { L.flows[r][j] = o; o += up256(L.flow_bytes * (size_t)L.chunk); }
    L.mask[r] = o;
    if (with_mask) o += up256(L.mask_bytes * (size_t)L.chunk);
  }
  L.prev_idx = o; o += up256(sizeof(int) * (size_t)P);
  L.cur_idx = o; o += up256(sizeof(int) * (size_t)P);
  L.vals = o; o += up256(sizeof(float) * (size_t)P);
  L.sums = o; o += up256(sizeof(double) * (size_t)P);
  L.scratch = o; o += up256(tclb200_scratch_bytes(L.chunk, H, W));
  L.total = o;
  return L;
}

int hfail(int code, const char* what, const char* detail = "") {
  char msg[512];
  snprintf(msg, sizeof(msg), "%s%s", what, detail);
  tcl::set_last_error(msg);
  return code;
}

// RAII for the per-call copy stream and ring events (destroying them is safe while work is pending: the runtime
// releases the resources once the work has completed)
struct Pipe {
  cudaStream_t copy = nullptr;
  cudaEvent_t ready[kRing] = {}, done[kRing] = {}, fork = nullptr;
  cudaError_t init() {
    cudaError_t e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < kRing; ++i) {
      if ((e = cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
  }
  ~Pipe() {
    for (int i = 0; i < kRing; ++i) {
      if (ready[i]) cudaEventDestroy(ready[i]);
      if (done[i]) cudaEventDestroy(done[i]);
    }
    if (fork) cudaEventDestroy(fork);
    if (copy) cudaStreamDestroy(copy);
  }
};

}  // namespace

extern "C" size_t tclb200_host_workspace_bytes(int P, int F, int C, int H, int W, int dtype, int chunk_pairs, int with_mask) {
  if (P <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return make_layout(P, F, C, H, W, dtype, chunk_pairs, with_mask != 0).total;
}

#define HOST_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) return hfail(TCLB200_ERR_CUDA, #expr ": ", cudaGetErrorString(e__)); \
  } while (0)

extern "C" int tclb200_tcl_forward_host(const tclb200_host_args* a, tclb200_stream_t stream) {
  if (!a) return hfail(TCLB200_ERR_INVALID, "args is NULL");
  if (a->P <= 0 || a->F <= 0 || a->C <= 0 || a->H <= 0 || a->W <= 0) return hfail(TCLB200_ERR_INVALID, "P, F, C, H, W must be positive");
  if (!a->bf || !a->frames || !a->prev_index || !a->cur_index) return hfail(TCLB200_ERR_INVALID, "bf, frames, prev_index and cur_index are required");
  if (!a->pair_vals && !a->pair_sums) return hfail(TCLB200_ERR_INVALID, "nothing to return: pair_vals and pair_sums are both NULL");
  if (a->dtype != TCLB200_F32 && a->dtype != TCLB200_BF16) return hfail(TCLB200_ERR_INVALID, "unknown dtype");
  for (int p = 0; p < a->P; ++p)
    if (a->prev_index[p] < 0 || a->prev_index[p] >= a->F || a->cur_index[p] < 0 || a->cur_index[p] >= a->F)
      return hfail(TCLB200_ERR_INVALID, "prev_index / cur_index outside [0, F)");
  const bool with_mask = !a->ff && a->mask_in;
  const Layout L = make_layout(a->P, a->F, a->C, a->H, a->W, a->dtype, a->chunk_pairs, with_mask);
  if (!a->workspace || a->workspace_bytes < L.total) return hfail(TCLB200_ERR_INVALID, "workspace missing or smaller than tclb200_host_workspace_bytes()");
  if ((reinterpret_cast<uintptr_t>(a->workspace) & 255u) != 0) return hfail(TCLB200_ERR_INVALID, "workspace must be 256-byte aligned");

  cudaStream_t comp = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(a->workspace);
  Pipe pipe;
  HOST_TRY(pipe.init());
  // the workspace may still be in use by earlier work on the caller's stream
  HOST_TRY(cudaEventRecord(pipe.fork, comp));
  HOST_TRY(cudaStreamWaitEvent(pipe.copy, pipe.fork, 0));
  HOST_TRY(cudaMemsetAsync(ws + L.scratch, 0, tclb200_scratch_bytes(L.chunk, a->H, a->W), comp));
  HOST_TRY(cudaMemcpyAsync(ws + L.prev_idx, a->prev_index, sizeof(int) * (size_t)a->P, cudaMemcpyHostToDevice, pipe.copy));
  HOST_TRY(cudaMemcpyAsync(ws + L.cur_idx, a->cur_index, sizeof(int) * (size_t)a->P, cudaMemcpyHostToDevice, pipe.copy));

  std::vector<char> on_device((size_t)a->F, 0);
  std::vector<int> want;
  const char* h_frames = reinterpret_cast<const char*>(a->frames);
  for (int s = 0, k = 0; s < a->P; s += L.chunk, ++k) {
    const int n = a->P - s < L.chunk ? a->P - s : L.chunk;
    const int slot = k % kRing;
    // frames this chunk needs and the device does not hold yet; runs of consecutive frames go as one copy
    want.clear();
    for (int p = s; p < s + n; ++p) {
      const int f2[2] = {a->prev_index[p], a->cur_index[p]};
      for (int j = 0; j < 2; ++j)
        if (!on_device[(size_t)f2[j]]) { on_device[(size_t)f2[j]] = 1; want.push_back(f2[j]); }
    }
    for (size_t i = 0; i < want.size();) {
      size_t j = i + 1;
      while (j < want.size() && want[j] == want[j - 1] + 1) ++j;
      const size_t off = (size_t)want[i] * L.frame_bytes;
      HOST_TRY(cudaMemcpyAsync(ws + L.frames + off, h_frames + off, (j - i) * L.frame_bytes, cudaMemcpyHostToDevice, pipe.copy));
      i = j;
    }
    if (k >= kRing) HOST_TRY(cudaStreamWaitEvent(pipe.copy, pipe.done[slot], 0));   // the slot's previous chunk has been consumed
    if (a->ff) HOST_TRY(cudaMemcpyAsync(ws + L.flows[slot][0], a->ff + (size_t)s * 2 * a->H * a->W, (size_t)n * L.flow_bytes, cudaMemcpyHostToDevice, pipe.copy));
    HOST_TRY(cudaMemcpyAsync(ws + L.flows[slot][1], a->bf + (size_t)s * 2 * a->H * a->W, (size_t)n * L.flow_bytes, cudaMemcpyHostToDevice, pipe.copy));
    if (with_mask) HOST_TRY(cudaMemcpyAsync(ws + L.mask[slot], a->mask_in + (size_t)s * a->H * a->W, (size_t)n * L.mask_bytes, cudaMemcpyHostToDevice, pipe.copy));
    HOST_TRY(cudaEventRecord(pipe.ready[slot], pipe.copy));
    HOST_TRY(cudaStreamWaitEvent(comp, pipe.ready[slot], 0));

    tclb200_tcl_args t;
    memset(&t, 0, sizeof(t));
    t.ff = a->ff ? reinterpret_cast<const float*>(ws + L.flows[slot][0]) : nullptr;
    t.bf = reinterpret_cast<const float*>(ws + L.flows[slot][1]);
    t.mask_in = with_mask ? reinterpret_cast<const float*>(ws + L.mask[slot]) : nullptr;
    t.prev = ws + L.frames; t.cur = ws + L.frames;
    t.prev_index = reinterpret_cast<const int*>(ws + L.prev_idx) + s;
    t.cur_index = reinterpret_cast<const int*>(ws + L.cur_idx) + s;
    t.n_prev_frames = a->F; t.n_cur_frames = a->F;
    t.pair_vals = reinterpret_cast<float*>(ws + L.vals) + s;
    t.pair_sums = reinterpret_cast<double*>(ws + L.sums) + s;
    t.scratch = ws + L.scratch; t.scratch_bytes = tclb200_scratch_bytes(L.chunk, a->H, a->W);
    t.B = n; t.C = a->C; t.H = a->H; t.W = a->W;
    t.dtype = a->dtype; t.flags = a->flags; t.loss = a->loss; t.finalize = a->finalize;
    const int rc = tclb200_tcl_forward(&t, stream);
    if (rc != TCLB200_OK) return rc;   // tclb200_last_error() already holds the message
    HOST_TRY(cudaEventRecord(pipe.done[slot], comp));
  }
  if (a->pair_vals) HOST_TRY(cudaMemcpyAsync(a->pair_vals, ws + L.vals, sizeof(float) * (size_t)a->P, cudaMemcpyDeviceToHost, comp));
  if (a->pair_sums) HOST_TRY(cudaMemcpyAsync(a->pair_sums, ws + L.sums, sizeof(double) * (size_t)a->P, cudaMemcpyDeviceToHost, comp));
  return TCLB200_OK;
}
