// Device-side greedy decoding bookkeeping (replaces the host loop of HF generate driven from myriad.py:447-450
// with prepare_inputs_for_generation modeling_llama.py:730-760 and StoppingCriteriaSub conversation.py:96-107).
// Everything that changes from step to step lives in one device struct so a decode step can be replayed as a
// CUDA graph: arg-max over the vocabulary, eos suppression for the first min_new_tokens, finished rows emit
// pad (= eos), row-0 stop sequences, and the position / KV-length / cache-slot counters of the next step.
#include "common.h"
#include "ptx.cuh"

namespace myr {

// first index of the maximum (torch.argmax tie rule), one CTA per row
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, long long ld, int V,
                                                      const int* __restrict__ state, int min_new, int eos,
                                                      int* __restrict__ raw) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sv[32];
  __shared__ int si[32];
  const int b = blockIdx.x;
  const int step = state[0];
  const float* row = logits + (size_t)b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float v = row[i];
    if (step < min_new && i == eos) v = -INFINITY;
    if (v > best || (v == best && i < bi)) {
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) raw[b] = bi;
  }
}

// state layout (int32): [0] step  [1] done  [2] cache_off  [3] reserved
//   [4 .. 4+B) unfinished   [4+B .. 4+2B) cur_tok   [4+2B .. 4+3B) kv_len   [4+3B .. 4+4B) pos
//   [4+4B .. 4+4B+B*max_new) tokens (row-major [B, max_new])
__global__ void greedy_update_kernel(int* __restrict__ state, const int* __restrict__ raw, int B, int max_new, int eos,
                                     const int* __restrict__ stops, int n_stops, int stop_max_len) {
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (state[1]) return;  // already done: decode steps replayed past the stop (the host polls every few steps) change nothing
  const int step = state[0];
  int* unfinished = state + 4;
  int* cur = state + 4 + B;
  int* kv_len = state + 4 + 2 * B;
  int* pos = state + 4 + 3 * B;
  int* toks = state + 4 + 4 * B;
  int any_unfinished = 0;
  for (int b = 0; b < B; ++b) {
    const int t = unfinished[b] ? raw[b] : eos;
    toks[b * max_new + step] = t;
    cur[b] = t;
    if (t == eos) unfinished[b] = 0;
    any_unfinished |= unfinished[b];
    kv_len[b] += 1;
    pos[b] += 1;
  }
  int stop = 0;
  for (int s = 0; s < n_stops; ++s) {  // row 0 only, as StoppingCriteriaSub does
    const int* seq = stops + s * stop_max_len;
    int len = 0;
    while (len < stop_max_len && seq[len] >= 0) ++len;
    if (len == 0 || step + 1 < len) continue;
    bool match = true;
    for (int k = 0; k < len; ++k) match = match && toks[step + 1 - len + k] == seq[k];
    stop |= match ? 1 : 0;
  }
  state[0] = step + 1;
  state[2] += 1;
  if (stop || !any_unfinished || step + 1 >= max_new) state[1] = 1;
}

}  // namespace myr

using namespace myr;

extern "C" int myr_greedy_step(const void* logits, int64_t ld_logits, int32_t B, int32_t V, void* state, void* scratch,
                               int32_t max_new_tokens, int32_t min_new_tokens, int32_t eos, const void* stop_seqs,
                               int32_t n_stops, int32_t stop_max_len, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(logits && state && scratch && B > 0 && V > 0 && max_new_tokens > 0, "greedy_step: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(argmax_kernel, dim3(B), dim3(1024), 0, stream, true, reinterpret_cast<const float*>(logits),
                               (long long)ld_logits, V, reinterpret_cast<const int*>(state), min_new_tokens, eos,
                               reinterpret_cast<int*>(scratch)));
  MYR_CHECK_LAUNCH();
  MYR_CHECK_CUDA(launch_kernel(greedy_update_kernel, dim3(1), dim3(32), 0, stream, true, reinterpret_cast<int*>(state),
                               reinterpret_cast<const int*>(scratch), B, max_new_tokens, eos,
                               reinterpret_cast<const int*>(stop_seqs), n_stops, stop_max_len));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
