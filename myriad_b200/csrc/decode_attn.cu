// Decode-step attention for one new token per sequence (HBM/latency-bound, CUDA cores): one CTA per (head, batch row) fuses
//   LoRA-B update of q and v  (peft, myriad.py:171-178)         q += s * B_q (A_q x),  v += s * B_v (A_v x)
//   RoPE of q and k           (modeling_llama.py:109-123)
//   KV-cache append           (replaces the torch.cat growth of modeling_llama.py:190-195)
//   softmax(q K^T / sqrt(dh)) V over the cache (modeling_llama.py:197-215; fp32 softmax)
// which the prefill path runs as myr_rope_cache + the tcgen05 flash kernel. With a single query row the 128-row MMA tile
// of the flash kernel is 99 % padding and its TMA -> MMA -> softmax -> MMA chain is pure latency; here every thread streams
// 16-byte pieces of K and V rows straight from the cache. Deterministic (fixed reduction order): CUDA-graph replays of the
// decode step reproduce eager launches bit for bit.
#include "decode_attn.cuh"

namespace myr {

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(const DecodeAttnParams p) {
  extern __shared__ float s_scores[];  // [Smax]
  __shared__ DecodeAttnSmem sm;
  pdl_wait();
  pdl_launch_dependents();
  decode_attn_task(p, blockIdx.x, blockIdx.y, threadIdx.x, sm, s_scores, [] { __syncthreads(); });
}

}  // namespace myr

using namespace myr;

extern "C" int myr_decode_attention(const myr_decode_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(a && a->qkv && a->pos && a->cos_table && a->sin_table && a->kcache && a->vcache && a->out, "decode_attention: null pointer");
  MYR_CHECK_ARG(a->B > 0 && a->H > 0 && a->dh == DA_DH, "decode_attention: head dim must be %d (got %d)", DA_DH, a->dh);
  MYR_CHECK_ARG(a->cache_len > 0 && a->cache_len <= 8192, "decode_attention: cache_len %d out of range", a->cache_len);
  MYR_CHECK_ARG(a->ldq % 8 == 0 && a->cache_token_stride % 8 == 0 && a->cache_batch_stride % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->kcache) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->vcache) & 15) == 0,
                "decode_attention: qkv / cache rows must be 16-byte aligned");
  MYR_CHECK_ARG(a->lora_r == 0 || a->lora_r == 8, "decode_attention: fused LoRA supports rank 8 only (got %d)", a->lora_r);
  DecodeAttnParams p;
  p.qkv = reinterpret_cast<const __half*>(a->qkv); p.ldq = a->ldq;
  p.B = a->B; p.H = a->H; p.Smax = a->cache_len;
  p.pos = reinterpret_cast<const int*>(a->pos);
  p.cos_t = reinterpret_cast<const float*>(a->cos_table); p.sin_t = reinterpret_cast<const float*>(a->sin_table);
  p.kcache = reinterpret_cast<__half*>(a->kcache); p.vcache = reinterpret_cast<__half*>(a->vcache);
  p.c_ts = a->cache_token_stride; p.c_bs = a->cache_batch_stride;
  p.cache_off = reinterpret_cast<const int*>(a->cache_off_dev); p.cache_off_host = a->cache_off;
  p.kv_len = reinterpret_cast<const int*>(a->kv_len);
  p.lora_bq = reinterpret_cast<const __half*>(a->lora_bq); p.lora_bv = reinterpret_cast<const __half*>(a->lora_bv);
  p.lora_r = (a->lora_bq && a->lora_bv) ? a->lora_r : 0; p.lora_scale = a->lora_scale;
  p.scale = a->scale;
  p.out = reinterpret_cast<__half*>(a->out); p.ldo = a->ldo;
  p.next_layer_stride = a->next_layer_stride;
  MYR_CHECK_CUDA(launch_kernel(decode_attn_kernel, dim3(a->H, a->B), dim3(DA_THREADS), (size_t)a->cache_len * sizeof(float), stream,
                               true, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
