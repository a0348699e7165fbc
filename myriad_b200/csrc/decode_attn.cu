// Decode-step attention for one new token per sequence (HBM/latency-bound, CUDA cores): one CTA per (head, batch row) fuses
//   LoRA-B update of q and v  (peft, myriad.py:171-178)         q += s * B_q (A_q x),  v += s * B_v (A_v x)
//   RoPE of q and k           (modeling_llama.py:109-123)
//   KV-cache append           (replaces the torch.cat growth of modeling_llama.py:190-195)
//   softmax(q K^T / sqrt(dh)) V over the cache (modeling_llama.py:197-215; fp32 softmax)
// which the prefill path runs as myr_rope_cache + the tcgen05 flash kernel. With a single query row the 128-row MMA tile
// of the flash kernel is 99 % padding and its TMA -> MMA -> softmax -> MMA chain is pure latency; here every thread streams
// 16-byte pieces of K and V rows straight from the cache. Deterministic (fixed reduction order): CUDA-graph replays of the
// decode step reproduce eager launches bit for bit.
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int DA_DH = 128;
constexpr int DA_THREADS = 128;

struct DecodeAttnParams {
  const __half* qkv; long long ldq;  // [B, ldq]: q | k | v | xa_q (r) | xa_v (r)
  int B, H, Smax;
  const int* pos;                    // [B] rotary position of the new token
  const float* cos_t; const float* sin_t;
  __half* kcache; __half* vcache; long long c_ts, c_bs;
  const int* cache_off; int cache_off_host;  // cache slot of the new token
  const int* kv_len;                 // [B] number of visible keys INCLUDING the new token
  const __half* lora_bq; const __half* lora_bv; int lora_r; float lora_scale;
  float scale;
  __half* out; long long ldo;        // [B, H * dh]
};

__device__ __forceinline__ void da_unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ float da_lora_dot(const __half* __restrict__ b, int row, const float (&xa)[8]) {
  float w[8];
  da_unpack8(__ldg(reinterpret_cast<const uint4*>(b + (size_t)row * 8)), w);
  float a = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) a = fmaf(w[r], xa[r], a);
  return a;
}

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(const DecodeAttnParams p) {
  extern __shared__ float s_scores[];  // [Smax]
  __shared__ float s_q[DA_DH];
  __shared__ __align__(16) __half s_k[DA_DH];
  __shared__ __align__(16) __half s_v[DA_DH];
  __shared__ float s_red[4];
  __shared__ float s_acc[4][DA_DH];
  pdl_wait();
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HD = p.H * DA_DH;
  const __half* row = p.qkv + (size_t)b * p.ldq;
  const int off = p.cache_off ? *p.cache_off : p.cache_off_host;
  int kvl = p.kv_len ? p.kv_len[b] : off + 1;
  if (kvl > p.Smax) kvl = p.Smax;
  __half* kbase = p.kcache + (size_t)b * p.c_bs + h * DA_DH;
  __half* vbase = p.vcache + (size_t)b * p.c_bs + h * DA_DH;

  // ---- phase 0: q / k / v of the new token: LoRA, rotation, cache append (one rotary pair per thread)
  if (tid < DA_DH / 2) {
    const int j = tid, half = DA_DH / 2;
    const int pos = p.pos[b];
    const float c = p.cos_t[(size_t)pos * half + j], sn = p.sin_t[(size_t)pos * half + j];
    float q1 = __half2float(row[h * DA_DH + j]), q2 = __half2float(row[h * DA_DH + half + j]);
    const float k1 = __half2float(row[HD + h * DA_DH + j]), k2 = __half2float(row[HD + h * DA_DH + half + j]);
    float v1 = __half2float(row[2 * HD + h * DA_DH + j]), v2 = __half2float(row[2 * HD + h * DA_DH + half + j]);
    if (p.lora_r) {
      float xq[8], xv[8];
      da_unpack8(*reinterpret_cast<const uint4*>(row + 3 * HD), xq);
      da_unpack8(*reinterpret_cast<const uint4*>(row + 3 * HD + 8), xv);
      q1 = fmaf(p.lora_scale, da_lora_dot(p.lora_bq, h * DA_DH + j, xq), q1);
      q2 = fmaf(p.lora_scale, da_lora_dot(p.lora_bq, h * DA_DH + half + j, xq), q2);
      v1 = fmaf(p.lora_scale, da_lora_dot(p.lora_bv, h * DA_DH + j, xv), v1);
      v2 = fmaf(p.lora_scale, da_lora_dot(p.lora_bv, h * DA_DH + half + j, xv), v2);
    }
    // fp16 rounding of the rotated q / k and of v: the values the prefill path stores (q in place, k / v in the cache)
    s_q[j] = round_f16(q1 * c - q2 * sn);
    s_q[half + j] = round_f16(q2 * c + q1 * sn);
    const __half ko1 = __float2half_rn(k1 * c - k2 * sn), ko2 = __float2half_rn(k2 * c + k1 * sn);
    const __half vo1 = __float2half_rn(v1), vo2 = __float2half_rn(v2);
    s_k[j] = ko1; s_k[half + j] = ko2;
    s_v[j] = vo1; s_v[half + j] = vo2;
    if (off >= 0 && off < p.Smax) {
      __half* kd = kbase + (size_t)off * p.c_ts;
      __half* vd = vbase + (size_t)off * p.c_ts;
      kd[j] = ko1; kd[half + j] = ko2;
      vd[j] = vo1; vd[half + j] = vo2;
    }
  }
  __syncthreads();

  // ---- phase 1: scores. A half-warp covers one key: lane l16 owns dims [8 * l16, 8 * l16 + 8)
  const int hw = lane >> 4, l16 = lane & 15;
  float qr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qr[i] = s_q[l16 * 8 + i];
  for (int j0 = warp * 2; j0 < kvl; j0 += 8) {  // trip count uniform across the warp: the shuffles below need all lanes
    const int j = j0 + hw;
    const bool valid = j < kvl;
    const __half* kp = (!valid || j == off) ? s_k : kbase + (size_t)j * p.c_ts;
    float kf[8];
    da_unpack8(*reinterpret_cast<const uint4*>(kp + l16 * 8), kf);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) d = fmaf(kf[i], qr[i], d);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (valid && l16 == 0) s_scores[j] = d * p.scale;
  }
  __syncthreads();

  // ---- softmax statistics (fp32)
  float m = -INFINITY;
  for (int j = tid; j < kvl; j += DA_THREADS) m = fmaxf(m, s_scores[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float l = 0.f;
  for (int j = tid; j < kvl; j += DA_THREADS) {
    const float e = __expf(s_scores[j] - m);
    s_scores[j] = e;
    l += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) s_red[warp] = l;
  __syncthreads();
  l = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);

  // ---- phase 2: O = P V
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j = warp * 2 + hw; j < kvl; j += 8) {
    const __half* vp = (j == off) ? s_v : vbase + (size_t)j * p.c_ts;
    float vf[8];
    da_unpack8(*reinterpret_cast<const uint4*>(vp + l16 * 8), vf);
    const float pj = s_scores[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, vf[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  if (hw == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[warp][l16 * 8 + i] = acc[i];
  }
  __syncthreads();
  {
    const float o = (s_acc[0][tid] + s_acc[1][tid]) + (s_acc[2][tid] + s_acc[3][tid]);
    p.out[(size_t)b * p.ldo + h * DA_DH + tid] = __float2half_rn(l > 0.f ? o / l : 0.f);
  }
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
  MYR_CHECK_CUDA(launch_kernel(decode_attn_kernel, dim3(a->H, a->B), dim3(DA_THREADS), (size_t)a->cache_len * sizeof(float), stream,
                               true, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
