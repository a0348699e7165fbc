// Decode-step attention body shared by the stand-alone kernel (decode_attn.cu) and the persistent decode-step kernel
// (decode_mega.cu). 128 threads cooperate on one (head, batch row); `sync` is the barrier over exactly those threads.
#pragma once
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
  long long next_layer_stride;       // elements from this layer's cache to the next layer's (0: none): L2 prefetch hint
  long long* trace;                  // debug: 6 x %globaltimer ns per CTA (myr_gemm_set_trace); stand-alone kernel only
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

// q . k over one 16-byte unit of the K row (8 halfs): four interleaved partial sums d[e % 4] - the same order in every SCALAR score
// loop (the register variant, the persistent decode kernel of decode_mega.cu and the MYR_DA_MMA=0 path of the TMA variant agree bit
// for bit; the default TMA / long-cache kernels compute the scores on mma.sync, another - equally deterministic - order). A single
// accumulator is a 128-deep dependent FMA chain (~0.3 us per key at 4 clk per FMA); four chains of 32 and 16-byte q reads cut the
// score phase of a 163-slot cache from ~1.3 us to about half.
__device__ __forceinline__ void da_dot8(const uint4& kraw, const float* __restrict__ q8, float (&d)[4]) {
  float kf[8];
  da_unpack8(kraw, kf);
  const float4 qa = *reinterpret_cast<const float4*>(q8), qb = *reinterpret_cast<const float4*>(q8 + 4);
  d[0] = fmaf(kf[0], qa.x, d[0]); d[1] = fmaf(kf[1], qa.y, d[1]); d[2] = fmaf(kf[2], qa.z, d[2]); d[3] = fmaf(kf[3], qa.w, d[3]);
  d[0] = fmaf(kf[4], qb.x, d[0]); d[1] = fmaf(kf[5], qb.y, d[1]); d[2] = fmaf(kf[6], qb.z, d[2]); d[3] = fmaf(kf[7], qb.w, d[3]);
}
__device__ __forceinline__ float da_dot_finish(const float (&d)[4]) { return (d[0] + d[1]) + (d[2] + d[3]); }

// shared-memory working set of one task (scores[] follows it: Smax floats)
struct DecodeAttnSmem {
  __align__(16) float q[DA_DH];
  __align__(16) __half k[DA_DH];
  __align__(16) __half v[DA_DH];
  float red[4];
  float acc[4][DA_DH];
};

template <class SyncF>
__device__ __forceinline__ void decode_attn_task(const DecodeAttnParams& p, int h, int b, int tid, DecodeAttnSmem& sm,
                                                 float* s_scores, SyncF sync) {
  float* s_q = sm.q;
  __half* s_k = sm.k;
  __half* s_v = sm.v;
  float* s_red = sm.red;
  float (*s_acc)[DA_DH] = sm.acc;
  const int warp = tid >> 5, lane = tid & 31;
  const int HD = p.H * DA_DH;
  const __half* row = p.qkv + (size_t)b * p.ldq;
  const int off = p.cache_off ? __ldcg(p.cache_off) : p.cache_off_host;
  int kvl = p.kv_len ? __ldcg(p.kv_len + b) : off + 1;
  if (kvl > p.Smax) kvl = p.Smax;
  __half* kbase = p.kcache + (size_t)b * p.c_bs + h * DA_DH;
  __half* vbase = p.vcache + (size_t)b * p.c_bs + h * DA_DH;

  // ---- phase 0: q / k / v of the new token: LoRA, rotation, cache append (one rotary pair per thread)
  if (tid < DA_DH / 2) {
    const int j = tid, half = DA_DH / 2;
    const int pos = __ldcg(p.pos + b);
    const float c = p.cos_t[(size_t)pos * half + j], sn = p.sin_t[(size_t)pos * half + j];
    // .cg loads: inside the persistent decode kernel the qkv row was written by other CTAs a moment ago (L1 is not coherent)
    float q1 = __half2float(__ldcg(row + h * DA_DH + j)), q2 = __half2float(__ldcg(row + h * DA_DH + half + j));
    const float k1 = __half2float(__ldcg(row + HD + h * DA_DH + j)), k2 = __half2float(__ldcg(row + HD + h * DA_DH + half + j));
    float v1 = __half2float(__ldcg(row + 2 * HD + h * DA_DH + j)), v2 = __half2float(__ldcg(row + 2 * HD + h * DA_DH + half + j));
    if (p.lora_r) {
      float xq[8], xv[8];
      da_unpack8(__ldcg(reinterpret_cast<const uint4*>(row + 3 * HD)), xq);
      da_unpack8(__ldcg(reinterpret_cast<const uint4*>(row + 3 * HD + 8)), xv);
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
  sync();

  // ---- phase 1: scores, one key per thread: the whole 256-byte K row is requested at once (16 independent 16-byte loads
  // in flight per thread), so a pass over the cache costs about one memory round trip instead of one per key
  for (int j = tid; j < kvl; j += DA_THREADS) {
    const uint4* kp = reinterpret_cast<const uint4*>((j == off) ? s_k : kbase + (size_t)j * p.c_ts);
    uint4 kv[DA_DH / 8];
#pragma unroll
    for (int i = 0; i < DA_DH / 8; ++i) kv[i] = kp[i];
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DA_DH / 8; ++i) da_dot8(kv[i], s_q + i * 8, d);
    s_scores[j] = da_dot_finish(d) * p.scale;
  }
  sync();

  // ---- softmax statistics (fp32)
  float m = -INFINITY;
  for (int j = tid; j < kvl; j += DA_THREADS) m = fmaxf(m, s_scores[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  sync();
  m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  sync();
  float l = 0.f;
  for (int j = tid; j < kvl; j += DA_THREADS) {
    const float e = __expf(s_scores[j] - m);
    s_scores[j] = e;
    l += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) s_red[warp] = l;
  sync();
  l = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);

  // ---- phase 2: O = P V. Warp w owns keys j = w (mod 4); lane l owns dims [4l, 4l + 4); 16 keys (16 independent 8-byte
  // loads per lane) in flight per step
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = warp; j0 < kvl; j0 += 64) {
    uint2 vv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = j0 + 4 * u;
      const __half* vp = (j < kvl) ? ((j == off) ? s_v : vbase + (size_t)j * p.c_ts) : s_v;
      vv[u] = *reinterpret_cast<const uint2*>(vp + lane * 4);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = j0 + 4 * u;
      if (j < kvl) {
        const float pj = s_scores[j];
        const __half2* h2 = reinterpret_cast<const __half2*>(&vv[u]);
        const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]);
        acc[0] = fmaf(pj, a.x, acc[0]);
        acc[1] = fmaf(pj, a.y, acc[1]);
        acc[2] = fmaf(pj, b.x, acc[2]);
        acc[3] = fmaf(pj, b.y, acc[3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) s_acc[warp][lane * 4 + i] = acc[i];
  if (p.next_layer_stride) {
    // The next layer's attention reads the same (head, row) slice of ITS cache ~100 us from now. At decode the cache is
    // touched once per step, so those lines sit in DRAM and, under the weight stream, a miss costs microseconds of a
    // latency-bound kernel: ask L2 for them now (weights are loaded evict-first, so the lines survive until then).
    const __half* kn = kbase + p.next_layer_stride;
    const __half* vn = vbase + p.next_layer_stride;
    for (int j = tid; j < 2 * kvl; j += DA_THREADS) {  // one 128-byte line = half a K or V row
      const size_t o = (size_t)(j >> 1) * p.c_ts + (j & 1) * 64;
      prefetch_l2(kn + o);
      prefetch_l2(vn + o);
    }
  }
  sync();
  {
    const float o = (s_acc[0][tid] + s_acc[1][tid]) + (s_acc[2][tid] + s_acc[3][tid]);
    p.out[(size_t)b * p.ldo + h * DA_DH + tid] = __float2half_rn(l > 0.f ? o / l : 0.f);
  }
}

}  // namespace myr
