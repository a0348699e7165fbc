// Backward-pass row / elementwise kernels (HBM-bound, coalesced, 16-byte vectorised where the layout allows):
// clamp-CE loss fwd+bwd, LayerNorm/RMSNorm backward, SwiGLU / GELU backward, RoPE backward, masked row softmax and
// its backward (attention backward = batched tcgen05 GEMMs around these), row gather/scatter, LoraAdaptorV2 wgrad,
// column sums (bias / base_prompts grads), fused AdamW over the flat parameter buffer.
#include <type_traits>
#include "common.h"
#include "ptx.cuh"

namespace myr {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block reductions for 256-thread blocks; result broadcast to all threads
__device__ __forceinline__ float block_sum_t(float v, float* red) {
  v = warp_sum_t(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_max_t(float v, float* red) {
  v = warp_max_t(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
  return t;
}

static inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------------
// clamp_CE_loss (modeling_llama.py:718-728): softmax -> clamp[1e-7, 1-1e-7] -> log -> NLL(mean over labels != -100)
// fwd: one CTA per row; stats[row] = (max, sumexp); row_loss[row] = -log(clamp(p_y)) or 0 when ignored
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits, long long ld, int V,
                                                     const long long* __restrict__ labels, float* __restrict__ row_loss,
                                                     float* __restrict__ stats) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float* row = logits + (size_t)r * ld;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, row[i]);
  m = block_max_t(m, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(row[i] - m);
  s = block_sum_t(s, red);
  if (threadIdx.x == 0) {
    stats[2 * r] = m;
    stats[2 * r + 1] = s;
    const long long y = labels[r];
    float l = 0.f;
    if (y >= 0) {
      float p = expf(row[y] - m) / s;
      p = fminf(fmaxf(p, 1e-7f), 1.f - 1e-7f);
      l = -logf(p);
    }
    row_loss[r] = l;
  }
}

// loss_out[0] = mean loss, loss_out[1] = number of supervised rows (fixed summation order: deterministic)
__global__ void __launch_bounds__(256) ce_reduce_kernel(const float* __restrict__ row_loss, const long long* __restrict__ labels,
                                                        int R, float* __restrict__ loss_out) {
  __shared__ float red[8];
  float s = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    if (labels[i] >= 0) {
      s += row_loss[i];
      c += 1.f;
    }
  }
  s = block_sum_t(s, red);
  c = block_sum_t(c, red);
  if (threadIdx.x == 0) {
    loss_out[0] = c > 0.f ? s / c : 0.f;
    loss_out[1] = c;
  }
}

// dlogits[r, j] = loss_scale / count * (p_j - [j == y]) if the row is supervised and p_y is strictly inside the clamp
// interval (the clamp has zero gradient outside), else 0. fp16 output (GEMM operand of the lm_head dgrad).
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, long long ld, int V,
                                                     const long long* __restrict__ labels, const float* __restrict__ stats,
                                                     const float* __restrict__ loss_out, float loss_scale,
                                                     __half* __restrict__ dlogits, long long ldd) {
  const int r = blockIdx.x;
  const float* row = logits + (size_t)r * ld;
  __half* drow = dlogits + (size_t)r * ldd;
  const long long y = labels[r];
  const float m = stats[2 * r], s = stats[2 * r + 1];
  float g = 0.f;
  if (y >= 0) {
    const float py = expf(row[y] - m) / s;
    if (py > 1e-7f && py < 1.f - 1e-7f) g = loss_scale / loss_out[1];
  }
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float d = 0.f;
    if (g != 0.f) d = g * (expf(row[i] - m) * inv - (i == y ? 1.f : 0.f));
    drow[i] = __float2half_rn(d);
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm backward w.r.t. the input (gamma/beta are frozen on this path):
//   xhat = (x - mean) * rstd;  gy = gamma * dy;  dx = rstd * (gy - mean(gy) - xhat * mean(gy * xhat))   (LN)
//                                                dx = rstd * (gy - xhat * mean(gy * xhat))               (RMS, mean = 0)
// x fp32 pre-norm rows (statistics recomputed), dy fp16 or fp32, optional `add` (fp32 residual-branch gradient)
// out32 = dx + add. One CTA per row, D <= 4096.
// ---------------------------------------------------------------------------------------------------
struct NormBwdParams {
  const float* x; long long ldx;
  const void* dy; int dy_dtype; long long lddy;
  const float* gamma; float eps; int rms; int D;
  const float* add; long long ldadd;
  float* out32; long long ldo;
  __half* out16; long long ldo16;
};

__global__ void __launch_bounds__(256) norm_bwd_kernel(const NormBwdParams p) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const float* x = p.x + (size_t)row * p.ldx;
  constexpr int MAXI = 16;
  float xv[MAXI], gv[MAXI];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int c = threadIdx.x + i * 256;
    xv[i] = c < p.D ? x[c] : 0.f;
    s += xv[i];
  }
  float mean = 0.f;
  if (!p.rms) mean = block_sum_t(s, red) / p.D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < p.D) q += (xv[i] - mean) * (xv[i] - mean);
  }
  const float rstd = rsqrtf(block_sum_t(q, red) / p.D + p.eps);
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int c = threadIdx.x + i * 256;
    gv[i] = 0.f;
    if (c < p.D) {
      const float dy = p.dy_dtype == MYR_F32 ? reinterpret_cast<const float*>(p.dy)[(size_t)row * p.lddy + c]
                                             : __half2float(reinterpret_cast<const __half*>(p.dy)[(size_t)row * p.lddy + c]);
      gv[i] = dy * p.gamma[c];
      xv[i] = (xv[i] - mean) * rstd;
      a += gv[i];
      b += gv[i] * xv[i];
    }
  }
  a = p.rms ? 0.f : block_sum_t(a, red) / p.D;
  b = block_sum_t(b, red) / p.D;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < p.D) {
      float dx = rstd * (gv[i] - a - xv[i] * b);
      if (p.add) dx += p.add[(size_t)row * p.ldadd + c];
      if (p.out32) p.out32[(size_t)row * p.ldo + c] = dx;
      if (p.out16) p.out16[(size_t)row * p.ldo16 + c] = __float2half_rn(dx);
    }
  }
}

// the same with 16-byte accesses (D and every row stride a multiple of 4, 16-byte aligned bases): one thread = up to four float4
// columns; the scalar version ran at a quarter of the HBM rate (26.8 us for the 47 MB of a 656 x 4096 fp32 launch)
__global__ void __launch_bounds__(256) norm_bwd_vec4_kernel(const NormBwdParams p) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const int D4 = p.D >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(p.x + (size_t)row * p.ldx);
  constexpr int MAXV = 4;
  float4 xv[MAXV], gv[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * 256;
    xv[i] = c < D4 ? x4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
  }
  float mean = 0.f;
  if (!p.rms) mean = block_sum_t(s, red) / p.D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < D4) {
      const float a0 = xv[i].x - mean, a1 = xv[i].y - mean, a2 = xv[i].z - mean, a3 = xv[i].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  }
  const float rstd = rsqrtf(block_sum_t(q, red) / p.D + p.eps);
  float a = 0.f, b = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(p.gamma);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * 256;
    gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D4) {
      float4 dy;
      if (p.dy_dtype == MYR_F32) {
        dy = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dy) + (size_t)row * p.lddy)[c];
      } else {
        const uint2 raw = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.dy) + (size_t)row * p.lddy)[c];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        dy = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      const float4 gm = g4[c];
      gv[i] = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
      xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd, (xv[i].w - mean) * rstd);
      a += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      b += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
    }
  }
  a = p.rms ? 0.f : block_sum_t(a, red) / p.D;
  b = block_sum_t(b, red) / p.D;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < D4) {
      float4 dx = make_float4(rstd * (gv[i].x - a - xv[i].x * b), rstd * (gv[i].y - a - xv[i].y * b), rstd * (gv[i].z - a - xv[i].z * b),
                              rstd * (gv[i].w - a - xv[i].w * b));
      if (p.add) {
        const float4 ad = reinterpret_cast<const float4*>(p.add + (size_t)row * p.ldadd)[c];
        dx.x += ad.x; dx.y += ad.y; dx.z += ad.z; dx.w += ad.w;
      }
      if (p.out32) reinterpret_cast<float4*>(p.out32 + (size_t)row * p.ldo)[c] = dx;
      if (p.out16) {
        const __half2 lo = __floats2half2_rn(dx.x, dx.y), hi = __floats2half2_rn(dx.z, dx.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(p.out16 + (size_t)row * p.ldo16)[c] = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// activation backward
// ---------------------------------------------------------------------------------------------------
// SwiGLU: dgu[:, :I] = dact * u * (sig + g * sig * (1 - sig)); dgu[:, I:] = dact * silu(g)
__global__ void swiglu_bwd_kernel(const __half* __restrict__ gu, long long ldg, const __half* __restrict__ dact, long long lda,
                                  __half* __restrict__ dgu, long long ldd, int T, int I) {
  const long long total = (long long)T * I;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / I;
    const int c = (int)(i % I);
    const float g = __half2float(gu[t * ldg + c]), u = __half2float(gu[t * ldg + I + c]), d = __half2float(dact[t * lda + c]);
    const float sig = 1.f / (1.f + expf(-g));
    dgu[t * ldd + c] = __float2half_rn(d * u * (sig + g * sig * (1.f - sig)));
    dgu[t * ldd + I + c] = __float2half_rn(d * g * sig);
  }
}
// the same with 16-byte accesses: one thread = 8 consecutive columns of one token (I and the row strides multiples of 8)
__global__ void swiglu_bwd_vec8_kernel(const __half* __restrict__ gu, long long ldg, const __half* __restrict__ dact, long long lda,
                                       __half* __restrict__ dgu, long long ldd, int T, int I) {
  const int iv = I >> 3;
  const long long total = (long long)T * iv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / iv;
    const int c = (int)(i - t * iv) << 3;
    const uint4 gr = *reinterpret_cast<const uint4*>(gu + t * ldg + c), ur = *reinterpret_cast<const uint4*>(gu + t * ldg + I + c);
    const uint4 dr = *reinterpret_cast<const uint4*>(dact + t * lda + c);
    const __half2* g2 = reinterpret_cast<const __half2*>(&gr);
    const __half2* u2 = reinterpret_cast<const __half2*>(&ur);
    const __half2* d2 = reinterpret_cast<const __half2*>(&dr);
    uint4 og, ou;
    __half2* og2 = reinterpret_cast<__half2*>(&og);
    __half2* ou2 = reinterpret_cast<__half2*>(&ou);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = __half22float2(g2[k]), u = __half22float2(u2[k]), d = __half22float2(d2[k]);
      const float s0 = 1.f / (1.f + expf(-g.x)), s1 = 1.f / (1.f + expf(-g.y));
      og2[k] = __floats2half2_rn(d.x * u.x * (s0 + g.x * s0 * (1.f - s0)), d.y * u.y * (s1 + g.y * s1 * (1.f - s1)));
      ou2[k] = __floats2half2_rn(d.x * g.x * s0, d.y * g.y * s1);
    }
    *reinterpret_cast<uint4*>(dgu + t * ldd + c) = og;
    *reinterpret_cast<uint4*>(dgu + t * ldd + I + c) = ou;
  }
}

// GELU(erf) fwd from a saved pre-activation, and its backward: dpre = dy * (Phi(x) + x * phi(x))
__global__ void gelu_fwd_kernel(const __half* __restrict__ pre, __half* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(gelu_erf(__half2float(pre[i])));
}
__global__ void gelu_bwd_kernel(const __half* __restrict__ pre, const __half* __restrict__ dy, __half* __restrict__ dpre,
                                long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = __half2float(pre[i]);
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    dpre[i] = __float2half_rn(__half2float(dy[i]) * (cdf + x * pdf));
  }
}

// RoPE backward on the q and k thirds of dqkv rows (inverse rotation), in place.
__global__ void rope_bwd_kernel(__half* __restrict__ dqkv, long long ld, int T, int H, int dh, const int* __restrict__ pos,
                                const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
  const int half = dh >> 1;
  const long long total = (long long)T * 2 * H * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % half);
    long long r = i / half;
    const int h = (int)(r % H);
    r /= H;
    const int which = (int)(r % 2);
    const long long t = r / 2;
    const float c = cos_t[(size_t)pos[t] * half + j], s = sin_t[(size_t)pos[t] * half + j];
    __half* p = dqkv + t * ld + (size_t)which * H * dh + h * dh;
    const float d1 = __half2float(p[j]), d2 = __half2float(p[half + j]);
    p[j] = __float2half_rn(d1 * c + d2 * s);
    p[half + j] = __float2half_rn(d2 * c - d1 * s);
  }
}

// 16-byte variant: one thread = 8 rotary pairs (dh / 2 a multiple of 8, row stride a multiple of 8, 16-byte aligned base)
__global__ void rope_bwd_vec8_kernel(__half* __restrict__ dqkv, long long ld, int T, int H, int dh, const int* __restrict__ pos,
                                     const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
  const int half = dh >> 1, hv = half >> 3;
  const long long total = (long long)T * 2 * H * hv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % hv) << 3;
    long long r = i / hv;
    const int h = (int)(r % H);
    r /= H;
    const int which = (int)(r % 2);
    const long long t = r / 2;
    const float4* cp = reinterpret_cast<const float4*>(cos_t + (size_t)pos[t] * half + j);
    const float4* sp = reinterpret_cast<const float4*>(sin_t + (size_t)pos[t] * half + j);
    const float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    __half* p = dqkv + t * ld + (size_t)which * H * dh + h * dh;
    uint4 r1 = *reinterpret_cast<const uint4*>(p + j), r2 = *reinterpret_cast<const uint4*>(p + half + j);
    __half2* a2 = reinterpret_cast<__half2*>(&r1);
    __half2* b2 = reinterpret_cast<__half2*>(&r2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 d1 = __half22float2(a2[k]), d2 = __half22float2(b2[k]);
      a2[k] = __floats2half2_rn(d1.x * c[2 * k] + d2.x * sn[2 * k], d1.y * c[2 * k + 1] + d2.y * sn[2 * k + 1]);
      b2[k] = __floats2half2_rn(d2.x * c[2 * k] - d1.x * sn[2 * k], d2.y * c[2 * k + 1] - d1.y * sn[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(p + j) = r1;
    *reinterpret_cast<uint4*>(p + half + j) = r2;
  }
}

// ---------------------------------------------------------------------------------------------------
// attention backward helpers on materialised score rows [n_rows = B*H*Sq, ld] (Sq, Skv <= a few hundred)
//   softmax: P = softmax(scale * S + mask) fp16; key j visible iff j < kv_len[b] && (!causal || j <= i); pad cols -> 0
//   bwd    : dS = scale * P * (dP - sum_j dP * P) fp16
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long lds, __half* __restrict__ P,
                                                           long long ldp, int H, int Sq, int Skv, int cols, float scale,
                                                           int causal, const int* __restrict__ kv_len) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const int i = (int)(row % Sq);
  const int b = (int)(row / ((long long)Sq * H));
  const int kvl = kv_len ? min(kv_len[b], Skv) : Skv;
  const int lim = causal ? min(kvl, i + 1) : kvl;
  const float* s = S + row * lds;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < lim; j += blockDim.x) m = fmaxf(m, s[j] * scale);
  m = block_max_t(m, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < lim; j += blockDim.x) sum += expf(s[j] * scale - m);
  sum = block_sum_t(sum, red);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x)
    P[row * ldp + j] = __float2half_rn(j < lim ? expf(s[j] * scale - m) * inv : 0.f);
}

__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const __half* __restrict__ P, long long ldp,
                                                               const float* __restrict__ dP, long long lddp,
                                                               __half* __restrict__ dS, long long lds, int cols, float scale) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  float acc = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    const float p = __half2float(P[row * ldp + j]);
    if (p != 0.f) acc += p * dP[row * lddp + j];  // masked / padded columns of dP may hold uninitialised values
  }
  acc = block_sum_t(acc, red);
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    const float p = __half2float(P[row * ldp + j]);
    dS[row * lds + j] = __float2half_rn(p != 0.f ? scale * p * (dP[row * lddp + j] - acc) : 0.f);
  }
}

// Short rows (the training sequences: cols <= 256): one WARP per row, the row in registers, shuffle reductions - a 256-thread block
// with two block-wide reductions per 164-column row spent its time in barriers (31 / 24 us per launch for 4 MB of scores).
__global__ void __launch_bounds__(256) softmax_rows_warp_kernel(const float* __restrict__ S, long long lds, __half* __restrict__ P,
                                                                long long ldp, long long n_rows, int H, int Sq, int Skv, int cols,
                                                                float scale, int causal, const int* __restrict__ kv_len) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int i = (int)(row % Sq);
  const int b = (int)(row / ((long long)Sq * H));
  const int kvl = kv_len ? min(kv_len[b], Skv) : Skv;
  const int lim = causal ? min(kvl, i + 1) : kvl;
  const float* s = S + row * lds;
  float v[8];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = lane + 32 * k;
    v[k] = j < lim ? s[j] * scale : -INFINITY;
    m = fmaxf(m, v[k]);
  }
  m = warp_max_t(m);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (lane + 32 * k < lim) ? expf(v[k] - m) : 0.f;
    sum += v[k];
  }
  sum = warp_sum_t(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = lane + 32 * k;
    if (j < cols) P[row * ldp + j] = __float2half_rn(v[k] * inv);
  }
}
__global__ void __launch_bounds__(256) softmax_bwd_rows_warp_kernel(const __half* __restrict__ P, long long ldp,
                                                                    const float* __restrict__ dP, long long lddp,
                                                                    __half* __restrict__ dS, long long lds, long long n_rows, int cols,
                                                                    float scale) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  float pv[8], dv[8];
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = lane + 32 * k;
    pv[k] = j < cols ? __half2float(P[row * ldp + j]) : 0.f;
    dv[k] = pv[k] != 0.f ? dP[row * lddp + j] : 0.f;  // masked / padded columns of dP may hold uninitialised values
    acc += pv[k] * dv[k];
  }
  acc = warp_sum_t(acc);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = lane + 32 * k;
    if (j < cols) dS[row * lds + j] = __float2half_rn(pv[k] != 0.f ? scale * pv[k] * (dv[k] - acc) : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------
// row gather / scatter with cast:  dst[r, :] = src[idx[r], :]   or   dst[idx[r], :] = src[r, :]
// ---------------------------------------------------------------------------------------------------
__global__ void index_rows_kernel(const void* __restrict__ src, int src_dtype, long long src_ld, void* __restrict__ dst,
                                  int dst_dtype, long long dst_ld, const int* __restrict__ idx, int R, int D, int scatter) {
  const long long total = (long long)R * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int c = (int)(i % D);
    const long long sr = scatter ? r : idx[r], dr = scatter ? idx[r] : r;
    const float v = src_dtype == MYR_F32 ? reinterpret_cast<const float*>(src)[sr * src_ld + c]
                                         : __half2float(reinterpret_cast<const __half*>(src)[sr * src_ld + c]);
    if (dst_dtype == MYR_F32)
      reinterpret_cast<float*>(dst)[dr * dst_ld + c] = v;
    else
      reinterpret_cast<__half*>(dst)[dr * dst_ld + c] = __float2half_rn(v);
  }
}

// column sums over grouped rows: out[c] (+)= scale * sum_{g < groups, r < rows} src[g * gs + r * ld + c]
// (bias gradients; base_prompts gradient with rows = 1 per prompt). Deterministic: one thread per column.
__global__ void colsum_kernel(const void* __restrict__ src, int src_dtype, long long ld, long long gs, int groups, int rows,
                              int D, float scale, float* __restrict__ out, int accumulate) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int g = 0; g < groups; ++g)
      for (int r = 0; r < rows; ++r) {
        const long long o = g * gs + r * ld + c;
        a += src_dtype == MYR_F32 ? reinterpret_cast<const float*>(src)[o] : __half2float(reinterpret_cast<const __half*>(src)[o]);
      }
    out[c] = (accumulate ? out[c] : 0.f) + scale * a;
  }
}

// Same sums with the rows spread over a block (the default; MYR_COLSUM_BLOCK=0 selects the one-thread-per-column kernel
// above). One block per `cpb` columns (32, or fewer where 32 would leave most SMs without a block: a 64-channel conv bias over
// 12 544 rows ran 235 us on two blocks), 1024 threads: thread (ry, cx) adds the rows ry, ry + 1024 / cpb, ... of column cx, the
// 1024 / cpb partial sums per column meet in shared memory and are added in ry order: deterministic.
__global__ void __launch_bounds__(1024) colsum_block_kernel(const void* __restrict__ src, int src_dtype, long long ld, long long gs, int groups,
                                                      int rows, int D, float scale, float* __restrict__ out, int accumulate, int cpb) {
  __shared__ float part[1024];
  const int cx = threadIdx.x % cpb, ry = threadIdx.x / cpb, nry = 1024 / cpb;
  const int c = blockIdx.x * cpb + cx;
  float a = 0.f;
  if (c < D) {
    const long long total = (long long)groups * rows;
    if (groups == 1 && src_dtype == MYR_F16) {  // the bias gradients: no 64-bit division per element
      const __half* s16 = reinterpret_cast<const __half*>(src) + c;
#pragma unroll 8
      for (int r = ry; r < rows; r += nry) a += __half2float(s16[(long long)r * ld]);
    } else {
      for (long long i = ry; i < total; i += nry) {
        const long long g = i / rows, r = i - g * rows;
        const long long o = g * gs + r * ld + c;
        a += src_dtype == MYR_F32 ? reinterpret_cast<const float*>(src)[o] : __half2float(reinterpret_cast<const __half*>(src)[o]);
      }
    }
  }
  part[threadIdx.x] = a;
  __syncthreads();
  if (ry == 0 && c < D) {
    float t = 0.f;
    for (int w = 0; w < nry; ++w) t += part[w * cpb + cx];
    out[c] = (accumulate ? out[c] : 0.f) + scale * t;
  }
}

// ---------------------------------------------------------------------------------------------------
// LoraAdaptorV2 backward (networks.py:81-93, y = x + W2 (W1 x)): weights only (x comes from the frozen ViT).
//   pass 1 (one warp per row): t = W1 x, dt = W2^T dy          -> td [rows, 2 * rank]
//   pass 2 (one thread per column i): dW2[i, r] = sum_rows dy[row, i] t[row, r]; dW1[r, i] = sum_rows dt[row, r] x[row, i]
// ---------------------------------------------------------------------------------------------------
__global__ void adaptor_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ w1,
                                        const float* __restrict__ w2, float* __restrict__ td, int rows, int D, int rank) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float t[4] = {0.f, 0.f, 0.f, 0.f}, dt[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = lane; i < D; i += 32) {
    const float xv = x[(size_t)warp * D + i], dv = dy[(size_t)warp * D + i];
    for (int r = 0; r < rank; ++r) {
      t[r] += xv * w1[(size_t)r * D + i];
      dt[r] += dv * w2[(size_t)i * rank + r];
    }
  }
  for (int r = 0; r < rank; ++r) {
    t[r] = warp_sum_t(t[r]);
    dt[r] = warp_sum_t(dt[r]);
  }
  if (lane == 0)
    for (int r = 0; r < rank; ++r) {
      td[(size_t)warp * 2 * rank + r] = t[r];
      td[(size_t)warp * 2 * rank + rank + r] = dt[r];
    }
}
// a block = 32 columns x 16 row slices; the slices meet in shared memory and are added in slice order (one thread per column walking
// all 1028 rows took 585 us for 11.6 MB)
__global__ void __launch_bounds__(512) adaptor_bwd_cols_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ td, float* __restrict__ dw1,
                                                              float* __restrict__ dw2, int rows, int D, int rank, float scale) {
  __shared__ float red[16][32][9];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  if (i < D) {
    const int per = (rows + 15) / 16;
    const int r0 = sl * per, r1 = min(rows, r0 + per);
#pragma unroll 4
    for (int row = r0; row < r1; ++row) {
      const float xv = __ldg(x + (size_t)row * D + i), dv = __ldg(dy + (size_t)row * D + i);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < rank) {
          a2[r] = fmaf(dv, __ldg(td + (size_t)row * 2 * rank + r), a2[r]);
          a1[r] = fmaf(__ldg(td + (size_t)row * 2 * rank + rank + r), xv, a1[r]);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    red[sl][cl][r] = a1[r];
    red[sl][cl][4 + r] = a2[r];
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    const int c = threadIdx.x >> 3, q = threadIdx.x & 7, r = q & 3;
    const int col = blockIdx.x * 32 + c;
    if (col < D && r < rank) {
      float tot = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < 16; ++s2) tot += red[s2][c][q];
      if (q < 4)
        dw1[(size_t)r * D + col] = scale * tot;
      else
        dw2[(size_t)col * rank + r] = scale * tot;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// fused AdamW over the flat fp32 parameter / gradient / moment buffers (runner_base.py:105-139 semantics:
// weight decay only where wd_mask != 0), with gradient unscaling (GradScaler.unscale_) folded in.
// found_inf (device int) != 0 skips the update (GradScaler.step).
// ---------------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             const unsigned char* __restrict__ wd_mask, long long n, float lr, float beta1, float beta2,
                             float eps, float wd, float bc1, float bc2, float inv_scale, const int* __restrict__ found_inf) {
  if (found_inf && *found_inf) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * inv_scale;
    float pv = p[i];
    if (wd_mask[i]) pv *= 1.f - lr * wd;
    const float mi = beta1 * m[i] + (1.f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    p[i] = pv - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  }
}
// four parameters per thread (16-byte accesses on the five streams: 1.8 GB per step for the 115 M trainable parameters)
__global__ void adamw_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                  const uchar4* __restrict__ wd_mask, long long n4, float lr, float beta1, float beta2, float eps,
                                  float wd, float bc1, float bc2, float inv_scale, const int* __restrict__ found_inf) {
  if (found_inf && *found_inf) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 gv = g[i];
    float4 pv = p[i], mv = m[i], vv = v[i];
    const uchar4 w = wd_mask[i];
    const float dec = 1.f - lr * wd;
    float* pp = reinterpret_cast<float*>(&pv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    const unsigned char wm[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gp[k] * inv_scale;
      float x = pp[k];
      if (wm[k]) x *= dec;
      const float mi = beta1 * mp[k] + (1.f - beta1) * gr;
      const float vi = beta2 * vp[k] + (1.f - beta2) * gr * gr;
      mp[k] = mi;
      vp[k] = vi;
      pp[k] = x - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    }
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
  }
}
__global__ void check_finite_kernel(const float* __restrict__ g, long long n, int* __restrict__ found_inf) {
  int bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!isfinite(g[i])) bad = 1;
  if (bad) atomicExch(found_inf, 1);
}


// ---------------------------------------------------------------------------------------------------
// LoRA dropout (peft: lora_dropout = 0.05 on the INPUT of each LoRA branch, myriad.py:171-178). Counter-based mask:
// element idx of a step's stream keeps its value iff hash(seed, idx) >= p * 2^32, kept values are scaled by 1 / (1 - p). The
// backward recomputes the same mask from (seed, offset), nothing is stored. The hash is a 32-bit multiply-xorshift mixer (two
// rounds) over the index folded with the seed: ~10 integer instructions per element. (The first version was the 64-bit splitmix
// finaliser, ~30 instructions: at 2 x 2.7 M elements per LoRA kernel the mask generation alone was ~18 us of each launch.)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t drop_hash(unsigned long long seed, unsigned long long idx) {
  uint32_t h = (uint32_t)idx * 0x9E3779B1u + ((uint32_t)(idx >> 32) + (uint32_t)seed) * 0x85EBCA77u + (uint32_t)(seed >> 32);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
__global__ void dropout_fwd_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ out, long long ldo, int rows, int D,
                                   uint32_t thresh, float keep_scale, unsigned long long seed, unsigned long long offset) {
  const long long n = (long long)rows * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / D;
    const int c = (int)(i - t * D);
    const bool keep = drop_hash(seed, offset + (unsigned long long)i) >= thresh;
    out[t * ldo + c] = keep ? __float2half_rn(__half2float(x[t * ldx + c]) * keep_scale) : __float2half_rn(0.f);
  }
}
__global__ void dropout_bwd_add_kernel(const float* __restrict__ g, long long ldg, float* __restrict__ acc, long long lda, int rows, int D,
                                       uint32_t thresh, float keep_scale, unsigned long long seed, unsigned long long offset) {
  const long long n = (long long)rows * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / D;
    const int c = (int)(i - t * D);
    if (drop_hash(seed, offset + (unsigned long long)i) >= thresh) acc[t * lda + c] += g[t * ldg + c] * keep_scale;
  }
}
__global__ void dropout_mask_kernel(unsigned char* __restrict__ out, long long n, uint32_t thresh, unsigned long long seed,
                                    unsigned long long offset) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = drop_hash(seed, offset + (unsigned long long)i) >= thresh ? 1 : 0;
}

}  // namespace myr

using namespace myr;

#define STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)

extern "C" int myr_clamp_ce_fwd(const void* logits, int64_t ld, int32_t R, int32_t V, const void* labels, void* row_loss,
                                void* stats, void* loss_out, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(logits && labels && row_loss && stats && loss_out && R > 0 && V > 0, "clamp_ce_fwd: bad arguments");
  ce_fwd_kernel<<<R, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), ld, V, reinterpret_cast<const long long*>(labels),
                                       reinterpret_cast<float*>(row_loss), reinterpret_cast<float*>(stats));
  MYR_CHECK_LAUNCH();
  ce_reduce_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const float*>(row_loss), reinterpret_cast<const long long*>(labels), R,
                                          reinterpret_cast<float*>(loss_out));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_clamp_ce_bwd(const void* logits, int64_t ld, int32_t R, int32_t V, const void* labels, const void* stats,
                                const void* loss_out, float loss_scale, void* dlogits, int64_t ldd, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(logits && labels && stats && loss_out && dlogits && R > 0 && V > 0, "clamp_ce_bwd: bad arguments");
  ce_bwd_kernel<<<R, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), ld, V, reinterpret_cast<const long long*>(labels),
                                       reinterpret_cast<const float*>(stats), reinterpret_cast<const float*>(loss_out), loss_scale,
                                       reinterpret_cast<__half*>(dlogits), ldd);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_norm_bwd(const void* x, int64_t ldx, const void* dy, int32_t dy_dtype, int64_t lddy, const void* gamma,
                            float eps, int32_t rms, int32_t rows, int32_t D, const void* add, int64_t ldadd, void* out32,
                            int64_t ldo, void* out16, int64_t ldo16, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(x && dy && gamma && rows > 0 && D > 0 && D <= 4096 && (out32 || out16), "norm_bwd: bad arguments (D <= 4096)");
  NormBwdParams p;
  p.x = reinterpret_cast<const float*>(x); p.ldx = ldx; p.dy = dy; p.dy_dtype = dy_dtype; p.lddy = lddy;
  p.gamma = reinterpret_cast<const float*>(gamma); p.eps = eps; p.rms = rms; p.D = D;
  p.add = reinterpret_cast<const float*>(add); p.ldadd = ldadd;
  p.out32 = reinterpret_cast<float*>(out32); p.ldo = ldo; p.out16 = reinterpret_cast<__half*>(out16); p.ldo16 = ldo16;
  auto al = [](const void* q, int a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  const bool vec = D % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && (add == nullptr || ldadd % 4 == 0) && (out32 == nullptr || ldo % 4 == 0) &&
                   (out16 == nullptr || ldo16 % 4 == 0) && al(x, 16) && al(dy, dy_dtype == MYR_F32 ? 16 : 8) && al(gamma, 16) && al(add, 16) &&
                   al(out32, 16) && al(out16, 8);
  if (vec)
    norm_bwd_vec4_kernel<<<rows, 256, 0, stream>>>(p);
  else
    norm_bwd_kernel<<<rows, 256, 0, stream>>>(p);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_swiglu_bwd(const void* gate_up, int64_t ld_gu, const void* dact, int64_t ld_da, void* dgu, int64_t ld_dgu,
                              int32_t T, int32_t I, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(gate_up && dact && dgu && T > 0 && I > 0, "swiglu_bwd: bad arguments");
  const bool vec = I % 8 == 0 && ld_gu % 8 == 0 && ld_da % 8 == 0 && ld_dgu % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(gate_up) | reinterpret_cast<uintptr_t>(dact) | reinterpret_cast<uintptr_t>(dgu)) & 15) == 0;
  if (vec)
    swiglu_bwd_vec8_kernel<<<grid_for((long long)T * (I / 8), 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(gate_up), ld_gu,
                                                                                     reinterpret_cast<const __half*>(dact), ld_da,
                                                                                     reinterpret_cast<__half*>(dgu), ld_dgu, T, I);
  else
    swiglu_bwd_kernel<<<grid_for((long long)T * I, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(gate_up), ld_gu,
                                                                          reinterpret_cast<const __half*>(dact), ld_da,
                                                                          reinterpret_cast<__half*>(dgu), ld_dgu, T, I);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

static inline uint32_t drop_thresh(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

extern "C" int myr_dropout_fwd(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t rows, int32_t D, float p, uint64_t seed,
                               uint64_t offset, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(x && out && rows > 0 && D > 0 && p >= 0.f && p < 1.f, "dropout_fwd: bad arguments");
  dropout_fwd_kernel<<<grid_for((long long)rows * D, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(x), ldx,
                                                                             reinterpret_cast<__half*>(out), ldo, rows, D, drop_thresh(p),
                                                                             1.0f / (1.0f - p), seed, offset);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_dropout_bwd_add(const void* g, int64_t ldg, void* acc, int64_t lda, int32_t rows, int32_t D, float p, uint64_t seed,
                                   uint64_t offset, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(g && acc && rows > 0 && D > 0 && p >= 0.f && p < 1.f, "dropout_bwd_add: bad arguments");
  dropout_bwd_add_kernel<<<grid_for((long long)rows * D, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(g), ldg,
                                                                                 reinterpret_cast<float*>(acc), lda, rows, D, drop_thresh(p),
                                                                                 1.0f / (1.0f - p), seed, offset);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_dropout_mask(void* out_u8, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(out_u8 && n > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  dropout_mask_kernel<<<grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<unsigned char*>(out_u8), n, drop_thresh(p), seed, offset);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_gelu_fwd(const void* pre, void* out, int64_t n, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(pre && out && n > 0, "gelu_fwd: bad arguments");
  gelu_fwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(pre), reinterpret_cast<__half*>(out), n);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_gelu_bwd(const void* pre, const void* dy, void* dpre, int64_t n, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(pre && dy && dpre && n > 0, "gelu_bwd: bad arguments");
  gelu_bwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(pre), reinterpret_cast<const __half*>(dy),
                                                        reinterpret_cast<__half*>(dpre), n);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_rope_bwd(void* dqkv, int64_t ld, int32_t T, int32_t H, int32_t dh, const void* pos, const void* cos_table,
                            const void* sin_table, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(dqkv && pos && cos_table && sin_table && T > 0 && dh % 2 == 0, "rope_bwd: bad arguments");
  if ((dh / 2) % 8 == 0 && ld % 8 == 0 && ((reinterpret_cast<uintptr_t>(dqkv) | reinterpret_cast<uintptr_t>(cos_table) |
                                           reinterpret_cast<uintptr_t>(sin_table)) & 15) == 0)
    rope_bwd_vec8_kernel<<<grid_for((long long)T * 2 * H * (dh / 16), 256), 256, 0, stream>>>(
        reinterpret_cast<__half*>(dqkv), ld, T, H, dh, reinterpret_cast<const int*>(pos), reinterpret_cast<const float*>(cos_table),
        reinterpret_cast<const float*>(sin_table));
  else
    rope_bwd_kernel<<<grid_for((long long)T * 2 * H * (dh / 2), 256), 256, 0, stream>>>(
        reinterpret_cast<__half*>(dqkv), ld, T, H, dh, reinterpret_cast<const int*>(pos), reinterpret_cast<const float*>(cos_table),
        reinterpret_cast<const float*>(sin_table));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_softmax_rows(const void* S, int64_t lds, void* P, int64_t ldp, int32_t B, int32_t H, int32_t Sq, int32_t Skv,
                                int32_t cols, float scale, int32_t causal, const void* kv_len, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(S && P && B > 0 && H > 0 && Sq > 0 && Skv > 0 && cols >= Skv, "softmax_rows: bad arguments");
  const long long n_rows = (long long)B * H * Sq;
  if (cols <= 256)
    softmax_rows_warp_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const float*>(S), lds,
                                                                               reinterpret_cast<__half*>(P), ldp, n_rows, H, Sq, Skv, cols,
                                                                               scale, causal, reinterpret_cast<const int*>(kv_len));
  else
    softmax_rows_kernel<<<B * H * Sq, 256, 0, stream>>>(reinterpret_cast<const float*>(S), lds, reinterpret_cast<__half*>(P), ldp, H,
                                                        Sq, Skv, cols, scale, causal, reinterpret_cast<const int*>(kv_len));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_softmax_bwd_rows(const void* P, int64_t ldp, const void* dP, int64_t lddp, void* dS, int64_t lds,
                                    int64_t n_rows, int32_t cols, float scale, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(P && dP && dS && n_rows > 0 && cols > 0, "softmax_bwd_rows: bad arguments");
  if (cols <= 256)
    softmax_bwd_rows_warp_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const __half*>(P), ldp,
                                                                                   reinterpret_cast<const float*>(dP), lddp,
                                                                                   reinterpret_cast<__half*>(dS), lds, n_rows, cols, scale);
  else
    softmax_bwd_rows_kernel<<<(unsigned)n_rows, 256, 0, stream>>>(reinterpret_cast<const __half*>(P), ldp,
                                                                 reinterpret_cast<const float*>(dP), lddp,
                                                                 reinterpret_cast<__half*>(dS), lds, cols, scale);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_index_rows(const void* src, int32_t src_dtype, int64_t src_ld, void* dst, int32_t dst_dtype, int64_t dst_ld,
                              const void* idx, int32_t R, int32_t D, int32_t scatter, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(src && dst && idx && R > 0 && D > 0, "index_rows: bad arguments");
  index_rows_kernel<<<grid_for((long long)R * D, 256), 256, 0, stream>>>(src, src_dtype, src_ld, dst, dst_dtype, dst_ld,
                                                                        reinterpret_cast<const int*>(idx), R, D, scatter);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_colsum(const void* src, int32_t src_dtype, int64_t ld, int64_t group_stride, int32_t groups, int32_t rows,
                          int32_t D, float scale, void* out, int32_t accumulate, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(src && out && groups > 0 && rows > 0 && D > 0, "colsum: bad arguments");
  static int block_variant = -1;
  if (block_variant < 0) {
    const char* e = getenv("MYR_COLSUM_BLOCK");
    block_variant = (e && e[0] == '0') ? 0 : 1;
  }
  if (block_variant)
  {
    int cpb = 32;
    while (cpb > 4 && ceil_div(D, cpb) < 64) cpb >>= 1;
    colsum_block_kernel<<<ceil_div(D, cpb), 1024, 0, stream>>>(src, src_dtype, ld, group_stride, groups, rows, D, scale,
                                                              reinterpret_cast<float*>(out), accumulate, cpb);
  }
  else
    colsum_kernel<<<ceil_div(D, 128), 128, 0, stream>>>(src, src_dtype, ld, group_stride, groups, rows, D, scale,
                                                       reinterpret_cast<float*>(out), accumulate);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_adaptor_bwd(const void* x, const void* dy, const void* w1, const void* w2, void* scratch, void* dw1,
                               void* dw2, int32_t rows, int32_t D, int32_t rank, float scale, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(x && dy && w1 && w2 && scratch && dw1 && dw2 && rows > 0 && D > 0 && rank > 0 && rank <= 4, "adaptor_bwd: bad arguments");
  adaptor_bwd_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, stream>>>(
      reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(w1),
      reinterpret_cast<const float*>(w2), reinterpret_cast<float*>(scratch), rows, D, rank);
  MYR_CHECK_LAUNCH();
  adaptor_bwd_cols_kernel<<<ceil_div(D, 32), 512, 0, stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(dy),
                                                               reinterpret_cast<const float*>(scratch), reinterpret_cast<float*>(dw1),
                                                               reinterpret_cast<float*>(dw2), rows, D, rank, scale);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_adamw_step(void* params, const void* grads, void* exp_avg, void* exp_avg_sq, const void* wd_mask, int64_t n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, float inv_scale,
                              void* found_inf, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && wd_mask && n > 0 && step > 0, "adamw_step: bad arguments");
  if (found_inf) {
    MYR_CHECK_CUDA(cudaMemsetAsync(found_inf, 0, sizeof(int), stream));
    check_finite_kernel<<<grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(grads), n, reinterpret_cast<int*>(found_inf));
    MYR_CHECK_LAUNCH();
  }
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  const bool vec = ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                     reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(wd_mask) & 3) == 0;
  const long long n4 = vec ? n / 4 : 0;
  if (n4 > 0)
    adamw_vec4_kernel<<<grid_for(n4, 256), 256, 0, stream>>>(reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
                                                             reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
                                                             reinterpret_cast<const uchar4*>(wd_mask), n4, lr, beta1, beta2, eps,
                                                             weight_decay, bc1, bc2, inv_scale, reinterpret_cast<const int*>(found_inf));
  if (n - 4 * n4 > 0)
    adamw_kernel<<<grid_for(n - 4 * n4, 256), 256, 0, stream>>>(
        reinterpret_cast<float*>(params) + 4 * n4, reinterpret_cast<const float*>(grads) + 4 * n4, reinterpret_cast<float*>(exp_avg) + 4 * n4,
        reinterpret_cast<float*>(exp_avg_sq) + 4 * n4, reinterpret_cast<const unsigned char*>(wd_mask) + 4 * n4, n - 4 * n4, lr, beta1,
        beta2, eps, weight_decay, bc1, bc2, inv_scale, reinterpret_cast<const int*>(found_inf));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_memset_zero(void* ptr, size_t bytes, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(ptr != nullptr, "memset_zero: null pointer");
  MYR_CHECK_CUDA(cudaMemsetAsync(ptr, 0, bytes, stream));
  return MYR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// LoRA branches of q_proj / v_proj (peft, myriad.py:171-178; y += s * B (A dropout(x))) as CUDA-core kernels. The rank is 8: as
// tcgen05 GEMMs (F = 8 or K = 8 against 128 x BN tiles) the ten launches per layer of the backward and the six of the forward
// cost 15-20 us each - 8 ms of a 44 ms training step for 0.03 % of its flops. Here: fp16 operands, fp32 accumulation, the dropout
// mask regenerated from (seed, offset + t * D + d) wherever the dropped input is needed (nothing stored), fixed summation orders.
// ---------------------------------------------------------------------------------------------------------------------
namespace myr {

constexpr int LR = 8;  // rank

__device__ __forceinline__ void lr_unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// dropped input as the GEMM path saw it: rn_f16(x * keep_scale) where kept, 0 elsewhere (p = 0: thresh = 0 keeps everything)
__device__ __forceinline__ float lr_drop(float x, uint32_t thresh, float keep_scale, unsigned long long seed, unsigned long long idx) {
  // branch-free, also for p = 0 (thresh = 0 keeps everything and keep_scale = 1 leaves the fp16 value as it is): a test per
  // element, even a uniform one, cuts the unrolled loops into 32 short dependent chains
  const float v = __half2float(__float2half_rn(x * keep_scale));
  const int keep = -(int)(drop_hash(seed, idx) >= thresh);
  return __int_as_float(__float_as_int(v) & keep);
}

// xa[t, j * 8 + k] = sum_d drop_j(x1[t, d]) * A[j * 8 + k, d]: one warp = TWO tokens, both branches, lane = 8 consecutive d per step;
// the four warps of a block walk A together (one pass over its 16 rows per 8 tokens: 10 MB of L2 traffic per launch instead of 84)
__global__ void __launch_bounds__(128) lora_xa_fwd_kernel(const __half* __restrict__ x1, long long ldx, const __half* __restrict__ A,
                                                         float* __restrict__ xa, int T, int D, uint32_t thresh, float keep_scale,
                                                         unsigned long long seed, unsigned long long off_q, unsigned long long off_v) {
  const int t0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 2, lane = threadIdx.x & 31;
  if (t0 >= T) return;
  const bool two = t0 + 1 < T;
  float acc[2][2 * LR];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int k = 0; k < 2 * LR; ++k) acc[u][k] = 0.f;
  for (int d0 = lane * 8; d0 < D; d0 += 256) {
    float xq[2][8], xv[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float x[8];
      lr_unpack8(__ldg(reinterpret_cast<const uint4*>(x1 + (size_t)min(t0 + u, T - 1) * ldx + d0)), x);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned long long idx = (unsigned long long)(t0 + u) * D + d0 + e;
        xq[u][e] = lr_drop(x[e], thresh, keep_scale, seed, off_q + idx);
        xv[u][e] = lr_drop(x[e], thresh, keep_scale, seed, off_v + idx);
      }
    }
#pragma unroll
    for (int k = 0; k < 2 * LR; ++k) {
      float a[8];
      lr_unpack8(__ldg(reinterpret_cast<const uint4*>(A + (size_t)k * D + d0)), a);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float* xs = k < LR ? xq[u] : xv[u];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][k] = fmaf(xs[e], a[e], acc[u][k]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int k = 0; k < 2 * LR; ++k) {
      const float v = warp_sum_t(acc[u][k]);
      if (lane == 0 && (u == 0 || two)) xa[(size_t)(t0 + u) * 2 * LR + k] = v;
    }
}

// qkv[t, col_j + d] = rn_f16(qkv[t, col_j + d] + s * sum_k xa[t, j * 8 + k] * B_j[d, k]): one thread = 8 consecutive d of one branch
// for EIGHT tokens (its 8 rows of B stay in registers)
__global__ void lora_b_apply_kernel(__half* __restrict__ qkv, long long ldq, const float* __restrict__ xa, const __half* __restrict__ bq,
                                    const __half* __restrict__ bv, int T, int D, float s, long long col_q, long long col_v) {
  const int dv = D >> 3;
  const int tg = (T + 7) >> 3;
  const long long total = (long long)tg * 2 * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d0 = (int)(i % dv) << 3;
    const long long r = i / dv;
    const int j = (int)(r & 1);
    const int tb = (int)(r >> 1) << 3;
    const __half* B = j ? bv : bq;
    float b[8][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) lr_unpack8(__ldg(reinterpret_cast<const uint4*>(B + (size_t)(d0 + e) * LR)), b[e]);
#pragma unroll 2
    for (int t = tb; t < min(T, tb + 8); ++t) {
      const float4 xa0 = *reinterpret_cast<const float4*>(xa + (size_t)t * 2 * LR + j * LR), xa1 = *reinterpret_cast<const float4*>(xa + (size_t)t * 2 * LR + j * LR + 4);
      const float xk[8] = {xa0.x, xa0.y, xa0.z, xa0.w, xa1.x, xa1.y, xa1.z, xa1.w};
      __half* dst = qkv + (size_t)t * ldq + (j ? col_v : col_q) + d0;
      float y[8];
      lr_unpack8(*reinterpret_cast<const uint4*>(dst), y);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) a = fmaf(xk[k], b[e][k], a);
        y[e] = fmaf(s, a, y[e]);
      }
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(y[2 * e], y[2 * e + 1]);
      *reinterpret_cast<uint4*>(dst) = o;
    }
  }
}

// d_xa[t, j * 8 + k] = s * sum_d dy_j[t, d] * B_j[d, k]: one warp = FOUR tokens of one branch (B is read once per four tokens)
__global__ void __launch_bounds__(128) lora_dxa_kernel(const __half* __restrict__ dqkv, long long ldq, const __half* __restrict__ bq,
                                                      const __half* __restrict__ bv, float* __restrict__ dxa, int T, int D, float s,
                                                      long long col_q, long long col_v) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int tq = (T + 3) >> 2;
  if (w >= 2 * tq) return;
  const int j = w & 1, t0 = (w >> 1) << 2;
  const __half* B = j ? bv : bq;
  float acc[4][LR];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < LR; ++k) acc[u][k] = 0.f;
  for (int d0 = lane * 8; d0 < D; d0 += 256) {
    float g[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      lr_unpack8(*reinterpret_cast<const uint4*>(dqkv + (size_t)min(t0 + u, T - 1) * ldq + (j ? col_v : col_q) + d0), g[u]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float b[8];
      lr_unpack8(__ldg(reinterpret_cast<const uint4*>(B + (size_t)(d0 + e) * LR)), b);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < LR; ++k) acc[u][k] = fmaf(g[u][e], b[k], acc[u][k]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < LR; ++k) {
      const float v = warp_sum_t(acc[u][k]);
      if (lane == 0 && t0 + u < T) dxa[(size_t)(t0 + u) * 2 * LR + j * LR + k] = s * v;
    }
}

// Column reductions over the tokens. A block = 64 columns (two per thread) x 16 token slices; the slices meet in shared memory in slice order.
//   dB_j[d, k] = alpha_b * sum_t dy_j[t, d] * xa[t, j * 8 + k]                     (which = 0, 1: branch q, v)
//   dA[j * 8 + k, d] = alpha_a * sum_t d_xa[t, j * 8 + k] * drop_j(x1[t, d])      (which = 2, 3)
__global__ void __launch_bounds__(512) lora_wgrad_kernel(const __half* __restrict__ dqkv, long long ldq, const float* __restrict__ xa,
                                                        const float* __restrict__ dxa, const __half* __restrict__ x1, long long ldx,
                                                        float* __restrict__ dbq, float* __restrict__ dbv, float* __restrict__ dA, int T, int D,
                                                        float alpha_b, float alpha_a, long long col_q, long long col_v, uint32_t thresh,
                                                        float keep_scale, unsigned long long seed, unsigned long long off_q,
                                                        unsigned long long off_v) {
  __shared__ float red[16][64][LR + 1];
  const int which = blockIdx.y, j = which & 1;
  const int dl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int d = blockIdx.x * 64 + 2 * dl;  // D is a multiple of 256: every block is full
  float acc[2][LR];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int k = 0; k < LR; ++k) acc[u][k] = 0.f;
  {
    const int per = (T + 15) / 16;
    const int t0 = sl * per, t1 = min(T, t0 + per);
    const float* coef = (which < 2 ? xa : dxa) + j * LR;
    const unsigned long long off = j ? off_v : off_q;
    const __half* src = which < 2 ? dqkv + (j ? col_v : col_q) + d : x1 + d;
    const long long ld = which < 2 ? ldq : ldx;
    auto body = [&](int t, auto drop) {
      const float2 raw = __half22float2(__ldg(reinterpret_cast<const __half2*>(src + (size_t)t * ld)));
      float v0 = raw.x, v1 = raw.y;
      if (decltype(drop)::value) {
        const unsigned long long idx = off + (unsigned long long)t * D + d;
        v0 = lr_drop(v0, thresh, keep_scale, seed, idx);
        v1 = lr_drop(v1, thresh, keep_scale, seed, idx + 1);
      }
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(coef + (size_t)t * 2 * LR)), c1 = __ldg(reinterpret_cast<const float4*>(coef + (size_t)t * 2 * LR + 4));
      const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int k = 0; k < LR; ++k) {
        acc[0][k] = fmaf(v0, c[k], acc[0][k]);
        acc[1][k] = fmaf(v1, c[k], acc[1][k]);
      }
    };
    // two loops, not one with the test inside: a branch in the body keeps the unrolled iterations' loads from being issued together
    if (which < 2) {
#pragma unroll 4
      for (int t = t0; t < t1; ++t) body(t, std::false_type{});
    } else {
#pragma unroll 4
      for (int t = t0; t < t1; ++t) body(t, std::true_type{});
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int k = 0; k < LR; ++k) red[sl][2 * dl + u][k] = acc[u][k];
  __syncthreads();
  // 512 (column, rank index) sums per block, one per thread: the 16 slices added in slice order
  {
    const int o = threadIdx.x;
    const int c = o >> 3, k = o & 7;
    float tot = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) tot += red[s2][c][k];
    const int dd = blockIdx.x * 64 + c;
    if (which < 2)
      (j ? dbv : dbq)[(size_t)dd * LR + k] = alpha_b * tot;
    else
      dA[(size_t)(j * LR + k) * D + dd] = alpha_a * tot;
  }
}

// d_x1[t, d] += sum_j keep_j(t, d) * keep_scale * sum_k d_xa[t, j * 8 + k] * A[j * 8 + k, d]: one thread = 4 consecutive d
__global__ void lora_dx_kernel(float* __restrict__ dx1, long long ldd, const float* __restrict__ dxa, const __half* __restrict__ A, int T, int D,
                               uint32_t thresh, float keep_scale, unsigned long long seed, unsigned long long off_q,
                               unsigned long long off_v) {
  const int dv = D >> 2;
  const long long total = (long long)T * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d0 = (int)(i % dv) << 2;
    const long long t = i / dv;
    float c[2 * LR];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(dxa + t * 2 * LR + 4 * q);
      c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
    float g[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int k = 0; k < 2 * LR; ++k) {
      const uint2 raw = *reinterpret_cast<const uint2*>(A + (size_t)k * D + d0);
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      float* gj = g[k / LR];
      gj[0] = fmaf(c[k], lo.x, gj[0]); gj[1] = fmaf(c[k], lo.y, gj[1]); gj[2] = fmaf(c[k], hi.x, gj[2]); gj[3] = fmaf(c[k], hi.y, gj[3]);
    }
    float4 acc = *reinterpret_cast<const float4*>(dx1 + t * ldd + d0);
    float* ap = reinterpret_cast<float*>(&acc);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned long long idx = (unsigned long long)t * D + d0 + e;
      // branch q first, then v: the order the two dropout_bwd_add launches of the GEMM path added them in
      // (selects, not branches; p = 0: thresh = 0 keeps everything and keep_scale = 1)
      const float nq = fmaf(g[0][e], keep_scale, ap[e]);
      ap[e] = drop_hash(seed, off_q + idx) >= thresh ? nq : ap[e];
      const float nv = fmaf(g[1][e], keep_scale, ap[e]);
      ap[e] = drop_hash(seed, off_v + idx) >= thresh ? nv : ap[e];
    }
    *reinterpret_cast<float4*>(dx1 + t * ldd + d0) = acc;
  }
}

}  // namespace myr

using namespace myr;

extern "C" int myr_lora_fwd(const void* x1, int64_t ldx, const void* A, const void* bq, const void* bv, void* xa, void* qkv, int64_t ldq,
                            int64_t col_q, int64_t col_v, int32_t T, int32_t D, int32_t r, float s, float p, uint64_t seed, uint64_t off_q,
                            uint64_t off_v, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(x1 && A && bq && bv && xa && qkv && T > 0 && D > 0 && D % 256 == 0 && r == LR && p >= 0.f && p < 1.f, "lora_fwd: bad arguments (rank 8, D %% 256 == 0)");
  MYR_CHECK_ARG(ldx % 8 == 0 && ldq % 8 == 0 && col_q % 8 == 0 && col_v % 8 == 0 &&
                    ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(bq) | reinterpret_cast<uintptr_t>(bv) |
                      reinterpret_cast<uintptr_t>(xa) | reinterpret_cast<uintptr_t>(qkv)) & 15) == 0,
                "lora_fwd: operands must be 16-byte aligned");
  const uint32_t th = p > 0.f ? drop_thresh(p) : 0u;
  lora_xa_fwd_kernel<<<ceil_div(T, 8), 128, 0, stream>>>(reinterpret_cast<const __half*>(x1), ldx, reinterpret_cast<const __half*>(A),
                                                         reinterpret_cast<float*>(xa), T, D, th, 1.0f / (1.0f - p), seed, off_q, off_v);
  MYR_CHECK_LAUNCH();
  lora_b_apply_kernel<<<grid_for((long long)ceil_div(T, 8) * 2 * (D / 8), 128), 128, 0, stream>>>(reinterpret_cast<__half*>(qkv), ldq, reinterpret_cast<const float*>(xa),
                                                                                    reinterpret_cast<const __half*>(bq), reinterpret_cast<const __half*>(bv), T, D,
                                                                                    s, col_q, col_v);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_lora_bwd(const void* dqkv, int64_t ldq, int64_t col_q, int64_t col_v, const void* xa, const void* A, const void* bq,
                            const void* bv, const void* x1, int64_t ldx, void* dxa_scratch, void* dbq, void* dbv, void* dA, void* dx1,
                            int64_t ldd, int32_t T, int32_t D, int32_t r, float s, float inv_scale, float p, uint64_t seed, uint64_t off_q,
                            uint64_t off_v, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(dqkv && xa && A && bq && bv && x1 && dxa_scratch && dbq && dbv && dA && dx1 && T > 0 && D > 0 && D % 256 == 0 && r == LR &&
                    p >= 0.f && p < 1.f,
                "lora_bwd: bad arguments (rank 8, D %% 256 == 0)");
  MYR_CHECK_ARG(ldq % 8 == 0 && ldx % 8 == 0 && ldd % 4 == 0 && col_q % 8 == 0 && col_v % 8 == 0 &&
                    ((reinterpret_cast<uintptr_t>(dqkv) | reinterpret_cast<uintptr_t>(xa) | reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(bq) |
                      reinterpret_cast<uintptr_t>(bv) | reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(dxa_scratch) |
                      reinterpret_cast<uintptr_t>(dx1)) & 15) == 0,
                "lora_bwd: operands must be 16-byte aligned");
  const uint32_t th = p > 0.f ? drop_thresh(p) : 0u;
  const float ks = 1.0f / (1.0f - p);
  lora_dxa_kernel<<<ceil_div(2 * ceil_div(T, 4), 4), 128, 0, stream>>>(reinterpret_cast<const __half*>(dqkv), ldq, reinterpret_cast<const __half*>(bq),
                                                          reinterpret_cast<const __half*>(bv), reinterpret_cast<float*>(dxa_scratch), T, D, s, col_q, col_v);
  MYR_CHECK_LAUNCH();
  lora_wgrad_kernel<<<dim3(D / 64, 4), 512, 0, stream>>>(reinterpret_cast<const __half*>(dqkv), ldq, reinterpret_cast<const float*>(xa),
                                                                  reinterpret_cast<const float*>(dxa_scratch), reinterpret_cast<const __half*>(x1), ldx,
                                                                  reinterpret_cast<float*>(dbq), reinterpret_cast<float*>(dbv), reinterpret_cast<float*>(dA), T, D,
                                                                  s * inv_scale, inv_scale, col_q, col_v, th, ks, seed, off_q, off_v);
  MYR_CHECK_LAUNCH();
  lora_dx_kernel<<<grid_for((long long)T * (D / 4), 256), 256, 0, stream>>>(reinterpret_cast<float*>(dx1), ldd, reinterpret_cast<const float*>(dxa_scratch),
                                                                           reinterpret_cast<const __half*>(A), T, D, th, ks, seed, off_q, off_v);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
