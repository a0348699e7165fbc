// Fused GEMM epilogue shared by the tcgen05 GEMM kernels (gemm.cu: one CTA per tile, gemm2.cu: CTA pairs):
// bias, fp16 rounding points, q-scale, GELU(erf) / ReLU, alpha, residual (fp16 / fp32), grouped-row store, SwiGLU.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int N_EPI_WARPS = 8;
constexpr int EPI_THREADS = 32 * N_EPI_WARPS;

struct Epilogue {
  const __half* bias;
  int act;
  int round_acc;
  int scale_cols;
  float scale;
  const void* res;
  int res_dtype;
  long long ldr;
  void* out;
  int out_dtype;
  long long ldo;
  float alpha;             // fp32 scale applied after the activation, before the residual (gradient unscale, LoRA alpha/r)
  int group_rows;          // 0 = plain rows; else out row t lives at (t / group_rows) * group_stride + (t % group_rows) * ldo
  long long group_stride;
  int vec;                 // 16-byte vector access to bias / res / out is legal (alignment checked on the host)
};

__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// erf with |abs error| <= 1.5e-7 (Abramowitz-Stegun 7.1.26): well below the fp16 rounding applied to every GELU output.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = __expf(-ax * ax);
  const float y = fmaf(-p * t, e, 1.0f);
  return copysignf(y, x);
}

// element-wise part of the epilogue, up to (not including) the residual add
__device__ __forceinline__ float epi_transform(const Epilogue& ep, float v, float bias_f, int f) {
  v += bias_f;
  if (ep.round_acc) v = round_f16(v);
  if (f < ep.scale_cols) v = round_f16(v * ep.scale);
  if (ep.act == MYR_ACT_GELU_ERF) v = round_f16(0.5f * v * (1.0f + fast_erf(v * 0.70710678118654752440f)));
  else if (ep.act == MYR_ACT_RELU) v = fmaxf(v, 0.f);
  return v * ep.alpha;
}

__device__ __forceinline__ long long out_row_offset(const Epilogue& ep, long long t) {
  return ep.group_rows ? (t / ep.group_rows) * ep.group_stride + (t % ep.group_rows) * ep.ldo : t * ep.ldo;
}

__device__ __forceinline__ void epi_store_scalar(const Epilogue& ep, float v, long long t, int f, long long boff) {
  if (ep.res) {
    v += (ep.res_dtype == MYR_F32) ? reinterpret_cast<const float*>(ep.res)[t * ep.ldr + f]
                                   : __half2float(reinterpret_cast<const __half*>(ep.res)[t * ep.ldr + f]);
  }
  const long long o = boff + out_row_offset(ep, t) + f;
  if (ep.out_dtype == MYR_F32)
    reinterpret_cast<float*>(ep.out)[o] = v;
  else
    reinterpret_cast<__half*>(ep.out)[o] = __float2half_rn(v);
}

__device__ __forceinline__ void unpack8h(const uint4 u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p = __half22float2(h[i]);
    f[2 * i] = p.x;
    f[2 * i + 1] = p.y;
  }
}
__device__ __forceinline__ uint4 pack8h(const float* f) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// lanes = tokens: this thread owns token t; v[0..16) are features f0 .. f0+15 of that token (already summed over k).
__device__ __forceinline__ void epi_row16(const Epilogue& ep, float* v, long long t, int f0, int F, long long obase) {
  if (ep.vec && f0 + 16 <= F) {
    float b[16];
    if (ep.bias) {
      const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + f0);
      unpack8h(__ldg(bp), b);
      unpack8h(__ldg(bp + 1), b + 8);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) b[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = epi_transform(ep, v[j], b[j], f0 + j);
    if (ep.res) {
      if (ep.res_dtype == MYR_F32) {
        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.res) + t * ep.ldr + f0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 r = rp[j];
          v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      } else {
        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(ep.res) + t * ep.ldr + f0);
        float r[16];
        unpack8h(rp[0], r);
        unpack8h(rp[1], r + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += r[j];
      }
    }
    if (ep.out_dtype == MYR_F32) {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + obase + f0);
#pragma unroll
      for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(ep.out) + obase + f0);
      op[0] = pack8h(v);
      op[1] = pack8h(v + 8);
    }
  } else {
    for (int j = 0; j < 16; ++j) {
      const int f = f0 + j;
      if (f >= F) break;
      float x = epi_transform(ep, v[j], ep.bias ? __half2float(ep.bias[f]) : 0.f, f);
      if (ep.res) {
        x += (ep.res_dtype == MYR_F32) ? reinterpret_cast<const float*>(ep.res)[t * ep.ldr + f]
                                       : __half2float(reinterpret_cast<const __half*>(ep.res)[t * ep.ldr + f]);
      }
      if (ep.out_dtype == MYR_F32)
        reinterpret_cast<float*>(ep.out)[obase + f] = x;
      else
        reinterpret_cast<__half*>(ep.out)[obase + f] = __float2half_rn(x);
    }
  }
}

// SwiGLU pair (modeling_llama.py:139-140) with the rounding points of the unfused path: gate / up rounded to fp16 first.
__device__ __forceinline__ float swiglu_pair(float g, float u) {
  g = round_f16(g);
  u = round_f16(u);
  return silu_f(g) * u;
}

// lanes = tokens, SwiGLU: g[16] / u[16] are gate / up features i0..i0+15 of token t; writes out[t, i0..i0+15] (fp16).
__device__ __forceinline__ void epi_row16_swiglu(const Epilogue& ep, const float* g, const float* u, int i0, int I, long long obase) {
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = swiglu_pair(g[j], u[j]);
  __half* op = reinterpret_cast<__half*>(ep.out) + obase + i0;
  if (ep.vec && i0 + 16 <= I) {
    reinterpret_cast<uint4*>(op)[0] = pack8h(o);
    reinterpret_cast<uint4*>(op)[1] = pack8h(o + 8);
  } else {
    for (int j = 0; j < 16 && i0 + j < I; ++j) op[j] = __float2half_rn(o[j]);
  }
}

}  // namespace myr
