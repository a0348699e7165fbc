// HBM-bound row kernels: LayerNorm / RMSNorm (optionally fused with the LoraAdaptorV2 rank-r residual),
// RoPE + KV-cache append, SwiGLU, embedding gather, strided row copy/cast, ViT token assembly.
// All are coalesced, 16-byte vectorised, one CTA per row (norms) or grid-stride (elementwise).
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAX_VEC = 4;  // float4 chunks per thread -> D <= 4096

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to N values; every thread gets the totals
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) red[warp * N + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += red[w * N + i];
    v[i] = t;
  }
}

__device__ __forceinline__ float4 load4(const void* base, int dtype, long long idx) {
  if (dtype == MYR_F32) return reinterpret_cast<const float4*>(base)[idx];
  const uint2 u = reinterpret_cast<const uint2*>(base)[idx];
  const __half2 a = *reinterpret_cast<const __half2*>(&u.x), b = *reinterpret_cast<const __half2*>(&u.y);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4_f16(__half* base, long long idx, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(base)[idx] = u;
}

struct NormParams {
  const void* x; int x_dtype; long long ldx;
  int rows, D;
  const float* gamma; const float* beta;  // beta null for RMSNorm
  float eps; int rms;
  const float* w1; const float* w2; int rank;  // optional LoraAdaptorV2: y = x + W2 (W1 x), W1 [rank, D], W2 [D, rank]
  __half* out16; long long ld16;
  float* out32; long long ld32;
  float* pre32; long long ldpre;  // optional: the pre-norm value (x or adaptor output) in fp32
  float* stats;                   // optional [rows, 2]: mean (0 for rms), rstd
};

// LayerNorm (blip2.py:119-125, eva_vit.py norm1/norm2, Qformer.py LayerNorm) / RMSNorm (modeling_llama.py:66-74),
// fp32 statistics, two-pass variance on register-resident data. Optional fused LoraAdaptorV2 (networks.py:81-93).
__global__ void __launch_bounds__(NORM_THREADS) norm_kernel(const NormParams p) {
  __shared__ float red[(NORM_THREADS / 32) * 4];
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;
  const int nvec = p.D >> 2;
  const char* xrow = reinterpret_cast<const char*>(p.x) + (size_t)row * p.ldx * (p.x_dtype == MYR_F32 ? 4 : 2);
  float4 v[NORM_MAX_VEC];
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = threadIdx.x + i * NORM_THREADS;
    v[i] = c < nvec ? load4(xrow, p.x_dtype, c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (p.rank > 0) {  // rank <= 4
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NORM_MAX_VEC; ++i) {
      const int c = threadIdx.x + i * NORM_THREADS;
      if (c < nvec) {
        for (int r = 0; r < p.rank; ++r) {
          const float4 w = reinterpret_cast<const float4*>(p.w1 + (size_t)r * p.D)[c];
          t[r] += v[i].x * w.x + v[i].y * w.y + v[i].z * w.z + v[i].w * w.w;
        }
      }
    }
    block_sum<4>(t, red);
#pragma unroll
    for (int i = 0; i < NORM_MAX_VEC; ++i) {
      const int c = threadIdx.x + i * NORM_THREADS;
      if (c < nvec) {
        float* e = reinterpret_cast<float*>(&v[i]);
        for (int k = 0; k < 4; ++k) {
          const float* w2 = p.w2 + (size_t)(c * 4 + k) * p.rank;
          float a = 0.f;
          for (int r = 0; r < p.rank; ++r) a += w2[r] * t[r];
          e[k] += a;
        }
      }
    }
  }
  if (p.pre32) {
#pragma unroll
    for (int i = 0; i < NORM_MAX_VEC; ++i) {
      const int c = threadIdx.x + i * NORM_THREADS;
      if (c < nvec) reinterpret_cast<float4*>(p.pre32 + (size_t)row * p.ldpre)[c] = v[i];
    }
  }
  float mean = 0.f;
  if (!p.rms) {
    float s[1] = {0.f};
#pragma unroll
    for (int i = 0; i < NORM_MAX_VEC; ++i) s[0] += v[i].x + v[i].y + v[i].z + v[i].w;
    block_sum<1>(s, red);
    mean = s[0] / p.D;
  }
  float q[1] = {0.f};
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = threadIdx.x + i * NORM_THREADS;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q[0] += a * a + b * b + cc * cc + d * d;
    }
  }
  block_sum<1>(q, red);
  const float rstd = rsqrtf(q[0] / p.D + p.eps);
  if (p.stats && threadIdx.x == 0) {
    p.stats[2 * row] = mean;
    p.stats[2 * row + 1] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = threadIdx.x + i * NORM_THREADS;
    if (c < nvec) {
      const float4 g = reinterpret_cast<const float4*>(p.gamma)[c];
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x;
      o.y = (v[i].y - mean) * rstd * g.y;
      o.z = (v[i].z - mean) * rstd * g.z;
      o.w = (v[i].w - mean) * rstd * g.w;
      if (p.beta) {
        const float4 bb = reinterpret_cast<const float4*>(p.beta)[c];
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      if (p.out16) store4_f16(p.out16 + (size_t)row * p.ld16, c, o);
      if (p.out32) reinterpret_cast<float4*>(p.out32 + (size_t)row * p.ld32)[c] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// RoPE (modeling_llama.py:109-123, non-interleaved halves) on q and k of a fused qkv row + KV-cache append.
// qkv: [T, 3*H*dh] fp16, T = B*S; q rotated in place; k rotated -> kcache[b, cache_off + s]; v -> vcache.
// One thread = 8 consecutive i < dh/2 of one (token, head): 16-byte loads/stores.
// ---------------------------------------------------------------------------------------------------
struct RopeParams {
  __half* qkv; long long ldq;
  int B, S, H, dh;
  const int* pos;       // [B*S] rotary position of each token
  const float* cos_t; const float* sin_t;  // [max_pos, dh/2] fp32
  __half* kcache; __half* vcache; long long c_ts, c_bs;  // cache element strides (token, batch)
  const int* cache_off;  // device scalar: first cache slot for s = 0 (null -> cache_off_host)
  int cache_off_host;
  // peft LoRA on q_proj / v_proj (myriad.py:171-178): xa = x A^T sits in columns [3*H*dh, 3*H*dh + 2r) of the qkv row
  // (A rows are appended to the fused qkv weight); q += scale * B_q xa[:r], v += scale * B_v xa[r:], before the rotation.
  const __half* lora_bq; const __half* lora_bv;  // [H*dh, r] fp16, r == 8
  int lora_r; float lora_scale;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// y[k] += scale * sum_r B[(row0 + k), r] * xa[r] for 8 consecutive output features (r == 8: one 16-byte row of B each)
__device__ __forceinline__ void lora_add8(float (&y)[8], const __half* __restrict__ b, int row0, const float (&xa)[8], float scale) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float w[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + (size_t)(row0 + k) * 8)), w);
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) a = fmaf(w[r], xa[r], a);
    y[k] = fmaf(scale, a, y[k]);
  }
}

__global__ void rope_cache_kernel(const RopeParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int half = p.dh >> 1, hv = half >> 3;  // 8-wide vectors per half
  const long long total = (long long)p.B * p.S * p.H * hv;
  const int off = p.cache_off ? *p.cache_off : p.cache_off_host;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int vi = (int)(i % hv);
    const int h = (int)((i / hv) % p.H);
    const long long t = i / ((long long)hv * p.H);
    const int b = (int)(t / p.S), s = (int)(t % p.S);
    const int pos = p.pos[t];
    float c[8], sn[8];
    {
      const float4* cp = reinterpret_cast<const float4*>(p.cos_t + (size_t)pos * half + vi * 8);
      const float4* sp = reinterpret_cast<const float4*>(p.sin_t + (size_t)pos * half + vi * 8);
      const float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
      c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
      sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
    }
    __half* row = p.qkv + t * p.ldq;
    const int HD = p.H * p.dh;
    __half* kdst = p.kcache + (long long)b * p.c_bs + (long long)(off + s) * p.c_ts + h * p.dh;
    __half* vdst = p.vcache + (long long)b * p.c_bs + (long long)(off + s) * p.c_ts + h * p.dh;
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: q (in place), 1: k (-> cache)
      __half* src = row + which * HD + h * p.dh;
      float x1[8], x2[8], o1[8], o2[8];
      unpack8(*reinterpret_cast<const uint4*>(src + vi * 8), x1);
      unpack8(*reinterpret_cast<const uint4*>(src + half + vi * 8), x2);
      if (which == 0 && p.lora_r) {
        float xa[8];
        unpack8(*reinterpret_cast<const uint4*>(row + 3 * HD), xa);
        lora_add8(x1, p.lora_bq, h * p.dh + vi * 8, xa, p.lora_scale);
        lora_add8(x2, p.lora_bq, h * p.dh + half + vi * 8, xa, p.lora_scale);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o1[k] = x1[k] * c[k] - x2[k] * sn[k];
        o2[k] = x2[k] * c[k] + x1[k] * sn[k];
      }
      __half* dst = which == 0 ? src : kdst;
      *reinterpret_cast<uint4*>(dst + vi * 8) = pack8(o1);
      *reinterpret_cast<uint4*>(dst + half + vi * 8) = pack8(o2);
    }
    const __half* vsrc = row + 2 * HD + h * p.dh;
    if (!p.lora_r) {
      *reinterpret_cast<uint4*>(vdst + vi * 8) = *reinterpret_cast<const uint4*>(vsrc + vi * 8);
      *reinterpret_cast<uint4*>(vdst + half + vi * 8) = *reinterpret_cast<const uint4*>(vsrc + half + vi * 8);
    } else {
      float xa[8], v1[8], v2[8];
      unpack8(*reinterpret_cast<const uint4*>(row + 3 * HD + 8), xa);
      unpack8(*reinterpret_cast<const uint4*>(vsrc + vi * 8), v1);
      unpack8(*reinterpret_cast<const uint4*>(vsrc + half + vi * 8), v2);
      lora_add8(v1, p.lora_bv, h * p.dh + vi * 8, xa, p.lora_scale);
      lora_add8(v2, p.lora_bv, h * p.dh + half + vi * 8, xa, p.lora_scale);
      *reinterpret_cast<uint4*>(vdst + vi * 8) = pack8(v1);
      *reinterpret_cast<uint4*>(vdst + half + vi * 8) = pack8(v2);
    }
  }
}

// SwiGLU (modeling_llama.py:139-140): out[t, i] = silu(gu[t, i]) * gu[t, I + i]
__global__ void swiglu_kernel(const __half* __restrict__ gu, long long ldg, __half* __restrict__ out, long long ldo, int T,
                              int I) {
  pdl_wait();
  pdl_launch_dependents();
  const int iv = I >> 3;
  const long long total = (long long)T * iv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / iv;
    const int c = (int)(i % iv) * 8;
    float g[8], u[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(gu + t * ldg + c), g);
    unpack8(*reinterpret_cast<const uint4*>(gu + t * ldg + I + c), u);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = silu_f(g[k]) * u[k];
    *reinterpret_cast<uint4*>(out + t * ldo + c) = pack8(o);
  }
}

// Embedding gather (myriad.py:308-311): out[r, :] = table[ids[r], :]  (fp16 table -> fp32 or fp16 rows)
__global__ void embed_kernel(const __half* __restrict__ table, int D, const long long* __restrict__ ids64,
                             const int* __restrict__ ids32, int n, void* out, int out_dtype, long long ldo) {
  pdl_wait();
  pdl_launch_dependents();
  const int dv = D >> 3;
  const long long total = (long long)n * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dv;
    const int c = (int)(i % dv) * 8;
    const long long id = ids64 ? ids64[r] : (long long)ids32[r];
    const uint4 u = *reinterpret_cast<const uint4*>(table + id * D + c);
    if (out_dtype == MYR_F16) {
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + r * ldo + c) = u;
    } else {
      float f[8];
      unpack8(u, f);
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + r * ldo + c);
      o[0] = make_float4(f[0], f[1], f[2], f[3]);
      o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}

// Strided row copy with dtype conversion: rows are numbered g * rows_per_group + r; used to place token groups
// into the concatenated buffers of myriad.py:249-266,372 without torch.cat.
__global__ void copy_rows_kernel(const void* src, int src_dtype, long long src_ld, long long src_gs, void* dst,
                                 int dst_dtype, long long dst_ld, long long dst_gs, int groups, int rows_per_group,
                                 int D) {
  const int dv = D >> 2;
  const long long total = (long long)groups * rows_per_group * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv);
    const long long rr = i / dv;
    const int r = (int)(rr % rows_per_group);
    const long long g = rr / rows_per_group;
    const size_t ses = src_dtype == MYR_F32 ? 4 : 2;
    const char* s = reinterpret_cast<const char*>(src) + (size_t)(g * src_gs + r * src_ld) * ses;
    const float4 v = load4(s, src_dtype, c);
    if (dst_dtype == MYR_F32)
      reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + g * dst_gs + r * dst_ld)[c] = v;
    else
      store4_f16(reinterpret_cast<__half*>(dst) + g * dst_gs + r * dst_ld, c, v);
  }
}

// ViT token assembly (eva_vit.py:326-331): x[b, 0] = cls + pos[0]; x[b, 1 + p] = patch[b, p] + pos[1 + p]  (fp32)
__global__ void vit_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                    const float* __restrict__ pos, float* __restrict__ x, int B, int N, int D) {
  const int dv = D >> 2;
  const long long total = (long long)B * N * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv);
    const long long t = i / dv;
    const int n = (int)(t % N);
    const long long b = t / N;
    const float4 pe = reinterpret_cast<const float4*>(pos + (size_t)n * D)[c];
    float4 v = n == 0 ? reinterpret_cast<const float4*>(cls)[c]
                      : reinterpret_cast<const float4*>(patch + (size_t)(b * (N - 1) + n - 1) * D)[c];
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    reinterpret_cast<float4*>(x + (size_t)t * D)[c] = v;
  }
}

static inline int ew_grid(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace myr

using namespace myr;

extern "C" int myr_norm_fwd(const myr_norm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(a && a->x && a->gamma, "norm: null pointer");
  MYR_CHECK_ARG(a->rows > 0 && a->D > 0 && a->D % 4 == 0 && a->D <= NORM_THREADS * NORM_MAX_VEC * 4,
                "norm: D=%d must be a multiple of 4 and <= %d", a->D, NORM_THREADS * NORM_MAX_VEC * 4);
  MYR_CHECK_ARG(a->rank >= 0 && a->rank <= 4, "norm: adaptor rank %d > 4", a->rank);
  MYR_CHECK_ARG(a->ldx % 4 == 0 && a->ld16 % 4 == 0 && a->ld32 % 4 == 0 && a->ldpre % 4 == 0, "norm: row strides must be multiples of 4");
  NormParams p;
  p.x = a->x; p.x_dtype = a->x_dtype; p.ldx = a->ldx; p.rows = a->rows; p.D = a->D;
  p.gamma = reinterpret_cast<const float*>(a->gamma); p.beta = reinterpret_cast<const float*>(a->beta);
  p.eps = a->eps; p.rms = a->rms;
  p.w1 = reinterpret_cast<const float*>(a->adaptor_w1); p.w2 = reinterpret_cast<const float*>(a->adaptor_w2);
  p.rank = (a->adaptor_w1 && a->adaptor_w2) ? a->rank : 0;
  p.out16 = reinterpret_cast<__half*>(a->out16); p.ld16 = a->ld16;
  p.out32 = reinterpret_cast<float*>(a->out32); p.ld32 = a->ld32;
  p.pre32 = reinterpret_cast<float*>(a->pre32); p.ldpre = a->ldpre;
  p.stats = reinterpret_cast<float*>(a->stats);
  MYR_CHECK_CUDA(launch_kernel(norm_kernel, dim3(a->rows), dim3(NORM_THREADS), 0, stream, true, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_rope_cache(const myr_rope_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(a && a->qkv && a->pos && a->cos_table && a->sin_table && a->kcache && a->vcache, "rope: null pointer");
  MYR_CHECK_ARG(a->dh % 16 == 0 && a->ldq % 8 == 0 && a->cache_token_stride % 8 == 0 && a->cache_batch_stride % 8 == 0,
                "rope: dh must be a multiple of 16 and strides 16-byte aligned");
  RopeParams p;
  p.qkv = reinterpret_cast<__half*>(a->qkv); p.ldq = a->ldq;
  p.B = a->B; p.S = a->S; p.H = a->H; p.dh = a->dh;
  p.pos = reinterpret_cast<const int*>(a->pos);
  p.cos_t = reinterpret_cast<const float*>(a->cos_table); p.sin_t = reinterpret_cast<const float*>(a->sin_table);
  p.kcache = reinterpret_cast<__half*>(a->kcache); p.vcache = reinterpret_cast<__half*>(a->vcache);
  p.c_ts = a->cache_token_stride; p.c_bs = a->cache_batch_stride;
  p.cache_off = reinterpret_cast<const int*>(a->cache_off_dev); p.cache_off_host = a->cache_off;
  p.lora_bq = reinterpret_cast<const __half*>(a->lora_bq); p.lora_bv = reinterpret_cast<const __half*>(a->lora_bv);
  p.lora_r = (a->lora_bq && a->lora_bv) ? a->lora_r : 0; p.lora_scale = a->lora_scale;
  MYR_CHECK_ARG(p.lora_r == 0 || p.lora_r == 8, "rope: fused LoRA supports rank 8 only (got %d)", p.lora_r);
  MYR_CHECK_ARG(p.lora_r == 0 || ((reinterpret_cast<uintptr_t>(a->lora_bq) | reinterpret_cast<uintptr_t>(a->lora_bv)) & 15) == 0,
                "rope: LoRA B matrices must be 16-byte aligned");
  const long long total = (long long)a->B * a->S * a->H * (a->dh / 16);
  MYR_CHECK_CUDA(launch_kernel(rope_cache_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, stream, true, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_swiglu(const void* gate_up, int64_t ld_gu, void* out, int64_t ld_out, int32_t T, int32_t I,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(gate_up && out && T > 0 && I > 0 && I % 8 == 0 && ld_gu % 8 == 0 && ld_out % 8 == 0, "swiglu: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(swiglu_kernel, dim3(ew_grid((long long)T * (I / 8), 256)), dim3(256), 0, stream, true,
                               reinterpret_cast<const __half*>(gate_up), (long long)ld_gu, reinterpret_cast<__half*>(out),
                               (long long)ld_out, T, I));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_embed(const void* table, int32_t D, const void* ids, int32_t ids_are_int64, int32_t n, void* out,
                         int32_t out_dtype, int64_t ld_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(table && ids && out && n > 0 && D % 8 == 0 && ld_out % 8 == 0, "embed: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(embed_kernel, dim3(ew_grid((long long)n * (D / 8), 256)), dim3(256), 0, stream, true,
                               reinterpret_cast<const __half*>(table), D,
                               ids_are_int64 ? reinterpret_cast<const long long*>(ids) : (const long long*)nullptr,
                               ids_are_int64 ? (const int*)nullptr : reinterpret_cast<const int*>(ids), n, out, out_dtype,
                               (long long)ld_out));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_copy_rows(const void* src, int32_t src_dtype, int64_t src_ld, int64_t src_group_stride, void* dst,
                             int32_t dst_dtype, int64_t dst_ld, int64_t dst_group_stride, int32_t groups,
                             int32_t rows_per_group, int32_t D, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(src && dst && groups > 0 && rows_per_group > 0 && D % 4 == 0, "copy_rows: bad arguments");
  MYR_CHECK_ARG(src_ld % 4 == 0 && dst_ld % 4 == 0 && src_group_stride % 4 == 0 && dst_group_stride % 4 == 0,
                "copy_rows: strides must be multiples of 4 elements");
  copy_rows_kernel<<<ew_grid((long long)groups * rows_per_group * (D / 4), 256), 256, 0, stream>>>(
      src, src_dtype, src_ld, src_group_stride, dst, dst_dtype, dst_ld, dst_group_stride, groups, rows_per_group, D);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_vit_assemble(const void* patch, const void* cls, const void* pos, void* x, int32_t B, int32_t N,
                                int32_t D, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(patch && cls && pos && x && B > 0 && N > 1 && D % 4 == 0, "vit_assemble: bad arguments");
  vit_assemble_kernel<<<ew_grid((long long)B * N * (D / 4), 256), 256, 0, stream>>>(
      reinterpret_cast<const float*>(patch), reinterpret_cast<const float*>(cls), reinterpret_cast<const float*>(pos),
      reinterpret_cast<float*>(x), B, N, D);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
