// Vision-expert heads: what adrefexpert.forward (adrefexpert_v2.py:245-301) does AFTER the ImageBind-Huge
// vision trunk (imagebind_model.py:486-504; the trunk itself runs on the ViT kernels: patch GEMM, LayerNorm, qkv GEMM, flash
// attention with dh = 80 padded to 96 by TMA, MLP GEMMs). All HBM-bound row / pixel kernels, fp32 arithmetic:
//   expert_tap       tokens[i].transpose(0,1)[:, 1:, :] (adrefexpert_v2.py:215-216,26-27): drop the class token of a tapped
//                    layer, fp32 stream -> fp16 GEMM operand; optionally divided by the row's L2 norm (the cosine branch)
//   expert_logits    zero-shot branch :285-286: 100 * (x / |x|) . text_k for the two text embeddings of the sample's class
//   expert_maps      :288-298: per tapped layer softmax over {normal, abnormal} at 16 x 16 (masks) and after bilinear up-sampling
//                    (align_corners = True) to 224 x 224 (maps), mean over the layers
//   expert_rowmax    k-shot branch :270-272: max over the reference patches of the cosine-similarity rows, mean over layers
//   expert_sim_maps  :274-278: simmask = 1 - sim at 16 x 16, anomaly map = 1 - bilinear(sim) at 224 x 224
#include "common.h"
#include "ptx.cuh"

namespace myr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per output row (b, i), i in [0, N - 1): source row b * N + 1 + i
__global__ void expert_tap_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int N, int D, int normalize) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B * (N - 1)) return;
  const int b = row / (N - 1), i = row - b * (N - 1);
  const float4* src = reinterpret_cast<const float4*>(x + ((size_t)b * N + 1 + i) * D);
  float ss = 0.f;
  if (normalize) {
    for (int c = lane; c < D / 4; c += 32) {
      const float4 v = src[c];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
  }
  const float s = normalize ? 1.0f / fmaxf(sqrtf(ss), 1e-8f) : 1.0f;  // F.cosine_similarity's eps
  __half2* dst = reinterpret_cast<__half2*>(out + (size_t)row * D);
  for (int c = lane; c < D / 4; c += 32) {
    const float4 v = src[c];
    dst[2 * c] = __floats2half2_rn(v.x * s, v.y * s);
    dst[2 * c + 1] = __floats2half2_rn(v.z * s, v.w * s);
  }
}

// one warp per patch row: logits[row][k] = scale * (x . text[b][k]) / |x|
__global__ void expert_logits_kernel(const float* __restrict__ tok, long long ld, const float* __restrict__ text, float* __restrict__ logits,
                                     int B, int P, int C, float scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B * P) return;
  const int b = row / P;
  const float* x = tok + (size_t)row * ld;
  const float* t0 = text + (size_t)b * 2 * C;
  const float* t1 = t0 + C;
  float ss = 0.f, d0 = 0.f, d1 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[c];
    ss = fmaf(v, v, ss);
    d0 = fmaf(v, t0[c], d0);
    d1 = fmaf(v, t1[c], d1);
  }
  ss = warp_sum(ss);
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  if (lane == 0) {
    const float inv = scale / sqrtf(ss);
    logits[(size_t)row * 2] = d0 * inv;
    logits[(size_t)row * 2 + 1] = d1 * inv;
  }
}

// bilinear sample of a G x G grid at output pixel (oy, ox) of an OUT x OUT image, align_corners = True
// (torch upsample_bilinear2d: src = dst * (G - 1) / (OUT - 1); i1 = i0 + (i0 < G - 1))
struct Bilin {
  int i00, i01, i10, i11;
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Bilin bilin_setup(int oy, int ox, int G, int OUT) {
  const float sc = OUT > 1 ? (float)(G - 1) / (float)(OUT - 1) : 0.f;
  const float fy = sc * oy, fx = sc * ox;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < G - 1 ? 1 : 0), x1 = x0 + (x0 < G - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0;
  Bilin r;
  r.i00 = y0 * G + x0; r.i01 = y0 * G + x1; r.i10 = y1 * G + x0; r.i11 = y1 * G + x1;
  r.w00 = (1.f - ly) * (1.f - lx); r.w01 = (1.f - ly) * lx; r.w10 = ly * (1.f - lx); r.w11 = ly * lx;
  return r;
}

// logits [L][B][G*G][2] -> maps [B][OUT*OUT] = mean_l softmax(bilinear(logits_l))[abnormal], masks [B][G*G] = mean_l softmax(logits_l)[abnormal]
__global__ void expert_maps_kernel(const float* __restrict__ logits, float* __restrict__ maps, float* __restrict__ masks, int L, int B,
                                   int G, int OUT) {
  const int b = blockIdx.y;
  const int P = G * G;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_map_blocks = (OUT * OUT + blockDim.x - 1) / blockDim.x;
  if ((int)blockIdx.x < n_map_blocks) {
    if (idx >= OUT * OUT) return;
    const Bilin s = bilin_setup(idx / OUT, idx % OUT, G, OUT);
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      const float2* lg = reinterpret_cast<const float2*>(logits + ((size_t)l * B + b) * P * 2);
      const float2 a = lg[s.i00], c = lg[s.i01], d = lg[s.i10], e = lg[s.i11];
      const float l0 = s.w00 * a.x + s.w01 * c.x + s.w10 * d.x + s.w11 * e.x;
      const float l1 = s.w00 * a.y + s.w01 * c.y + s.w10 * d.y + s.w11 * e.y;
      const float m = fmaxf(l0, l1);
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      acc += e1 / (e0 + e1);
    }
    maps[(size_t)b * OUT * OUT + idx] = acc / L;
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) {
        const float2 v = reinterpret_cast<const float2*>(logits + ((size_t)l * B + b) * P * 2)[i];
        const float m = fmaxf(v.x, v.y);
        const float e0 = expf(v.x - m), e1 = expf(v.y - m);
        acc += e1 / (e0 + e1);
      }
      masks[(size_t)b * P + i] = acc / L;
    }
  }
}

// one warp per row: acc[row] (+)= weight * max_r S[row][r]
__global__ void expert_rowmax_kernel(const float* __restrict__ S, long long ld, float* __restrict__ acc, int rows, int R, float weight,
                                     int accumulate) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = S + (size_t)row * ld;
  float m = -INFINITY;
  for (int r = lane; r < R; r += 32) m = fmaxf(m, s[r]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) acc[row] = (accumulate ? acc[row] : 0.f) + weight * m;
}

// sim [B][G*G] -> simmask [B][G*G] = 1 - sim, maps [B][OUT*OUT] = 1 - bilinear(sim)
__global__ void expert_sim_maps_kernel(const float* __restrict__ sim, float* __restrict__ maps, float* __restrict__ simmask, int B, int G,
                                       int OUT) {
  const int b = blockIdx.y;
  const int P = G * G;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_map_blocks = (OUT * OUT + blockDim.x - 1) / blockDim.x;
  const float* sb = sim + (size_t)b * P;
  if ((int)blockIdx.x < n_map_blocks) {
    if (idx >= OUT * OUT) return;
    const Bilin s = bilin_setup(idx / OUT, idx % OUT, G, OUT);
    maps[(size_t)b * OUT * OUT + idx] = 1.0f - (s.w00 * sb[s.i00] + s.w01 * sb[s.i01] + s.w10 * sb[s.i10] + s.w11 * sb[s.i11]);
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) simmask[(size_t)b * P + i] = 1.0f - sb[i];
  }
}

}  // namespace myr

using namespace myr;

extern "C" int myr_expert_tap(const void* x, void* out16, int32_t B, int32_t N, int32_t D, int32_t normalize, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(x && out16 && B > 0 && N > 1 && D > 0 && D % 4 == 0, "expert_tap: bad arguments");
  const int rows = B * (N - 1);
  MYR_CHECK_CUDA(launch_kernel(expert_tap_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, stream, false, reinterpret_cast<const float*>(x),
                               reinterpret_cast<__half*>(out16), (int)B, (int)N, (int)D, (int)normalize));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_expert_logits(const void* tokens, int64_t ld, const void* text, void* logits, int32_t B, int32_t P, int32_t C,
                                 float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(tokens && text && logits && B > 0 && P > 0 && C > 0 && ld >= C, "expert_logits: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(expert_logits_kernel, dim3(ceil_div(B * P, 8)), dim3(256), 0, stream, false,
                               reinterpret_cast<const float*>(tokens), (long long)ld, reinterpret_cast<const float*>(text),
                               reinterpret_cast<float*>(logits), (int)B, (int)P, (int)C, scale));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_expert_maps(const void* logits, void* maps, void* masks, int32_t L, int32_t B, int32_t G, int32_t OUT, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(logits && maps && masks && L > 0 && B > 0 && G > 1 && OUT > 1, "expert_maps: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(expert_maps_kernel, dim3(ceil_div(OUT * OUT, 256) + 1, B), dim3(256), 0, stream, false,
                               reinterpret_cast<const float*>(logits), reinterpret_cast<float*>(maps), reinterpret_cast<float*>(masks), (int)L,
                               (int)B, (int)G, (int)OUT));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_expert_rowmax(const void* S, int64_t ld, void* acc, int32_t rows, int32_t R, float weight, int32_t accumulate,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(S && acc && rows > 0 && R > 0 && ld >= R, "expert_rowmax: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(expert_rowmax_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, stream, false, reinterpret_cast<const float*>(S),
                               (long long)ld, reinterpret_cast<float*>(acc), (int)rows, (int)R, weight, (int)accumulate));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_expert_sim_maps(const void* sim, void* maps, void* simmask, int32_t B, int32_t G, int32_t OUT, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(sim && maps && simmask && B > 0 && G > 1 && OUT > 1, "expert_sim_maps: bad arguments");
  MYR_CHECK_CUDA(launch_kernel(expert_sim_maps_kernel, dim3(ceil_div(OUT * OUT, 256) + 1, B), dim3(256), 0, stream, false,
                               reinterpret_cast<const float*>(sim), reinterpret_cast<float*>(maps), reinterpret_cast<float*>(simmask), (int)B,
                               (int)G, (int)OUT));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
