// Expert-prior conv stacks (VEInstructorV2 / VETokenizer, networks.py:95-197) in NHWC fp16.
//   layers 1-3 (1->4->16->64 channels, 224->28): direct conv3x3(pad 1) + bias + ReLU + maxpool2 fused, fp32 math
//   layers 4-5 (64->256->1024) and the heads (1x1 -> 768, 5x5 -> 4096): im2col (this file) + the tcgen05 GEMM
//   (bias + ReLU epilogue) + maxpool2 (this file). Weights are pre-permuted to [Cout, kh, kw, Cin] on the host.
#include "common.h"
#include "ptx.cuh"

namespace myr {

// one thread = one pooled output (b, py, px, co); co fastest so a warp shares its 4x4xCin input window.
__global__ void conv3x3_relu_pool_kernel(const void* __restrict__ in, int in_is_f32, const float* __restrict__ w,
                                         const float* __restrict__ bias, __half* __restrict__ out, int B, int H, int W,
                                         int Cin, int Cout) {
  const int OH = H >> 1, OW = W >> 1;
  const long long total = (long long)B * OH * OW * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long r = i / Cout;
    const int px = (int)(r % OW);
    r /= OW;
    const int py = (int)(r % OH);
    const int b = (int)(r / OH);
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
    const float* wc = w + (size_t)co * 9 * Cin;
    const int y0 = 2 * py - 1, x0 = 2 * px - 1;  // top-left of the 4x4 input window
    for (int dy = 0; dy < 4; ++dy) {
      const int y = y0 + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = 0; dx < 4; ++dx) {
        const int x = x0 + dx;
        if (x < 0 || x >= W) continue;
        const size_t base = ((size_t)(b * H + y) * W + x) * Cin;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = in_is_f32 ? reinterpret_cast<const float*>(in)[base + ci]
                                    : __half2float(reinterpret_cast<const __half*>(in)[base + ci]);
          // window element (dy, dx) feeds conv output (oy, ox) with kernel tap (dy - oy, dx - ox)
          if (dy <= 2 && dx <= 2) a00 = fmaf(v, wc[(dy * 3 + dx) * Cin + ci], a00);
          if (dy <= 2 && dx >= 1) a01 = fmaf(v, wc[(dy * 3 + dx - 1) * Cin + ci], a01);
          if (dy >= 1 && dx <= 2) a10 = fmaf(v, wc[((dy - 1) * 3 + dx) * Cin + ci], a10);
          if (dy >= 1 && dx >= 1) a11 = fmaf(v, wc[((dy - 1) * 3 + dx - 1) * Cin + ci], a11);
        }
      }
    }
    const float m = fmaxf(fmaxf(a00, a01), fmaxf(a10, a11)) + bias[co];
    out[i] = __float2half_rn(fmaxf(m, 0.f));
  }
}

// im2col on NHWC fp16: out[(b, oy, ox), (ky, kx, c)] = in[b, oy - pad + ky, ox - pad + kx, c] (0 outside)
__global__ void im2col_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C, int KH,
                              int KW, int pad, int OH, int OW) {
  const int cv = C >> 3;
  const long long total = (long long)B * OH * OW * KH * KW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int kx = (int)(r % KW);
    r /= KW;
    const int ky = (int)(r % KH);
    r /= KH;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const int y = oy - pad + ky, x = ox - pad + kx;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && x >= 0 && x < W) v = *reinterpret_cast<const uint4*>(in + ((size_t)(b * H + y) * W + x) * C + c);
    const size_t orow = ((size_t)b * OH + oy) * OW + ox;
    *reinterpret_cast<uint4*>(out + orow * ((size_t)KH * KW * C) + (size_t)(ky * KW + kx) * C + c) = v;
  }
}

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void maxpool2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C) {
  const int OH = H >> 1, OW = W >> 1, cv = C >> 3;
  const long long total = (long long)B * OH * OW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int px = (int)(r % OW);
    r /= OW;
    const int py = (int)(r % OH);
    const int b = (int)(r / OH);
    const __half* p00 = in + ((size_t)(b * H + 2 * py) * W + 2 * px) * C + c;
    const uint4 a = *reinterpret_cast<const uint4*>(p00), bq = *reinterpret_cast<const uint4*>(p00 + C);
    const uint4 cq = *reinterpret_cast<const uint4*>(p00 + (size_t)W * C), d = *reinterpret_cast<const uint4*>(p00 + (size_t)W * C + C);
    uint4 o;
    o.x = hmax2_u32(hmax2_u32(a.x, bq.x), hmax2_u32(cq.x, d.x));
    o.y = hmax2_u32(hmax2_u32(a.y, bq.y), hmax2_u32(cq.y, d.y));
    o.z = hmax2_u32(hmax2_u32(a.z, bq.z), hmax2_u32(cq.z, d.z));
    o.w = hmax2_u32(hmax2_u32(a.w, bq.w), hmax2_u32(cq.w, d.w));
    *reinterpret_cast<uint4*>(out + (((size_t)b * OH + py) * OW + px) * C + c) = o;
  }
}

__global__ void maxpool2_scalar_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C) {
  const int OH = H >> 1, OW = W >> 1;
  const long long total = (long long)B * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int px = (int)(r % OW);
    r /= OW;
    const int py = (int)(r % OH);
    const int b = (int)(r / OH);
    const __half* p00 = in + ((size_t)(b * H + 2 * py) * W + 2 * px) * C + c;
    const float m = fmaxf(fmaxf(__half2float(p00[0]), __half2float(p00[C])),
                          fmaxf(__half2float(p00[(size_t)W * C]), __half2float(p00[(size_t)W * C + C])));
    out[i] = __float2half_rn(m);
  }
}

// ViT patchify (PatchEmbed conv 14x14 stride 14 as a GEMM, eva_vit.py:196-203): image fp32 NCHW ->
// patches fp16 [B * g * g, ldp], column order (c, ky, kx) == proj.weight.reshape(D, 3 * P * P); pad columns zeroed.
__global__ void patchify_kernel(const float* __restrict__ img, __half* __restrict__ out, int B, int C, int HW, int P,
                                int ldp) {
  const int g = HW / P;
  const long long total = (long long)B * g * g * C * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ky = (int)(i % P);
    long long r = i / P;
    const int c = (int)(r % C);
    r /= C;
    const int px = (int)(r % g);
    r /= g;
    const int py = (int)(r % g);
    const int b = (int)(r / g);
    const float* src = img + (((size_t)b * C + c) * HW + (py * P + ky)) * HW + px * P;
    __half* dst = out + ((size_t)(b * g + py) * g + px) * ldp + (c * P + ky) * P;
    for (int kx = 0; kx < P; ++kx) dst[kx] = __float2half_rn(src[kx]);
    if (c == C - 1 && ky == P - 1)
      for (int k = C * P * P; k < ldp; ++k) out[((size_t)(b * g + py) * g + px) * ldp + k] = __float2half_rn(0.f);
  }
}

static inline int ew_grid2(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace myr

using namespace myr;

extern "C" int myr_conv3x3_relu_pool(const void* in, int32_t in_dtype, const void* w, const void* bias, void* out,
                                     int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(in && w && bias && out && B > 0 && H % 2 == 0 && W % 2 == 0 && Cin > 0 && Cout > 0, "conv3x3: bad arguments");
  const long long total = (long long)B * (H / 2) * (W / 2) * Cout;
  conv3x3_relu_pool_kernel<<<ew_grid2(total, 128), 128, 0, stream>>>(in, in_dtype == MYR_F32, reinterpret_cast<const float*>(w),
                                                                     reinterpret_cast<const float*>(bias),
                                                                     reinterpret_cast<__half*>(out), B, H, W, Cin, Cout);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_im2col(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t KH, int32_t KW,
                          int32_t pad, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(in && out && B > 0 && C % 8 == 0, "im2col: bad arguments (C must be a multiple of 8)");
  const int OH = H + 2 * pad - KH + 1, OW = W + 2 * pad - KW + 1;
  MYR_CHECK_ARG(OH > 0 && OW > 0, "im2col: empty output");
  const long long total = (long long)B * OH * OW * KH * KW * (C / 8);
  im2col_kernel<<<ew_grid2(total, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out), B,
                                                         H, W, C, KH, KW, pad, OH, OW);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_maxpool2(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(in && out && B > 0 && H % 2 == 0 && W % 2 == 0 && C > 0, "maxpool2: bad arguments");
  if (C % 8 == 0) {
    const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
    maxpool2_kernel<<<ew_grid2(total, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out),
                                                             B, H, W, C);
  } else {
    const long long total = (long long)B * (H / 2) * (W / 2) * C;
    maxpool2_scalar_kernel<<<ew_grid2(total, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(in),
                                                                    reinterpret_cast<__half*>(out), B, H, W, C);
  }
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_patchify(const void* image, void* patches, int32_t B, int32_t C, int32_t HW, int32_t P, int32_t ldp,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(image && patches && B > 0 && HW % P == 0 && ldp >= C * P * P && ldp % 8 == 0, "patchify: bad arguments");
  const long long total = (long long)B * (HW / P) * (HW / P) * C * P;
  patchify_kernel<<<ew_grid2(total, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(image),
                                                           reinterpret_cast<__half*>(patches), B, C, HW, P, ldp);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
