// Training-side kernels of the expert-prior conv stacks (networks.py:95-197): unfused conv3x3+ReLU forward that keeps
// the pre-pool activation, maxpool+ReLU backward, direct wgrad/dgrad for the small-channel layers, col2im for the
// im2col/GEMM layers. NHWC, filters [Cout, kh, kw, Cin]; fp32 math on fp16 activations.
#include "common.h"
#include "ptx.cuh"

namespace myr {

static inline int grid_c(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ float ld_act(const void* p, int is_f32, size_t i) {
  return is_f32 ? reinterpret_cast<const float*>(p)[i] : __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// y[b, y, x, co] = relu(bias[co] + sum in[b, y+ky-1, x+kx-1, ci] * w[co, ky, kx, ci])   (no pooling)
__global__ void conv3x3_relu_kernel(const void* __restrict__ in, int in_f32, const float* __restrict__ w,
                                    const float* __restrict__ bias, __half* __restrict__ out, int B, int H, int W, int Cin,
                                    int Cout) {
  const long long total = (long long)B * H * W * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long r = i / Cout;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float a = bias[co];
    const float* wc = w + (size_t)co * 9 * Cin;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const size_t base = ((size_t)(b * H + yy) * W + xx) * Cin;
        for (int ci = 0; ci < Cin; ++ci) a = fmaf(ld_act(in, in_f32, base + ci), wc[(ky * 3 + kx) * Cin + ci], a);
      }
    }
    out[i] = __float2half_rn(fmaxf(a, 0.f));
  }
}

// dy[b, 2py+oy, 2px+ox, c] = dpool[b, py, px, c] at the arg-max of the 2x2 window (first maximum in row-major order,
// as torch's max_pool2d backward) if that post-ReLU value is > 0, else 0. y: post-ReLU pre-pool fp16.
__global__ void pool_relu_bwd_kernel(const __half* __restrict__ y, const void* __restrict__ dpool, int dp_f32,
                                     __half* __restrict__ dy, int B, int H, int W, int C) {
  const int OH = H >> 1, OW = W >> 1;
  const long long total = (long long)B * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int px = (int)(r % OW);
    r /= OW;
    const int py = (int)(r % OH);
    const int b = (int)(r / OH);
    const size_t p00 = ((size_t)(b * H + 2 * py) * W + 2 * px) * C + c;
    const size_t offs[4] = {p00, p00 + C, p00 + (size_t)W * C, p00 + (size_t)W * C + C};
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = __half2float(y[offs[k]]);
      if (v > best) {
        best = v;
        bi = k;
      }
    }
    const float g = best > 0.f ? ld_act(dpool, dp_f32, i) : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) dy[offs[k]] = __float2half_rn(k == bi ? g : 0.f);
  }
}

// direct weight/bias gradient: combos = Cout * (9 * Cin + 1); each block owns a chunk of positions, each thread a set
// of combos; block partials are added to dw / db with one atomic per (block, combo).
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const void* __restrict__ in, int in_f32, const __half* __restrict__ dy,
                                                            float* __restrict__ dw, float* __restrict__ db, int B, int H, int W,
                                                            int Cin, int Cout, int chunk, float scale) {
  const long long npos = (long long)B * H * W;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = p0 + chunk < npos ? p0 + chunk : npos;
  const int per_co = 9 * Cin + 1;
  const int combos = Cout * per_co;
  for (int cb = threadIdx.x; cb < combos; cb += blockDim.x) {
    const int co = cb / per_co, k = cb % per_co;
    float a = 0.f;
    if (k == 9 * Cin) {
      for (long long p = p0; p < p1; ++p) a += __half2float(dy[p * Cout + co]);
      atomicAdd(&db[co], scale * a);
    } else {
      const int tap = k / Cin, ci = k % Cin, ky = tap / 3, kx = tap % 3;
      for (long long p = p0; p < p1; ++p) {
        const int x = (int)(p % W), y = (int)((p / W) % H);
        const int yy = y + ky - 1, xx = x + kx - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const float g = __half2float(dy[p * Cout + co]);
        a = fmaf(g, ld_act(in, in_f32, (size_t)(p + (long long)(ky - 1) * W + (kx - 1)) * Cin + ci), a);
      }
      atomicAdd(&dw[(size_t)co * 9 * Cin + k], scale * a);
    }
  }
}

// The same gradient for the layers whose (output channel, tap, input channel) combinations fit one block (1 -> 4 channels at
// 224 x 224: 40 combinations; 4 -> 16 at 112 x 112: 592). A block walks image rows: it stages the row of dy and the three input
// rows around it (zero halo, so the inner loop has no bounds test and no index division) in shared memory; thread = one
// combination x one slice of the row's x range (`ngroups` slices when the combinations are few), accumulating in a register over
// all its rows; the slices meet in shared memory in slice order, then one atomic per (block, combination) as above.
// The position-chunk kernel above needed 162 / 487 us for these two layers (8 M / 29 M multiply-adds).
__global__ void __launch_bounds__(1024) conv3x3_wgrad_rows_kernel(const void* __restrict__ in, int in_f32, const __half* __restrict__ dy,
                                                                 float* __restrict__ dw, float* __restrict__ db, int B, int H, int W,
                                                                 int Cin, int Cout, int ngroups, float scale) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  float* in_s = reinterpret_cast<float*>(wg_smem);                      // [3][W + 2][Cin]
  __half* dy_s = reinterpret_cast<__half*>(in_s + 3 * (W + 2) * Cin);  // [W][Cout]
  const int per_co = 9 * Cin + 1, combos = Cout * per_co;
  const int tid = threadIdx.x;
  const int grp = tid / combos, cb = tid - grp * combos;
  const bool active = grp < ngroups;
  const int co = cb / per_co, k = cb % per_co;
  const bool is_bias = k == 9 * Cin;
  const int tap = is_bias ? 0 : k / Cin, ci = is_bias ? 0 : k % Cin, ky = tap / 3, kx = tap % 3;
  const int xw = (W + ngroups - 1) / ngroups;
  const int x0 = grp * xw, x1 = min(W, x0 + xw);
  const int n_in = 3 * (W + 2) * Cin, n_dy = W * Cout;
  float a = 0.f;
  for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
    const int b = row / H, y = row - b * H;
    __syncthreads();  // the previous row's readers are done
    for (int i = tid; i < n_in; i += blockDim.x) {
      const int c = i % Cin, r = i / Cin;
      const int xx = r % (W + 2) - 1, yy = y + r / (W + 2) - 1;
      in_s[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? ld_act(in, in_f32, ((size_t)(b * H + yy) * W + xx) * Cin + c) : 0.f;
    }
    for (int i = tid; i < n_dy; i += blockDim.x) dy_s[i] = dy[(size_t)row * n_dy + i];
    __syncthreads();
    if (active) {
      if (is_bias) {
        for (int x = x0; x < x1; ++x) a += __half2float(dy_s[x * Cout + co]);
      } else {
        const float* ip = in_s + (ky * (W + 2) + kx) * Cin + ci;
#pragma unroll 4
        for (int x = x0; x < x1; ++x) a = fmaf(__half2float(dy_s[x * Cout + co]), ip[x * Cin], a);
      }
    }
  }
  __syncthreads();
  float* red = in_s;  // >= blockDim floats (the launch sizes the buffer)
  red[tid] = active ? a : 0.f;
  __syncthreads();
  if (tid < combos) {
    float t = 0.f;
    for (int g = 0; g < ngroups; ++g) t += red[g * combos + tid];
    if (is_bias)
      atomicAdd(&db[co], scale * t);
    else
      atomicAdd(&dw[(size_t)co * 9 * Cin + k], scale * t);
  }
}

// direct input gradient: din[b, y, x, ci] = sum_{co, ky, kx} dy[b, y-ky+1, x-kx+1, co] * w[co, ky, kx, ci]
__global__ void conv3x3_dgrad_kernel(const __half* __restrict__ dy, const float* __restrict__ w, __half* __restrict__ din, int B,
                                     int H, int W, int Cin, int Cout) {
  const long long total = (long long)B * H * W * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    long long r = i / Cin;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float a = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y - ky + 1;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x - kx + 1;
        if (xx < 0 || xx >= W) continue;
        const __half* g = dy + ((size_t)(b * H + yy) * W + xx) * Cout;
        for (int co = 0; co < Cout; ++co) a = fmaf(__half2float(g[co]), w[((size_t)co * 9 + ky * 3 + kx) * Cin + ci], a);
      }
    }
    din[i] = __float2half_rn(a);
  }
}

// col2im (gather form): din[b, y, x, c] = sum_{ky, kx} dcols[(b, y+pad-ky, x+pad-kx), (ky, kx, c)] over valid outputs
__global__ void col2im_kernel(const __half* __restrict__ dcols, __half* __restrict__ din, int B, int H, int W, int C, int KH,
                              int KW, int pad, int OH, int OW) {
  const long long total = (long long)B * H * W * C;
  const size_t ldc = (size_t)KH * KW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float a = 0.f;
    for (int ky = 0; ky < KH; ++ky) {
      const int oy = y + pad - ky;
      if (oy < 0 || oy >= OH) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int ox = x + pad - kx;
        if (ox < 0 || ox >= OW) continue;
        a += __half2float(dcols[(((size_t)b * OH + oy) * OW + ox) * ldc + (size_t)(ky * KW + kx) * C + c]);
      }
    }
    din[i] = __float2half_rn(a);
  }
}

}  // namespace myr

using namespace myr;

#define STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)

extern "C" int myr_conv3x3_relu(const void* in, int32_t in_dtype, const void* w, const void* bias, void* out, int32_t B, int32_t H,
                                int32_t W, int32_t Cin, int32_t Cout, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(in && w && bias && out && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_relu: bad arguments");
  conv3x3_relu_kernel<<<grid_c((long long)B * H * W * Cout, 128), 128, 0, stream>>>(
      in, in_dtype == MYR_F32, reinterpret_cast<const float*>(w), reinterpret_cast<const float*>(bias), reinterpret_cast<__half*>(out),
      B, H, W, Cin, Cout);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_pool_relu_bwd(const void* y, const void* dpool, int32_t dpool_dtype, void* dy, int32_t B, int32_t H, int32_t W,
                                 int32_t C, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(y && dpool && dy && B > 0 && H % 2 == 0 && W % 2 == 0 && C > 0, "pool_relu_bwd: bad arguments");
  pool_relu_bwd_kernel<<<grid_c((long long)B * (H / 2) * (W / 2) * C, 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(y), dpool, dpool_dtype == MYR_F32, reinterpret_cast<__half*>(dy), B, H, W, C);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_conv3x3_wgrad(const void* in, int32_t in_dtype, const void* dy, void* dw, void* db, int32_t B, int32_t H,
                                 int32_t W, int32_t Cin, int32_t Cout, float scale, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(in && dy && dw && db && B > 0 && Cin > 0 && Cout > 0, "conv3x3_wgrad: bad arguments");
  const long long npos = (long long)B * H * W;
  MYR_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * 9 * Cin, stream));
  MYR_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, stream));
  const int combos = Cout * (9 * Cin + 1);
  static int rows_variant = -1;
  if (rows_variant < 0) {
    const char* e = getenv("MYR_WGRAD_ROWS");
    rows_variant = (e && e[0] == '0') ? 0 : 1;
  }
  if (combos <= 1024 && rows_variant) {
    int ngroups = 512 / combos;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > 16) ngroups = 16;
    const int threads = (combos * ngroups + 31) / 32 * 32;
    size_t smem = sizeof(float) * 3 * (size_t)(W + 2) * Cin + sizeof(__half) * (size_t)W * Cout;
    if (smem < sizeof(float) * threads) smem = sizeof(float) * threads;
    if (smem <= 48 * 1024) {
      int grid = 2 * sm_count();
      if (grid > B * H) grid = B * H;
      conv3x3_wgrad_rows_kernel<<<grid, threads, smem, stream>>>(in, in_dtype == MYR_F32, reinterpret_cast<const __half*>(dy),
                                                                 reinterpret_cast<float*>(dw), reinterpret_cast<float*>(db), B, H, W, Cin,
                                                                 Cout, ngroups, scale);
      MYR_CHECK_LAUNCH();
      return MYR_OK;
    }
  }
  int chunk = 512;
  while (npos / chunk > 4096) chunk *= 2;
  conv3x3_wgrad_kernel<<<(unsigned)((npos + chunk - 1) / chunk), 256, 0, stream>>>(
      in, in_dtype == MYR_F32, reinterpret_cast<const __half*>(dy), reinterpret_cast<float*>(dw), reinterpret_cast<float*>(db), B, H,
      W, Cin, Cout, chunk, scale);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_conv3x3_dgrad(const void* dy, const void* w, void* din, int32_t B, int32_t H, int32_t W, int32_t Cin,
                                 int32_t Cout, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(dy && w && din && B > 0 && Cin > 0 && Cout > 0, "conv3x3_dgrad: bad arguments");
  conv3x3_dgrad_kernel<<<grid_c((long long)B * H * W * Cin, 128), 128, 0, stream>>>(
      reinterpret_cast<const __half*>(dy), reinterpret_cast<const float*>(w), reinterpret_cast<__half*>(din), B, H, W, Cin, Cout);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

extern "C" int myr_col2im(const void* dcols, void* din, int32_t B, int32_t H, int32_t W, int32_t C, int32_t KH, int32_t KW,
                          int32_t pad, void* stream_) {
  STREAM;
  MYR_CHECK_ARG(dcols && din && B > 0 && C > 0, "col2im: bad arguments");
  const int OH = H + 2 * pad - KH + 1, OW = W + 2 * pad - KW + 1;
  MYR_CHECK_ARG(OH > 0 && OW > 0, "col2im: empty output");
  col2im_kernel<<<grid_c((long long)B * H * W * C, 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(dcols),
                                                                          reinterpret_cast<__half*>(din), B, H, W, C, KH, KW, pad,
                                                                          OH, OW);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
