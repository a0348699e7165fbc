// tcgen05 + TMA GEMM for sm_100a:  out[t, f] = epilogue( sum_k X[t, k] * W[f, k] ).
//
// Layout choice ("features on lanes"): the weight tile (128 features) is the UMMA A operand, so each of
// the 128 TMEM lanes holds one output feature; the token tile (BN = 16..256 tokens) is the UMMA N side.
// The same kernel therefore covers decode (T <= 32: weight streaming, HBM-bound, split-K over the SMs)
// and prefill / ViT (T in the thousands: tensor-bound, BN up to 256 => 128x256x16 MMAs at full rate).
//
// Persistent, warp-specialised CTA (192 threads, 1 CTA / SM):
//   warp 0   : TMA producer   (global -> 128B-swizzled smem ring, up to 8 stages, mbarrier complete_tx)
//   warp 1   : MMA issuer     (one thread issues tcgen05.mma; tcgen05.commit frees smem stages / publishes TMEM)
//   warps 2-5: epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> coalesced global stores)
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Replaces the cuBLAS calls behind nn.Linear in the reference hot path (see include/myriad_b200.h).
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int BM = 128;       // features per tile (UMMA M)
constexpr int BK = 64;        // halfs per k-block = one 128-byte swizzle row
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int SMEM_TILE_BUDGET = 192 * 1024;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STAGE_COLS = 256;

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
};

struct GemmKernelParams {
  int T, F, K;
  int BN, n_tt, n_ft, ksplit, kb_total, kb_per_split;
  int num_stages, stage_bytes;
  int x_mn, w_mn;
  uint32_t idesc;
  float* partial;  // [ksplit][T][F] fp32 when ksplit > 1
  int nb1, nbatch;                 // batched mode: batch index bidx -> (b0 = bidx / nb1, b1 = bidx % nb1)
  long long o_bs0, o_bs1;          // output element strides of the two batch dims
  Epilogue ep;
};

__device__ __forceinline__ void epilogue_store(const Epilogue& ep, float v, float bias_f, long long t, int f,
                                               long long boff = 0) {
  v += bias_f;
  if (ep.round_acc) v = round_f16(v);
  if (f < ep.scale_cols) v = round_f16(v * ep.scale);
  if (ep.act == MYR_ACT_GELU_ERF) v = round_f16(gelu_erf(v));
  else if (ep.act == MYR_ACT_RELU) v = fmaxf(v, 0.f);
  v *= ep.alpha;
  if (ep.res) {
    float r = (ep.res_dtype == MYR_F32) ? reinterpret_cast<const float*>(ep.res)[t * ep.ldr + f]
                                        : __half2float(reinterpret_cast<const __half*>(ep.res)[t * ep.ldr + f]);
    v += r;
  }
  const long long o = boff + (ep.group_rows ? (t / ep.group_rows) * ep.group_stride + (t % ep.group_rows) * ep.ldo + f
                                            : t * ep.ldo + f);
  if (ep.out_dtype == MYR_F32)
    reinterpret_cast<float*>(ep.out)[o] = v;
  else
    reinterpret_cast<__half*>(ep.out)[o] = __float2half_rn(v);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
               const GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.num_stages * p.stage_bytes);
  uint64_t* full = bars;                      // [MAX_STAGES]  TMA -> MMA
  uint64_t* empty = bars + MAX_STAGES;        // [MAX_STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * MAX_STAGES;    // [2]           MMA -> epilogue
  uint64_t* tempty = tfull + 2;               // [2]           epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.n_tt * p.n_ft * p.ksplit * p.nbatch;
  const uint32_t tx_bytes = A_STAGE_BYTES + p.BN * BK * 2;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tt = u % p.n_tt;
        const int ft = (u / p.n_tt) % p.n_ft;
        const int ks = (u / (p.n_tt * p.n_ft)) % p.ksplit;
        const int bidx = u / (p.n_tt * p.n_ft * p.ksplit);
        const int b0 = bidx / p.nb1, b1 = bidx % p.nb1;
        const int f0 = ft * BM, t0 = tt * p.BN;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          if (!p.w_mn) {
            tma_load_4d(sa, &tmW, &full[stage], kb * BK, f0, b1, b0);
          } else {
            tma_load_4d(sa, &tmW, &full[stage], f0, kb * BK, b1, b0);
            tma_load_4d(sa + 8192, &tmW, &full[stage], f0 + 64, kb * BK, b1, b0);
          }
          if (!p.x_mn) {
            tma_load_4d(sb, &tmX, &full[stage], kb * BK, t0, b1, b0);
          } else {
            for (int i = 0; i < p.BN / 64; ++i) tma_load_4d(sb + i * 8192, &tmX, &full[stage], t0 + 64 * i, kb * BK, b1, b0);
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t a_lbo = p.w_mn ? 8192 : 16, b_lbo = p.x_mn ? 8192 : 16;
      const uint32_t a_kstep = p.w_mn ? 2048 : 32, b_kstep = p.x_mn ? 2048 : 32;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int ks = (u / (p.n_tt * p.n_ft)) % p.ksplit;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STAGE_COLS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = make_smem_desc(sa + kk * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + kk * b_kstep, b_lbo, 1024);
            tc_mma_f16(d_tmem, da, db, p.idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);  // frees this smem stage once the MMAs above have read it
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull[as]);  // accumulator complete -> epilogue
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tt = u % p.n_tt;
      const int ft = (u / p.n_tt) % p.n_ft;
      const int ks = (u / (p.n_tt * p.n_ft)) % p.ksplit;
      const int bidx = u / (p.n_tt * p.n_ft * p.ksplit);
      const long long boff = (long long)(bidx / p.nb1) * p.o_bs0 + (long long)(bidx % p.nb1) * p.o_bs1;
      const int f = ft * BM + q * 32 + lane;
      const int t0 = tt * p.BN;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * ACC_STAGE_COLS;
      const bool f_ok = f < p.F;
      float bias_f = 0.f;
      if (p.ksplit == 1 && p.ep.bias && f_ok) bias_f = __half2float(p.ep.bias[f]);
      const int nchunks = min(p.BN, p.T - t0 + 15) / 16;  // chunks with at least one valid token
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[16];
        tmem_ld16(taddr + c * 16, r);
        tmem_ld_wait();
        if (f_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int t = t0 + c * 16 + j;
            if (t < p.T) {
              const float v = __uint_as_float(r[j]);
              if (p.ksplit == 1)
                epilogue_store(p.ep, v, bias_f, t, f, boff);
              else
                p.partial[((long long)ks * p.T + t) * p.F + f] = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// split-K: sum fp32 partials in fixed order (deterministic) and apply the epilogue.
__global__ void gemm_splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, int T, int F, Epilogue ep) {
  const long long n = (long long)T * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long t = i / F;
    float v = 0.f;
    for (int s = 0; s < ksplit; ++s) v += partial[(long long)s * n + i];
    const float b = ep.bias ? __half2float(ep.bias[f]) : 0.f;
    epilogue_store(ep, v, b, t, f);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tile-shape heuristics
// ---------------------------------------------------------------------------------------------------
struct Plan {
  int BN, n_tt, n_ft, ksplit, kb_total, kb_per_split, num_stages, stage_bytes;
};

static Plan make_plan(int T, int F, int K, int x_mn, int bn_hint, int ksplit_hint) {
  Plan pl;
  const int sms = sm_count();
  pl.n_ft = ceil_div(F, BM);
  pl.kb_total = ceil_div(K, BK);
  const int gran = x_mn ? 64 : 16;
  int best_bn = 0;
  if (bn_hint > 0) {
    best_bn = bn_hint;
  } else {
    const int t_pad = ceil_div(T, gran) * gran;
    if (t_pad <= 256) {
      best_bn = t_pad;
    } else {
      // cost model (cycles per k-step of 16): MMA = BN/2, smem operand reads = 32 + BN/4; rounds of persistent CTAs.
      double best = 1e30;
      for (int bn = 256; bn >= 64; bn -= gran) {
        const int n_tt = ceil_div(T, bn);
        const long long tiles = (long long)n_tt * pl.n_ft;
        const long long rounds = (tiles + sms - 1) / sms;
        const double per = (bn / 2.0 > 32 + bn / 4.0) ? bn / 2.0 : 32 + bn / 4.0;
        const double cost = rounds * (per + 6.0);
        if (cost < best - 1e-9) {
          best = cost;
          best_bn = bn;
        }
      }
    }
  }
  pl.BN = best_bn;
  pl.n_tt = ceil_div(T, pl.BN);
  const int tiles = pl.n_tt * pl.n_ft;
  int ks = 1;
  if (ksplit_hint > 0) {
    ks = ksplit_hint;
  } else if (tiles < sms && pl.kb_total >= 16) {
    // weight-streaming regime: spread k-blocks over idle SMs; pick the split with the best wave efficiency.
    double best = 0;
    for (int cand = 1; cand <= 16; ++cand) {
      const int per = ceil_div(pl.kb_total, cand);
      if (per < 8 && cand > 1) break;
      const int eff_ks = ceil_div(pl.kb_total, per);
      const long long units = (long long)tiles * eff_ks;
      const long long rounds = (units + sms - 1) / sms;
      const double eff = (double)units / (double)(rounds * sms);
      if (eff > best + 0.02) {
        best = eff;
        ks = eff_ks;
      }
    }
  }
  pl.kb_per_split = ceil_div(pl.kb_total, ks);
  pl.ksplit = ceil_div(pl.kb_total, pl.kb_per_split);
  pl.stage_bytes = A_STAGE_BYTES + pl.BN * BK * 2;
  int st = SMEM_TILE_BUDGET / pl.stage_bytes;
  pl.num_stages = st > MAX_STAGES ? MAX_STAGES : st;
  return pl;
}

}  // namespace myr

using namespace myr;

extern "C" size_t myr_gemm_workspace_bytes(int32_t T, int32_t F, int32_t K) {
  (void)K;
  // upper bound: 16-way split-K is only chosen when tiles < #SMs, i.e. small T*F
  return (size_t)16 * (size_t)T * (size_t)F * sizeof(float);
}

extern "C" int myr_gemm_f16(const myr_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(a != nullptr, "gemm: null args");
  MYR_CHECK_ARG(a->T > 0 && a->F > 0 && a->K > 0, "gemm: bad shape T=%d F=%d K=%d", a->T, a->F, a->K);
  const int nb0 = a->nb0 > 0 ? a->nb0 : 1, nb1 = a->nb1 > 0 ? a->nb1 : 1;
  const int nbatch = nb0 * nb1;
  MYR_CHECK_ARG(nbatch == 1 || (a->res == nullptr && a->out_group_rows == 0), "gemm: batched mode supports no residual / row groups");
  MYR_CHECK_ARG(a->x_bs0 % 8 == 0 && a->x_bs1 % 8 == 0 && a->w_bs0 % 8 == 0 && a->w_bs1 % 8 == 0,
                "gemm: batch strides must be multiples of 8 elements");
  MYR_CHECK_ARG(a->ldx % 8 == 0 && a->ldw % 8 == 0, "gemm: ldx=%lld ldw=%lld must be multiples of 8", (long long)a->ldx,
                (long long)a->ldw);
  MYR_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
                "gemm: x/w must be 16-byte aligned");
  MYR_CHECK_ARG(a->out != nullptr && a->x != nullptr && a->w != nullptr, "gemm: null pointer");
  MYR_CHECK_ARG(a->bn_hint == 0 || (a->bn_hint % 16 == 0 && a->bn_hint >= 16 && a->bn_hint <= 256),
                "gemm: bn_hint=%d must be a multiple of 16 in [16,256]", a->bn_hint);
  MYR_CHECK_ARG(!a->x_mn_major || a->bn_hint % 64 == 0, "gemm: MN-major x needs bn_hint multiple of 64");

  Plan pl = make_plan(a->T, a->F, a->K, a->x_mn_major, a->bn_hint, nbatch > 1 ? 1 : a->ksplit_hint);
  if (pl.ksplit > 1) {
    const size_t need = (size_t)pl.ksplit * a->T * a->F * sizeof(float);
    if (a->workspace == nullptr || a->workspace_bytes < need) {
      if (a->ksplit_hint > 0) {
        set_error("gemm: split-K=%d needs %zu workspace bytes, got %zu", pl.ksplit, need, a->workspace_bytes);
        return MYR_ERR_WORKSPACE;
      }
      pl = make_plan(a->T, a->F, a->K, a->x_mn_major, a->bn_hint, 1);  // silently valid: no split, same result order
    }
  }

  CUtensorMap tmW, tmX;
  {
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    dims[2] = (uint64_t)nb1; dims[3] = (uint64_t)nb0; box[2] = 1; box[3] = 1;
    if (!a->w_mn_major) {
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)a->F; box[0] = BK; box[1] = BM;
    } else {
      dims[0] = (uint64_t)a->F; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)a->ldw * 2;
    strides[1] = nb1 > 1 ? (uint64_t)a->w_bs1 * 2 : strides[0] * dims[1];
    strides[2] = nb0 > 1 ? (uint64_t)a->w_bs0 * 2 : strides[1] * dims[2];
    int rc = make_tmap_f16(&tmW, a->w, 4, dims, strides, box);
    if (rc) return rc;
    if (!a->x_mn_major) {
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)a->T; box[0] = BK; box[1] = (uint32_t)pl.BN;
    } else {
      dims[0] = (uint64_t)a->T; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)a->ldx * 2;
    strides[1] = nb1 > 1 ? (uint64_t)a->x_bs1 * 2 : strides[0] * dims[1];
    strides[2] = nb0 > 1 ? (uint64_t)a->x_bs0 * 2 : strides[1] * dims[2];
    rc = make_tmap_f16(&tmX, a->x, 4, dims, strides, box);
    if (rc) return rc;
  }

  GemmKernelParams p;
  p.T = a->T; p.F = a->F; p.K = a->K;
  p.BN = pl.BN; p.n_tt = pl.n_tt; p.n_ft = pl.n_ft; p.ksplit = pl.ksplit;
  p.kb_total = pl.kb_total; p.kb_per_split = pl.kb_per_split;
  p.num_stages = pl.num_stages; p.stage_bytes = pl.stage_bytes;
  p.x_mn = a->x_mn_major; p.w_mn = a->w_mn_major;
  p.idesc = make_idesc_f16(BM, pl.BN, a->w_mn_major, a->x_mn_major);
  p.partial = reinterpret_cast<float*>(a->workspace);
  p.nb1 = nb1; p.nbatch = nbatch; p.o_bs0 = a->o_bs0; p.o_bs1 = a->o_bs1;
  p.ep.bias = reinterpret_cast<const __half*>(a->bias);
  p.ep.act = a->act; p.ep.round_acc = a->round_acc;
  p.ep.scale_cols = a->scale_cols; p.ep.scale = a->scale;
  p.ep.res = a->res; p.ep.res_dtype = a->res_dtype; p.ep.ldr = a->ldr;
  p.ep.out = a->out; p.ep.out_dtype = a->out_dtype; p.ep.ldo = a->ldo;
  p.ep.group_rows = a->out_group_rows; p.ep.group_stride = a->out_group_stride;
  p.ep.alpha = a->alpha_set ? a->alpha : 1.0f;

  const size_t smem_bytes = (size_t)pl.num_stages * pl.stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int units = pl.n_tt * pl.n_ft * pl.ksplit * nbatch;
  const int grid = units < sm_count() ? units : sm_count();
  gemm_tc_kernel<<<grid, GEMM_THREADS, smem_bytes, stream>>>(tmW, tmX, p);
  MYR_CHECK_LAUNCH();
  if (pl.ksplit > 1) {
    const long long n = (long long)a->T * a->F;
    int rgrid = (int)((n + 255) / 256);
    if (rgrid > sm_count() * 8) rgrid = sm_count() * 8;
    gemm_splitk_reduce_kernel<<<rgrid, 256, 0, stream>>>(p.partial, pl.ksplit, a->T, a->F, p.ep);
    MYR_CHECK_LAUNCH();
  }
  return MYR_OK;
}
