// Weight streaming for 5 <= T <= 32 tokens on sm_100a:  out[t, f] = epilogue( sum_k x[t, k] * W[f, k] ).
//
// Greedy decode of the throughput sweep's batches of 16 / 32 sequences (BASELINE.json configs[4]; one token per sequence and
// step, modeling_llama.py:730-760) still multiplies every LLaMA weight matrix by a handful of activation rows: 2 T flop per
// weight byte, HBM streaming. gemv.cu covers T <= 4 by keeping the whole [4, K] activation block in shared memory; at 32
// tokens that block is 256 - 700 KB, and the tcgen05 arrangement of gemm.cu pays a TMEM drain and a split-tile fix-up per
// launch (0.3 - 0.4 of the HBM rate in profiles/r2_sweep.json). This kernel keeps gemv.cu's structure - work items of 8 weight
// rows reduced entirely inside one CTA, one TMA producer thread, 8 mma.sync consumer warps, static + atomic work split, weights
// requested ahead of the programmatic-dependent-launch wait - and streams the ACTIVATIONS through the same ring instead:
//   * a ring stage = 512 k of up to 32 weight rows + the same 512 k of the token block (8 NT tokens, NT = ceil(T / 8), rows >= T
//     zero-filled by TMA), each ONE 128-byte-swizzled tensor box [k-block][row][64 halfs] (a copy costs the producer ~130 clk
//     whatever its size: 2 - 3 copies per 32 KB of weights). The token block comes out of L2 once per 32-row group, i.e. <= 1
//     byte of L2 traffic per byte of HBM traffic (the SM's ingest, ~43 B/clk, carries both);
//   * consumer warp w owns k-block w (64 k) of every stage: 2 A fragments (rows 0-15, 16-31) and NT B fragments per 16 k by
//     ldmatrix, 2 NT mma.sync.m16n8k16 with fp32 accumulators held over the whole K walk; the 8 warps' partial sums meet in
//     shared memory once per group and are added in warp order (deterministic: CUDA-graph replay == eager);
//   * epilogue: bias, fp16 / fp32 residual (in place, requested a whole group ahead), fp16 / fp32 store, or SwiGLU over the
//     64-row interleaved gate / up weights (modeling_llama.py:139-140; a group is then 16 gate rows + the 16 matching up rows).
// Replaces, for 5 <= T <= 32: nn.Linear of q/k/v/o/gate/up/down/lm_head (modeling_llama.py:139-140,168-231,629-716) + the peft
// LoRA-A rows riding on the qkv weight (myriad.py:171-178).
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int MT_CWARPS = 8;
constexpr int MT_CTHREADS = MT_CWARPS * 32;
constexpr int MT_THREADS = MT_CTHREADS + 32;       // + producer warp
constexpr int MT_SK = 512;                         // k elements per stage: one 64-k block per consumer warp
constexpr int MT_BOX_BYTES = 8 * MT_SK * 2;        // 8 rows x 512 k: 8 KiB
constexpr int MT_W_BYTES = 4 * MT_BOX_BYTES;       // up to 32 weight rows per stage
struct GemvMtMaps {
  CUtensorMap w[4];  // weights (64 k, F rows, K / 64 k-blocks) with boxes of 8 / 16 / 24 / 32 rows x 8 k-blocks
  CUtensorMap x;     // tokens  (64 k, T rows, K / 64) with one box of 8 NT rows x 8 k-blocks
};
constexpr int MT_MAX_STAGES = 6;
constexpr int MT_SMEM_BUDGET = 227 * 1024;

struct GemvMtParams {
  int F, K, T;
  const __half* bias;
  const void* res; int res_dtype; long long ldr;
  void* out; int out_dtype; long long ldo;
  int swiglu;
  int n_units;   // units in total: 8 output rows (SwiGLU: 8 gate / up pairs = 8 + 8 weight rows)
  int gsz;       // units per row group: 4 (plain: 32 rows) or 2 (SwiGLU: 16 pairs)
  int u_static;  // units [0, u_static) are split evenly by CTA index, the rest go through `counter`
  int* counter;  // zero on entry, left at zero (null: u_static == n_units)
  int n_kc;      // stages per row group = ceil(K / 512)
  int stages;
  int w_static;
};

__device__ __forceinline__ void mt_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mt_ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mt_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MT_CTHREADS) : "memory"); }

// first weight row of the row group that starts at unit u0 (SwiGLU: of its gate rows; the matching up rows follow 64 rows on:
// weights interleaved in blocks of 64, [gate 0..63 | up 0..63 | gate 64..127 | ...]; a two-unit group starts at an even unit,
// so its 16 pairs never straddle a block)
__device__ __forceinline__ int mt_row0(const GemvMtParams& p, int u0) {
  if (!p.swiglu) return u0 * 8;
  const int i0 = u0 * 8;
  return ((i0 >> 6) << 7) + (i0 & 63);
}

template <int NT>
__global__ void __launch_bounds__(MT_THREADS) gemv_mt_kernel(const __grid_constant__ GemvMtMaps tm, const GemvMtParams p) {
  constexpr int STAGE_BYTES = MT_W_BYTES + NT * MT_BOX_BYTES;
  constexpr int RED_FLOATS = MT_CWARPS * 2 * NT * 128;  // [warp][fragment][token slice][lane][4]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  uint8_t* ring = smem;
  float* red = reinterpret_cast<float*>(ring + (size_t)p.stages * STAGE_BYTES);
  int* s_gid = reinterpret_cast<int*>(red + RED_FLOATS);     // [stages] 8 * first unit + units of the stage's row group (-1: end)
  uint64_t* full = reinterpret_cast<uint64_t*>(s_gid + 8);   // [stages] producer -> consumers
  uint64_t* empty = full + MT_MAX_STAGES;                    // [stages] consumers -> producer

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MT_CWARPS);
    }
    fence_mbar_init();
    for (int i = 0; i < (p.swiglu ? 2 : 4); ++i) tma_prefetch_desc(&tm.w[i]);
    tma_prefetch_desc(&tm.x);
  }
  __syncthreads();
  pdl_launch_dependents();

  if (warp == MT_CWARPS) {
    // ---------------- producer: one thread; per stage one weight box (SwiGLU: gate + up) + one token box ----------------
    if (lane == 0) {
      const uint64_t pol_w = l2_policy_evict_first(), pol_x = l2_policy_evict_last();
      // Weights are static: the stages of the first ring pass are requested before the programmatic-dependent-launch wait; the
      // token boxes of those stages (written by the previous kernel) follow right after it.
      bool waited = false;
      int n_pend = 0, pend_st[MT_MAX_STAGES], pend_kc[MT_MAX_STAGES];
      auto x_boxes = [&](int st, int kc) {
        tma_load_3d_hint(smem_u32(ring) + st * STAGE_BYTES + MT_W_BYTES, &tm.x, smem_u32(&full[st]), 0, 0, kc * (MT_SK / 64), pol_x);
      };
      auto flush = [&]() {
        if (waited) return;
        pdl_wait();
        waited = true;
        for (int i = 0; i < n_pend; ++i) x_boxes(pend_st[i], pend_kc[i]);
        n_pend = 0;
      };
      if (!p.w_static) flush();
      // work list of this CTA: its slice of the evenly split units [0, u_static), walked in row groups of gsz units, then row
      // groups from the shared pool [u_static, n_units)
      const int s1 = (int)(((long long)(blockIdx.x + 1) * p.u_static) / gridDim.x);
      const int su = (int)(((long long)blockIdx.x * p.u_static) / gridDim.x);
      const int pool = (p.n_units - p.u_static + p.gsz - 1) / p.gsz;
      auto grab = [&]() {  // -> first unit of a pool group, or n_units
        if (p.counter == nullptr) return p.n_units;
        flush();  // the counter is shared with the previous launch, which leaves it at zero when it completes
        const int v = atomicAdd(p.counter, 1);
        if (v == pool + (int)gridDim.x - 1) *reinterpret_cast<volatile int*>(p.counter) = 0;  // the last grab of this launch
        return v < pool ? p.u_static + v * p.gsz : p.n_units;
      };
      int st = 0, n_issued = 0;
      uint32_t phase = 0;
      int u0 = su < s1 ? su : grab();
      while (u0 < p.n_units) {
        const int lim = u0 < p.u_static ? s1 : p.n_units;
        const int nu = (p.swiglu && (u0 & 1)) ? 1 : min(p.gsz, lim - u0);
        int next = -1;
        if (u0 < p.u_static && u0 + nu < s1) next = u0 + nu;
        else if (waited) next = grab();  // fetched while this group is being requested: the atomic's latency hides
        const int row0 = mt_row0(p, u0);
        const CUtensorMap* mw = &tm.w[nu - 1];
        for (int kc = 0; kc < p.n_kc; ++kc) {
          if (n_issued == p.stages) flush();  // the ring is full of weight-only stages: nothing drains until the tokens land
          mbar_wait(&empty[st], phase ^ 1);
          s_gid[st] = u0 * 8 + nu;
          mbar_arrive_expect_tx(&full[st], (uint32_t)(((p.swiglu ? 2 * nu : nu) + NT) * MT_BOX_BYTES));
          const uint32_t dst = smem_u32(ring) + st * STAGE_BYTES, bar = smem_u32(&full[st]);
          tma_load_3d_hint(dst, mw, bar, 0, row0, kc * (MT_SK / 64), pol_w);
          if (p.swiglu) tma_load_3d_hint(dst + 2 * MT_BOX_BYTES, mw, bar, 0, row0 + 64, kc * (MT_SK / 64), pol_w);
          if (waited) {
            x_boxes(st, kc);
          } else {
            pend_st[n_pend] = st;
            pend_kc[n_pend] = kc;
            ++n_pend;
          }
          ++n_issued;
          if (++st == p.stages) {
            st = 0;
            phase ^= 1;
          }
        }
        if (next < 0) next = grab();
        u0 = next;
      }
      flush();
      // end marker
      mbar_wait(&empty[st], phase ^ 1);
      s_gid[st] = -1;
      mbar_arrive(&full[st]);
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  pdl_wait();
  // A fragments: a box of R rows lies as [k-block][R rows][128 B], 16-byte unit u of row r at unit u ^ (r % 8). ldmatrix.x4 lanes
  // 0-7 / 8-15 / 16-23 / 24-31 address (rows 0-7, k 0-7) (rows 8-15, k 0-7) (rows 0-7, k 8-15) (rows 8-15, k 8-15) of a 16 x 16
  // fragment. Plain: fragment f = rows 16 f .. 16 f + 15 of the one weight box (R = 8 units-of-the-group rows); SwiGLU: fragment 0
  // = the gate box, fragment 1 = the up box (16 KiB on). Rows a short group does not have read stale bytes of the stage:
  // garbage that is dropped.
  const int a_r = lane & 7, a_rh = (lane >> 3) & 1, a_hi = lane >> 4;
  const uint32_t a_lane = smem_u32(ring) + (a_rh * 8 + a_r) * 128;
  const uint32_t a_f1 = p.swiglu ? 2 * MT_BOX_BYTES : 16 * 128;
  // B fragments: the token box [k-block][8 NT tokens][128 B]; ldmatrix.x4 lanes 0-7 / 8-15 address (tokens 0-7, k 0-7) / (k 8-15)
  // of token slice n, lanes 16-31 the same of slice n + 1: registers {b0, b1} of slice n, {b0, b1} of slice n + 1
  const int b_r = lane & 7, b_hi = (lane >> 3) & 1, b_n = lane >> 4;
  const uint32_t b_lane = smem_u32(ring) + MT_W_BYTES + warp * (NT * 1024) + (b_n * 8 + b_r) * 128;
  const bool k_ok_all = (p.K % MT_SK) == 0;

  float d[2][NT][4];
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) d[f][n][j] = 0.f;

  // epilogue roles. plain: row e_r = tid % 32, tokens e_t0 + 8 i; SwiGLU: pair e_r = tid % 16, tokens e_t0 + 16 i
  const int e_r = p.swiglu ? (tid & 15) : (tid & 31);
  const int e_t0 = p.swiglu ? (tid >> 4) : (tid >> 5);
  constexpr int E_N = NT;  // outputs per thread (plain: NT tokens; SwiGLU: ceil(NT / 2))
  float pre_res[E_N], pre_bias = 0.f;
#pragma unroll
  for (int i = 0; i < E_N; ++i) pre_res[i] = 0.f;

  int kc = 0, st = 0;
  uint32_t phase = 0;
  for (;;) {
    mbar_wait(&full[st], phase);
    const int g = s_gid[st];
    if (g < 0) break;
    if (kc == 0 && !p.swiglu && (p.res || p.bias)) {
      // epilogue operands of this thread's outputs: requested now, a whole row group of streaming ahead of their use
      const int row = (g >> 3) * 8 + e_r;
      const bool ok = e_r < (g & 7) * 8 && row < p.F;
      pre_bias = (ok && p.bias) ? __half2float(__ldg(p.bias + row)) : 0.f;
#pragma unroll
      for (int i = 0; i < E_N; ++i) {
        const int t = e_t0 + 8 * i;
        pre_res[i] = 0.f;
        if (ok && p.res && t < p.T)
          pre_res[i] = (p.res_dtype == MYR_F32) ? __ldcg(reinterpret_cast<const float*>(p.res) + (long long)t * p.ldr + row)
                                                : __half2float(__ldcg(reinterpret_cast<const __half*>(p.res) + (long long)t * p.ldr + row));
      }
    }
    if (k_ok_all || kc * MT_SK + warp * 64 < p.K) {  // K is a multiple of 64: a warp's k-block is whole or absent
      const uint32_t a_st = a_lane + (uint32_t)st * STAGE_BYTES + (uint32_t)(warp * (g & 7)) * 1024;  // k-block stride = rows * 128
      const uint32_t b_st = b_lane + (uint32_t)st * STAGE_BYTES;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0[4], a1[4];
        const uint32_t a_off = ((((ks << 1) + a_hi) ^ a_r) << 4);
        mt_ldsm_x4(a_st + a_off, a0);
        mt_ldsm_x4(a_st + a_f1 + a_off, a1);
        const uint32_t b_off = ((((ks << 1) + b_hi) ^ b_r) << 4);
#pragma unroll
        for (int n = 0; n + 1 < NT; n += 2) {
          uint32_t b[4];
          mt_ldsm_x4(b_st + n * 1024 + b_off, b);
          mt_mma(d[0][n], a0, b[0], b[1]);
          mt_mma(d[1][n], a1, b[0], b[1]);
          mt_mma(d[0][n + 1], a0, b[2], b[3]);
          mt_mma(d[1][n + 1], a1, b[2], b[3]);
        }
        if (NT & 1) {
          uint32_t b[2];
          // .x2 takes its addresses from lanes 0-15 (b_n == 0 there); lanes 16-31 pass a valid address that is ignored
          mt_ldsm_x2(b_st - b_n * 1024 + (NT - 1) * 1024 + b_off, b);
          mt_mma(d[0][NT - 1], a0, b[0], b[1]);
          mt_mma(d[1][NT - 1], a1, b[0], b[1]);
        }
      }
    }
    __syncwarp();  // every lane's ldmatrix of this stage is done before the stage is handed back
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == p.stages) {
      st = 0;
      phase ^= 1;
    }
    if (++kc < p.n_kc) continue;

    // ---- group complete: the 8 warps' partial sums meet in shared memory ----
    kc = 0;
    mt_consumer_sync();  // the previous group's epilogue has read `red` (long ago: this barrier does not wait in practice)
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        *reinterpret_cast<float4*>(red + ((warp * 2 + f) * NT + n) * 128 + lane * 4) = make_float4(d[f][n][0], d[f][n][1], d[f][n][2], d[f][n][3]);
        d[f][n][0] = d[f][n][1] = d[f][n][2] = d[f][n][3] = 0.f;
      }
    mt_consumer_sync();
    // D fragment of (fragment f, slice n): lane l holds (row l / 4, tokens 8 n + 2 (l % 4) + {0, 1}) in [0..1] and (row l / 4 + 8,
    // same tokens) in [2..3]
    auto total = [&](int f, int rr, int t) {
      const int l = (rr & 7) * 4 + ((t & 7) >> 1), j = (t & 1) + 2 * (rr >> 3), n = t >> 3;
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < MT_CWARPS; ++w) a += red[((w * 2 + f) * NT + n) * 128 + l * 4 + j];
      return a;
    };
    const int u0 = g >> 3, nu = g & 7;
    if (!p.swiglu) {
      const int row = u0 * 8 + e_r;
      if (e_r < nu * 8 && row < p.F) {
#pragma unroll
        for (int i = 0; i < E_N; ++i) {
          const int t = e_t0 + 8 * i;
          if (t < p.T) {
            float a = total(e_r >> 4, e_r & 15, t);
            a += pre_bias;
            a += pre_res[i];
            if (p.out_dtype == MYR_F32)
              reinterpret_cast<float*>(p.out)[(long long)t * p.ldo + row] = a;
            else
              reinterpret_cast<__half*>(p.out)[(long long)t * p.ldo + row] = __float2half_rn(a);
          }
        }
      }
    } else if (e_r < nu * 8) {
      // SwiGLU (modeling_llama.py:139-140) with the rounding points of the unfused path: gate / up rounded to fp16 first;
      // pair e_r of the group: gate row = fragment 0 row e_r, up row = fragment 1 row e_r
#pragma unroll
      for (int i = 0; i < (NT + 1) / 2; ++i) {
        const int t = e_t0 + 16 * i;
        if (t < p.T) {
          const float a = round_f16(total(0, e_r, t)), b = round_f16(total(1, e_r, t));
          reinterpret_cast<__half*>(p.out)[(long long)t * p.ldo + u0 * 8 + e_r] = __float2half_rn(silu_f(a) * b);
        }
      }
    }
  }
}

}  // namespace myr

using namespace myr;

struct GemvMtPlan {
  int nt, n_units, gsz, grid, u_static, use_counter, stages, n_kc;
  size_t smem;
};
static GemvMtPlan gemv_mt_plan(int T, int F, int K, bool swiglu, int sms, bool have_counter) {
  GemvMtPlan pl;
  pl.nt = ceil_div(T, 8);
  pl.n_kc = ceil_div(K, MT_SK);
  pl.n_units = swiglu ? F / 16 : ceil_div(F, 8);
  pl.gsz = swiglu ? 2 : 4;
  const int n_groups = ceil_div(pl.n_units, pl.gsz);
  pl.grid = pl.n_units < sms ? pl.n_units : sms;
  pl.use_counter = (have_counter && n_groups >= 4 * pl.grid) ? 1 : 0;
  pl.u_static = pl.use_counter ? (int)((long long)pl.n_units * 3 / 4) / pl.gsz * pl.gsz : pl.n_units;
  const size_t stage = (size_t)MT_W_BYTES + (size_t)pl.nt * MT_BOX_BYTES;
  const size_t fixed = (size_t)MT_CWARPS * 2 * pl.nt * 128 * 4 + 8 * 4 + 2 * MT_MAX_STAGES * 8 + 64 + 1024;
  int stages = (int)((MT_SMEM_BUDGET - fixed) / stage);
  if (stages > MT_MAX_STAGES) stages = MT_MAX_STAGES;
  pl.stages = stages;
  pl.smem = (size_t)stages * stage + fixed;
  return pl;
}

// out6: nt, n_units, gsz, grid, u_static, stages (host arithmetic, exported for the CPU tests)
extern "C" int myr_gemv_mt_plan(int32_t T, int32_t F, int32_t K, int32_t act, int32_t sms, int32_t have_counter, int32_t* out6) {
  MYR_CHECK_ARG(out6 != nullptr && T >= 1 && T <= 32 && F > 0 && K > 0 && K % 128 == 0 && sms > 0, "gemv_mt_plan: bad arguments");
  const GemvMtPlan pl = gemv_mt_plan(T, F, K, act == MYR_ACT_SWIGLU, sms, have_counter != 0);
  out6[0] = pl.nt; out6[1] = pl.n_units; out6[2] = pl.gsz; out6[3] = pl.grid; out6[4] = pl.u_static; out6[5] = pl.stages;
  return MYR_OK;
}

template <int NT>
static int gemv_mt_launch(const GemvMtPlan& pl, const GemvMtMaps& tm, const GemvMtParams& p, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemv_mt_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BUDGET));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemv_mt_kernel<NT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  MYR_CHECK_CUDA(launch_kernel(gemv_mt_kernel<NT>, dim3((unsigned)pl.grid), dim3(MT_THREADS), pl.smem, stream, pdl, tm, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

// Called by myr_gemm_f16 for 5 <= T <= 32 (see gemv_mt_eligible in gemm.cu). `counter`: one zero-initialised int of the
// caller's workspace (or null). Returns MYR_OK or an error code.
int myr_gemv_mt_dispatch(const myr_gemm_args* a, cudaStream_t stream, int* counter) {
  GemvMtParams p;
  p.F = a->F; p.K = a->K; p.T = a->T;
  p.bias = reinterpret_cast<const __half*>(a->bias);
  p.res = a->res; p.res_dtype = a->res_dtype; p.ldr = a->ldr;
  p.out = a->out; p.out_dtype = a->out_dtype; p.ldo = a->ldo;
  p.swiglu = a->act == MYR_ACT_SWIGLU;
  p.w_static = a->w_static;
  const GemvMtPlan pl = gemv_mt_plan(a->T, a->F, a->K, p.swiglu != 0, sm_count(), counter != nullptr);
  MYR_CHECK_ARG(pl.stages >= 2, "gemm: the multi-token streaming path found no room for its ring");
  p.n_kc = pl.n_kc;
  p.stages = pl.stages;
  p.n_units = pl.n_units;
  p.gsz = pl.gsz;
  p.u_static = pl.u_static;
  p.counter = pl.use_counter ? counter : nullptr;
  // weights / tokens as 3-D tensors (64 k, rows, K / 64 k-blocks), 128-byte swizzled boxes of R rows x 8 k-blocks
  GemvMtMaps tm;
  for (int i = 0; i < 4; ++i) {
    const uint64_t dims[3] = {64, (uint64_t)a->F, (uint64_t)(a->K / 64)};
    const uint64_t strides[2] = {(uint64_t)a->ldw * 2, 128};
    const uint32_t box[3] = {64, (uint32_t)(8 * (i + 1)), MT_SK / 64};
    const int rc = make_tmap_f16(&tm.w[i], a->w, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {64, (uint64_t)a->T, (uint64_t)(a->K / 64)};
    const uint64_t strides[2] = {(uint64_t)a->ldx * 2, 128};
    const uint32_t box[3] = {64, (uint32_t)(8 * pl.nt), MT_SK / 64};
    const int rc = make_tmap_f16(&tm.x, a->x, 3, dims, strides, box);
    if (rc) return rc;
  }
  switch (pl.nt) {
    case 1: return gemv_mt_launch<1>(pl, tm, p, stream, a->pdl != 0);
    case 2: return gemv_mt_launch<2>(pl, tm, p, stream, a->pdl != 0);
    case 3: return gemv_mt_launch<3>(pl, tm, p, stream, a->pdl != 0);
    default: return gemv_mt_launch<4>(pl, tm, p, stream, a->pdl != 0);
  }
}
