// Small-batch weight streaming for sm_100a:  out[t, f] = epilogue( sum_k x[t, k] * W[f, k] ),  T <= 4 tokens.
//
// Greedy decode at the reference's batch sizes (Myriad.generate: one token per sequence and step, modeling_llama.py:730-760)
// multiplies every LLaMA weight matrix by 1..4 activation rows: 2 flop per weight byte, pure HBM streaming. The tcgen05 GEMM
// (gemm.cu) streams at the HBM rate in its steady state, but every launch pays for a TMEM accumulator drain, a split-tile
// fix-up through global memory (32 weight tiles of 128 rows do not fill 148 SMs without splitting K) and a separate
// RMSNorm launch in front of it: ~45 of 113 us per layer with the HBM idle (profiles/r1_decode_timeline.md).
// This kernel removes those phases instead of hiding them:
//   * the unit of work is an ITEM of 8 output rows (SwiGLU: 8 gate/up pairs) x the whole K range, reduced entirely inside one
//     CTA: no split tiles, no partial tiles in global memory, no fix-up. One CTA per SM takes ~3/4 of its share of the items
//     by index and the rest from an atomic counter, because SMs do not stream equally fast (the first and the last CTA of a
//     statically balanced launch finish 2-7 us apart on B200);
//   * a producer thread streams weights with TMA through a 3-D view of W (64 k, rows, K/64): one 8 KiB box = 8 rows x 512 k,
//     128-byte swizzled, into a ring of 16 / 32 KiB stages (128-192 KB in flight per SM, evict-first in L2; the
//     items assigned by index are requested before the PDL wait because weights are static). (16-byte cp.async tops
//     out at ~4.2 TB/s, 1-D bulk copies of 2 KiB rows at ~4.4 TB/s on B200; 8 KiB tensor boxes reach 6.9 TB/s.)
//   * the 4 x K activation block comes in by TMA bulk copy, one 1024-k chunk per mbarrier, so the first group starts on
//     chunk 0 while the rest is in flight (global loads issued under a saturated TMA stream take 2-3 us per round trip;
//     TMA transfers do not queue behind it);
//   * 8 consumer warps split each 1024-k stage; mma.sync.m16n8k16 (A = 16 weight rows via ldmatrix, B = the activation rows,
//     fp32 accumulate in registers over the k stages of a group) does the dot products; the 8 warps' partial sums meet in
//     shared memory once per group and are added in warp order (deterministic, CUDA-graph replay == eager). tcgen05 would
//     need a TMEM drain per 16 rows and buys nothing at 2 flop/byte;
//   * epilogue per group: bias, fp16/fp32 residual (in place), fp16/fp32 store, SwiGLU over 64-row interleaved gate/up
//     weights; LlamaRMSNorm (modeling_llama.py:66-74) in front of q/k/v, gate/up and lm_head is split between two launches:
//     the projection that produces the fp32 stream (o_proj / down_proj) also writes rn_f16(h * gamma_next) and each item's sum
//     of squares (post_*), the consumer multiplies its dot products by rsqrt(sum / K + eps) (norm_ss) - the scale is a
//     per-token scalar and the projection is linear in x, so no norm launch sits in the dependent chain.
// Replaces, for T <= 4: nn.Linear of q/k/v/o/gate/up/down/lm_head (modeling_llama.py:139-140,168-231,629-716) + peft LoRA-A
// rows riding on the qkv weight (myriad.py:171-178) + LlamaRMSNorm in front of them.
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int GV_CWARPS = 8;                      // consumer (MMA) warps
constexpr int GV_CTHREADS = GV_CWARPS * 32;
constexpr int GV_THREADS = GV_CTHREADS + 64;      // + producer warp (weight TMA) + stager warp (activation TMA, norm scale)
constexpr int GV_MAX_KC = 32;                     // k stages per row group: K <= 32768
constexpr int GV_T = 4;                           // activation rows held in shared memory (the N side holds 8; rows >= T are zero)
constexpr int GV_WK = 128;                        // k elements per consumer warp and stage
constexpr int GV_SK = GV_CWARPS * GV_WK;          // k elements per stage (1024)
constexpr int GV_BOX_K = 512;                     // k elements per TMA box: 8 k-blocks of 64 halfs (one 128-byte swizzle row each)
constexpr int GV_BOX_BYTES = 8 * GV_BOX_K * 2;    // 8 rows: 8 KiB
constexpr int GV_STAGE_BYTES = 4 * GV_BOX_BYTES;  // 16 rows x 1024 k: boxes [k half][row half]
constexpr int GV_MAX_STAGES = 6;
constexpr int GV_SMEM_BUDGET = 227 * 1024;        // max dynamic shared memory per CTA on sm_100
constexpr int GV_XPAD = 8;                        // halfs of padding per activation row (bank spread)
constexpr int GV_RED_FLOATS = GV_CWARPS * 16 * 4; // one work item's D fragments: [warp][16 lanes][4]

struct GemvParams {
  int F, K, T;
  const __half* x; long long ldx;                 // fp16 activations [T, K]
  const float* in_ss; float eps;                  // optional: x lacks the RMSNorm scale; per-unit sums of squares written by its producer
  const float* post_gamma; __half* post16; long long post_ld; float* post_ss;  // producer side of that hand-over
  const __half* bias;
  const void* res; int res_dtype; long long ldr;
  void* out; int out_dtype; long long ldo;
  int swiglu;
  int n_units;                                    // units in total: 8 output rows (SwiGLU: 8 gate/up pairs = 8 + 8 weight rows)
  int gsz;                                        // units per MMA row group: 2 (plain: 16 rows) or 1 (SwiGLU)
  int u_static;                                   // units [0, u_static) are split evenly by CTA index, the rest go through `counter`
  int* counter;                                   // zero on entry, left at zero (null: u_static == n_units)
  int n_kc, kp;                                   // k stages per row group = ceil(K / 1024), K padded to 128
  int stages;
  int w_static;
  long long* trace;                               // debug: 6 x %globaltimer ns per CTA (myr_gemm_set_trace)
};
__device__ __forceinline__ long long gv_time() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// 1-D bulk copy global -> shared (TMA engine, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16 x 8, fp32) += A(16 x 16 fp16, row) * B(16 x 8 fp16, col): A = 16 weight rows, B = 8 activation rows
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GV_CTHREADS) : "memory"); }

// first weight row of the low (rows 0..7) / high (rows 8..15) 8-row box of the row group that starts at unit u0:
//   plain  : units u0 (low) and u0 + 1 (high; absent in a one-unit group): rows 8 u0 .. 8 u0 + 15
//   SwiGLU : low = gate rows of pairs 8 u0 .. 8 u0 + 7, high = the matching up rows
//            (weights interleaved in blocks of 64: [gate 0..63 | up 0..63 | gate 64..127 | ...])
__device__ __forceinline__ int gv_half_row0(const GemvParams& p, int u0, int half) {
  if (!p.swiglu) return (u0 + half) * 8;
  const int i0 = u0 * 8;
  return ((i0 >> 6) << 7) + (i0 & 63) + (half << 6);
}

__global__ void __maxnreg__(96) gemv_kernel(const __grid_constant__ CUtensorMap tmW, const GemvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xld = p.kp + GV_XPAD;

  uint8_t* ring = smem;
  __half* xs = reinterpret_cast<__half*>(ring + (size_t)p.stages * GV_STAGE_BYTES);       // [GV_T][xld]
  float* red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + (size_t)GV_T * xld * 2);  // [2][warp][16 lanes][4]
  float* s_part = red + 2 * GV_RED_FLOATS;                                                  // [2][GV_T] (spare)
  float* s_rstd = s_part + 2 * GV_T;                                                        // [GV_T]
  int* s_gid = reinterpret_cast<int*>(s_rstd + GV_T);                                       // [stages] 2 * first unit + (two units) of the stage's row group (-1: end)
  uint64_t* full = reinterpret_cast<uint64_t*>(s_gid + 16);                                 // [stages] producer -> consumers
  uint64_t* empty = full + GV_MAX_STAGES;                                                   // [stages] consumers -> producer
  uint64_t* xbar = empty + GV_MAX_STAGES;                                                   // [n_kc] activation chunk kc has landed
  uint64_t* rbar = xbar + GV_MAX_KC;                                                        // RMSNorm scale ready

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GV_CWARPS);
    }
    for (int kc = 0; kc < p.n_kc; ++kc) mbar_init(&xbar[kc], 1);
    mbar_init(rbar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
  }
  __syncthreads();
  pdl_launch_dependents();
  if (p.trace && tid == 0) p.trace[blockIdx.x * 6 + 0] = gv_time();  // CTA start

  if (warp == GV_CWARPS) {
    // ------------------------------ producer: one thread, 4 boxes of 8 KiB per stage ------------------------------
    if (lane == 0) {
      if (!p.w_static) pdl_wait();
      const uint64_t pol = l2_policy_evict_first();
      // work list of this CTA: its slice of the evenly split units [0, u_static), walked in row groups of gsz units (the last one
      // may be a single unit), then row groups from the shared pool [u_static, n_units)
      const int s1 = (int)(((long long)(blockIdx.x + 1) * p.u_static) / gridDim.x);
      int su = (int)(((long long)blockIdx.x * p.u_static) / gridDim.x);
      const int pool = (p.n_units - p.u_static + p.gsz - 1) / p.gsz;
      bool waited = p.w_static == 0;
      auto grab = [&]() {  // -> first unit of a pool group, or n_units
        if (p.counter == nullptr) return p.n_units;
        if (!waited) {  // the counter is shared with the previous launch, which leaves it at zero when it completes
          pdl_wait();
          waited = true;
        }
        const int v = atomicAdd(p.counter, 1);
        if (v == pool + (int)gridDim.x - 1) *reinterpret_cast<volatile int*>(p.counter) = 0;  // the last grab of this launch
        return v < pool ? p.u_static + v * p.gsz : p.n_units;
      };
      int st = 0;
      uint32_t phase = 0;
      int u0 = su < s1 ? su : grab();
      while (u0 < p.n_units) {
        const int lim = u0 < p.u_static ? s1 : p.n_units;
        const int nu = min(p.gsz, lim - u0);
        // The group after this one is fetched while this one is being requested, so the atomic's latency hides - except
        // for the first fetch, which has to wait for the previous kernel: then this group's stages go out first.
        int next = -1;
        if (u0 < p.u_static && u0 + nu < s1) next = u0 + nu;
        else if (waited) next = grab();
        const bool hi = p.swiglu || nu == 2;
        const int r_lo = gv_half_row0(p, u0, 0), r_hi = gv_half_row0(p, u0, 1);
        for (int kc = 0; kc < p.n_kc; ++kc) {
          const int k0 = kc * GV_SK;
          const int n_kh = (k0 + GV_BOX_K < p.K) ? 2 : 1;   // second 512-k half absent at the end of K
          mbar_wait(&empty[st], phase ^ 1);
          s_gid[st] = u0 * 2 + (hi ? 1 : 0);
          mbar_arrive_expect_tx(&full[st], (uint32_t)(n_kh * (hi ? 2 : 1) * GV_BOX_BYTES));
          const uint32_t dst = smem_u32(ring) + st * GV_STAGE_BYTES, bar = smem_u32(&full[st]);
          for (int kh = 0; kh < n_kh; ++kh) {
            tma_load_3d_hint(dst + (kh * 2) * GV_BOX_BYTES, &tmW, bar, 0, r_lo, (k0 + kh * GV_BOX_K) >> 6, pol);
            if (hi) tma_load_3d_hint(dst + (kh * 2 + 1) * GV_BOX_BYTES, &tmW, bar, 0, r_hi, (k0 + kh * GV_BOX_K) >> 6, pol);
          }
          if (++st == p.stages) {
            st = 0;
            phase ^= 1;
          }
        }
        if (next < 0) next = grab();
        u0 = next;
      }
      // end marker
      mbar_wait(&empty[st], phase ^ 1);
      s_gid[st] = -1;
      mbar_arrive(&full[st]);
    }
    return;
  }

  if (warp == GV_CWARPS + 1) {
    // ------------------ stager: activations -> shared fp16 [GV_T][xld] by TMA, one 1024-k chunk per barrier ------------------
    // rows >= T are never copied: zero them once
    for (int i = lane; i < (GV_T - p.T) * (xld >> 3); i += 32)
      *reinterpret_cast<uint4*>(xs + (size_t)p.T * xld + i * 8) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    pdl_wait();
    if (lane == 0) {
      fence_proxy_async_smem();
      for (int kc = 0; kc < p.n_kc; ++kc) {
        const int k0 = kc * GV_SK;
        const uint32_t bytes = (uint32_t)min(GV_SK, p.K - k0) * 2;
        mbar_arrive_expect_tx(&xbar[kc], bytes * p.T);
        for (int t = 0; t < p.T; ++t)
          bulk_g2s(smem_u32(xs + (size_t)t * xld + k0), p.x + (long long)t * p.ldx + k0, bytes, smem_u32(&xbar[kc]));
      }
    }
    if (p.in_ss) {
      // RMSNorm scale from the producer's per-item sums of squares; needed by the first item epilogue only
      // lane l sums the items l / 4, l / 4 + 8, ... of token l % 4, then a fixed shuffle tree: deterministic
      const int parts = (int)__ldcg(p.in_ss);
      float tot = 0.f;
      for (int c = lane >> 2; c < parts; c += 8) tot += __ldcg(p.in_ss + 4 + c * GV_T + (lane & 3));
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane < GV_T) s_rstd[lane] = rsqrtf(tot / p.K + p.eps);
      __syncwarp();
      if (lane == 0) mbar_arrive(rbar);
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  pdl_wait();
  if (p.trace && tid == 0) p.trace[blockIdx.x * 6 + 1] = gv_time();  // predecessor released

  // Stage layout = 4 TMA boxes [k half][row half], each [8 k-blocks][8 rows][128 bytes] with the 128-byte swizzle (16-byte
  // unit u of row r sits at unit u ^ r): ldmatrix.x4 fetches (rows 0-7, k 0-7) (rows 8-15, k 0-7) (rows 0-7, k 8-15)
  // (rows 8-15, k 8-15) of a 16 x 16 A fragment without bank conflicts. A one-unit group leaves the high boxes stale: the
  // fragment's upper half then holds garbage and its results are dropped.
  // B fragments (activations): b0 = x[n][k .. k+1], b1 = x[n][k+8 .. k+9] with n = lane / 4, k = 2 * (lane % 4); n >= 4 -> 0
  const int a_r = lane & 7, a_rh = (lane >> 3) & 1, a_hi = lane >> 4;
  // this warp's 128 k = k-blocks 2 * (warp % 4), + 1 of k half warp / 4
  const uint32_t a_lane = smem_u32(ring) + ((warp >> 2) * 2 + a_rh) * GV_BOX_BYTES + (warp & 3) * 2 * 1024 + a_r * 128;
  // D fragment: lane l holds (row l/4, tokens 2 * (l % 4) + {0, 1}) in d[0..1] and (row l/4 + 8, same tokens) in d[2..3]:
  // only lanes with l % 4 < 2 carry tokens 0..3
  const bool d_lane = (lane & 3) < 2;
  const int d_slot = (warp * 16 + (lane >> 2) * 2 + (lane & 1)) * 4;
  const int b_n = lane >> 2, b_k = (lane & 3) * 2;
  const __half* xb = xs + (size_t)(b_n < GV_T ? b_n : 0) * xld + warp * GV_WK + b_k;
  // epilogue role: one (row, token) of a row group per thread for the first 64 (plain) / 32 (SwiGLU: pair, token) threads
  const int e_t = tid & 3, e_r = tid >> 2;
  const int e_l = (e_r & 7) * 2 + (e_t >> 1), e_j = (e_t & 1) + 2 * ((e_r >> 3) & 1);  // lane slot / component of (e_r, e_t)
  int kc = 0, st = 0, n_done = 0;
  uint32_t phase = 0;
  float pre_res = 0.f, pre_gam = 0.f, pre_bias = 0.f;
  bool have_scale = p.in_ss == nullptr;
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  for (;;) {
    mbar_wait(&full[st], phase);
    const int g = s_gid[st];
    if (g < 0) break;
    if (p.trace && tid == 0 && n_done == 0 && kc == 0) p.trace[blockIdx.x * 6 + 3] = gv_time();  // first weight stage landed
    if (kc == 0 && !p.swiglu && tid < 16 * GV_T) {
      // epilogue operands of this thread's (row, token): requested now, a whole row group of streaming ahead of their use
      // (a global load issued under the saturated weight stream takes 2-3 us)
      const int row = (g >> 1) * 8 + e_r;
      const bool ok = e_r < ((g & 1) ? 16 : 8) && e_t < p.T && row < p.F;
      pre_res = 0.f;
      pre_gam = 0.f;
      pre_bias = 0.f;
      if (ok && p.res) {
        pre_res = (p.res_dtype == MYR_F32) ? __ldcg(reinterpret_cast<const float*>(p.res) + (long long)e_t * p.ldr + row)
                                           : __half2float(__ldcg(reinterpret_cast<const __half*>(p.res) + (long long)e_t * p.ldr + row));
      }
      if (ok && p.post16) pre_gam = __ldg(p.post_gamma + row);
      if (ok && p.bias) pre_bias = __half2float(__ldg(p.bias + row));
    }
    if (n_done == 0) {
      mbar_wait(&xbar[kc], 0);  // the first group walks K while the later activation chunks are still landing
      if (p.trace && tid == 0 && kc == 0) p.trace[blockIdx.x * 6 + 2] = gv_time();  // first activation chunk landed
    }
    if (kc * GV_SK + warp * GV_WK < p.K) {  // K is a multiple of 128: a warp's slice is whole or absent
      const uint32_t a_st = a_lane + (uint32_t)st * GV_STAGE_BYTES;
      const __half* xk = xb + kc * GV_SK;
#pragma unroll
      for (int ks = 0; ks < GV_WK / 16; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a_st + (ks >> 2) * 1024 + (((((ks & 3) << 1) + a_hi) ^ a_r) << 4), a);
        const uint32_t b0 = b_n < GV_T ? *reinterpret_cast<const uint32_t*>(xk + ks * 16) : 0u;
        const uint32_t b1 = b_n < GV_T ? *reinterpret_cast<const uint32_t*>(xk + ks * 16 + 8) : 0u;
        mma_16816(d, a, b0, b1);
      }
    }
    __syncwarp();  // every lane's ldmatrix of this stage is done before the stage is handed back
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == p.stages) {
      st = 0;
      phase ^= 1;
    }
    if (++kc < p.n_kc) continue;

    // ---- group complete: the 8 warps' partial sums meet in shared memory (double-buffered: one barrier per group) ----
    kc = 0;
    float* rb = red + (n_done & 1) * GV_RED_FLOATS;
    ++n_done;
    if (d_lane) *reinterpret_cast<float4*>(rb + d_slot) = make_float4(d[0], d[1], d[2], d[3]);
    d[0] = d[1] = d[2] = d[3] = 0.f;
    if (!have_scale) {
      mbar_wait(rbar, 0);
      have_scale = true;
    }
    consumer_sync();
    auto total = [&](int l, int j) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < GV_CWARPS; ++w) a += rb[(w * 16 + l) * 4 + j];
      return a;
    };
    const int u0 = g >> 1;
    if (!p.swiglu) {
      const int row = u0 * 8 + e_r;
      float sq = 0.f;
      if (tid < ((g & 1) ? 16 : 8) * GV_T && e_t < p.T && row < p.F) {
        float a = total(e_l, e_j);
        if (p.in_ss) a *= s_rstd[e_t];  // RMSNorm: the per-token scale factored out of the dot product
        a += pre_bias;
        a += pre_res;
        if (p.out_dtype == MYR_F32)
          reinterpret_cast<float*>(p.out)[(long long)e_t * p.ldo + row] = a;
        else
          reinterpret_cast<__half*>(p.out)[(long long)e_t * p.ldo + row] = __float2half_rn(a);
        if (p.post16) {
          // hand-over to the next projection's RMSNorm: its activations without the per-token scale + this CTA's sum of squares
          p.post16[(long long)e_t * p.post_ld + row] = __float2half_rn(a * pre_gam);
          sq = a * a;
        }
      }
      if (p.post16 && warp < 1 + (g & 1)) {
        // sum of squares per token of each 8-row unit (warp 0: rows 0-7, warp 1: rows 8-15; lanes with the same token, fixed
        // shuffle tree -> deterministic and independent of which CTA ran the unit)
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane < GV_T) p.post_ss[4 + (u0 + warp) * GV_T + lane] = sq;
        if (u0 == 0 && tid == 0) p.post_ss[0] = (float)p.n_units;
      }
    } else if (tid < 8 * GV_T && e_t < p.T) {
      // SwiGLU (modeling_llama.py:139-140) with the rounding points of the unfused path: gate / up rounded to fp16 first;
      // pair e_r of the group: gate row = slot e_r, up row = slot e_r + 8 (component + 2 of the same lane slot)
      float a = total(e_l, e_j), b = total(e_l, e_j + 2);
      if (p.in_ss) {
        a *= s_rstd[e_t];
        b *= s_rstd[e_t];
      }
      a = round_f16(a);
      b = round_f16(b);
      reinterpret_cast<__half*>(p.out)[(long long)e_t * p.ldo + u0 * 8 + e_r] = __float2half_rn(silu_f(a) * b);
    }
  }
  if (p.trace && tid == 0) p.trace[blockIdx.x * 6 + 4] = gv_time();  // last stage consumed
  if (p.trace && tid == 0) p.trace[blockIdx.x * 6 + 5] = gv_time();  // done
}

}  // namespace myr

using namespace myr;

// Work split and ring depth of one launch (pure host arithmetic; also exported as myr_gemv_plan for the CPU tests).
struct GemvPlan {
  int n_units, gsz, grid, u_static, use_counter, stages, n_kc, kp;
  size_t fixed_smem;
};
static GemvPlan gemv_plan(int F, int K, bool swiglu, int sms, bool have_counter) {
  GemvPlan pl;
  pl.n_kc = ceil_div(K, GV_SK);
  pl.kp = ceil_div(K, GV_WK) * GV_WK;
  pl.n_units = swiglu ? F / 16 : ceil_div(F, 8);
  pl.gsz = swiglu ? 1 : 2;
  const int n_groups = ceil_div(pl.n_units, pl.gsz);
  pl.grid = n_groups < sms ? n_groups : sms;
  static int dyn = -1;
  if (dyn < 0) {
    const char* e = getenv("MYR_GEMV_DYNAMIC");
    dyn = (e && e[0] == '0') ? 0 : 1;
  }
  // ~3/4 of the units by index (requested before the PDL wait), the rest from the counter - when a CTA has enough row groups
  // for that to be finer than the even split (o_proj / down_proj have < 2 groups per CTA: evenly split units only)
  pl.use_counter = (dyn && have_counter && n_groups >= 4 * pl.grid) ? 1 : 0;
  pl.u_static = pl.use_counter ? (int)((long long)pl.n_units * 3 / 4) / pl.gsz * pl.gsz : pl.n_units;
  pl.fixed_smem = (size_t)GV_T * (pl.kp + GV_XPAD) * 2 + (2 * GV_RED_FLOATS + 3 * GV_T + 16) * 4 +
                  (2 * GV_MAX_STAGES + GV_MAX_KC + 1) * 8 + 1024;
  // leave ~8 KB of the SM's shared memory to a small co-resident CTA of the next kernel unless that would cost a ring stage
  // of an already shallow ring
  int stages = (int)(((long long)GV_SMEM_BUDGET - 8192 - (long long)pl.fixed_smem) / GV_STAGE_BYTES);
  if (stages < 4) stages = (int)(((long long)GV_SMEM_BUDGET - (long long)pl.fixed_smem) / GV_STAGE_BYTES);
  if (stages > GV_MAX_STAGES) stages = GV_MAX_STAGES;
  pl.stages = stages;
  return pl;
}

extern "C" int myr_gemv_plan(int32_t F, int32_t K, int32_t act, int32_t sms, int32_t have_counter, int32_t* out8) {
  MYR_CHECK_ARG(out8 != nullptr && F > 0 && K > 0 && K % 128 == 0 && sms > 0, "gemv_plan: bad arguments");
  const GemvPlan pl = gemv_plan(F, K, act == MYR_ACT_SWIGLU, sms, have_counter != 0);
  out8[0] = pl.n_units; out8[1] = pl.gsz; out8[2] = pl.grid; out8[3] = pl.u_static;
  out8[4] = pl.use_counter; out8[5] = pl.stages; out8[6] = pl.n_kc; out8[7] = (int32_t)pl.fixed_smem;
  return MYR_OK;
}

// Called by myr_gemm_f16 for T <= 4 (see gemv_eligible in gemm.cu). `counter`: one zero-initialised int of the caller's
// workspace (or null). Returns MYR_OK or an error code.
int myr_gemv_dispatch(const myr_gemm_args* a, cudaStream_t stream, int* counter) {
  GemvParams p;
  p.F = a->F; p.K = a->K; p.T = a->T;
  p.x = reinterpret_cast<const __half*>(a->x); p.ldx = a->ldx;
  p.in_ss = reinterpret_cast<const float*>(a->norm_ss); p.eps = a->norm_eps;
  p.post_gamma = reinterpret_cast<const float*>(a->post_gamma);
  p.post16 = reinterpret_cast<__half*>(a->post_out16); p.post_ld = a->post_ld;
  p.post_ss = reinterpret_cast<float*>(a->post_ss);
  MYR_CHECK_ARG(p.post16 == nullptr || (p.post_gamma && p.post_ss && !(a->act == MYR_ACT_SWIGLU) && a->post_ld > 0),
                "gemm: post_out16 needs post_gamma, post_ss, a row stride and a plain (non-SwiGLU) epilogue");
  p.bias = reinterpret_cast<const __half*>(a->bias);
  p.res = a->res; p.res_dtype = a->res_dtype; p.ldr = a->ldr;
  p.out = a->out; p.out_dtype = a->out_dtype; p.ldo = a->ldo;
  p.swiglu = a->act == MYR_ACT_SWIGLU;
  p.w_static = a->w_static;
  p.trace = next_trace_slot();
  const GemvPlan pl = gemv_plan(a->F, a->K, p.swiglu != 0, sm_count(), counter != nullptr);
  p.n_kc = pl.n_kc;
  p.kp = pl.kp;
  MYR_CHECK_ARG(p.n_kc <= GV_MAX_KC, "gemm: K=%d exceeds the small-batch path (K <= %d)", a->K, GV_MAX_KC * GV_SK);
  // weights as a 3-D tensor: (64 k, F rows, K / 64 k-blocks); one box = 64 x 8 rows x 8 k-blocks = 8 KiB, 128-byte swizzled
  CUtensorMap tmW;
  {
    const uint64_t dims[3] = {64, (uint64_t)a->F, (uint64_t)(a->K / 64)};
    const uint64_t strides[2] = {(uint64_t)a->ldw * 2, 128};
    const uint32_t box[3] = {64, 8, GV_BOX_K / 64};
    const int rc = make_tmap_f16(&tmW, a->w, 3, dims, strides, box);
    if (rc) return rc;
  }
  p.n_units = pl.n_units;
  p.gsz = pl.gsz;
  p.counter = pl.use_counter ? counter : nullptr;
  p.u_static = pl.u_static;
  const int grid = pl.grid;
  const size_t fixed = pl.fixed_smem;
  const int stages = pl.stages;
  if (stages < 2) {
    set_error("gemm: K=%d does not fit the small-batch path", a->K);
    return MYR_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  static bool attr_set = false;
  if (!attr_set) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // always the largest shared-memory carve-out, also for launches with a shallow ring: an SM hosts CTAs of two kernels at
    // once only under one carve-out, and switching it waits for the SM to drain
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  MYR_CHECK_CUDA(launch_kernel(gemv_kernel, dim3((unsigned)grid), dim3(GV_THREADS), (size_t)stages * GV_STAGE_BYTES + fixed, stream,
                               a->pdl != 0, tmW, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
