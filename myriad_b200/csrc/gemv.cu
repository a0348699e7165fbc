// Small-batch weight streaming for sm_100a:  out[t, f] = epilogue( sum_k x[t, k] * W[f, k] ),  T <= 4 tokens.
//
// Greedy decode at the reference's batch sizes (Myriad.generate: one token per sequence and step, modeling_llama.py:730-760)
// multiplies every LLaMA weight matrix by 1..4 activation rows: 2 flop per weight byte, pure HBM streaming. The tcgen05 GEMM
// (gemm.cu) streams at the HBM rate in its steady state, but every launch pays for a TMEM accumulator drain, a split-tile
// fix-up through global memory (32 weight tiles of 128 rows do not fill 148 SMs without splitting K) and a separate
// RMSNorm launch in front of it: ~45 of 113 us per layer with the HBM idle (profiles/r1_decode_timeline.md).
// This kernel removes those phases instead of hiding them:
//   * one CTA per SM owns a contiguous slice of OUTPUT ROWS (balanced to 8 rows), so a row's whole K range is reduced
//     inside the CTA: no split tiles, no workspace, no atomics, no fix-up;
//   * a producer thread streams weights with TMA through a 3-D view of W (64 k, rows, K/64): one 8 KiB box = 8 rows x 512 k,
//     128-byte swizzled, into a ring of up to 6 stages x 32 KiB (128-192 KB in flight per SM, evict-first in L2, filled
//     before the PDL wait because weights are static). (16-byte cp.async tops out at ~4.2 TB/s, 1-D bulk copies of 2 KiB
//     rows at ~4.4 TB/s on B200; 8 KiB tensor boxes keep the TMA engine's per-instruction cost off the critical path.)
//   * the 4 x K activation block lives in shared memory as fp16; for the projections that follow an RMSNorm
//     (modeling_llama.py:66-74: q/k/v, gate/up, lm_head) the CTA computes the norm itself from the fp32 residual stream,
//     which removes the norm launches from the dependent chain;
//   * 8 consumer warps split each 1024-k stage; mma.sync.m16n8k16 (A = 16 weight rows via ldmatrix, B = the activation rows,
//     fp32 accumulate) does the dot products, so no shuffles; warps are summed in fixed order (deterministic, CUDA-graph
//     replay == eager). tcgen05 would need a TMEM drain per 16 rows and buys nothing at 2 flop/byte;
//   * epilogues: bias, fp16/fp32 residual (in place), fp16/fp32 store, SwiGLU over 64-row interleaved gate/up weights.
// Replaces, for T <= 4: nn.Linear of q/k/v/o/gate/up/down/lm_head (modeling_llama.py:139-140,168-231,629-716) + peft LoRA-A
// rows riding on the qkv weight (myriad.py:171-178) + LlamaRMSNorm in front of them.
#include "common.h"
#include "ptx.cuh"

namespace myr {

constexpr int GV_CWARPS = 8;                      // consumer (MMA) warps
constexpr int GV_CTHREADS = GV_CWARPS * 32;
constexpr int GV_SWARPS = 2;                      // activation-staging warps
constexpr int GV_STHREADS = GV_SWARPS * 32;
constexpr int GV_THREADS = GV_CTHREADS + 32 + GV_STHREADS;  // consumers + one producer warp (TMA issue) + stagers
constexpr int GV_MAX_KC = 32;                     // k stages per row: K <= 32768
constexpr int GV_HK = 512;                        // k per TMA transfer of the fp32 residual rows (fused RMSNorm)
constexpr int GV_HBUF_BYTES = (4 + 1) * GV_HK * 4;  // GV_T rows + the gamma slice: 10 KiB
constexpr int GV_UR = 8;                          // weight rows per unit = rows of one TMA box
constexpr int GV_T = 4;                           // activation rows held in shared memory (the N side holds 8; rows >= T are zero)
constexpr int GV_WK = 128;                        // k elements per consumer warp and stage
constexpr int GV_SK = GV_CWARPS * GV_WK;          // k elements per stage (1024)
constexpr int GV_BOX_K = 512;                     // k elements per TMA box: 8 k-blocks of 64 halfs (one 128-byte swizzle row each)
constexpr int GV_BOX_BYTES = GV_UR * GV_BOX_K * 2;  // 8 KiB
constexpr int GV_STAGE_BYTES = 4 * GV_BOX_BYTES;  // 16 rows x 1024 k: boxes [k half][row half]
constexpr int GV_MAX_STAGES = 6;
constexpr int GV_SMEM_BUDGET = 227 * 1024;        // max dynamic shared memory per CTA on sm_100
constexpr int GV_XPAD = 8;                        // halfs of padding per activation row (bank spread)

struct GemvParams {
  int F, K, T;
  const __half* x; long long ldx;                 // fp16 activations [T, K] ...
  const float* h32; long long ldh;                // ... or the fp32 residual stream to RMS-normalise on the fly
  const float* gamma; float eps;
  const float* in_ss;                             // ... or x was pre-scaled by the producer (post_*) and only the RMSNorm scale is missing
  const float* post_gamma; __half* post16; long long post_ld; float* post_ss;  // producer side of that hand-over
  const __half* bias;
  const void* res; int res_dtype; long long ldr;
  void* out; int out_dtype; long long ldo;
  int swiglu;
  int units;                                      // 8-row units in total (SwiGLU: units of 8 gate/up pairs = 8 + 8 rows)
  int n_kc, kp;                                   // k stages per row group = ceil(K / 1024), K padded to 128
  int max_groups;                                 // row groups of the largest CTA slice (sizes the partial-sum buffer)
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

// A row group is 16 weight rows = two 8-row halves, each one TMA box per 512 k:
//   plain  : group g of the slice = units ub + 2g (rows 0..7) and ub + 2g + 1 (rows 8..15, absent in an odd tail)
//   SwiGLU : group g = unit ub + g: rows 0..7 = gate rows of pairs 8(ub+g) .. +7, rows 8..15 = the matching up rows
//            (weights interleaved in blocks of 64: [gate 0..63 | up 0..63 | gate 64..127 | ...])
__device__ __forceinline__ int gv_half_row0(const GemvParams& p, int ub, int n_units, int g, int half) {
  if (!p.swiglu) return (2 * g + half < n_units) ? (ub + 2 * g + half) * GV_UR : -1;
  const int i0 = (ub + g) * GV_UR;
  return ((i0 >> 6) << 7) + (i0 & 63) + (half << 6);
}

__global__ void __maxnreg__(96) gemv_kernel(const __grid_constant__ CUtensorMap tmW, const GemvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // balanced contiguous slices of 8-row units: sizes differ by at most one
  const int ub = (int)(((long long)blockIdx.x * p.units) / gridDim.x);
  const int n_units = (int)(((long long)(blockIdx.x + 1) * p.units) / gridDim.x) - ub;
  const int n_groups = p.swiglu ? n_units : (n_units + 1) / 2;
  const int n_it = n_groups * p.n_kc;                     // iteration = (row group outer, k stage inner): 16 rows x 1024 k
  const int xld = p.kp + GV_XPAD;

  uint8_t* ring = smem;
  __half* xs = reinterpret_cast<__half*>(ring + (size_t)p.stages * GV_STAGE_BYTES);       // [GV_T][xld]
  float* s_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + (size_t)GV_T * xld * 2);  // [warps][GV_T]
  float4* red = reinterpret_cast<float4*>(s_part + GV_CWARPS * GV_T);                       // [warps][max_groups][16 lanes]
  float* hbuf = reinterpret_cast<float*>(red + (size_t)GV_CWARPS * p.max_groups * 16);      // [GV_T][GV_HK] fp32 (fused RMSNorm only)
  uint64_t* full = reinterpret_cast<uint64_t*>(hbuf + (p.h32 ? (GV_T + 1) * GV_HK : 0));          // [stages] producer -> consumers
  uint64_t* empty = full + GV_MAX_STAGES;                                                     // [stages] consumers -> producer
  uint64_t* xbar = empty + GV_MAX_STAGES;                                                     // [n_kc] activation chunk kc staged
  uint64_t* rbar = xbar + GV_MAX_KC;                                                          // RMSNorm scale ready
  uint64_t* hbar = rbar + 1;                                                                  // staging buffer filled
  float* s_rstd = reinterpret_cast<float*>(hbar + 1);                                         // [GV_T]

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GV_CWARPS);
    }
    for (int kc = 0; kc < p.n_kc; ++kc) mbar_init(&xbar[kc], p.h32 ? GV_STHREADS : 1);
    mbar_init(rbar, 1);
    mbar_init(hbar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
  }
  __syncthreads();
  pdl_launch_dependents();
  if (p.trace && tid == 0) p.trace[blockIdx.x * 6 + 0] = gv_time();  // CTA start

  if (warp == GV_CWARPS) {
    // ------------------------------ producer: one thread, up to 4 boxes of 8 KiB per stage ------------------------------
    // Weights do not depend on the previous kernel, so the whole ring is filled before anybody waits for it.
    if (lane == 0) {
      if (!p.w_static) pdl_wait();
      const uint64_t pol = l2_policy_evict_first();
      int st = 0, g = 0, kc = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_it; ++it) {
        const int r_lo = gv_half_row0(p, ub, n_units, g, 0), r_hi = gv_half_row0(p, ub, n_units, g, 1);
        const int k0 = kc * GV_SK;
        const int n_kh = (k0 + GV_BOX_K < p.K) ? 2 : 1;   // second 512-k half absent at the end of K
        mbar_wait(&empty[st], phase ^ 1);
        mbar_arrive_expect_tx(&full[st], (uint32_t)(n_kh * ((r_hi >= 0) ? 2 : 1) * GV_BOX_BYTES));
        const uint32_t dst = smem_u32(ring) + st * GV_STAGE_BYTES, bar = smem_u32(&full[st]);
        for (int kh = 0; kh < n_kh; ++kh) {
          tma_load_3d_hint(dst + (kh * 2) * GV_BOX_BYTES, &tmW, bar, 0, r_lo, (k0 + kh * GV_BOX_K) >> 6, pol);
          if (r_hi >= 0) tma_load_3d_hint(dst + (kh * 2 + 1) * GV_BOX_BYTES, &tmW, bar, 0, r_hi, (k0 + kh * GV_BOX_K) >> 6, pol);
        }
        if (++st == p.stages) {
          st = 0;
          phase ^= 1;
        }
        if (++g == n_groups) {
          g = 0;
          ++kc;
        }
      }
    }
    return;
  }

  if (warp > GV_CWARPS) {
    // ------------------------------ stagers: activations -> shared fp16 [GV_T][xld], one 1024-k chunk at a time ------------------------------
    // The consumers walk K in the same order (chunk outer, row groups inner), so they start on chunk 0 while the rest is
    // still being staged: the ~2 TB/s at which L2 serves the same activation lines to all 148 CTAs stays off the critical path.
    const int stid = tid - GV_CTHREADS - 32;
    if (p.h32 == nullptr) {
      // rows >= T are never copied: zero them once
      for (int i = stid; i < (GV_T - p.T) * (xld >> 3); i += GV_STHREADS)
        *reinterpret_cast<uint4*>(xs + (size_t)p.T * xld + i * 8) = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("bar.sync 2, %0;" ::"n"(GV_STHREADS) : "memory");
      if (stid == 0) {
        fence_proxy_async_smem();
        pdl_wait();
        for (int kc = 0; kc < p.n_kc; ++kc) {
          const int k0 = kc * GV_SK;
          const uint32_t bytes = (uint32_t)min(GV_SK, p.K - k0) * 2;
          mbar_arrive_expect_tx(&xbar[kc], bytes * p.T);
          for (int t = 0; t < p.T; ++t)
            bulk_g2s(smem_u32(xs + (size_t)t * xld + k0), p.x + (long long)t * p.ldx + k0, bytes, smem_u32(&xbar[kc]));
        }
      }
      if (p.in_ss) {
        // RMSNorm scale from the producer's per-CTA sums of squares (summed in CTA order: deterministic); needed by the
        // epilogue only, so its latency is off the critical path
        if (stid != 0) pdl_wait();
        if (stid < GV_T) {
          const int parts = (int)__ldcg(p.in_ss);
          float tot = 0.f;
          for (int c = 0; c < parts; ++c) tot += __ldcg(p.in_ss + 4 + c * GV_T + stid);
          s_rstd[stid] = rsqrtf(tot / p.K + p.eps);
        }
        asm volatile("bar.sync 2, %0;" ::"n"(GV_STHREADS) : "memory");
        if (stid == 0) mbar_arrive(rbar);
      }
    } else {
      // LlamaRMSNorm (modeling_llama.py:66-74): x = h * rsqrt(mean(h^2) + eps) * gamma. The per-token scale is a scalar, so
      // it is applied in the epilogue (the GEMV is linear in x) and the chunks of rn_f16(h * gamma) can be staged while the
      // sum of squares is still being accumulated.
      // The fp32 rows arrive by TMA bulk copy, 512 k at a time, in an 8 KiB staging buffer (LSU loads issued under the
      // saturated weight stream take ~2.5 us per round trip on B200; the TMA path does not queue behind it).
      float ss[GV_T] = {0.f, 0.f, 0.f, 0.f};
      const float4* hb4 = reinterpret_cast<const float4*>(hbuf);
      const float4* gb4 = hb4 + GV_T * (GV_HK / 4);   // gamma slice rides along in the same transfer group
      const int n_half = (p.K + GV_HK - 1) / GV_HK;
      pdl_wait();
      if (p.trace && stid == 0) p.trace[blockIdx.x * 6 + 1] = gv_time();  // predecessor released
      uint32_t hph = 0;
      for (int hh = 0; hh < n_half; ++hh) {
        const int k0 = hh * GV_HK;
        const int nk = min(GV_HK, p.K - k0);
        if (stid == 0) {
          mbar_arrive_expect_tx(hbar, (uint32_t)((p.T + 1) * nk * 4));
          for (int t = 0; t < p.T; ++t) bulk_g2s(smem_u32(hbuf + t * GV_HK), p.h32 + (long long)t * p.ldh + k0, (uint32_t)nk * 4, smem_u32(hbar));
          bulk_g2s(smem_u32(hbuf + GV_T * GV_HK), p.gamma + k0, (uint32_t)nk * 4, smem_u32(hbar));
        }
        mbar_wait(hbar, hph);
        hph ^= 1;
#pragma unroll
        for (int t = 0; t < GV_T; ++t) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = stid + j * GV_STHREADS;
            if (c * 4 < nk) {
              const float4 v = t < p.T ? hb4[t * (GV_HK / 4) + c] : make_float4(0.f, 0.f, 0.f, 0.f);
              ss[t] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
              const float4 gm = gb4[c];
              const __half2 a = __floats2half2_rn(v.x * gm.x, v.y * gm.y);
              const __half2 b = __floats2half2_rn(v.z * gm.z, v.w * gm.w);
              uint2 u;
              u.x = *reinterpret_cast<const uint32_t*>(&a);
              u.y = *reinterpret_cast<const uint32_t*>(&b);
              *reinterpret_cast<uint2*>(xs + (size_t)t * xld + k0 + c * 4) = u;
            }
          }
        }
        if ((hh & 1) || hh == n_half - 1) {
          mbar_arrive(&xbar[hh >> 1]);  // both halves of the 1024-k chunk are staged (this thread's part)
          if (p.trace && stid == 0 && hh <= 1) p.trace[blockIdx.x * 6 + 3] = gv_time();  // fused RMSNorm: first 1024-k chunk staged
          if (p.trace && stid == 0 && hh == n_half - 1) p.trace[blockIdx.x * 6 + 4] = gv_time();  // fused RMSNorm: all chunks staged
        }
        asm volatile("bar.sync 2, %0;" ::"n"(GV_STHREADS) : "memory");  // staging buffer free for the next copy
      }
#pragma unroll
      for (int t = 0; t < GV_T; ++t) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss[t] += __shfl_xor_sync(0xffffffffu, ss[t], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int t = 0; t < GV_T; ++t) s_part[(warp - GV_CWARPS - 1) * GV_T + t] = ss[t];
      }
      asm volatile("bar.sync 2, %0;" ::"n"(GV_STHREADS) : "memory");
      if (stid == 0) {
#pragma unroll
        for (int t = 0; t < GV_T; ++t) {
          float tot = 0.f;
          for (int w = 0; w < GV_SWARPS; ++w) tot += s_part[w * GV_T + t];
          s_rstd[t] = rsqrtf(tot / p.K + p.eps);
        }
        mbar_arrive(rbar);
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  pdl_wait();
  if (p.trace && tid == 0 && !p.h32) p.trace[blockIdx.x * 6 + 1] = gv_time();  // predecessor released

  // ---- main loop: per iteration this warp multiplies its 16 x 128 slice of the stage with its 128 x 8 slice of x ----
  // Stage layout = 4 TMA boxes [k half][row half], each [8 k-blocks][8 rows][128 bytes] with the 128-byte swizzle (16-byte
  // unit u of row r sits at unit u ^ r): ldmatrix.x4 fetches (rows 0-7, k 0-7) (rows 8-15, k 0-7) (rows 0-7, k 8-15)
  // (rows 8-15, k 8-15) of a 16 x 16 A fragment without bank conflicts.
  // B fragments (activations): b0 = x[n][k .. k+1], b1 = x[n][k+8 .. k+9] with n = lane / 4, k = 2 * (lane % 4); n >= 4 -> 0
  const int a_r = lane & 7, a_rh = (lane >> 3) & 1, a_hi = lane >> 4;
  // this warp's 128 k = k-blocks 2 * (warp % 4), + 1 of k half warp / 4
  const uint32_t a_lane = smem_u32(ring) + ((warp >> 2) * 2 + a_rh) * GV_BOX_BYTES + (warp & 3) * 2 * 1024 + a_r * 128;
  // D fragment: lane l holds tokens 2 * (l % 4) + {0, 1}: only lanes with l % 4 < 2 carry tokens 0..3
  const bool d_lane = (lane & 3) < 2;
  float4* red_w = red + (size_t)warp * p.max_groups * 16 + (lane >> 2) * 2 + (lane & 1);
  const int b_n = lane >> 2, b_k = (lane & 3) * 2;
  const __half* xb = xs + (size_t)(b_n < GV_T ? b_n : 0) * xld + warp * GV_WK + b_k;
  int g = 0, kc = 0, st = 0;
  uint32_t phase = 0;
  uint32_t bfrag[GV_WK / 16][2];
  bool mine = false;
  for (int it = 0; it < n_it; ++it) {
    if (g == 0) {
      // next 1024-k chunk of the activations: wait until it is staged, then keep this warp's 128 x 8 slice in registers
      mine = kc * GV_SK + warp * GV_WK < p.K;  // K is a multiple of 128: a warp's slice is whole or absent
      mbar_wait(&xbar[kc], 0);
      if (p.trace && tid == 0 && it == 0) p.trace[blockIdx.x * 6 + 2] = gv_time();  // first activation chunk staged
      if (mine) {
#pragma unroll
        for (int ks = 0; ks < GV_WK / 16; ++ks) {
          bfrag[ks][0] = b_n < GV_T ? *reinterpret_cast<const uint32_t*>(xb + kc * GV_SK + ks * 16) : 0u;
          bfrag[ks][1] = b_n < GV_T ? *reinterpret_cast<const uint32_t*>(xb + kc * GV_SK + ks * 16 + 8) : 0u;
        }
      }
    }
    mbar_wait(&full[st], phase);
    if (p.trace && tid == 0 && it == 0 && !p.h32) p.trace[blockIdx.x * 6 + 3] = gv_time();  // first stage landed
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (mine) {
      const uint32_t a_st = a_lane + (uint32_t)st * GV_STAGE_BYTES;
#pragma unroll
      for (int ks = 0; ks < GV_WK / 16; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a_st + (ks >> 2) * 1024 + (((((ks & 3) << 1) + a_hi) ^ a_r) << 4), a);
        mma_16816(d, a, bfrag[ks][0], bfrag[ks][1]);
      }
    }
    __syncwarp();  // every lane's ldmatrix of this stage is done before the stage is handed back
    if (lane == 0) mbar_arrive(&empty[st]);
    // running sum over the k chunks of this (warp, row group), in chunk order
    if (d_lane) {
      float4 acc = make_float4(d[0], d[1], d[2], d[3]);
      if (kc > 0) {
        const float4 o = red_w[g * 16];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      red_w[g * 16] = acc;
    }
    if (++st == p.stages) {
      st = 0;
      phase ^= 1;
    }
    if (++g == n_groups) {
      g = 0;
      ++kc;
    }
  }
  if (p.h32 || p.in_ss) mbar_wait(rbar, 0);
  consumer_sync();
  if (p.trace && tid == 0 && !p.h32) p.trace[blockIdx.x * 6 + 4] = gv_time();  // last stage consumed

  // ---- cross-warp sum in warp order + epilogue ----
  // D fragment: lane l holds (row l/4, token 2*(l%4) + {0,1}) in .x/.y and (row l/4 + 8, same tokens) in .z/.w
  const float* redf = reinterpret_cast<const float*>(red);
  auto total = [&](int gg, int rr, int t) {
    const int l = (rr & 7) * 2 + (t >> 1), j = (t & 1) + 2 * (rr >> 3);
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < GV_CWARPS; ++w) a += redf[(((size_t)w * p.max_groups + gg) * 16 + l) * 4 + j];
    return (p.h32 || p.in_ss) ? a * s_rstd[t] : a;  // fused RMSNorm: the per-token scale factored out of the dot product
  };
  if (!p.swiglu) {
    const int row0 = ub * GV_UR, n_rows = min(n_units * GV_UR, p.F - row0);
    float ssq = 0.f;  // this thread's token is tid % GV_T in every iteration (GV_CTHREADS is a multiple of GV_T)
    for (int i = tid; i < n_rows * GV_T; i += GV_CTHREADS) {
      const int u = i / GV_T, t = i % GV_T;
      if (t >= p.T) continue;
      const int row = row0 + u;
      float a = total(u >> 4, u & 15, t);
      if (p.bias) a += __half2float(p.bias[row]);
      if (p.res) {
        a += (p.res_dtype == MYR_F32) ? reinterpret_cast<const float*>(p.res)[(long long)t * p.ldr + row]
                                      : __half2float(reinterpret_cast<const __half*>(p.res)[(long long)t * p.ldr + row]);
      }
      if (p.out_dtype == MYR_F32)
        reinterpret_cast<float*>(p.out)[(long long)t * p.ldo + row] = a;
      else
        reinterpret_cast<__half*>(p.out)[(long long)t * p.ldo + row] = __float2half_rn(a);
      if (p.post16) {
        // hand-over to the next projection's RMSNorm: its activations without the per-token scale, and this slice's sum of squares
        p.post16[(long long)t * p.post_ld + row] = __float2half_rn(a * p.post_gamma[row]);
        ssq = fmaf(a, a, ssq);
      }
    }
    if (p.post16) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);  // lanes with the same token
      if (lane < GV_T) s_part[warp * GV_T + lane] = ssq;
      consumer_sync();
      if (tid < GV_T) {
        float tot = 0.f;
        for (int w = 0; w < GV_CWARPS; ++w) tot += s_part[w * GV_T + tid];
        p.post_ss[4 + blockIdx.x * GV_T + tid] = tot;
        if (blockIdx.x == 0 && tid == 0) p.post_ss[0] = (float)gridDim.x;
      }
    }
  } else {
    // SwiGLU (modeling_llama.py:139-140) with the rounding points of the unfused path: gate / up rounded to fp16 first
    for (int i = tid; i < n_units * GV_UR * GV_T; i += GV_CTHREADS) {
      const int u = i / GV_T, t = i % GV_T;
      if (t >= p.T) continue;
      const float a = round_f16(total(u >> 3, u & 7, t)), b = round_f16(total(u >> 3, (u & 7) + 8, t));
      reinterpret_cast<__half*>(p.out)[(long long)t * p.ldo + ub * GV_UR + u] = __float2half_rn(silu_f(a) * b);
    }
  }
  if (p.trace && tid == 0) p.trace[blockIdx.x * 6 + 5] = gv_time();  // epilogue done
}

}  // namespace myr

using namespace myr;

// Called by myr_gemm_f16 for T <= 4 (see gemv_eligible in gemm.cu). Returns MYR_OK or an error code.
int myr_gemv_dispatch(const myr_gemm_args* a, cudaStream_t stream) {
  GemvParams p;
  p.F = a->F; p.K = a->K; p.T = a->T;
  p.x = reinterpret_cast<const __half*>(a->x); p.ldx = a->ldx;
  p.h32 = reinterpret_cast<const float*>(a->norm_h32); p.ldh = a->norm_ldh;
  p.gamma = reinterpret_cast<const float*>(a->norm_gamma); p.eps = a->norm_eps;
  p.in_ss = reinterpret_cast<const float*>(a->norm_ss);
  p.post_gamma = reinterpret_cast<const float*>(a->post_gamma);
  p.post16 = reinterpret_cast<__half*>(a->post_out16); p.post_ld = a->post_ld;
  p.post_ss = reinterpret_cast<float*>(a->post_ss);
  MYR_CHECK_ARG(!(p.in_ss && p.h32), "gemm: norm_h32 and norm_ss are alternatives");
  MYR_CHECK_ARG(p.post16 == nullptr || (p.post_gamma && p.post_ss && !(a->act == MYR_ACT_SWIGLU) && a->post_ld > 0),
                "gemm: post_out16 needs post_gamma, post_ss, a row stride and a plain (non-SwiGLU) epilogue");
  p.bias = reinterpret_cast<const __half*>(a->bias);
  p.res = a->res; p.res_dtype = a->res_dtype; p.ldr = a->ldr;
  p.out = a->out; p.out_dtype = a->out_dtype; p.ldo = a->ldo;
  p.swiglu = a->act == MYR_ACT_SWIGLU;
  p.w_static = a->w_static;
  p.trace = next_trace_slot();
  p.n_kc = ceil_div(a->K, GV_SK);
  p.kp = ceil_div(a->K, GV_WK) * GV_WK;
  MYR_CHECK_ARG(p.n_kc <= GV_MAX_KC, "gemm: K=%d exceeds the small-batch path (K <= %d)", a->K, GV_MAX_KC * GV_SK);
  if (p.h32) {
    MYR_CHECK_ARG(p.gamma != nullptr && a->norm_ldh % 4 == 0 && (reinterpret_cast<uintptr_t>(p.h32) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.gamma) & 15) == 0,
                  "gemm: fused RMSNorm needs 16-byte aligned fp32 rows");
  }
  // weights as a 3-D tensor: (64 k, F rows, K / 64 k-blocks); one box = 64 x 8 rows x 8 k-blocks = 8 KiB, 128-byte swizzled
  CUtensorMap tmW;
  {
    const uint64_t dims[3] = {64, (uint64_t)a->F, (uint64_t)(a->K / 64)};
    const uint64_t strides[2] = {(uint64_t)a->ldw * 2, 128};
    const uint32_t box[3] = {64, GV_UR, GV_BOX_K / 64};
    const int rc = make_tmap_f16(&tmW, a->w, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int sms = sm_count();
  p.units = p.swiglu ? a->F / 16 : ceil_div(a->F, GV_UR);
  const size_t x_bytes = (size_t)GV_T * (p.kp + GV_XPAD) * 2;
  // one CTA per SM; more (several waves) only when a slice's partial sums would crowd the weight ring out of shared memory
  for (int grid = p.units < sms ? p.units : sms;; grid *= 2) {
    if (grid > p.units) grid = p.units;
    const int upc = ceil_div(p.units, grid);
    p.max_groups = p.swiglu ? upc : (upc + 1) / 2;
    const size_t red_bytes = (size_t)GV_CWARPS * p.max_groups * 16 * sizeof(float4);
    const size_t fixed = x_bytes + red_bytes + (p.h32 ? GV_HBUF_BYTES : 0) + GV_CWARPS * GV_T * 4 + (2 * GV_MAX_STAGES + GV_MAX_KC + 2) * 8 +
                         GV_T * 4 + 1024;
    // leave ~8 KB of the SM's shared memory to a small co-resident CTA of the next kernel (decode attention pre-loads its K rows
    // while this kernel streams) unless that would cost a ring stage of an already shallow ring
    static long long budget = -1;
    if (budget < 0) {
      const char* e = getenv("MYR_GEMV_SMEM_KB");
      budget = e ? atoll(e) * 1024 : (long long)GV_SMEM_BUDGET - 8192;
    }
    int stages = (int)((budget - (long long)fixed) / GV_STAGE_BYTES);
    if (stages < 4) stages = (int)(((long long)GV_SMEM_BUDGET - (long long)fixed) / GV_STAGE_BYTES);
    if (stages > GV_MAX_STAGES) stages = GV_MAX_STAGES;
    if (stages >= 3 || (stages >= 2 && grid == p.units)) {
      p.stages = stages;
      static bool attr_set = false;
      if (!attr_set) {
        MYR_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
      }
      MYR_CHECK_CUDA(launch_kernel(gemv_kernel, dim3((unsigned)grid), dim3(GV_THREADS), (size_t)stages * GV_STAGE_BYTES + fixed, stream,
                                   a->pdl != 0, tmW, p));
      MYR_CHECK_LAUNCH();
      return MYR_OK;
    }
    if (grid == p.units) break;
  }
  set_error("gemm: K=%d does not fit the small-batch path", a->K);
  return MYR_ERR_UNSUPPORTED;
}
