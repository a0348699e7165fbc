// CTA-pair tcgen05 GEMM for the tensor-bound shapes (T > 64: ViT blocks, Q-Former, LLaMA prefill / training forward):
//   out[t, f] = epilogue( sum_k X[t, k] * W[f, k] )        X, W fp16 K-major, fp32 accumulation in TMEM.
//
// Why pairs: one CTA issuing 128 x 256 UMMAs reads 48 KB of operands from its shared memory per 64-wide k-block while TMA
// writes the next 48 KB into it — 192 B/clk against the 128 B/clk an SM's shared memory moves, so gemm.cu's kernel tops
// out near 60 % of the tensor peak. With tcgen05.mma.cta_group::2 two SMs compute one 256 x BN tile: each holds its 128
// rows of A and only HALF of the N side (32 KB per k-block at BN = 256), the halves are exchanged inside the MMA.
// Clusters of 2P CTAs (P pairs) additionally share the A tile: every CTA fetches 1/P of its 128 A rows and TMA-multicasts
// the slice to the same-rank CTAs of the other pairs, which cuts the L2 -> SM request traffic that binds next.
//
// Two operand arrangements, chosen per shape by the host plan below:
//   row_mode 1 "lanes = tokens":   A = X (256 tokens per pair), N side = W (BN features). Epilogue thread = one token row,
//                                  16 consecutive features per TMEM load, 16-byte vector loads / stores.
//   row_mode 0 "lanes = features": A = W (256 features per pair), N side = X (BN tokens). Used when T is not a good
//                                  multiple of 256 (T = 524 -> 3 x 176, T = 1028 -> 6 x 176, T = 656 -> 3 x 224): BN is any
//                                  multiple of 16, so the token padding drops from 20-30 % to 1-3 %.
// Roles per CTA (320 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA of the pair only), warps 2-9 epilogue
// (TMEM -> registers -> fused epilogue -> global), double-buffered accumulators (2 x 256 TMEM columns).
// Persistent: cluster c takes work units c, c + n_clusters, ...; a unit = one A tile x P adjacent N tiles, ordered so that
// clusters running side by side read the same weight tile (the weights stream from HBM once, L2 serves the other readers).
#include "common.h"
#include "gemm_epi.cuh"
#include "ptx.cuh"

namespace myr {

constexpr int G2_BK = 64;                         // halfs per k-block = one 128-byte swizzle row
constexpr int G2_A_ROWS = 128;                    // rows of A per CTA (UMMA M = 256 over the pair)
constexpr int G2_A_BYTES = G2_A_ROWS * G2_BK * 2;  // 16 KiB
constexpr int G2_MAX_STAGES = 8;
constexpr int G2_EPI_WARPS = 16;  // 4 per TMEM lane quarter: the epilogue is latency-bound (global loads / stores, erf), more warps hide it
constexpr int G2_EPI_THREADS = 32 * G2_EPI_WARPS;
constexpr int G2_THREADS = 64 + G2_EPI_THREADS;
constexpr int G2_SMEM_TILE_BUDGET = 192 * 1024;
constexpr int G2_SWIGLU_STAGE_BYTES = 4 * 2 * 16 * 64 * 4;  // lanes = features SwiGLU: [csel][buffer][16 tokens][64 up rows] fp32

struct Gemm2Params {
  int T, F, K;
  int row_mode;
  int BN, P;
  int n_mt, n_nt, n_ng, kb_total, n_units;
  int num_stages, stage_bytes;
  int S, kb_per;       // split-K: S consecutive units share a tile, each covers kb_per k-blocks; partial sums are added in place
  int* flags;          // [tile][rank in pair]: splits of a tile that have written their contribution (ordered => deterministic)
  int dbg;             // timing ablations (MYR_G2_DBG; results are wrong): 1 no MMAs, 2 no N-side TMA, 4 no A TMA, 8 MMAs with N = 16
  int pf_dist;         // L2 prefetch distance of the weight operand in k-blocks (0 = off)
  int ka;              // 64-wide k-atoms per ring stage (1: 2-D tensor maps; 2 / 4: 3-D maps (64 k, rows, atoms), K % 64 == 0, P == 1)
  uint32_t idesc;
  long long* trace;  // debug (myr_gemm_set_trace): per CTA 6 x %globaltimer ns
  Epilogue ep;
};

__device__ __forceinline__ void g2_unit(const Gemm2Params& p, int u, int& mt, int& ng, int& ks) {
  const int uu = u / p.S;
  ks = u - uu * p.S;
  u = uu;
  if (p.row_mode) {  // weights are the N side: neighbouring units walk the token tiles of one weight group
    ng = u / p.n_mt;
    mt = u - ng * p.n_mt;
  } else {           // weights are the A side: neighbouring units walk the token groups of one weight tile
    mt = u / p.n_ng;
    ng = u - mt * p.n_ng;
  }
}

__device__ __forceinline__ void g2_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA2,
             const Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + p.num_stages * p.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);  // [G2_MAX_STAGES] TMA (both CTAs) -> MMA, leader's copy is the live one
  uint64_t* empty = full + G2_MAX_STAGES;               // [G2_MAX_STAGES] MMA commits of all P pairs -> this CTA's producer
  uint64_t* tfull = empty + G2_MAX_STAGES;              // [2] MMA -> epilogue of both CTAs of the pair
  uint64_t* tempty = tfull + 2;                         // [2] epilogue warps of both CTAs -> MMA (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* gu_stage = reinterpret_cast<float*>(tail + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int r = crank & 1;    // rank inside the pair: 0 leads
  const int pr = crank >> 1;  // pair index inside the cluster
  const int csize = 2 * p.P;
  const int n_clusters = (int)gridDim.x / csize;
  const int cluster_id = (int)blockIdx.x / csize;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (uint32_t)p.P);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * G2_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anything is multicast at them / committed to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 6 + 0] = gtime_ns();  // set-up done

  const int a_slice_rows = G2_A_ROWS / p.P;
  const uint32_t b_half_bytes = (uint32_t)(p.BN / 2) * G2_BK * 2;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs of every pair) ------------------------------
    if (lane == 0) {
      uint16_t mask_rank = 0;
      for (int q = 0; q < p.P; ++q) mask_rank |= (uint16_t)(1u << (2 * q + r));
      const uint32_t tx_pair = 2u * (uint32_t)p.ka * (((p.dbg & 4) ? 0u : (uint32_t)G2_A_BYTES) + ((p.dbg & 2) ? 0u : b_half_bytes));  // bytes landing in both CTAs of a pair per stage
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();
      if (p.trace) p.trace[blockIdx.x * 6 + 1] = gtime_ns();  // predecessor done
      for (int u = cluster_id; u < p.n_units; u += n_clusters) {
        int mt, ng, ks;
        g2_unit(p, u, mt, ng, ks);
        const int a_row0 = mt * 256 + r * G2_A_ROWS + pr * a_slice_rows;
        const int b_row0 = (ng * p.P + pr) * p.BN + r * (p.BN / 2);
        const int kb0 = ks * p.kb_per, kb1 = min(p.kb_total, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (p.pf_dist > 0) {
            // weights arrive from HBM: ask L2 for them pf_dist k-blocks before the load that needs them, so the ring's
            // bytes in flight cover the L2 latency instead of the DRAM latency (the ring cannot grow: shared memory is full)
            if (kb == kb0) {
              for (int d = 0; d < p.pf_dist && kb0 + d < kb1; ++d) {
                if (p.row_mode) tma_prefetch_2d(&tmB, (kb0 + d) * G2_BK, b_row0);
                else tma_prefetch_2d(&tmA, (kb0 + d) * G2_BK, a_row0);
              }
            }
            if (kb + p.pf_dist < kb1) {
              if (p.row_mode) tma_prefetch_2d(&tmB, (kb + p.pf_dist) * G2_BK, b_row0);
              else tma_prefetch_2d(&tmA, (kb + p.pf_dist) * G2_BK, a_row0);
            }
          }
          mbar_wait(&empty[stage], phase ^ 1);
          if ((p.dbg & 16) && p.trace && blockIdx.x < 2 && u == cluster_id && kb < 64) p.trace[888 + (blockIdx.x ? 128 : 0) + kb] = gtime_ns();
          if (r == 0) mbar_arrive_expect_tx(&full[stage], tx_pair);
          uint8_t* sa = smem + stage * p.stage_bytes;
          if (p.ka > 1) {
            // one copy per operand and stage brings ka k-atoms (a copy costs the issuing thread the same whatever its size);
            // with P pairs per cluster this CTA's 1/P slice of the A rows goes to the same-rank CTA of every pair, one copy per atom
            // (the slices of one atom must lie side by side: [atom][128 rows][128 B])
            if (p.dbg & 4) {
            } else if (p.P > 1) {
              for (int at = 0; at < p.ka; ++at)
                tma_load_2d_pair_multicast(sa + at * G2_A_BYTES + pr * a_slice_rows * (G2_BK * 2), &tmA2, &full[stage],
                                           (kb * p.ka + at) * G2_BK, a_row0, mask_rank);
            } else {
              tma_load_3d_pair(sa, &tmA, &full[stage], 0, a_row0, kb * p.ka);
            }
            if (!(p.dbg & 2)) tma_load_3d_pair(sa + p.ka * G2_A_BYTES, &tmB, &full[stage], 0, b_row0, kb * p.ka);
          } else if (p.dbg & 4) {
          } else if (p.P > 1)
            tma_load_2d_pair_multicast(sa + pr * a_slice_rows * (G2_BK * 2), &tmA, &full[stage], kb * G2_BK, a_row0, mask_rank);
          else
            tma_load_2d_pair(sa, &tmA, &full[stage], kb * G2_BK, a_row0);
          if (p.ka == 1 && !(p.dbg & 2)) tma_load_2d_pair(sa + G2_A_BYTES, &tmB, &full[stage], kb * G2_BK, b_row0);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (lane == 0 && r == 0) {
      const uint16_t mask_all = (uint16_t)((1u << csize) - 1u);
      const uint16_t mask_pair = (uint16_t)(3u << (2 * pr));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int u = cluster_id; u < p.n_units; u += n_clusters) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        const int ks = u % p.S;
        const int kb0 = ks * p.kb_per, kb1 = min(p.kb_total, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if ((p.dbg & 16) && p.trace && blockIdx.x == 0 && u == cluster_id && kb < 64) p.trace[888 + 64 + kb] = gtime_ns();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + p.ka * G2_A_BYTES;
          for (int at = 0; at < p.ka; ++at) {  // k-atom at of the stage: [atom][row][128 B] per operand
            const uint32_t sa_at = sa + at * G2_A_BYTES, sb_at = sb + at * b_half_bytes;
#pragma unroll
            for (int kk = 0; kk < G2_BK / 16; ++kk) {
              const uint64_t da = make_smem_desc(sa_at + kk * 32, 16, 1024);
              const uint64_t db = make_smem_desc(sb_at + kk * 32, 16, 1024);
              if (!(p.dbg & 1)) tc_mma_f16_pair(d_tmem, da, db, p.idesc, (kb > kb0 || at > 0 || kk > 0) ? 1u : 0u);
            }
          }
          tc_commit_pair(&empty[stage], mask_all);  // frees this stage in every CTA that multicasts into the pair
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit_pair(&tfull[as], mask_pair);
        if (p.trace) p.trace[blockIdx.x * 6 + 2] = gtime_ns();  // last MMA of the (last) unit issued
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..17 of both CTAs) ------------------------------
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int csel = (warp - 2) >> 2;  // which quarter of the column chunks this warp handles (chunks csel, csel + 4, ...)
    const int lrow = q * 32 + lane;    // TMEM lane == row of this CTA's half of the A tile
    const bool swiglu = p.ep.act == MYR_ACT_SWIGLU;
    const int nchunks = p.BN / 16;
    int as = 0;
    uint32_t aphase = 0;
    int it = 0;  // SwiGLU hand-over buffer parity: runs on across units so consecutive chunks never share a buffer
    pdl_wait();  // residual / output may still be in use by the previous kernel
    for (int u = cluster_id; u < p.n_units; u += n_clusters) {
      int mt, ng, ks;
      g2_unit(p, u, mt, ng, ks);
      const int m0 = mt * 256 + r * G2_A_ROWS;
      const int n0 = (ng * p.P + pr) * p.BN;
      // fp32 residual epilogue (o_proj / down_proj / ViT proj / fc2): this warp's first two chunks of residual values are
      // requested BEFORE the accumulator is awaited, so their L2 latency hides behind the unit's MMAs
      float pre[2][16];
      bool pre_ok = false;
      if (!p.row_mode && !swiglu && ks == 0 && p.ep.res != nullptr && p.ep.res_dtype == MYR_F32 && p.ep.out_dtype == MYR_F32 &&
          p.ep.act == MYR_ACT_NONE && !p.ep.round_acc && p.ep.scale_cols == 0 && p.ep.group_rows == 0 && m0 + lrow < p.F) {
        pre_ok = true;
        const float* rbase = reinterpret_cast<const float*>(p.ep.res) + (m0 + lrow);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int t0 = n0 + (csel + 4 * h2) * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) pre[h2][j] = (t0 + j < p.T && (csel + 4 * h2) * 16 + j < p.BN) ? __ldcg(rbase + (long long)(t0 + j) * p.ep.ldr) : 0.f;
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (p.trace && threadIdx.x == 64) p.trace[blockIdx.x * 6 + 3] = gtime_ns();  // accumulator of the (last) unit ready
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * 256;
      if (p.row_mode) {
        const long long t = (long long)m0 + lrow;
        const bool t_ok = t < p.T;
        const long long obase = t_ok ? out_row_offset(p.ep, t) : 0;
        if (!swiglu) {
          for (int ch = csel; ch < nchunks; ch += 4) {
            const int f0 = n0 + ch * 16;
            if (f0 >= p.F) break;
            uint32_t rr[16];
            tmem_ld16(taddr + ch * 16, rr);
            tmem_ld_wait();
            if (t_ok) {
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
              epi_row16(p.ep, v, t, f0, p.F, obase);
            }
          }
        } else {
          // weight rows are interleaved in blocks of 64: [gate 0..63 | up 0..63 | gate 64..127 | ...]; BN % 128 == 0
          const int units = p.BN / 32;
          for (int uu = csel; uu < units; uu += 4) {
            const int blk = uu >> 2, cg = uu & 3;
            const int gcol = blk * 128 + cg * 16;
            if (n0 + gcol >= p.F) break;
            uint32_t rg[16], ru[16];
            tmem_ld16(taddr + gcol, rg);
            tmem_ld16(taddr + gcol + 64, ru);
            tmem_ld_wait();
            if (t_ok) {
              float gv[16], uv[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                gv[j] = __uint_as_float(rg[j]);
                uv[j] = __uint_as_float(ru[j]);
              }
              epi_row16_swiglu(p.ep, gv, uv, (n0 >> 1) + blk * 64 + cg * 16, p.F >> 1, obase);
            }
          }
        }
      } else if (!swiglu) {
        // lanes = features: this thread owns feature f; a 16-column TMEM chunk = 16 consecutive tokens. For one token the 32
        // lanes of a warp touch 32 consecutive features: one 64 B (fp16) / 128 B (fp32) segment per instruction.
        const int f = m0 + lrow;
        const bool f_ok = f < p.F;
        const Epilogue& ep = p.ep;
        const float bias_f = (ep.bias && f_ok && ks == 0) ? __half2float(ep.bias[f]) : 0.f;
        const int cols = min(p.BN, p.T - n0);           // valid token columns of this tile (<= 0: padding tile of a group)
        const int nch = cols > 0 ? (cols + 15) >> 4 : 0;
        const bool plain = !ep.round_acc && ep.scale_cols == 0 && ep.group_rows == 0;
        const bool fast16 = plain && ep.out_dtype == MYR_F16 && ep.res == nullptr && ep.alpha == 1.0f && ep.act != MYR_ACT_SWIGLU;
        const bool fast32 = plain && ep.out_dtype == MYR_F32 && (ep.res == nullptr || ep.res_dtype == MYR_F32) && ep.act == MYR_ACT_NONE;
        const bool has_res = ep.res != nullptr || ks > 0;
        const int tile_flag = ((mt * p.n_nt + ng * p.P + pr) << 1) + r;
        if (p.S > 1 && ks > 0) {  // ordered in-place accumulation: wait until the splits below this one have landed
          if (threadIdx.x == 64) {
            while (ld_acquire(p.flags + tile_flag) != ks) __nanosleep(64);
          }
          g2_bar_sync(1, G2_EPI_THREADS);
        }
        auto do_chunk = [&](const uint32_t (&rr)[16], int ch, const float* have) {
          const int t0 = n0 + ch * 16;
          const int nc = min(16, p.T - t0);
          if (!f_ok) return;
          if (fast32 && have != nullptr) {
            float* op = reinterpret_cast<float*>(ep.out) + (long long)t0 * ep.ldo + f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nc) op[j * ep.ldo] = (__uint_as_float(rr[j]) + bias_f) * ep.alpha + have[j];
          } else if (fast32) {
            // out = (acc + bias) * alpha + res: all residual loads are issued before the first store (res may alias out)
            const float* rp = (ks == 0 ? reinterpret_cast<const float*>(ep.res) + (long long)t0 * ep.ldr
                                       : reinterpret_cast<const float*>(ep.out) + (long long)t0 * ep.ldo) + f;
            const long long ldr = ks == 0 ? ep.ldr : ep.ldo;
            float* op = reinterpret_cast<float*>(ep.out) + (long long)t0 * ep.ldo + f;
            float rv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) rv[j] = (has_res && j < nc) ? __ldcg(rp + j * ldr) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nc) op[j * ep.ldo] = (__uint_as_float(rr[j]) + bias_f) * ep.alpha + rv[j];
          } else if (fast16) {
            __half* op = reinterpret_cast<__half*>(ep.out) + (long long)t0 * ep.ldo + f;
            if (ep.act == MYR_ACT_GELU_ERF) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float v = __uint_as_float(rr[j]) + bias_f;
                if (j < nc) op[j * ep.ldo] = __float2half_rn(0.5f * v * (1.0f + fast_erf(v * 0.70710678118654752440f)));
              }
            } else if (ep.act == MYR_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nc) op[j * ep.ldo] = __float2half_rn(fmaxf(__uint_as_float(rr[j]) + bias_f, 0.f));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nc) op[j * ep.ldo] = __float2half_rn(__uint_as_float(rr[j]) + bias_f);
            }
          } else {
            Epilogue e2 = ep;
            if (ks > 0) {  // later splits add onto what the earlier ones stored
              e2.res = ep.out; e2.res_dtype = ep.out_dtype; e2.ldr = ep.ldo;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nc) epi_store_scalar(e2, epi_transform(e2, __uint_as_float(rr[j]), bias_f, f), t0 + j, f, 0);
          }
        };
        // two TMEM loads in flight per wait: 32 independent columns of epilogue work behind every tcgen05.wait::ld
        for (int ch = csel; ch < nch; ch += 8) {
          uint32_t ra[16], rb[16];
          const bool two = ch + 4 < nch;
          tmem_ld16(taddr + ch * 16, ra);
          if (two) tmem_ld16(taddr + (ch + 4) * 16, rb);
          tmem_ld_wait();
          const bool first = pre_ok && ch == csel;
          if (first) {
            do_chunk(ra, ch, pre[0]);
            if (two) do_chunk(rb, ch + 4, pre[1]);
          } else {
            do_chunk(ra, ch, nullptr);
            if (two) do_chunk(rb, ch + 4, nullptr);
          }
        }
        if (p.S > 1) {
          __threadfence();
          g2_bar_sync(1, G2_EPI_THREADS);
          if (threadIdx.x == 64) st_release(p.flags + tile_flag, ks == p.S - 1 ? 0 : ks + 1);
        }
      } else {
        // lanes = features with SwiGLU: lanes 0..63 of this CTA hold 64 gate rows, lanes 64..127 the matching up rows.
        // The up half hands its values over through shared memory (double-buffered, one named barrier per chunk).
        const int i = (m0 >> 1) + (lrow & 63);
        const bool i_ok = i < (p.F >> 1);
        float* buf0 = gu_stage + csel * (2 * 16 * 64);
        for (int ch = csel; ch < nchunks; ch += 4, ++it) {
          if (n0 + ch * 16 >= p.T) break;
          uint32_t rr[16];
          tmem_ld16(taddr + ch * 16, rr);
          tmem_ld_wait();
          float* buf = buf0 + (it & 1) * (16 * 64);
          if (q >= 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) buf[j * 64 + (lrow - 64)] = __uint_as_float(rr[j]);
          }
          g2_bar_sync(2 + csel, 128);
          if (q < 2 && i_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = n0 + ch * 16 + j;
              if (t < p.T)
                reinterpret_cast<__half*>(p.ep.out)[out_row_offset(p.ep, t) + i] =
                    __float2half_rn(swiglu_pair(__uint_as_float(rr[j]), buf[j * 64 + lrow]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  if (p.trace && threadIdx.x == 64) p.trace[blockIdx.x * 6 + 4] = gtime_ns();  // epilogue of this CTA done
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peers may still multicast into / signal this CTA's shared memory, and the pair frees TMEM together
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct Plan2 {
  int row_mode, BN, P, S, n_mt, n_nt, n_ng, n_units, n_clusters, num_stages, stage_bytes, ka, swiglu;
  double cost;
};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

static bool g2_init() {
  static int state = 0;  // 0: not tried, 1: ok, -1: failed
  if (state == 0) {
    state = (cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess) ? 1 : -1;
    if (state < 0) cudaGetLastError();
  }
  return state > 0;
}

static int g2_max_clusters(int P, size_t smem_bytes) {
  static int cache[5] = {0, 0, 0, 0, 0};
  if (cache[P] > 0) return cache[P];
  g2_init();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sm_count() / (2 * P) * (2 * P)));
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)(2 * P);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm2_kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = sm_count() / (2 * P) - (P > 1 ? 1 : 0);
  }
  cache[P] = n;
  return n;
}

static void g2_fill(Plan2& pl, int T, int F, int K) {
  const int rows_a = pl.row_mode ? T : F, rows_b = pl.row_mode ? F : T;
  pl.n_mt = ceil_div(rows_a, 256);
  pl.n_nt = ceil_div(rows_b, pl.BN);
  pl.n_ng = ceil_div(pl.n_nt, pl.P);
  pl.n_units = pl.n_mt * pl.n_ng * pl.S;
  // ring stage = ka k-atoms of 64: a TMA copy costs its issuing thread ~0.15 us whatever its size, and at one atom per stage
  // (two copies per 64 k) the main loop ran at the copy-issue rate, not at the tensor or the memory rate (profiles/r2_gemm2_sweep.md)
  pl.ka = (K % G2_BK == 0) ? env_int("MYR_G2_KA", 2) : 1;
  if (pl.ka < 1 || pl.ka > 4) pl.ka = 1;
  pl.stage_bytes = pl.ka * (G2_A_BYTES + (pl.BN / 2) * G2_BK * 2);
  // the SwiGLU hand-over buffer is only there for that epilogue: the other launches give its 8 KB (and the slack) to the ring
  const int budget = pl.swiglu ? G2_SMEM_TILE_BUDGET : (227 * 1024 - 1024 - 256 - 512);
  int st = budget / pl.stage_bytes;
  pl.num_stages = st > G2_MAX_STAGES ? G2_MAX_STAGES : st;
  const int f_st = env_int("MYR_G2_STAGES", 0);
  if (f_st > 0 && f_st < pl.num_stages) pl.num_stages = f_st;
}

static size_t g2_smem_bytes(const Plan2& pl) {
  return (size_t)pl.num_stages * pl.stage_bytes + 1024 /*align*/ + 256 /*barriers*/ + (pl.swiglu ? G2_SWIGLU_STAGE_BYTES : 0);
}

// split-K adds partial sums in place, in split order: only for fp32 out = fp32 residual + X W^T (o_proj / down_proj / ViT
// proj / fc2: few output tiles, long K) in the lanes = features arrangement
static bool g2_split_ok(const myr_gemm_args* a) {
  return a->out_dtype == MYR_F32 && a->res != nullptr && a->res_dtype == MYR_F32 && a->act == MYR_ACT_NONE && !a->round_acc &&
         a->scale_cols == 0 && a->out_group_rows == 0 && a->workspace != nullptr && a->workspace_bytes >= 64 * 1024 &&
         (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0;
}

// Cost model in SM cycles (relative; calibrated on B200, profiles/r2_gemm2_sweep.md): per 64-wide k-block a pair needs
// 2 * BN tensor cycles, 256 + BN cycles of shared-memory traffic (TMA writes + operand reads of (128 + BN / 2) x 128 B at
// 128 B/clk) and its L2 -> SM bytes at ~l2_bw B/clk/SM; a unit adds its epilogue where the next unit's MMAs do not hide
// it; units run in waves over the CTA pairs.
static Plan2 g2_plan(const myr_gemm_args* a) {
  const int T = a->T, F = a->F, K = a->K;
  const bool swiglu = a->act == MYR_ACT_SWIGLU;
  const int kb = ceil_div(K, G2_BK);
  const int f_mode = env_int("MYR_G2_MODE", -1), f_bn = env_int("MYR_G2_BN", 0), f_p = env_int("MYR_G2_P", 0), f_s = env_int("MYR_G2_S", 0);
  const double l2_bw = (double)env_int("MYR_G2_L2BW", 48);
  const bool split_ok = g2_split_ok(a);
  Plan2 best;
  best.cost = 1e30;
  best.BN = 0;
  for (int mode = 0; mode <= 1; ++mode) {
    if (f_mode >= 0 && mode != f_mode) continue;
    if (f_mode < 0 && mode == 1 && !swiglu) continue;  // lanes = tokens needs a transposing epilogue to store coalesced: not built
    const int gran = (mode && swiglu) ? 128 : 16;
    for (int bn = gran; bn <= 256; bn += gran) {
      if (f_bn > 0 && bn != f_bn) continue;
      if (bn < 128 && bn != f_bn) continue;  // narrower tiles move more operand bytes per flop than an SM can take in
      for (int P = 1; P <= 4; P *= 2) {
        if (f_p > 0 ? P != f_p : P != 1) continue;  // multicast across pairs measured no gain on B200 (L2 already merges)
        for (int S = 1; S <= 4; ++S) {
          // ordered split-K serialises the epilogues of a tile's splits: measured a loss, except where one split per tile leaves
          // at least half of the CTA pairs idle and K is long (ViT fc2 at batch 4: 36 tiles x K = 6144 -> 33 us against 39 us)
          const int base_units = ceil_div(mode ? T : F, 256) * ceil_div(ceil_div(mode ? F : T, bn), P);
          const bool two_way = S == 2 && split_ok && mode == 0 && 2 * base_units <= sm_count() / 2 && kb >= 64;
          if (f_s > 0 ? S != f_s : (S != 1 && !two_way)) continue;
          if (S > 1 && (!split_ok || mode != 0 || (S - 1) * ceil_div(kb, S) >= kb || kb / S < 8)) continue;
          Plan2 pl;
          pl.row_mode = mode;
          pl.swiglu = swiglu ? 1 : 0;
          pl.BN = bn;
          pl.P = P;
          pl.S = S;
          g2_fill(pl, T, F, K);
          if (P > 1 && pl.n_nt < P && f_p == 0) continue;
          if (S > 1 && pl.n_mt * pl.n_nt * 2 > 16000) continue;
          pl.n_clusters = g2_max_clusters(P, g2_smem_bytes(pl));
          if (pl.n_clusters > pl.n_units) pl.n_clusters = pl.n_units;
          const int waves = ceil_div(pl.n_units, pl.n_clusters);
          const double mma = 2.0 * bn;
          const double sm = 256.0 + bn;
          const double l2 = (128.0 / P + bn / 2.0) * 128.0 / l2_bw;
          double per = mma > sm ? mma : sm;
          if (l2 > per) per = l2;
          const double epi = 1500.0 + 20.0 * bn;  // one tile's epilogue (cycles), exposed once per CTA and where MMAs are shorter
          const double unit_mma = ceil_div(kb, S) * per + 500.0;
          const double unit = unit_mma > epi ? unit_mma : epi;
          pl.cost = waves * unit + epi + (S > 1 ? 800.0 : 0.0);
          if (pl.cost < best.cost) best = pl;
        }
      }
    }
  }
  return best;
}

bool gemm2_eligible(const myr_gemm_args* a, int nbatch) {
  static int on = -1;
  if (on < 0) on = env_int("MYR_G2", 1);
  if (!on) return false;
  const int min_t = env_int("MYR_G2_MIN_T", 128);
  if (a->T < min_t || a->x_mn_major || a->w_mn_major || nbatch != 1 || a->ksplit_hint > 1 || a->bn_hint > 0) return false;
  if (a->K < 256 || a->F < 128) return false;
  if (a->norm_ss != nullptr || a->post_out16 != nullptr) return false;
  return true;
}

int gemm2_launch(const myr_gemm_args* a, cudaStream_t stream, int* handled) {
  const Plan2 pl = g2_plan(a);
  *handled = 1;
  if (pl.BN == 0) {
    set_error("gemm2: no plan for T=%d F=%d K=%d", a->T, a->F, a->K);
    return MYR_ERR_UNSUPPORTED;
  }
  // few output tiles (o_proj / down_proj / ViT proj / fc2 at batch 4: 36-48 tiles of 256 x 176 for 74 CTA pairs): the stream-K
  // kernel of gemm.cu, which splits K over all 148 SMs, is faster there (profiles/r2_gemm2_sweep.md)
  // ... and so is a launch whose last wave of tiles would leave most CTA pairs idle (ViT qkv at batch 4: 102 tiles = 1.38 waves)
  const int waves = ceil_div(pl.n_units, pl.n_clusters);
  const double wave_eff = (double)pl.n_units / ((double)waves * pl.n_clusters);
  // ... and so is a SHORT launch inside a sequence of other kernels: a pair-kernel launch (cluster of 2, all 512 TMEM columns, 200+ KB of
  // shared memory) overlaps worse with its neighbours' tails than the single-CTA kernel; measured in situ, the ViT qkv GEMM (12 GFLOP, 136 tiles)
  // cost the 39-block encoder +1.3 ms on this kernel although it is 1.6 us faster back to back (profiles/r2_gemm2_sweep.md §5)
  const double gflop = 2.0 * a->T * (double)a->F * a->K * 1e-9;
  if ((pl.n_units < env_int("MYR_G2_MIN_UNITS", 60) || wave_eff < 0.8 || gflop < env_int("MYR_G2_MIN_GFLOP", 15)) && env_int("MYR_G2_MODE", -1) < 0) {
    *handled = 0;
    return MYR_OK;
  }
  const void* pa = pl.row_mode ? a->x : a->w;
  const void* pb = pl.row_mode ? a->w : a->x;
  const int64_t lda = pl.row_mode ? a->ldx : a->ldw, ldb = pl.row_mode ? a->ldw : a->ldx;
  const int rows_a = pl.row_mode ? a->T : a->F, rows_b = pl.row_mode ? a->F : a->T;
  CUtensorMap tmA, tmB, tmA2;
  memset(&tmA2, 0, sizeof(tmA2));
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)rows_a; box[0] = G2_BK; box[1] = (uint32_t)(G2_A_ROWS / pl.P);
    strides[0] = (uint64_t)lda * 2;
    int rc = MYR_OK;
    if (pl.ka == 1) {
      rc = make_tmap_f16(&tmA, pa, 2, dims, strides, box);
      if (rc) return rc;
      dims[1] = (uint64_t)rows_b; box[1] = (uint32_t)(pl.BN / 2);
      strides[0] = (uint64_t)ldb * 2;
      rc = make_tmap_f16(&tmB, pb, 2, dims, strides, box);
      if (rc) return rc;
    } else {
      // (64 k, rows, K / 64 atoms): a box of ka atoms lands as [atom][row][128 B], each atom a canonical 128-byte-swizzled tile
      uint64_t d3[3] = {G2_BK, (uint64_t)rows_a, (uint64_t)(a->K / G2_BK)}, s3[2] = {(uint64_t)lda * 2, G2_BK * 2};
      uint32_t b3[3] = {G2_BK, G2_A_ROWS, (uint32_t)pl.ka};
      rc = make_tmap_f16(&tmA, pa, 3, d3, s3, b3);
      if (rc) return rc;
      if (pl.P > 1) {  // per-atom multicast slices of the A tile
        rc = make_tmap_f16(&tmA2, pa, 2, dims, strides, box);
        if (rc) return rc;
      }
      d3[1] = (uint64_t)rows_b; s3[0] = (uint64_t)ldb * 2; b3[1] = (uint32_t)(pl.BN / 2);
      rc = make_tmap_f16(&tmB, pb, 3, d3, s3, b3);
      if (rc) return rc;
    }
  }
  Gemm2Params p;
  p.T = a->T; p.F = a->F; p.K = a->K;
  p.row_mode = pl.row_mode; p.BN = pl.BN; p.P = pl.P;
  p.n_mt = pl.n_mt; p.n_nt = pl.n_nt; p.n_ng = pl.n_ng; p.kb_total = ceil_div(a->K, G2_BK * pl.ka); p.n_units = pl.n_units;
  p.num_stages = pl.num_stages; p.stage_bytes = pl.stage_bytes;
  p.dbg = env_int("MYR_G2_DBG", 0);
  p.idesc = make_idesc_f16(256, (p.dbg & 8) ? 16 : pl.BN, 0, 0);
  p.S = pl.S;
  p.kb_per = ceil_div(p.kb_total, pl.S);
  p.flags = reinterpret_cast<int*>(a->workspace);  // head of the workspace: zero on entry, left at zero (include/myriad_b200.h)
  p.ka = pl.ka;
  p.pf_dist = pl.ka > 1 ? 0 : env_int("MYR_G2_PF", 0);  // measured: prefetching ahead of the ring costs 10-15 % on B200 (profiles/r2_gemm2_sweep.md)
  p.trace = next_trace_slot();
  p.ep.bias = reinterpret_cast<const __half*>(a->bias);
  p.ep.act = a->act; p.ep.round_acc = a->round_acc;
  p.ep.scale_cols = a->scale_cols; p.ep.scale = a->scale;
  p.ep.res = a->res; p.ep.res_dtype = a->res_dtype; p.ep.ldr = a->ldr;
  p.ep.out = a->out; p.ep.out_dtype = a->out_dtype; p.ep.ldo = a->ldo;
  p.ep.group_rows = a->out_group_rows; p.ep.group_stride = a->out_group_stride;
  p.ep.alpha = a->alpha_set ? a->alpha : 1.0f;
  {
    const int oq = a->out_dtype == MYR_F32 ? 4 : 8, rq = a->res_dtype == MYR_F32 ? 4 : 8;
    bool v = (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->ldo % oq == 0 && a->out_group_stride % oq == 0;
    if (a->res) v = v && (reinterpret_cast<uintptr_t>(a->res) & 15) == 0 && a->ldr % rq == 0;
    if (a->bias) v = v && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
    p.ep.vec = v ? 1 : 0;
  }
  if (!g2_init()) {
    set_error("gemm2: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
    return MYR_ERR_CUDA;
  }
  static int dbg = -1;
  if (dbg < 0) dbg = env_int("MYR_G2_DEBUG", 0);
  static long long last_key = -1;
  const long long key = ((((long long)a->T * 65536 + a->F) * 65536 + a->K) * 2 + pl.row_mode) * 4096 + pl.BN * 8 + pl.P + pl.S * 2048;
  if (dbg && key != last_key && ((last_key = key), true))
    fprintf(stderr, "[gemm2] T=%d F=%d K=%d -> mode=%d BN=%d P=%d S=%d units=%d clusters=%d stages=%d cost=%.0f\n", a->T, a->F, a->K,
            pl.row_mode, pl.BN, pl.P, pl.S, pl.n_units, pl.n_clusters, pl.num_stages, pl.cost);
  const int grid = pl.n_clusters * 2 * pl.P;
  MYR_CHECK_CUDA(launch_kernel_cluster(gemm2_kernel, dim3((unsigned)grid), dim3(G2_THREADS), g2_smem_bytes(pl), stream, a->pdl != 0,
                                       2 * pl.P, tmA, tmB, tmA2, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}

}  // namespace myr
