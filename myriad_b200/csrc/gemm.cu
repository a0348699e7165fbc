// tcgen05 + TMA GEMM for sm_100a:  out[t, f] = epilogue( sum_k X[t, k] * W[f, k] ).
//
// One persistent, warp-specialised kernel (320 threads, 1 CTA / SM) in two operand arrangements:
//   * "lanes = features" (T <= 64: decode / weight streaming, HBM-bound): the 128-row UMMA A tile is a weight tile, the
//     token tile (BN = 16..64) is the UMMA N side. Every byte of W is read exactly once.
//   * "lanes = tokens"   (T  > 64: prefill / ViT / training, tensor-bound): A = 128 tokens, B = BN <= 256 features, so an
//     epilogue thread owns one token row and 16 consecutive features per TMEM load: bias / residual / output move as
//     16-byte vectors.
// Warp roles:
//   warp 0    : TMA producer   (global -> 128B-swizzled smem ring, up to 8 stages, mbarrier complete_tx)
//   warp 1    : MMA issuer     (one thread issues tcgen05.mma; tcgen05.commit frees smem stages / publishes TMEM)
//   warps 2-9 : epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> global), two warps per TMEM lane quarter
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of one segment overlap the MMAs of the next.
//
// Work decomposition ("stream-K"): the (tile, k-block) space is cut into one contiguous range of `per` k-blocks per CTA.
// `per` a multiple of the k-blocks per tile gives the classic data-parallel schedule; otherwise ranges cross tile
// borders, every CTA gets the same number of k-blocks (no wave quantisation when 172 weight tiles meet 148 SMs) and a
// tile covered by several CTAs is finished by whichever CTA arrives last: partial accumulators go to the workspace, an
// arrival counter per tile elects the finisher, which sums the partials in segment order (deterministic) and runs the
// fused epilogue. The counters live at the head of the workspace, start at zero and are reset by the finisher.
//
// Programmatic dependent launch: the kernel releases its dependents at once and waits for its own predecessor only
// before it touches activations; with `w_static` the producer streams the first ring of WEIGHT tiles before that wait, so
// weight prefetch overlaps the tail of the previous kernel.
//
// Replaces the cuBLAS calls behind nn.Linear in the reference hot path (see include/myriad_b200.h).
#include "common.h"
#include "ptx.cuh"
#include "gemm_epi.cuh"

namespace myr {

constexpr int BM = 128;       // rows of the A tile (UMMA M)
constexpr int BK = 64;        // halfs per k-block = one 128-byte swizzle row
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int SMEM_TILE_BUDGET = 192 * 1024;      // 1 CTA / SM (lanes = tokens)
constexpr int SMEM_TILE_BUDGET_OCC2 = 108 * 1024;  // 2 CTAs / SM (lanes = features): the next kernel's CTA moves in early
constexpr int GEMM_THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STAGE_COLS = 256;
constexpr size_t COUNTER_BYTES = 64 * 1024;  // head of the workspace: int32 arrival counters, one per split tile
constexpr int MAX_COUNTERS = (int)(COUNTER_BYTES / 4);


struct GemmKernelParams {
  int T, F, K;
  int row_mode;            // 1: lanes = tokens (A = X), 0: lanes = features (A = W)
  int BN, n_mt, n_nt, kb_total;
  long long total_kb;      // n_tiles * kb_total
  int per;                 // k-blocks per CTA range
  int max_seg;             // partial slots per tile
  int num_stages, stage_bytes;
  int a_mn, b_mn;
  uint32_t idesc;
  float* partial;
  int* counters;
  int nb1, nbatch;                 // batched mode: batch index bidx -> (b0 = bidx / nb1, b1 = bidx % nb1)
  long long o_bs0, o_bs1;          // output element strides of the two batch dims
  int w_static;
  int tmem_cols, acc_stride;       // TMEM columns allocated by this CTA, columns between the two accumulator stages
  int n_mp;                        // cluster mode: pairs of token tiles (a "tile" index then names a pair x one weight tile)
  long long* trace;                // debug (MYR_GEMM_TRACE): per CTA 6 x %globaltimer ns, see myr_gemm_trace_read
  Epilogue ep;
};


struct Seg {
  int tile, kb0, kb1;
};
// next segment of the contiguous k-block range [g, g1): stays inside one tile
__device__ __forceinline__ Seg next_seg(long long& g, long long g1, int kb_total) {
  Seg s;
  s.tile = (int)(g / kb_total);
  s.kb0 = (int)(g - (long long)s.tile * kb_total);
  const long long room = g1 - g;
  s.kb1 = (room < (long long)(kb_total - s.kb0)) ? s.kb0 + (int)room : kb_total;
  g += s.kb1 - s.kb0;
  return s;
}

struct TileCoord {
  int m0, n0, b0, b1, bidx;
};
// cluster of 2 CTAs: both work on the same weight (B) tile and on adjacent token tiles; each loads half of the B tile
// and multicasts it to the pair, halving the L2 -> SM traffic that bounds the lanes = tokens arrangement
__device__ __forceinline__ TileCoord tile_coord_pair(const GemmKernelParams& p, int tile, int rank) {
  TileCoord c;
  const int nt = tile / p.n_mp;
  const int mp = tile - nt * p.n_mp;
  c.bidx = 0;
  c.b0 = 0;
  c.b1 = 0;
  c.m0 = (2 * mp + rank) * BM;
  c.n0 = nt * p.BN;
  return c;
}

__device__ __forceinline__ TileCoord tile_coord(const GemmKernelParams& p, int tile) {
  TileCoord c;
  const int per_batch = p.n_mt * p.n_nt;
  c.bidx = tile / per_batch;
  const int rem = tile - c.bidx * per_batch;
  int mt, nt;
  if (p.row_mode) {  // CTAs running side by side share the weight (B) tile and walk the token tiles
    nt = rem / p.n_mt;
    mt = rem - nt * p.n_mt;
  } else {
    mt = rem / p.n_nt;
    nt = rem - mt * p.n_nt;
  }
  c.m0 = mt * BM;
  c.n0 = nt * p.BN;
  c.b0 = c.bidx / p.nb1;
  c.b1 = c.bidx - c.b0 * p.nb1;
  return c;
}

// kOcc = 2 (weight streaming): <= 96 registers, <= ~110 KB smem and only the TMEM columns it needs, so the CTA of the NEXT
// GEMM in the stream (programmatic dependent launch) is resident and has its weight ring full before this one drains.
template <int kOcc, int kCluster = 1>
__global__ void __launch_bounds__(GEMM_THREADS, kOcc)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.num_stages * p.stage_bytes);
  uint64_t* full = bars;                      // [MAX_STAGES]  TMA -> MMA
  uint64_t* empty = bars + MAX_STAGES;        // [MAX_STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * MAX_STAGES;    // [2]           MMA -> epilogue
  uint64_t* tempty = tfull + 2;               // [2]           epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile int* s_last = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kCluster);  // cluster mode: a stage is free once BOTH CTAs' MMAs have read it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], N_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 6 + 0] = gtime_ns();
  const int crank = kCluster > 1 ? (int)cluster_ctarank() : 0;
  const int unit = kCluster > 1 ? (int)(blockIdx.x / kCluster) : (int)blockIdx.x;  // both CTAs of a cluster walk the same range

  const long long g0 = (long long)unit * p.per;
  const long long g1 = (g0 + p.per < p.total_kb) ? g0 + p.per : p.total_kb;
  const uint32_t b_bytes = (uint32_t)p.BN * BK * 2;
  const uint32_t tx_bytes = A_STAGE_BYTES + b_bytes;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // lanes = features: A is a weight matrix read exactly once -> evict-first keeps activations / partial tiles in L2
      const uint64_t pol_a = p.row_mode ? 0ull : l2_policy_evict_first();
      // weight (A) tiles of the first ring fill do not depend on the previous kernel: issue them before the PDL wait
      int pre = 0;
      if (p.w_static) {
        long long g = g0;
        while (g < g1 && pre < p.num_stages) {
          const Seg s = next_seg(g, g1, p.kb_total);
          const TileCoord c = tile_coord(p, s.tile);
          for (int kb = s.kb0; kb < s.kb1 && pre < p.num_stages; ++kb, ++pre) {
            uint8_t* sa = smem + pre * p.stage_bytes;
            mbar_arrive_expect_tx(&full[pre], tx_bytes);
            tma_load_4d_hint(sa, &tmA, &full[pre], kb * BK, c.m0, c.b1, c.b0, pol_a);
          }
        }
      }
      pdl_wait();
      if (p.trace) p.trace[blockIdx.x * 6 + 1] = gtime_ns();
      int it = 0;
      long long g = g0;
      while (g < g1) {
        const Seg s = next_seg(g, g1, p.kb_total);
        const TileCoord c = kCluster > 1 ? tile_coord_pair(p, s.tile, crank) : tile_coord(p, s.tile);
        for (int kb = s.kb0; kb < s.kb1; ++kb, ++it) {
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (it >= pre) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], tx_bytes);
            if (!p.a_mn) {
              if (p.row_mode)
                tma_load_4d(sa, &tmA, &full[stage], kb * BK, c.m0, c.b1, c.b0);
              else
                tma_load_4d_hint(sa, &tmA, &full[stage], kb * BK, c.m0, c.b1, c.b0, pol_a);
            } else {
              tma_load_4d(sa, &tmA, &full[stage], c.m0, kb * BK, c.b1, c.b0);
              tma_load_4d(sa + 8192, &tmA, &full[stage], c.m0 + 64, kb * BK, c.b1, c.b0);
            }
          }
          if constexpr (kCluster > 1) {
            // my half of the weight tile, delivered to both CTAs of the pair
            const int half_rows = p.BN / 2;
            tma_load_4d_multicast(sb + crank * half_rows * (BK * 2), &tmB, &full[stage], kb * BK, c.n0 + crank * half_rows, 0, 0,
                                  (uint16_t)0x3);
          } else if (!p.b_mn) {
            tma_load_4d(sb, &tmB, &full[stage], kb * BK, c.n0, c.b1, c.b0);
          } else {
            for (int i = 0; i < p.BN / 64; ++i) tma_load_4d(sb + i * 8192, &tmB, &full[stage], c.n0 + 64 * i, kb * BK, c.b1, c.b0);
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
      const uint32_t a_lbo = p.a_mn ? 8192 : 16, b_lbo = p.b_mn ? 8192 : 16;
      const uint32_t a_kstep = p.a_mn ? 2048 : 32, b_kstep = p.b_mn ? 2048 : 32;
      long long g = g0;
      while (g < g1) {
        const Seg s = next_seg(g, g1, p.kb_total);
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = make_smem_desc(sa + kk * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + kk * b_kstep, b_lbo, 1024);
            tc_mma_f16(d_tmem, da, db, p.idesc, (kb > s.kb0 || kk > 0) ? 1u : 0u);
          }
          if constexpr (kCluster > 1)
            tc_commit_multicast(&empty[stage], (uint16_t)0x3);  // the peer multicasts into this stage too
          else
            tc_commit(&empty[stage]);  // frees this smem stage once the MMAs above have read it
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull[as]);  // accumulator complete -> epilogue
        if (p.trace) p.trace[blockIdx.x * 6 + 2] = gtime_ns();  // last MMA of the segment issued
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int hsel = (warp - 2) >> 2;     // which half of the column chunks this warp handles
    const int epi_tid = threadIdx.x - 64;
    const int lrow = q * 32 + lane;       // TMEM lane == row of the A tile
    const bool swiglu = p.ep.act == MYR_ACT_SWIGLU;
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();  // residual / output / workspace may still be in use by the previous kernel
    long long g = g0;
    while (g < g1) {
      const Seg s = next_seg(g, g1, p.kb_total);
      const TileCoord c = kCluster > 1 ? tile_coord_pair(p, s.tile, crank) : tile_coord(p, s.tile);
      const long long boff = (long long)c.b0 * p.o_bs0 + (long long)c.b1 * p.o_bs1;
      // which CTAs cover this tile
      const long long tb = (long long)s.tile * p.kb_total;
      const int first = (int)(tb / p.per), last = (int)((tb + p.kb_total - 1) / p.per);
      const int n_seg = last - first + 1;
      const int seg = unit - first;
      const bool via_ws = n_seg > 1 || (swiglu && !p.row_mode);
      const int nchunks = p.BN / 16;
      float* ws = p.partial + ((size_t)s.tile * p.max_seg + seg) * ((size_t)p.BN * BM);


      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (p.trace && epi_tid == 0) p.trace[blockIdx.x * 6 + 3] = gtime_ns();  // accumulator of the (last) segment ready
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * p.acc_stride;

      if (via_ws) {
        for (int ch = hsel; ch < nchunks; ch += 2) {
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) ws[(size_t)(ch * 16 + j) * BM + lrow] = __uint_as_float(r[j]);
        }
      } else if (p.row_mode) {
        const long long t = (long long)c.m0 + lrow;
        const bool t_ok = t < p.T;
        const long long obase = boff + (t_ok ? out_row_offset(p.ep, t) : 0);
        if (!swiglu) {
          for (int ch = hsel; ch < nchunks; ch += 2) {
            const int f0 = c.n0 + ch * 16;
            if (f0 >= p.F) break;
            uint32_t r[16];
            tmem_ld16(taddr + ch * 16, r);
            tmem_ld_wait();
            if (t_ok) {
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
              epi_row16(p.ep, v, t, f0, p.F, obase);
            }
          }
        } else {
          // weight rows are interleaved in blocks of 64: [gate 0..63 | up 0..63 | gate 64..127 | ...]
          const int units = p.BN / 32;  // 16-wide gate chunks in this tile
          for (int u = hsel; u < units; u += 2) {
            const int blk = u >> 2, cg = u & 3;
            const int gcol = blk * 128 + cg * 16;
            if (c.n0 + gcol >= p.F) break;
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
              epi_row16_swiglu(p.ep, gv, uv, (c.n0 >> 1) + blk * 64 + cg * 16, p.F >> 1, obase);
            }
          }
        }
      } else {
        // lanes = features, single segment: straight from TMEM
        const int f = c.m0 + lrow;
        const bool f_ok = f < p.F;
        const float bias_f = (p.ep.bias && f_ok) ? __half2float(p.ep.bias[f]) : 0.f;
        for (int ch = hsel; ch < nchunks; ch += 2) {
          if (c.n0 + ch * 16 >= p.T) break;
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
          if (f_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = c.n0 + ch * 16 + j;
              if (t < p.T) epi_store_scalar(p.ep, epi_transform(p.ep, __uint_as_float(r[j]), bias_f, f), t, f, boff);
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

      if (via_ws) {
        // publish the partial, elect the finisher (last CTA to arrive at this tile)
        __threadfence();
        epi_bar_sync();
        if (epi_tid == 0) {
          int lastf = 1;
          if (n_seg > 1) {
            const int old = atomicAdd(&p.counters[s.tile], 1);
            lastf = (old == n_seg - 1);
            if (lastf) p.counters[s.tile] = 0;  // every segment has arrived: leave the counter clean for the next launch
          }
          *s_last = lastf;
        }
        epi_bar_sync();
        const int lastf = *s_last;
        if (lastf) {
          __threadfence();
          const float* w0 = p.partial + (size_t)s.tile * p.max_seg * ((size_t)p.BN * BM);
          const size_t seg_stride = (size_t)p.BN * BM;
          if (p.row_mode) {
            const long long t = (long long)c.m0 + lrow;
            if (t < p.T) {
              const long long obase = boff + out_row_offset(p.ep, t);
              if (!swiglu) {
                for (int ch = hsel; ch < nchunks; ch += 2) {
                  const int f0 = c.n0 + ch * 16;
                  if (f0 >= p.F) break;
                  float v[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float a = 0.f;
                    for (int sg = 0; sg < n_seg; ++sg) a += __ldcg(w0 + sg * seg_stride + (size_t)(ch * 16 + j) * BM + lrow);
                    v[j] = a;
                  }
                  epi_row16(p.ep, v, t, f0, p.F, obase);
                }
              } else {
                const int units = p.BN / 32;
                for (int u = hsel; u < units; u += 2) {
                  const int blk = u >> 2, cg = u & 3;
                  const int gcol = blk * 128 + cg * 16;
                  if (c.n0 + gcol >= p.F) break;
                  float gv[16], uv[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float a = 0.f, b = 0.f;
                    for (int sg = 0; sg < n_seg; ++sg) {
                      a += __ldcg(w0 + sg * seg_stride + (size_t)(gcol + j) * BM + lrow);
                      b += __ldcg(w0 + sg * seg_stride + (size_t)(gcol + 64 + j) * BM + lrow);
                    }
                    gv[j] = a;
                    uv[j] = b;
                  }
                  epi_row16_swiglu(p.ep, gv, uv, (c.n0 >> 1) + blk * 64 + cg * 16, p.F >> 1, obase);
                }
              }
            }
          } else if (!swiglu) {
            const int f = c.m0 + lrow;
            if (f < p.F) {
              const float bias_f = p.ep.bias ? __half2float(p.ep.bias[f]) : 0.f;
              for (int col = hsel; col < p.BN; col += 2) {
                const int t = c.n0 + col;
                if (t >= p.T) break;
                float a = 0.f;
                for (int sg = 0; sg < n_seg; ++sg) a += __ldcg(w0 + sg * seg_stride + (size_t)col * BM + lrow);
                epi_store_scalar(p.ep, epi_transform(p.ep, a, bias_f, f), t, f, boff);
              }
            }
          } else if (lrow < 64) {
            // lanes = features with SwiGLU: lane l holds gate row l, lane l + 64 the matching up row
            const int i = (c.m0 >> 1) + lrow;
            if (i < (p.F >> 1)) {
              for (int col = hsel; col < p.BN; col += 2) {
                const int t = c.n0 + col;
                if (t >= p.T) break;
                float a = 0.f, b = 0.f;
                for (int sg = 0; sg < n_seg; ++sg) {
                  a += __ldcg(w0 + sg * seg_stride + (size_t)col * BM + lrow);
                  b += __ldcg(w0 + sg * seg_stride + (size_t)col * BM + lrow + 64);
                }
                reinterpret_cast<__half*>(p.ep.out)[boff + out_row_offset(p.ep, t) + i] = __float2half_rn(swiglu_pair(a, b));
              }
            }
          }
        }
      }
    }
  }

  if (p.trace && threadIdx.x == 64) p.trace[blockIdx.x * 6 + 4] = gtime_ns();  // epilogue / fix-up of this CTA done
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();  // the peer may still multicast into / signal this CTA's shared memory
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// host side: operand arrangement, tile shape and range length
// ---------------------------------------------------------------------------------------------------
struct Plan {
  int row_mode, BN, n_mt, n_nt, kb_total, per, grid, max_seg, num_stages, stage_bytes;
  int cluster, n_mp;  // cluster = 2: CTA pairs share a multicast weight tile (tiles are then (token-tile pair, weight tile))
  long long total_kb;
  int n_tiles;
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static Plan make_plan(const myr_gemm_args* a, int nbatch, bool allow_split, size_t ws_floats) {
  Plan pl;
  const int sms = sm_count();
  const int T = a->T, F = a->F, K = a->K;
  const bool swiglu = a->act == MYR_ACT_SWIGLU;
  pl.kb_total = ceil_div(K, BK);
  pl.row_mode = (T > 64) ? 1 : 0;
  pl.cluster = 1;
  pl.n_mp = 0;
  int gran;
  if (pl.row_mode) {
    gran = swiglu ? 128 : (a->w_mn_major ? 64 : 16);
    pl.n_mt = ceil_div(T, BM);
    static int cluster_env = -1;
    if (cluster_env < 0) {
      const char* e = getenv("MYR_GEMM_CLUSTER");
      cluster_env = (e && e[0] == '0') ? 0 : 1;
    }
    // lanes = tokens is bounded by L2 -> SM bandwidth (~42 B/clk/SM): pairs of CTAs on adjacent token tiles share the
    // weight tile through TMA multicast, which halves the weight traffic per CTA
    // (measured on B200: +30 % at T = 8224, F = 6144, K = 1408; a loss below ~32 token tiles, where pairing costs a wave)
    if (cluster_env && !a->x_mn_major && !a->w_mn_major && nbatch == 1 && pl.n_mt >= 32 && a->ksplit_hint <= 1 && sms % 2 == 0)
      pl.cluster = 2;
    const int m_units = pl.cluster == 2 ? ceil_div(pl.n_mt, 2) : pl.n_mt;
    const int slots = pl.cluster == 2 ? sms / 2 : sms;
    int best_bn = 0;
    if (a->bn_hint > 0) {
      best_bn = round_up(a->bn_hint, gran);
      if (best_bn > 256) best_bn = 256;
    } else {
      // cost (cycles per 16-wide k-step): MMA = BN/2, smem operand reads = 32 + BN/4; x rounds of persistent CTAs
      double best = 1e30;
      const int f_cap = round_up(F, gran);
      for (int bn = 256; bn >= (gran > 32 ? gran : 32); bn -= gran) {
        if (bn > f_cap && bn != gran) continue;
        const long long tiles = (long long)m_units * ceil_div(F, bn) * nbatch;
        const long long rounds = (tiles + slots - 1) / slots;
        // cycles per 64-wide k-block: tensor pipe 2 * bn; L2 -> SM operand traffic at ~36 B/clk/SM (the binding one for
        // bn <= 256: measured 12 TB/s chip-wide on the T = 524 gate/up GEMM); multicast halves the weight-tile bytes
        const double mma = 2.0 * bn;
        const double l2 = (16384.0 + 128.0 * bn / pl.cluster) / 36.0;
        const double per = mma > l2 ? mma : l2;
        const double cost = rounds * (pl.kb_total * per + 400.0 + 3.0 * bn);
        if (cost < best - 1e-9) {
          best = cost;
          best_bn = bn;
        }
      }
      if (best_bn == 0) best_bn = gran > 32 ? gran : 32;
    }
    pl.BN = best_bn;
    pl.n_nt = ceil_div(F, pl.BN);
  } else {
    gran = a->x_mn_major ? 64 : 16;
    pl.BN = round_up(T, gran);
    if (a->bn_hint > 0 && round_up(a->bn_hint, gran) < pl.BN) pl.BN = round_up(a->bn_hint, gran);
    pl.n_mt = ceil_div(F, BM);
    pl.n_nt = ceil_div(T, pl.BN);
  }
  pl.n_tiles = pl.n_mt * pl.n_nt * nbatch;
  if (pl.cluster == 2) {
    pl.n_mp = ceil_div(pl.n_mt, 2);
    pl.n_tiles = pl.n_mp * pl.n_nt;
  }
  pl.total_kb = (long long)pl.n_tiles * pl.kb_total;
  pl.stage_bytes = A_STAGE_BYTES + pl.BN * BK * 2;
  int st = (pl.row_mode ? SMEM_TILE_BUDGET : SMEM_TILE_BUDGET_OCC2) / pl.stage_bytes;
  pl.num_stages = st > MAX_STAGES ? MAX_STAGES : st;

  // data-parallel default: whole tiles per CTA (per CTA pair in cluster mode)
  const int units = pl.cluster == 2 ? sms / 2 : sms;
  const int tiles_per_cta = ceil_div(pl.n_tiles, units);
  pl.per = tiles_per_cta * pl.kb_total;
  pl.grid = ceil_div(pl.n_tiles, tiles_per_cta) * pl.cluster;
  pl.max_seg = 1;
  if (pl.cluster == 2) return pl;
  bool want_split = false;
  int per = pl.per;
  if (a->ksplit_hint > 1) {
    per = ceil_div(pl.kb_total, a->ksplit_hint);
    want_split = true;
  } else if (a->ksplit_hint == 0 && allow_split && pl.total_kb >= 2LL * sms) {
    const double dp_eff = (double)pl.n_tiles / ((double)tiles_per_cta * sms);
    const int sk_per = (int)((pl.total_kb + sms - 1) / sms);
    if (!pl.row_mode) {
      // weight streaming: balance bytes across all SMs unless the data-parallel schedule is already balanced
      want_split = dp_eff < 0.95 && sk_per >= 2;
    } else {
      // tensor-bound: partial tiles cost L2 traffic, so split only when few tiles would leave most SMs idle
      want_split = dp_eff < 0.7 && pl.n_tiles < sms && pl.kb_total >= 16 && sk_per >= 4;
    }
    per = sk_per;
  }
  if (want_split && nbatch == 1) {
    const int max_seg = (pl.kb_total + per - 1) / per + 1;
    const size_t need = (size_t)pl.n_tiles * max_seg * pl.BN * BM;
    if (pl.n_tiles < MAX_COUNTERS && need <= ws_floats) {
      pl.per = per;
      pl.grid = (int)((pl.total_kb + per - 1) / per);
      pl.max_seg = max_seg;
    } else if (a->ksplit_hint > 1) {
      pl.grid = -1;  // explicit request that cannot be honoured
    }
  }
  if (!pl.row_mode && swiglu && pl.max_seg == 1) {
    // lanes = features SwiGLU always finishes through the workspace (gate / up rows sit on different lanes)
    const size_t need = (size_t)pl.n_tiles * pl.BN * BM;
    if (need > ws_floats) pl.grid = -1;
  }
  return pl;
}

}  // namespace myr

using namespace myr;

/* debug: every following GEMM launch writes 148 x 6 timestamps at `buf` and advances it (NULL stops tracing) */
extern "C" void myr_gemm_set_trace(void* buf) { set_trace_buffer(buf); }

extern "C" size_t myr_gemm_workspace_bytes(int32_t T, int32_t F, int32_t K) {
  (void)K;
  // counters + partial tiles: a split tile keeps at most ceil(kb / per) + 1 <= 4 partials in the shapes the heuristics pick
  const size_t tiles = (size_t)ceil_div(T > 64 ? T : F, BM) * (size_t)ceil_div(T > 64 ? F : T, T > 64 ? 256 : 64);
  return COUNTER_BYTES + tiles * 4 * 256 * BM * sizeof(float);
}

int myr_gemv_dispatch(const myr_gemm_args* a, cudaStream_t stream, int* counter);     // gemv.cu
int myr_gemv_mt_dispatch(const myr_gemm_args* a, cudaStream_t stream, int* counter);  // gemv_mt.cu
namespace myr {
bool gemm2_eligible(const myr_gemm_args* a, int nbatch);  // gemm2.cu: CTA-pair kernel for the tensor-bound shapes
int gemm2_launch(const myr_gemm_args* a, cudaStream_t stream, int* handled);
}

static int g_gemv = -1;  // < 0: not read from the environment yet
static int gemv_enabled() {
  if (g_gemv < 0) {
    const char* e = getenv("MYR_GEMV");
    g_gemv = (e && e[0] == '0') ? 0 : 1;
  }
  return g_gemv;
}
extern "C" int32_t myr_set_gemv(int32_t enabled) {
  const int old = gemv_enabled();
  g_gemv = enabled ? 1 : 0;
  return old;
}

static bool gemv_eligible(const myr_gemm_args* a, int nbatch) {
  return a->T <= 4 && !a->x_mn_major && !a->w_mn_major && nbatch == 1 && a->scale_cols == 0 && !a->round_acc && !a->alpha_set &&
         a->out_group_rows == 0 && (a->act == MYR_ACT_NONE || a->act == MYR_ACT_SWIGLU) && a->K % 128 == 0 && a->K <= 32768 && a->bn_hint == 0 &&
         a->ksplit_hint == 0;
}

// 5 <= T <= 32 (decode steps of the throughput sweep's batches): weights AND tokens streamed through one ring (gemv_mt.cu)
static bool gemv_mt_eligible(const myr_gemm_args* a, int nbatch) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MYR_GEMV_MT");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on && a->T >= 5 && a->T <= 32 && !a->x_mn_major && !a->w_mn_major && nbatch == 1 && a->scale_cols == 0 && !a->round_acc &&
         !a->alpha_set && a->out_group_rows == 0 && (a->act == MYR_ACT_NONE || a->act == MYR_ACT_SWIGLU) && a->K % 128 == 0 &&
         a->bn_hint == 0 && a->ksplit_hint == 0 && a->norm_ss == nullptr && a->post_out16 == nullptr && a->F >= 1024;
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
  MYR_CHECK_ARG(a->bn_hint >= 0 && a->bn_hint <= 256, "gemm: bn_hint=%d must be in [0,256]", a->bn_hint);
  const bool swiglu = a->act == MYR_ACT_SWIGLU;
  if (swiglu) {
    MYR_CHECK_ARG(a->F % 128 == 0 && a->bias == nullptr && a->res == nullptr && a->out_dtype == MYR_F16 && nbatch == 1 &&
                      !a->w_mn_major && a->scale_cols == 0,
                  "gemm: SwiGLU epilogue needs F %% 128 == 0 (64-row interleaved gate/up), fp16 out, no bias/residual");
  }
  if (gemv_enabled() && gemv_eligible(a, nbatch)) {
    MYR_CHECK_ARG((a->res == nullptr || (a->ldr > 0)) && a->ldo > 0, "gemm: bad leading dimensions");
    // the last arrival counter of the workspace head hands out row groups (zero on entry, left at zero)
    int* ctr = nullptr;
    if (a->workspace != nullptr && a->workspace_bytes >= COUNTER_BYTES && (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0)
      ctr = reinterpret_cast<int*>(a->workspace) + (MAX_COUNTERS - 1);
    return myr_gemv_dispatch(a, stream, ctr);
  }
  if (gemv_enabled() && gemv_mt_eligible(a, nbatch)) {
    MYR_CHECK_ARG((a->res == nullptr || (a->ldr > 0)) && a->ldo > 0, "gemm: bad leading dimensions");
    int* ctr = nullptr;
    if (a->workspace != nullptr && a->workspace_bytes >= COUNTER_BYTES && (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0)
      ctr = reinterpret_cast<int*>(a->workspace) + (MAX_COUNTERS - 1);
    return myr_gemv_mt_dispatch(a, stream, ctr);
  }
  if (a->norm_ss != nullptr || a->post_out16 != nullptr) {
    set_error("gemm: the RMSNorm hand-over exists on the small-batch path only (T <= 4, plain K-major operands, no hints)");
    return MYR_ERR_UNSUPPORTED;
  }

  if (gemm2_eligible(a, nbatch)) {
    int handled = 0;
    const int rc = gemm2_launch(a, stream, &handled);
    if (handled || rc != MYR_OK) return rc;
  }

  size_t ws_floats = 0;
  int* counters = nullptr;
  float* partial = nullptr;
  if (a->workspace != nullptr && a->workspace_bytes > COUNTER_BYTES && (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0) {
    counters = reinterpret_cast<int*>(a->workspace);
    partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + COUNTER_BYTES);
    ws_floats = (a->workspace_bytes - COUNTER_BYTES) / sizeof(float);
  }
  Plan pl = make_plan(a, nbatch, true, ws_floats);
  if (pl.grid < 0) {
    set_error("gemm: split-K / SwiGLU fix-up needs a larger workspace (got %zu bytes)", a->workspace_bytes);
    return MYR_ERR_WORKSPACE;
  }

  // operand roles: A = the 128-row side, B = the BN side
  const void* pa = pl.row_mode ? a->x : a->w;
  const void* pb = pl.row_mode ? a->w : a->x;
  const int64_t lda = pl.row_mode ? a->ldx : a->ldw, ldb = pl.row_mode ? a->ldw : a->ldx;
  const int a_mn = pl.row_mode ? a->x_mn_major : a->w_mn_major, b_mn = pl.row_mode ? a->w_mn_major : a->x_mn_major;
  const int rows_a = pl.row_mode ? a->T : a->F, rows_b = pl.row_mode ? a->F : a->T;
  const int64_t a_bs0 = pl.row_mode ? a->x_bs0 : a->w_bs0, a_bs1 = pl.row_mode ? a->x_bs1 : a->w_bs1;
  const int64_t b_bs0 = pl.row_mode ? a->w_bs0 : a->x_bs0, b_bs1 = pl.row_mode ? a->w_bs1 : a->x_bs1;
  MYR_CHECK_ARG(!b_mn || pl.BN % 64 == 0, "gemm: MN-major operand on the N side needs a tile multiple of 64 (BN=%d)", pl.BN);

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    dims[2] = (uint64_t)nb1; dims[3] = (uint64_t)nb0; box[2] = 1; box[3] = 1;
    if (!a_mn) {
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)rows_a; box[0] = BK; box[1] = BM;
    } else {
      dims[0] = (uint64_t)rows_a; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)lda * 2;
    strides[1] = nb1 > 1 ? (uint64_t)a_bs1 * 2 : strides[0] * dims[1];
    strides[2] = nb0 > 1 ? (uint64_t)a_bs0 * 2 : strides[1] * dims[2];
    int rc = make_tmap_f16(&tmA, pa, 4, dims, strides, box);
    if (rc) return rc;
    if (!b_mn) {
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)rows_b; box[0] = BK; box[1] = (uint32_t)(pl.BN / pl.cluster);
    } else {
      dims[0] = (uint64_t)rows_b; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)ldb * 2;
    strides[1] = nb1 > 1 ? (uint64_t)b_bs1 * 2 : strides[0] * dims[1];
    strides[2] = nb0 > 1 ? (uint64_t)b_bs0 * 2 : strides[1] * dims[2];
    rc = make_tmap_f16(&tmB, pb, 4, dims, strides, box);
    if (rc) return rc;
  }

  GemmKernelParams p;
  p.T = a->T; p.F = a->F; p.K = a->K;
  p.row_mode = pl.row_mode;
  p.BN = pl.BN; p.n_mt = pl.n_mt; p.n_nt = pl.n_nt; p.kb_total = pl.kb_total;
  p.total_kb = pl.total_kb; p.per = pl.per; p.max_seg = pl.max_seg;
  p.num_stages = pl.num_stages; p.stage_bytes = pl.stage_bytes;
  p.a_mn = a_mn; p.b_mn = b_mn;
  p.idesc = make_idesc_f16(BM, pl.BN, a_mn, b_mn);
  p.partial = partial; p.counters = counters;
  p.nb1 = nb1; p.nbatch = nbatch; p.o_bs0 = a->o_bs0; p.o_bs1 = a->o_bs1;
  p.n_mp = pl.n_mp;
  p.trace = next_trace_slot();  // next launch, next slot
  // weights may be prefetched ahead of the dependency only when they are the K-major A operand and the caller says so
  p.w_static = (a->w_static && !pl.row_mode && !a_mn) ? 1 : 0;
  if (pl.row_mode) {
    p.tmem_cols = TMEM_COLS;
    p.acc_stride = ACC_STAGE_COLS;
  } else {
    int cols = 32;
    while (cols < 2 * pl.BN) cols <<= 1;
    p.tmem_cols = cols;
    p.acc_stride = cols / 2;
  }
  p.ep.bias = reinterpret_cast<const __half*>(a->bias);
  p.ep.act = a->act; p.ep.round_acc = a->round_acc;
  p.ep.scale_cols = a->scale_cols; p.ep.scale = a->scale;
  p.ep.res = a->res; p.ep.res_dtype = a->res_dtype; p.ep.ldr = a->ldr;
  p.ep.out = a->out; p.ep.out_dtype = a->out_dtype; p.ep.ldo = a->ldo;
  p.ep.group_rows = a->out_group_rows; p.ep.group_stride = a->out_group_stride;
  p.ep.alpha = a->alpha_set ? a->alpha : 1.0f;
  {
    const int oq = a->out_dtype == MYR_F32 ? 4 : 8, rq = a->res_dtype == MYR_F32 ? 4 : 8;
    bool v = (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->ldo % oq == 0 && a->out_group_stride % oq == 0 &&
             a->o_bs0 % oq == 0 && a->o_bs1 % oq == 0;
    if (a->res) v = v && (reinterpret_cast<uintptr_t>(a->res) & 15) == 0 && a->ldr % rq == 0;
    if (a->bias) v = v && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
    p.ep.vec = v ? 1 : 0;
  }

  const size_t smem_bytes = (size_t)pl.num_stages * pl.stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr_set = true;
  }
  if (pl.cluster == 2)
    MYR_CHECK_CUDA(launch_kernel_cluster(gemm_tc_kernel<1, 2>, dim3((unsigned)pl.grid), dim3(GEMM_THREADS), smem_bytes, stream,
                                         a->pdl != 0, 2, tmA, tmB, p));
  else if (pl.row_mode)
    MYR_CHECK_CUDA(launch_kernel(gemm_tc_kernel<1>, dim3((unsigned)pl.grid), dim3(GEMM_THREADS), smem_bytes, stream, a->pdl != 0,
                                 tmA, tmB, p));
  else
    MYR_CHECK_CUDA(launch_kernel(gemm_tc_kernel<2>, dim3((unsigned)pl.grid), dim3(GEMM_THREADS), smem_bytes, stream, a->pdl != 0,
                                 tmA, tmB, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
