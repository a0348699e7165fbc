// Fused attention backward for SHORT sequences (the training step: LLaMA S = 164 causal with per-row visible lengths, Q-Former
// self-attention over its 81 + text rows): dQ, dK, dV from Q, K, V, dO in ONE launch.
//
// Reference semantics: the autograd backward of softmax(Q K^T * scale + mask) V as modeling_llama.py:196-222 and
// Qformer.py:180-260 compute it eagerly. training.py::_attn_bwd ran it as five batched GEMM launches and two row kernels with
// the S x S scores in HBM (~100 us per LLaMA layer for 4.4 GFLOP: every launch is latency-bound); here the scores never leave
// the SM:
//   * one CTA per (row b, head h, 64-wide slice of the head dimension): K and V of the head stay in shared memory, the queries
//     walk through in tiles of 16 rows;
//   * per query tile (mma.sync m16n8k16, fp16 operands, fp32 accumulation - 82 MFLOP per head, far too little for a tcgen05
//     pipeline to amortise its set-up): S = Q K^T and dP = dO V^T over the FULL head dimension (both dim slices recompute them:
//     cheaper than exchanging them), masked softmax with the same rounding points as the row kernels it replaces (P rounded to
//     fp16 before it is used; dS = scale * P * (dP - sum_j dP P) rounded to fp16), then dQ[:, slice] = dS K[:, slice] (stored),
//     dV[:, slice] += P^T dO[:, slice] and dK[:, slice] += dS^T Q[:, slice] (registers, stored after the last tile);
//   * every reduction has a fixed order: deterministic.
// Limits: dh in {64, 128}, Skv <= 256 (two 16-key accumulator tiles per warp); longer caches keep the GEMM path.
#include "common.h"
#include "ptx.cuh"

namespace myr {

struct AttnBwdParams {
  const __half *q, *k, *v, *dO;
  __half *dq, *dk, *dv;
  long long q_ts, q_bs, k_ts, k_bs, v_ts, v_bs, do_ts, do_bs, dq_ts, dq_bs, dk_ts, dk_bs, dv_ts, dv_bs;
  int B, H, Sq, Skv, SkP;
  float scale;
  int causal;
  const int* kv_len;
};

constexpr int AB_QT = 16;       // query rows per tile
constexpr int AB_THREADS = 256;  // 8 warps
constexpr int AB_DC = 64;        // head-dimension slice of a CTA's outputs

__device__ __forceinline__ void ab_ldm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ab_ldm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ab_ldm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ab_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int DH>
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_small_kernel(const AttnBwdParams p) {
  constexpr int LD = DH + 8;  // halfs per staged row: the 16-byte pad keeps ldmatrix rows on distinct banks
  extern __shared__ __align__(16) uint8_t ab_smem[];
  const int SkP = p.SkP, LDP = SkP + 8;
  __half* sK = reinterpret_cast<__half*>(ab_smem);  // [SkP][LD]
  __half* sV = sK + (size_t)SkP * LD;               // [SkP][LD]
  __half* sQ = sV + (size_t)SkP * LD;               // [2][16][LD]  (double-buffered: the next tile arrives by cp.async)
  __half* sdO = sQ + 2 * AB_QT * LD;                // [2][16][LD]
  __half* sP = sdO + 2 * AB_QT * LD;                // [16][LDP]
  __half* sdS = sP + AB_QT * LDP;                   // [16][LDP]
  __shared__ float s_max[AB_QT][8], s_sum[AB_QT][8], s_del[AB_QT][8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;  // accumulator fragment: rows g, g + 8; columns 2 tq, 2 tq + 1
  const int c0 = blockIdx.x * AB_DC, h = blockIdx.y, b = blockIdx.z;
  const int Sq = p.Sq, Skv = p.Skv;
  const int kvl = p.kv_len ? min(p.kv_len[b], Skv) : Skv;

  const __half* gq = p.q + (size_t)b * p.q_bs + (size_t)h * DH;
  const __half* gk = p.k + (size_t)b * p.k_bs + (size_t)h * DH;
  const __half* gv = p.v + (size_t)b * p.v_bs + (size_t)h * DH;
  const __half* gdo = p.dO + (size_t)b * p.do_bs + (size_t)h * DH;

  // ---- K and V of this head (rows beyond Skv zero)
  {
    constexpr int UPR = DH / 8;  // 16-byte units per row
    for (int i = tid; i < SkP * UPR; i += AB_THREADS) {
      const int r = i / UPR, u = i - r * UPR;
      uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
      if (r < Skv) {
        kk = *reinterpret_cast<const uint4*>(gk + (size_t)r * p.k_ts + u * 8);
        vv = *reinterpret_cast<const uint4*>(gv + (size_t)r * p.v_ts + u * 8);
      }
      *reinterpret_cast<uint4*>(sK + (size_t)r * LD + u * 8) = kk;
      *reinterpret_cast<uint4*>(sV + (size_t)r * LD + u * 8) = vv;
    }
  }

  // 8-key score tiles are dealt round-robin to the warps (<= 4 each), 16-key accumulator tiles of dK / dV too (<= 2 each)
  float accV[2][8][4], accK[2][8][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) accV[m][n][e] = accK[m][n][e] = 0.f;

  const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_base = smem_u32(sQ), sdO_base = smem_u32(sdO), sP_a = smem_u32(sP),
                 sdS_a = smem_u32(sdS);
  // ldmatrix lane roles
  const int l_r16 = lane & 15, l_c8 = (lane >> 4) << 3;  // A operand (row-major [16][k]): row, k offset
  const int l_r8 = lane & 7, l_q = lane >> 3;             // 8 x 8 matrix index q = 0 .. 3 and the row inside it

  // query tile -> shared memory by 16-byte asynchronous copies (rows beyond Sq: zero fill)
  auto load_tile = [&](int i0, int buf) {
    constexpr int UPR = DH / 8;
    for (int i = tid; i < AB_QT * UPR; i += AB_THREADS) {
      const int r = i / UPR, u = i - r * UPR;
      const int row = min(i0 + r, Sq - 1);
      const uint32_t n = i0 + r < Sq ? 16u : 0u;
      const uint32_t dq_ = smem_u32(sQ + (buf * AB_QT + r) * LD + u * 8), dd_ = smem_u32(sdO + (buf * AB_QT + r) * LD + u * 8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dq_), "l"(gq + (size_t)row * p.q_ts + u * 8), "r"(n) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dd_), "l"(gdo + (size_t)row * p.do_ts + u * 8), "r"(n) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_tile(0, 0);
  const int kv16 = (kvl + 15) & ~15;  // keys beyond the visible length are masked for every query

  for (int i0 = 0, buf = 0; i0 < Sq; i0 += AB_QT, buf ^= 1) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // this tile has landed (first pass: K / V too); the previous tile's readers of sQ / sdO / sP / sdS are done
    if (i0 + AB_QT < Sq) load_tile(i0 + AB_QT, buf ^ 1);
    const uint32_t sQ_a = sQ_base + buf * AB_QT * LD * 2, sdO_a = sdO_base + buf * AB_QT * LD * 2;
    // keys this tile can see at all: causal rows end at their own position (Sq == Skv), every row at the visible length
    const int kmax = min(kv16, p.causal ? min(SkP, i0 + AB_QT) : SkP);
    const int n_nt = kmax >> 3, n_mt = kmax >> 4;

    // ---- S = Q K^T, dP = dO V^T for this warp's key tiles
    float s[4][4], dp[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[n][e] = dp[n][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ks += 2) {
      uint32_t aq0[4], aq1[4], ad0[4], ad1[4];
      ab_ldm_x4(sQ_a + (l_r16 * LD + ks * 16 + l_c8) * 2, aq0);
      ab_ldm_x4(sQ_a + (l_r16 * LD + ks * 16 + 16 + l_c8) * 2, aq1);
      ab_ldm_x4(sdO_a + (l_r16 * LD + ks * 16 + l_c8) * 2, ad0);
      ab_ldm_x4(sdO_a + (l_r16 * LD + ks * 16 + 16 + l_c8) * 2, ad1);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int nt = warp + 8 * n;
        if (nt < n_nt) {
          // B operand [k][n] lives as K[n][k]: four 8 x 8 matrices = (keys, k 0-7) (keys, k 8-15) (keys, k 16-23) (keys, k 24-31)
          uint32_t bk[4], bv[4];
          const uint32_t off = ((nt * 8 + l_r8) * LD + ks * 16 + l_q * 8) * 2;
          ab_ldm_x4(sK_a + off, bk);
          ab_ldm_x4(sV_a + off, bv);
          ab_mma(s[n], aq0, bk[0], bk[1]);
          ab_mma(s[n], aq1, bk[2], bk[3]);
          ab_mma(dp[n], ad0, bv[0], bv[1]);
          ab_mma(dp[n], ad1, bv[2], bv[3]);
        }
      }
    }

    // ---- masked softmax over the keys (rows g and g + 8 of the tile), fixed-order reductions through shared memory
    const int ra = i0 + g, rb = i0 + g + 8;
    const int lim_a = p.causal ? min(kvl, ra + 1) : kvl, lim_b = p.causal ? min(kvl, rb + 1) : kvl;
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int j = (warp + 8 * n) * 8 + 2 * tq;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] = (j + e < lim_a) ? s[n][e] * p.scale : -INFINITY;
        s[n][2 + e] = (j + e < lim_b) ? s[n][2 + e] * p.scale : -INFINITY;
        ma = fmaxf(ma, s[n][e]);
        mb = fmaxf(mb, s[n][2 + e]);
      }
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    if (tq == 0) {
      s_max[g][warp] = ma;
      s_max[g + 8][warp] = mb;
    }
    __syncthreads();
    ma = mb = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      ma = fmaxf(ma, s_max[g][w]);
      mb = fmaxf(mb, s_max[g + 8][w]);
    }
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] = s[n][e] > -INFINITY ? __expf(s[n][e] - ma) : 0.f;
        s[n][2 + e] = s[n][2 + e] > -INFINITY ? __expf(s[n][2 + e] - mb) : 0.f;
        sa += s[n][e];
        sb += s[n][2 + e];
      }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1);
    sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb += __shfl_xor_sync(0xffffffffu, sb, 1);
    sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    if (tq == 0) {
      s_sum[g][warp] = sa;
      s_sum[g + 8][warp] = sb;
    }
    __syncthreads();
    sa = sb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      sa += s_sum[g][w];
      sb += s_sum[g + 8][w];
    }
    const float inv_a = sa > 0.f ? 1.f / sa : 0.f, inv_b = sb > 0.f ? 1.f / sb : 0.f;
    float da = 0.f, db = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] = round_f16(s[n][e] * inv_a);  // P as the MMAs below see it
        s[n][2 + e] = round_f16(s[n][2 + e] * inv_b);
        if (s[n][e] != 0.f) da = fmaf(s[n][e], dp[n][e], da);
        if (s[n][2 + e] != 0.f) db = fmaf(s[n][2 + e], dp[n][2 + e], db);
      }
    da += __shfl_xor_sync(0xffffffffu, da, 1);
    da += __shfl_xor_sync(0xffffffffu, da, 2);
    db += __shfl_xor_sync(0xffffffffu, db, 1);
    db += __shfl_xor_sync(0xffffffffu, db, 2);
    if (tq == 0) {
      s_del[g][warp] = da;
      s_del[g + 8][warp] = db;
    }
    __syncthreads();
    da = db = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      da += s_del[g][w];
      db += s_del[g + 8][w];
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int nt = warp + 8 * n;
      if (nt < n_nt) {
        const int j = nt * 8 + 2 * tq;
        float ds[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          ds[e] = s[n][e] != 0.f ? p.scale * s[n][e] * (dp[n][e] - da) : 0.f;
          ds[2 + e] = s[n][2 + e] != 0.f ? p.scale * s[n][2 + e] * (dp[n][2 + e] - db) : 0.f;
        }
        *reinterpret_cast<__half2*>(sP + g * LDP + j) = __floats2half2_rn(s[n][0], s[n][1]);
        *reinterpret_cast<__half2*>(sP + (g + 8) * LDP + j) = __floats2half2_rn(s[n][2], s[n][3]);
        *reinterpret_cast<__half2*>(sdS + g * LDP + j) = __floats2half2_rn(ds[0], ds[1]);
        *reinterpret_cast<__half2*>(sdS + (g + 8) * LDP + j) = __floats2half2_rn(ds[2], ds[3]);
      }
    }
    __syncthreads();

    // ---- dQ[tile, c0 + 8 warp ..] = dS K[:, slice]: A = dS [16][keys], B[k = key][n = dim] is K itself ([k][n] row-major: .trans)
    {
      float dq[4] = {0.f, 0.f, 0.f, 0.f};
      const int dcol = c0 + warp * 8;
      for (int ks = 0; ks < n_mt; ++ks) {
        uint32_t a[4], bb[2];
        ab_ldm_x4(sdS_a + (l_r16 * LDP + ks * 16 + l_c8) * 2, a);
        ab_ldm_x2_t(sK_a + ((ks * 16 + (lane & 15)) * LD + dcol) * 2, bb);
        ab_mma(dq, a, bb[0], bb[1]);
      }
      __half* o = p.dq + (size_t)b * p.dq_bs + (size_t)h * DH + dcol + 2 * tq;
      if (ra < Sq) *reinterpret_cast<__half2*>(o + (size_t)ra * p.dq_ts) = __floats2half2_rn(dq[0], dq[1]);
      if (rb < Sq) *reinterpret_cast<__half2*>(o + (size_t)rb * p.dq_ts) = __floats2half2_rn(dq[2], dq[3]);
    }

    // ---- dV[keys, slice] += P^T dO[:, slice]; dK[keys, slice] += dS^T Q[:, slice]   (K = the 16 query rows of the tile)
    {
      // B operands [k = query][n = dim]: dO / Q rows as staged ([k][n] row-major: .trans), two 8-dim tiles per ldmatrix.x4
      uint32_t bo[4][4], bq[4][4];
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        const uint32_t off = (((l_q & 1) * 8 + l_r8) * LD + c0 + n2 * 16 + (l_q >> 1) * 8) * 2;
        ab_ldm_x4_t(sdO_a + off, bo[n2]);
        ab_ldm_x4_t(sQ_a + off, bq[n2]);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int mt = warp + 8 * m;
        if (mt < n_mt) {
          // A operands [m = key][k = query] live as P[query][key] / dS[query][key] ([k][m] row-major: .trans);
          // matrices: (m 0-7, k 0-7) (m 8-15, k 0-7) (m 0-7, k 8-15) (m 8-15, k 8-15)
          uint32_t ap[4], as[4];
          const uint32_t off = (((l_q >> 1) * 8 + l_r8) * LDP + mt * 16 + (l_q & 1) * 8) * 2;
          ab_ldm_x4_t(sP_a + off, ap);
          ab_ldm_x4_t(sdS_a + off, as);
#pragma unroll
          for (int n2 = 0; n2 < 4; ++n2) {
            ab_mma(accV[m][2 * n2], ap, bo[n2][0], bo[n2][1]);
            ab_mma(accV[m][2 * n2 + 1], ap, bo[n2][2], bo[n2][3]);
            ab_mma(accK[m][2 * n2], as, bq[n2][0], bq[n2][1]);
            ab_mma(accK[m][2 * n2 + 1], as, bq[n2][2], bq[n2][3]);
          }
        }
      }
    }
  }

  // ---- dK, dV of this head slice
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int mt = warp + 8 * m;
    if (mt >= (SkP >> 4)) continue;
    const int ka = mt * 16 + g, kb = ka + 8;
    __half* ok = p.dk + (size_t)b * p.dk_bs + (size_t)h * DH + c0 + 2 * tq;
    __half* ov = p.dv + (size_t)b * p.dv_bs + (size_t)h * DH + c0 + 2 * tq;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (ka < Skv) {
        *reinterpret_cast<__half2*>(ok + (size_t)ka * p.dk_ts + n * 8) = __floats2half2_rn(accK[m][n][0], accK[m][n][1]);
        *reinterpret_cast<__half2*>(ov + (size_t)ka * p.dv_ts + n * 8) = __floats2half2_rn(accV[m][n][0], accV[m][n][1]);
      }
      if (kb < Skv) {
        *reinterpret_cast<__half2*>(ok + (size_t)kb * p.dk_ts + n * 8) = __floats2half2_rn(accK[m][n][2], accK[m][n][3]);
        *reinterpret_cast<__half2*>(ov + (size_t)kb * p.dv_ts + n * 8) = __floats2half2_rn(accV[m][n][2], accV[m][n][3]);
      }
    }
  }
}

static size_t ab_smem_bytes(int dh, int SkP) {
  const size_t LD = dh + 8, LDP = SkP + 8;
  return 2 * ((size_t)2 * SkP * LD + 4 * AB_QT * LD + 2 * AB_QT * LDP);
}

}  // namespace myr

using namespace myr;

extern "C" int myr_attn_bwd_small_supported(int32_t Sq, int32_t Skv, int32_t dh) {
  return (dh == 64 || dh == 128) && Sq > 0 && Skv > 0 && Skv <= 256 ? 1 : 0;
}

extern "C" int myr_attn_bwd_small(const void* q, int64_t q_ts, int64_t q_bs, const void* k, int64_t k_ts, int64_t k_bs, const void* v,
                                  int64_t v_ts, int64_t v_bs, const void* dO, int64_t do_ts, int64_t do_bs, void* dq, int64_t dq_ts,
                                  int64_t dq_bs, void* dk, int64_t dk_ts, int64_t dk_bs, void* dv, int64_t dv_ts, int64_t dv_bs,
                                  int32_t B, int32_t H, int32_t Sq, int32_t Skv, int32_t dh, float scale, int32_t causal,
                                  const void* kv_len, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(q && k && v && dO && dq && dk && dv && B > 0 && H > 0, "attn_bwd_small: bad arguments");
  MYR_CHECK_ARG(myr_attn_bwd_small_supported(Sq, Skv, dh), "attn_bwd_small: dh must be 64 or 128 and Skv <= 256 (got dh %d, Skv %d)", dh, Skv);
  MYR_CHECK_ARG(!causal || Sq == Skv, "attn_bwd_small: the causal mask assumes Sq == Skv");
  const int64_t strides[] = {q_ts, q_bs, k_ts, k_bs, v_ts, v_bs, do_ts, do_bs, dq_ts, dq_bs, dk_ts, dk_bs, dv_ts, dv_bs};
  for (int64_t s : strides) MYR_CHECK_ARG(s % 8 == 0, "attn_bwd_small: strides must be multiples of 8 elements");
  const void* ptrs[] = {q, k, v, dO, dq, dk, dv};
  for (const void* ptr : ptrs) MYR_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "attn_bwd_small: operands must be 16-byte aligned");
  AttnBwdParams p;
  p.q = reinterpret_cast<const __half*>(q); p.k = reinterpret_cast<const __half*>(k); p.v = reinterpret_cast<const __half*>(v);
  p.dO = reinterpret_cast<const __half*>(dO);
  p.dq = reinterpret_cast<__half*>(dq); p.dk = reinterpret_cast<__half*>(dk); p.dv = reinterpret_cast<__half*>(dv);
  p.q_ts = q_ts; p.q_bs = q_bs; p.k_ts = k_ts; p.k_bs = k_bs; p.v_ts = v_ts; p.v_bs = v_bs; p.do_ts = do_ts; p.do_bs = do_bs;
  p.dq_ts = dq_ts; p.dq_bs = dq_bs; p.dk_ts = dk_ts; p.dk_bs = dk_bs; p.dv_ts = dv_ts; p.dv_bs = dv_bs;
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.SkP = (Skv + 15) / 16 * 16;
  p.scale = scale; p.causal = causal; p.kv_len = reinterpret_cast<const int*>(kv_len);
  const size_t smem = ab_smem_bytes(dh, p.SkP);
  static bool attr = false;
  if (!attr) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const dim3 grid((unsigned)(dh / AB_DC), (unsigned)H, (unsigned)B);
  if (dh == 64)
    attn_bwd_small_kernel<64><<<grid, AB_THREADS, smem, stream>>>(p);
  else
    attn_bwd_small_kernel<128><<<grid, AB_THREADS, smem, stream>>>(p);
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
