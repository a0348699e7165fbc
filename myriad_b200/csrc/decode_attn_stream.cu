// Decode-step attention over LONG caches (the throughput sweep's S = 1024 / 2048 and its batches of 16 / 32 rows): HBM-bound.
//
// One new token per sequence reads the whole K / V cache of its row once per layer: 2 * B * S * 8 KiB (4.3 GB per step at
// B = 4, S = 2048; 34 GB at B = 32), more than the weights. decode_attn.cu's kernels give one CTA one (head, row): 128 CTAs
// at B = 4, each walking its 2 x S x 256 bytes through three dependent memory round trips (1.3 TB/s measured). Here:
//   * the work list is FLAT: the 128-key chunks of all (row, head) pairs in (row, head, chunk) order, N chunks in total (N
//     follows the rows' visible lengths, read from the device step state); a persistent grid of min(#SMs, N) CTAs cuts it into
//     equal contiguous ranges, so every SM streams the same number of bytes whatever B, H and the lengths are;
//   * a producer thread walks the range and keeps a ring of 3 chunks (K and V of 128 keys = 64 KiB by four 128-byte-swizzled
//     TMA boxes; the last chunk of a row only its visible keys, in boxes of 32) in flight, across (row, head) boundaries and
//     ahead of the programmatic-dependent-launch wait (every cached row was written at least one decode step ago);
//   * four consumer warps run an online softmax over the chunks of a (row, head) segment: q.k on the tensor cores (mma.sync
//     m16n8k16 straight from the swizzled K tile by ldmatrix, q as a one-column B operand: 16 MMAs per warp and chunk instead of
//     128 FFMAs + 128 conversions per thread - with the scalar loop the consumers, not HBM, set the pace), running maximum M with
//     exp(M_old - M_new) rescaling of the per-warp P.V accumulators and per-lane partial sums - two 128-thread barriers per chunk,
//     all reductions in a fixed order;
//   * a segment that holds all chunks of its (row, head) writes the output row; otherwise it leaves (M, L, O[128]) in the
//     workspace and the LAST segment of that (row, head) to arrive (one atomic per segment) combines them in segment order -
//     deterministic: CUDA-graph replay == eager launches.
// The new token's q / k / v (LoRA-B + RoPE, peft + modeling_llama.py:109-123) are computed at every segment start (a few
// loads); the segment whose range holds the token's cache slot appends k / v (the torch.cat of modeling_llama.py:190-195) and
// patches that row into the staged tile (the producer may have fetched the chunk before the append). Arithmetic: fp32 softmax (modeling_llama.py:214), fp32 accumulation.
#include "decode_attn.cuh"

namespace myr {

constexpr int DS_CH = 128;                   // keys per chunk
constexpr int DS_TILE = DS_CH * 128;         // one [128 keys][64 halfs] box: 16 KiB
constexpr int DS_STAGE = 4 * DS_TILE;        // K lo / K hi / V lo / V hi
constexpr int DS_STAGES = 3;
constexpr int DS_THREADS = DA_THREADS + 32;  // 4 consumer warps + producer warp
constexpr int DS_PART = 2 + DA_DH;           // floats per partial result

struct DecodeAttnStreamParams {
  DecodeAttnParams a;
  float* part;    // [B][H][max_parts][DS_PART]
  int* counters;  // [B][H], zero on entry, left at zero
  int max_parts;
};

__device__ __forceinline__ void ds_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DA_THREADS) : "memory"); }

// flat chunk index -> (row, head, chunk); nc[b] = chunks of row b. `base` = flat index of (b, head 0, chunk 0).
struct DsPos {
  int b, h, c, base;
};
__device__ __forceinline__ DsPos ds_locate(int idx, const int* nc, int B, int H) {
  DsPos r;
  int base = 0, b = 0;
  while (b + 1 < B && idx >= base + H * nc[b]) {
    base += H * nc[b];
    ++b;
  }
  r.b = b;
  r.base = base;
  const int rem = idx - base;
  r.h = rem / nc[b];
  r.c = rem - r.h * nc[b];
  return r;
}

struct DsMaps {
  CUtensorMap k, v;      // boxes of 128 keys x 64 halfs
  CUtensorMap k32, v32;  // boxes of 32 keys x 64 halfs (tail chunks)
};

__global__ void __launch_bounds__(DS_THREADS) decode_attn_stream_kernel(const __grid_constant__ DsMaps tm,
                                                                        const DecodeAttnStreamParams pp) {
  const DecodeAttnParams& p = pp.a;
  extern __shared__ uint8_t ds_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ds_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ DecodeAttnSmem sm;
  __shared__ float s_p[2][DS_CH];
  __shared__ float s_red[2][4];
  __shared__ int s_nc[128];  // chunks per row
  __shared__ int s_kvl[128]; // visible keys per row
  __shared__ int s_pos[128]; // rotary position of the new token per row
  __shared__ uint64_t full[DS_STAGES], empty[DS_STAGES];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HD = p.H * DA_DH;

  if (tid == 0) {
    for (int s = 0; s < DS_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);
    }
    fence_mbar_init();
  }
  // step state (cache slot, visible lengths) is written by the greedy-update kernel at the END of the previous decode step,
  // i.e. before the kernels in front of this one could start: safe to read ahead of the wait
  const int off_g = p.cache_off ? __ldcg(p.cache_off) : p.cache_off_host;
  if (tid < p.B) {
    int kvl = p.kv_len ? __ldcg(p.kv_len + tid) : off_g + 1;
    kvl = min(max(kvl, 1), p.Smax);
    s_kvl[tid] = kvl;
    s_nc[tid] = (kvl + DS_CH - 1) / DS_CH;
    s_pos[tid] = __ldcg(p.pos + tid);
  }
  __syncthreads();
  pdl_launch_dependents();
  int N = 0;
  for (int b = 0; b < p.B; ++b) N += p.H * s_nc[b];
  const int G = min((int)gridDim.x, N);
  if ((int)blockIdx.x >= G) return;
  const int lo = (int)(((long long)blockIdx.x * N) / G), hi = (int)(((long long)(blockIdx.x + 1) * N) / G);

  if (warp == 4) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      DsPos q = ds_locate(lo, s_nc, p.B, p.H);
      int st = 0;
      uint32_t phase = 0;
      for (int idx = lo; idx < hi; ++idx) {
        mbar_wait(&empty[st], phase ^ 1);
        uint8_t* dst = ring + st * DS_STAGE;
        const int n = s_kvl[q.b] - q.c * DS_CH;  // visible keys from this chunk's first slot on
        if (n > 96) {
          mbar_arrive_expect_tx(&full[st], (uint32_t)DS_STAGE);
          for (int hf = 0; hf < 2; ++hf) {
            tma_load_4d(dst + hf * DS_TILE, &tm.k, &full[st], hf * 64, q.h, q.c * DS_CH, q.b);
            tma_load_4d(dst + (2 + hf) * DS_TILE, &tm.v, &full[st], hf * 64, q.h, q.c * DS_CH, q.b);
          }
        } else {
          // tail of a row: only the 32-key boxes that hold visible keys (the tile layout [key][128 B] is the same)
          const int nb = (n + 31) >> 5;
          mbar_arrive_expect_tx(&full[st], (uint32_t)(nb * 4 * 32 * 128));
          for (int i = 0; i < nb; ++i)
            for (int hf = 0; hf < 2; ++hf) {
              tma_load_4d(dst + hf * DS_TILE + i * 4096, &tm.k32, &full[st], hf * 64, q.h, q.c * DS_CH + i * 32, q.b);
              tma_load_4d(dst + (2 + hf) * DS_TILE + i * 4096, &tm.v32, &full[st], hf * 64, q.h, q.c * DS_CH + i * 32, q.b);
            }
        }
        if (++st == DS_STAGES) {
          st = 0;
          phase ^= 1;
        }
        if (++q.c == s_nc[q.b]) {  // next (row, head)
          q.c = 0;
          if (++q.h == p.H) {
            q.h = 0;
            ++q.b;
          }
        }
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  DsPos q = ds_locate(lo, s_nc, p.B, p.H);
  int st = 0;
  uint32_t phase = 0;
  int idx = lo;
  // inputs of a segment's prologue (threads 0 .. 63: one rotary pair each), requested one segment ahead
  struct {
    float cs, sn;
    uint4 lb[4], xq, xv;
    __half q1, q2, k1, k2, v1, v2;
  } pro = {};
  auto pro_load = [&](int b, int h) {
    if (tid >= DA_DH / 2) return;
    const int pos = s_pos[b], j = tid, half = DA_DH / 2;
    pro.cs = p.cos_t[(size_t)pos * half + j];
    pro.sn = p.sin_t[(size_t)pos * half + j];
    const __half* row = p.qkv + (size_t)b * p.ldq;
    if (p.lora_r) {
      pro.lb[0] = __ldg(reinterpret_cast<const uint4*>(p.lora_bq + (size_t)(h * DA_DH + j) * 8));
      pro.lb[1] = __ldg(reinterpret_cast<const uint4*>(p.lora_bq + (size_t)(h * DA_DH + half + j) * 8));
      pro.lb[2] = __ldg(reinterpret_cast<const uint4*>(p.lora_bv + (size_t)(h * DA_DH + j) * 8));
      pro.lb[3] = __ldg(reinterpret_cast<const uint4*>(p.lora_bv + (size_t)(h * DA_DH + half + j) * 8));
      pro.xq = __ldcg(reinterpret_cast<const uint4*>(row + 3 * HD));
      pro.xv = __ldcg(reinterpret_cast<const uint4*>(row + 3 * HD + 8));
    }
    pro.q1 = __ldcg(row + h * DA_DH + j);
    pro.q2 = __ldcg(row + h * DA_DH + half + j);
    pro.k1 = __ldcg(row + HD + h * DA_DH + j);
    pro.k2 = __ldcg(row + HD + h * DA_DH + half + j);
    pro.v1 = __ldcg(row + 2 * HD + h * DA_DH + j);
    pro.v2 = __ldcg(row + 2 * HD + h * DA_DH + half + j);
  };
  pdl_wait();  // qkv of this step
  pro_load(q.b, q.h);
  while (idx < hi) {
    // ---- segment: chunks [c0, c1) of (row b, head h)
    const int b = q.b, h = q.h, c0 = q.c;
    const int ncb = s_nc[b];
    const int c1 = min(ncb, c0 + (hi - idx));
    const int kvl = s_kvl[b];
    __half* kbase = p.kcache + (size_t)b * p.c_bs + h * DA_DH;
    __half* vbase = p.vcache + (size_t)b * p.c_bs + h * DA_DH;
    const bool owns_new = off_g >= c0 * DS_CH && off_g < c1 * DS_CH && off_g < p.Smax;
    {
      // the new token's q / k / v: LoRA, rotation (one rotary pair per thread), cache append by the owning segment. Its inputs
      // (`pro`) were requested one segment ago (below), so no memory round trip sits between two segments.
      ds_sync();  // the previous segment is done with sm.q / sm.k / sm.v / sm.acc
      if (tid < DA_DH / 2) {
        const int j = tid, half = DA_DH / 2;
        const float cs = pro.cs, sn = pro.sn;
        float q1 = __half2float(pro.q1), q2 = __half2float(pro.q2);
        const float k1 = __half2float(pro.k1), k2 = __half2float(pro.k2);
        float v1 = __half2float(pro.v1), v2 = __half2float(pro.v2);
        if (p.lora_r) {
          float xq[8], xv[8];
          da_unpack8(pro.xq, xq);
          da_unpack8(pro.xv, xv);
          auto dot8 = [](const uint4& wrow, const float (&xa)[8]) {  // same order of operations as da_lora_dot
            float w[8];
            da_unpack8(wrow, w);
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) acc = fmaf(w[r], xa[r], acc);
            return acc;
          };
          q1 = fmaf(p.lora_scale, dot8(pro.lb[0], xq), q1);
          q2 = fmaf(p.lora_scale, dot8(pro.lb[1], xq), q2);
          v1 = fmaf(p.lora_scale, dot8(pro.lb[2], xv), v1);
          v2 = fmaf(p.lora_scale, dot8(pro.lb[3], xv), v2);
        }
        // fp16 rounding of the rotated q / k and of v: the values the prefill path stores (q in place, k / v in the cache)
        sm.q[j] = round_f16(q1 * cs - q2 * sn);
        sm.q[half + j] = round_f16(q2 * cs + q1 * sn);
        const __half ko1 = __float2half_rn(k1 * cs - k2 * sn), ko2 = __float2half_rn(k2 * cs + k1 * sn);
        const __half vo1 = __float2half_rn(v1), vo2 = __float2half_rn(v2);
        sm.k[j] = ko1; sm.k[half + j] = ko2;
        sm.v[j] = vo1; sm.v[half + j] = vo2;
        if (owns_new) {
          __half* kd = kbase + (size_t)off_g * p.c_ts;
          __half* vd = vbase + (size_t)off_g * p.c_ts;
          kd[j] = ko1; kd[half + j] = ko2;
          vd[j] = vo1; vd[half + j] = vo2;
        }
      }
      ds_sync();
      // request the NEXT segment's inputs now: they arrive under this segment's chunks
      if (c1 == ncb && idx + (c1 - c0) < hi) {
        const bool wrap = h + 1 == p.H;
        pro_load(wrap ? b + 1 : b, wrap ? 0 : h + 1);
      }
    }
    // q as the B operand of the score MMAs (mma.sync.m16n8k16: A = 16 keys x 16 dims of the K tile, B = 16 dims x 8 columns of which
    // only column 0 - lanes 0 .. 3 - carries q; sm.q holds fp16-representable values, so the cast is exact)
    uint32_t qb[DA_DH / 16][2];
#pragma unroll
    for (int ks = 0; ks < DA_DH / 16; ++ks) {
      qb[ks][0] = qb[ks][1] = 0u;
      if (lane < 4) {
        const __half2 lo = __floats2half2_rn(sm.q[ks * 16 + 2 * lane], sm.q[ks * 16 + 2 * lane + 1]);
        const __half2 hi = __floats2half2_rn(sm.q[ks * 16 + 8 + 2 * lane], sm.q[ks * 16 + 8 + 2 * lane + 1]);
        qb[ks][0] = *reinterpret_cast<const uint32_t*>(&lo);
        qb[ks][1] = *reinterpret_cast<const uint32_t*>(&hi);
      }
    }
    float M = -INFINITY, l_t = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int v_half = lane >> 4, v_unit = (lane & 15) >> 1, v_sub = (lane & 1) * 8;
    for (int c = c0; c < c1; ++c, ++idx) {
      const int par = c & 1;
      const int j_lo = c * DS_CH;
      const int n = min(DS_CH, kvl - j_lo);
      const int off = off_g - j_lo;  // local slot of the new token (outside [0, n): not in this chunk)
      mbar_wait(&full[st], phase);
      const uint8_t* sK = ring + st * DS_STAGE;
      const uint8_t* sV = sK + 2 * DS_TILE;
      // the new token's K / V row is not in the tile (it was appended after the producer ran ahead): patch it in
      if (off >= 0 && off < n) {
        if (tid < 32) {
          const int u = tid & 15;
          const uint4 val = tid < 16 ? reinterpret_cast<const uint4*>(sm.k)[u] : reinterpret_cast<const uint4*>(sm.v)[u];
          uint8_t* dst = const_cast<uint8_t*>(tid < 16 ? sK : sV) + (u >> 3) * DS_TILE + off * 128 + (((u & 7) ^ (off & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = val;
        }
        ds_sync();
      }
      // ---- scores on the tensor cores: warp w takes keys [32 w, 32 w + 32) as two 16-key tiles; the results land in the lanes
      // with lane % 4 == 0 (fragment column 0): keys 32 w + 16 t + lane / 4 (+ 8). Before: one key per thread, 128 FFMAs and 128
      // half -> float conversions each - the four consumer warps, not HBM, set the pace (1.85 us per 64 KB chunk).
      float sc4[4];
      {
        const uint32_t sK_a = smem_u32(sK);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          const int row = warp * 32 + t * 16 + (lane & 15);
#pragma unroll
          for (int ks = 0; ks < DA_DH / 16; ++ks) {
            const int unit = (ks & 3) * 2 + (lane >> 4);  // 16-byte unit inside the 64-dim tile; rows are 128-byte-swizzled
            uint32_t a[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                         : "r"(sK_a + (ks >> 2) * DS_TILE + row * 128 + ((unit ^ (row & 7)) << 4)));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(qb[ks][0]), "r"(qb[ks][1]));
          }
          const int k0 = warp * 32 + t * 16 + (lane >> 2);
          sc4[2 * t] = ((lane & 3) == 0 && k0 < n) ? c[0] * p.scale : -INFINITY;
          sc4[2 * t + 1] = ((lane & 3) == 0 && k0 + 8 < n) ? c[2] * p.scale : -INFINITY;
        }
      }
      float m = fmaxf(fmaxf(sc4[0], sc4[1]), fmaxf(sc4[2], sc4[3]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) s_red[par][warp] = m;
      ds_sync();
      const float Mn = fmaxf(M, fmaxf(fmaxf(s_red[par][0], s_red[par][1]), fmaxf(s_red[par][2], s_red[par][3])));
      const float sc = __expf(M - Mn);  // 0 on the first chunk (M = -inf)
      float p_sum = 0.f;
      if ((lane & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pj = sc4[i] > -INFINITY ? __expf(sc4[i] - Mn) : 0.f;
          s_p[par][warp * 32 + (i >> 1) * 16 + (lane >> 2) + (i & 1) * 8] = pj;
          p_sum += pj;
        }
      }
      l_t = fmaf(l_t, sc, p_sum);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] *= sc;
      M = Mn;
      ds_sync();
      // ---- P V: warp w owns keys j = w (mod 4) in increasing order; lane l owns dims [4 l, 4 l + 4)
#pragma unroll 4
      for (int j = warp; j < n; j += 4) {
        const float pj = s_p[par][j];
        const uint2 raw = *reinterpret_cast<const uint2*>(sV + v_half * DS_TILE + j * 128 + ((v_unit ^ (j & 7)) << 4) + v_sub);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
        const float2 a = __half22float2(h2[0]), bb = __half22float2(h2[1]);
        acc[0] = fmaf(pj, a.x, acc[0]);
        acc[1] = fmaf(pj, a.y, acc[1]);
        acc[2] = fmaf(pj, bb.x, acc[2]);
        acc[3] = fmaf(pj, bb.y, acc[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == DS_STAGES) {
        st = 0;
        phase ^= 1;
      }
    }
    // ---- segment result: L = sum of the per-thread partial sums, O = sum of the four warps' accumulators (fixed order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l_t += __shfl_xor_sync(0xffffffffu, l_t, o);
    if (lane == 0) s_red[0][warp] = l_t;  // the chunk loop's last reads of s_red lie behind its second barrier
#pragma unroll
    for (int i = 0; i < 4; ++i) sm.acc[warp][lane * 4 + i] = acc[i];
    ds_sync();
    const float L = (s_red[0][0] + s_red[0][1]) + (s_red[0][2] + s_red[0][3]);
    const float o_c = (sm.acc[0][tid] + sm.acc[1][tid]) + (sm.acc[2][tid] + sm.acc[3][tid]);
    if (c0 == 0 && c1 == ncb) {
      p.out[(size_t)b * p.ldo + h * DA_DH + tid] = __float2half_rn(L > 0.f ? o_c / L : 0.f);
    } else {
      // contributors of this (row, head): the CTAs whose ranges meet its chunks [g0, g1); slot = CTA index - first contributor
      const int g0 = q.base + h * ncb, g1 = g0 + ncb;
      const int i_first = (int)((((long long)(g0 + 1)) * G + N - 1) / N) - 1;
      const int i_last = (int)(((long long)g1 * G + N - 1) / N) - 1;
      const int cnt = i_last - i_first + 1;
      float* base = pp.part + ((size_t)(b * p.H + h) * pp.max_parts) * DS_PART;
      float* mine = base + ((int)blockIdx.x - i_first) * DS_PART;
      mine[2 + tid] = o_c;
      if (tid == 0) {
        mine[0] = M;
        mine[1] = L;
      }
      __threadfence();
      ds_sync();
      if (tid == 0) {
        int* ctr = pp.counters + b * p.H + h;
        const int old = atomicAdd(ctr, 1);
        s_last = (old == cnt - 1);
        if (s_last) *ctr = 0;  // every segment has arrived: leave the counter clean for the next launch / graph replay
      }
      ds_sync();
      if (s_last) {
        __threadfence();
        float Mx = -INFINITY;
        for (int cc = 0; cc < cnt; ++cc) Mx = fmaxf(Mx, __ldcg(base + cc * DS_PART));
        float Ls = 0.f, Os = 0.f;
        for (int cc = 0; cc < cnt; ++cc) {
          const float w = __expf(__ldcg(base + cc * DS_PART) - Mx);
          Ls = fmaf(w, __ldcg(base + cc * DS_PART + 1), Ls);
          Os = fmaf(w, __ldcg(base + cc * DS_PART + 2 + tid), Os);
        }
        p.out[(size_t)b * p.ldo + h * DA_DH + tid] = __float2half_rn(Ls > 0.f ? Os / Ls : 0.f);
      }
    }
    // next segment
    q.c = c1;
    if (q.c == ncb) {
      q.c = 0;
      if (++q.h == p.H) {
        q.h = 0;
        q.base += p.H * ncb;
        ++q.b;
      }
    }
  }
}

}  // namespace myr

using namespace myr;

// workspace bytes the split needs: counters + (chunks of the longest cache + 1) partial slots per (row, head)
size_t myr_decode_attn_stream_ws(int B, int H, int cache_len) {
  const size_t ctr = ((size_t)B * H * 4 + 15) & ~size_t(15);
  return ctr + (size_t)B * H * ((cache_len + DS_CH - 1) / DS_CH + 1) * DS_PART * 4;
}

// called by myr_decode_attention (decode_attn.cu) for caches beyond 256 slots when the caller passed a workspace
int myr_decode_attn_stream_launch(const DecodeAttnParams& p, void* ws, size_t ws_bytes, const void* kcache, const void* vcache,
                                  cudaStream_t stream) {
  MYR_CHECK_ARG(p.B <= 128, "decode_attention: the long-cache path holds at most 128 rows (got %d)", p.B);
  MYR_CHECK_ARG(ws_bytes >= myr_decode_attn_stream_ws(p.B, p.H, p.Smax), "decode_attention: split workspace too small");
  DecodeAttnStreamParams pp;
  pp.a = p;
  pp.a.trace = nullptr;
  pp.counters = reinterpret_cast<int*>(ws);
  pp.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + (((size_t)p.B * p.H * 4 + 15) & ~size_t(15)));
  pp.max_parts = (p.Smax + DS_CH - 1) / DS_CH + 1;
  DsMaps tm;
  {
    const void* ptrs[4] = {kcache, vcache, kcache, vcache};
    CUtensorMap* maps[4] = {&tm.k, &tm.v, &tm.k32, &tm.v32};
    for (int i = 0; i < 4; ++i) {
      const uint64_t dims[4] = {(uint64_t)DA_DH, (uint64_t)p.H, (uint64_t)p.Smax, (uint64_t)p.B};
      const uint64_t strides[3] = {(uint64_t)DA_DH * 2, (uint64_t)p.c_ts * 2, (uint64_t)p.c_bs * 2};
      const uint32_t box[4] = {64, 1, (uint32_t)(i < 2 ? DS_CH : 32), 1};
      const int rc = make_tmap_f16(maps[i], ptrs[i], 4, dims, strides, box);
      if (rc) return rc;
    }
  }
  const size_t smem = (size_t)DS_STAGES * DS_STAGE + 1024;
  static bool attr = false;
  if (!attr) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  const int max_chunks = p.B * p.H * ((p.Smax + DS_CH - 1) / DS_CH);
  const int grid = max_chunks < sm_count() ? max_chunks : sm_count();
  MYR_CHECK_CUDA(launch_kernel(decode_attn_stream_kernel, dim3((unsigned)grid), dim3(DS_THREADS), smem, stream, true, tm, pp));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
