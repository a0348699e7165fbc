// Persistent decode-step kernel: ONE launch runs a whole greedy-decode step of the LLaMA stack (modeling_llama.py:466-716
// for a single new token per sequence): embedding gather, 32 x [RMSNorm, qkv (+ LoRA A) projection, LoRA-B + RoPE +
// KV-cache append + attention, o_proj + residual, RMSNorm, gate/up projection + SwiGLU, down_proj + residual], final
// RMSNorm and lm_head.
//
// Why: a decode step is ~13.5 GB of weights streamed once (HBM-bound). As separate kernels (even graph-captured, with
// programmatic dependent launch) every one of the ~230 kernel boundaries drains the memory pipeline and costs ~4 us of
// dependency latency, i.e. as much time as the streaming itself. Here the 148 CTAs stay resident for the whole step:
//   * warp 0 (TMA producer) walks the op list and keeps an 11-stage ring of 16 KB WEIGHT tiles full. Weights never depend
//     on the step's data, so it runs up to 26 MB ahead, across op boundaries; only the 2 KB activation tile of a k-block
//     waits for the flag of the op that produces it;
//   * warp 1 issues tcgen05.mma (M = 128 weight rows, N = 16 tokens, fp32 accumulators in TMEM, two stages);
//   * warps 2-9 drain TMEM, finish split tiles (stream-K: every CTA owns an equal contiguous range of the op's k-blocks;
//     partials go to a workspace, the last CTA to arrive at a tile sums them in fixed order and applies the epilogue:
//     store / fp32 residual add / SwiGLU), count finished tiles per op, and the CTA that finishes an op's LAST tile runs
//     the op's tail (RMSNorm of the residual stream into the next GEMM's fp16 operand) and raises the op's flag.
//   * the attention op runs on the same warps, one (head, batch row) task per CTA, after the qkv op's flag.
// All cross-CTA hand-offs are release/acquire through global memory (threadfence + flag); operands produced inside the
// kernel are re-read with .cg loads or TMA (L2), never through the non-coherent L1. Reductions have a fixed order, so a
// replayed step reproduces the multi-kernel path's tokens bit for bit where the arithmetic is the same.
#include <string.h>

#include "decode_attn.cuh"

namespace myr {

constexpr int MG_BM = 128;
constexpr int MG_BK = 64;
constexpr int MG_BN = 16;       // UMMA N (token columns in TMEM)
constexpr int MG_TMAX = 8;      // sequences per step this kernel handles (register budget of the hand-off code)
constexpr int MG_A_BYTES = MG_BM * MG_BK * 2;  // 16 KiB
constexpr int MG_B_BYTES = MG_BN * MG_BK * 2;  //  2 KiB
constexpr int MG_STAGE_BYTES = MG_A_BYTES + MG_B_BYTES;
constexpr int MG_STAGES = 11;
constexpr int MG_EPI_WARPS = 8;
constexpr int MG_EPI_THREADS = 32 * MG_EPI_WARPS;
constexpr int MG_THREADS = 64 + MG_EPI_THREADS;
constexpr int MG_TMEM_COLS = 32;
constexpr int MG_MAX_SCORES = 4096;  // floats of smem for attention scores (cache length limit of the fused path)

enum { MG_GEMM = 0, MG_ATTN = 1, MG_EMBED = 2 };
enum { MG_EPI_F16 = 0, MG_EPI_RES32 = 1, MG_EPI_SWIGLU = 2, MG_EPI_F32 = 3 };

// device-side op record (built on the host by myr_mega_plan, copied verbatim)
struct MegaOp {
  int kind, epi;
  int T, F, K;
  int n_tiles, kb_total, per, max_seg;
  long long total_kb;
  long long partial_off;   // floats into the partial workspace
  int counter_off;         // first tile counter of this op
  int n_tasks;             // ATTN: heads * batch rows
  int dep_idx, dep_target; // this op may read its input once sync[dep_idx] >= dep_target (dep_idx < 0: no dependency)
  int has_tail;            // the op ends with an RMSNorm tail: completion is published through flags[] by the closing CTA
  int total_units;         // tiles (GEMM) / tasks (ATTN) / 1 (EMBED)
  void* out; long long ldo;
  // tail, run by the CTA that completes the op: dst16[t, :] = rmsnorm(src32[t, :]) * gamma  (norm_rows = 0: none)
  const float* norm_src; __half* norm_dst; const float* gamma; float eps; int D; int norm_rows;
  // MG_EMBED: h32[t, :] = table[ids[t], :]
  const __half* table; const int* ids; float* h32;
  DecodeAttnParams attn;
};

struct MegaParams {
  const MegaOp* ops; int n_ops;
  const CUtensorMap* maps;  // [2 * n_ops]: (weights, activations) of GEMM ops
  float* partial;
  int* tile_counters;       // arrival counters of split tiles (self-resetting)
  int* op_done;             // [n_ops] finished tiles / tasks (reset by the step's last finisher)
  int* flags;               // [n_ops] op complete incl. tail; flags == sync, op_done == sync + n_ops
  int inflight;             // weight tiles a CTA may have in flight (issued, not landed): bounds memory latency under load
  long long* trace;         // optional [3 * n_ops] globaltimer ns: op complete | CTA 0 producer saw its input | CTA 0 first tile drained
                            // followed by [n_ops][grid][2]: per CTA last drain of the op | producer saw the op's input
};
#define MG_TR(slot)                                                                                              \
  do {                                                                                                           \
    if (p.trace && epi_tid == 0) p.trace[3 * p.n_ops + (size_t)p.n_ops * gridDim.x * 2 + (size_t)oi * 8 + (slot)] = gtime(); \
  } while (0)
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void mg_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MG_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void mg_attn_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// spin until sync[i] >= target (watchdog: a protocol bug traps instead of hanging the GPU)
__device__ __forceinline__ void wait_dep(const int* flags, int i, int target) {
  if (i < 0) return;
  if (ld_acquire(flags + i) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire(flags + i) < target) {
    if (clock64() - t0 > 4000000000LL) {
      printf("[myriad_b200] decode-step watchdog: block %d thread %d waits for sync[%d] >= %d\n", blockIdx.x, threadIdx.x, i, target);
      __trap();
    }
  }
}

// position of one CTA inside the k-block space of the op list (GEMM ops only)
struct Cursor {
  int op;
  long long g, g1;
  int tile, kb, kb_end, kb_total;
  bool valid;
};
__device__ __forceinline__ void cur_load_seg(Cursor& c) {
  c.tile = (int)(c.g / c.kb_total);
  c.kb = (int)(c.g - (long long)c.tile * c.kb_total);
  const long long room = c.g1 - c.g;
  c.kb_end = (room < (long long)(c.kb_total - c.kb)) ? c.kb + (int)room : c.kb_total;
}
__device__ __forceinline__ void cur_seek(Cursor& c, const MegaParams& p) {
  while (c.op < p.n_ops) {
    const MegaOp& o = p.ops[c.op];
    if (o.kind == MG_GEMM) {
      const long long g0 = (long long)blockIdx.x * o.per;
      if (g0 < o.total_kb) {
        c.g = g0;
        c.g1 = (g0 + o.per < o.total_kb) ? g0 + o.per : o.total_kb;
        c.kb_total = o.kb_total;
        cur_load_seg(c);
        c.valid = true;
        return;
      }
    }
    ++c.op;
  }
  c.valid = false;
}
__device__ __forceinline__ void cur_init(Cursor& c, const MegaParams& p) {
  c.op = 0;
  cur_seek(c, p);
}
// move past `n` k-blocks of the current segment (n = 1: next k-block; n = kb_end - kb: next segment)
__device__ __forceinline__ void cur_advance(Cursor& c, const MegaParams& p, int n) {
  c.kb += n;
  c.g += n;
  if (c.kb < c.kb_end) return;
  if (c.g < c.g1) {
    cur_load_seg(c);
  } else {
    ++c.op;
    cur_seek(c, p);
  }
}

__device__ __forceinline__ float mg_swiglu(float g, float u) {
  g = round_f16(g);
  u = round_f16(u);
  return silu_f(g) * u;
}

// epilogue of one finished output element block: lane = weight row (feature), col = token
__device__ __forceinline__ void mg_store(const MegaOp& o, float v, int t, int f) {
  if (o.epi == MG_EPI_RES32) {
    float* h = reinterpret_cast<float*>(o.out) + (size_t)t * o.ldo + f;
    *h = __ldcg(h) + v;
  } else if (o.epi == MG_EPI_F32) {
    reinterpret_cast<float*>(o.out)[(size_t)t * o.ldo + f] = v;
  } else {
    reinterpret_cast<__half*>(o.out)[(size_t)t * o.ldo + f] = __float2half_rn(v);
  }
}

__global__ void __launch_bounds__(MG_THREADS, 1) decode_step_kernel(const MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MG_STAGES * MG_STAGE_BYTES);
  uint64_t* full = bars;                  // [MG_STAGES] weight tile landed
  uint64_t* empty = bars + MG_STAGES;     // [MG_STAGES]
  uint64_t* bfull = bars + 2 * MG_STAGES; // [MG_STAGES] activation tile landed
  uint64_t* tfull = bars + 3 * MG_STAGES; // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile int* s_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 4);  // [8]
  DecodeAttnSmem* s_attn = reinterpret_cast<DecodeAttnSmem*>(reinterpret_cast<uint8_t*>(bars) + 512);
  float* s_scores = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_attn) + ((sizeof(DecodeAttnSmem) + 15) & ~size_t(15)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&bfull[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], MG_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, MG_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      Cursor ca, cb;
      cur_init(ca, p);
      cur_init(cb, p);
      // weights are read exactly once per step: evict-first, so the 13 GB stream does not flush the step's small working
      // set (residual stream, partial tiles, norm weights, KV cache) out of L2 into DRAM latency
      const uint64_t pol_stream = l2_policy_evict_first();
      int ia = 0, ib = 0;         // k-blocks whose weight / activation tile has been issued
      int ready_op = -1;          // last op whose input flag has been observed
      bool waited_prev = false;   // griddepcontrol.wait executed (before the first activation access)
      long long t0 = clock64();
      while (cb.valid) {
        bool progress = false;
        if (ca.valid && ia - ib < MG_STAGES) {
          const int st = ia % MG_STAGES;
          // throttle: at most `inflight` weight tiles issued and not landed. Requests queue in the memory system, so a
          // deeper queue buys no bandwidth but delays every hand-off round trip of the epilogue warps behind it.
          bool room = true;
          if (ia >= p.inflight) {
            const int n = ia - p.inflight;
            room = mbar_try_wait(&full[n % MG_STAGES], ((uint32_t)(n / MG_STAGES)) & 1u);
          }
          if (room && mbar_try_wait(&empty[st], (((uint32_t)(ia / MG_STAGES)) & 1u) ^ 1u)) {
            mbar_arrive_expect_tx(&full[st], MG_A_BYTES);
            tma_load_2d_hint(smem + st * MG_STAGE_BYTES, &p.maps[2 * ca.op], &full[st], ca.kb * MG_BK, ca.tile * MG_BM, pol_stream);
            ++ia;
            cur_advance(ca, p, 1);
            progress = true;
          }
        }
        if (ib < ia) {
          bool ok = cb.op == ready_op;
          if (!ok) {
            if (!waited_prev) {
              pdl_wait();
              waited_prev = true;
            }
            const MegaOp& o = p.ops[cb.op];
            ok = o.dep_idx < 0 || ld_acquire(p.flags + o.dep_idx) >= o.dep_target;
            if (ok) {
              ready_op = cb.op;
              fence_proxy_async_all();  // the activations were written through the generic proxy by other CTAs
              if (p.trace && blockIdx.x == 0) p.trace[3 * cb.op + 1] = gtime();
              if (p.trace) p.trace[3 * p.n_ops + ((size_t)cb.op * gridDim.x + blockIdx.x) * 2 + 1] = gtime();
            }
          }
          if (ok) {
            const int st = ib % MG_STAGES;
            mbar_arrive_expect_tx(&bfull[st], MG_B_BYTES);
            tma_load_2d(smem + st * MG_STAGE_BYTES + MG_A_BYTES, &p.maps[2 * cb.op + 1], &bfull[st], cb.kb * MG_BK, 0);
            ++ib;
            cur_advance(cb, p, 1);
            progress = true;
          }
        }
        if (progress) {
          t0 = clock64();
        } else if (clock64() - t0 > 4000000000LL) {
          printf("[myriad_b200] decode-step watchdog: producer of block %d stuck at op %d (a %d b %d)\n", blockIdx.x, cb.op, ia, ib);
          __trap();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(MG_BM, MG_BN, 0, 0);
      Cursor c;
      cur_init(c, p);
      int i = 0, as = 0;
      uint32_t aphase = 0;
      while (c.valid) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * MG_BN;
        const int n = c.kb_end - c.kb;
        for (int k = 0; k < n; ++k, ++i) {
          const int st = i % MG_STAGES;
          mbar_wait(&full[st], ((uint32_t)(i / MG_STAGES)) & 1u);
          mbar_wait(&bfull[st], ((uint32_t)(i / MG_STAGES)) & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + st * MG_STAGE_BYTES);
          const uint32_t sb = sa + MG_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < MG_BK / 16; ++kk) {
            const uint64_t da = make_smem_desc(sa + kk * 32, 16, 1024);
            const uint64_t db = make_smem_desc(sb + kk * 32, 16, 1024);
            tc_mma_f16(d_tmem, da, db, idesc, (k > 0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(&empty[st]);
        }
        tc_commit(&tfull[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
        cur_advance(c, p, n);
      }
    }
  } else {
    // ------------------------------ epilogue / attention / op tails (warps 2..9) ------------------------------
    const int q = warp & 3;               // TMEM lane quarter
    const int hsel = (warp - 2) >> 2;     // 0: warps 2-5, 1: warps 6-9
    const int epi_tid = threadIdx.x - 64;
    const int lrow = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    Cursor c;
    cur_init(c, p);
    pdl_wait();

    for (int oi = 0; oi < p.n_ops; ++oi) {
      // by value: fields read through a reference to global memory would be re-loaded after every store / barrier (the
      // compiler must assume aliasing), which serialises the latency-critical hand-off code below
      const MegaOp o = p.ops[oi];
      bool finished_op = false;  // this CTA completed the op's last tile / task (uniform over the epilogue threads)

      if (o.kind == MG_EMBED) {
        if (blockIdx.x == 0) {
          // h32[t, :] = table[ids[t], :]   (myriad.py:308-311 / modeling_llama.py:505)
          for (int t = 0; t < o.T; ++t) {
            const int id = __ldcg(o.ids + t);
            const __half* src = o.table + (size_t)id * o.D;
            for (int c8 = epi_tid * 8; c8 < o.D; c8 += MG_EPI_THREADS * 8) {
              float f[8];
              da_unpack8(__ldg(reinterpret_cast<const uint4*>(src + c8)), f);
              float4* dst = reinterpret_cast<float4*>(o.h32 + (size_t)t * o.D + c8);
              dst[0] = make_float4(f[0], f[1], f[2], f[3]);
              dst[1] = make_float4(f[4], f[5], f[6], f[7]);
            }
          }
          __threadfence();
          mg_epi_bar();
          finished_op = true;
        }
      } else if (o.kind == MG_ATTN) {
        if ((int)blockIdx.x < o.n_tasks) {
          if (epi_tid == 0) wait_dep(p.flags, o.dep_idx, o.dep_target);
          mg_epi_bar();
          for (int task = blockIdx.x; task < o.n_tasks; task += gridDim.x) {
            if (epi_tid < 128) {
              mg_attn_bar();  // the previous task's shared-memory working set is no longer in use
              decode_attn_task(o.attn, task % o.attn.H, task / o.attn.H, epi_tid, *s_attn, s_scores, [] { mg_attn_bar(); });  // o is a local copy
            }
            __threadfence();
            mg_epi_bar();
            if (epi_tid == 0) red_release_add(&p.op_done[oi], 1);  // consumers poll the task count itself
          }
        }
      } else {
        // ---------------- GEMM op: every segment of this CTA's range ----------------
        if (oi + 1 < p.n_ops && p.ops[oi + 1].kind == MG_ATTN) {
          // the attention op that follows reads this CTA's (head, row) slice of the KV cache: start pulling it from DRAM
          // into L2 now, while the qkv weights stream (a DRAM miss under full streaming load costs several microseconds)
          const DecodeAttnParams& ap = p.ops[oi + 1].attn;
          const int task = blockIdx.x;
          if (task < ap.B * ap.H) {
            const int hh = task % ap.H, bb = task / ap.H;
            int kvl = __ldcg(ap.kv_len + bb);
            if (kvl > ap.Smax) kvl = ap.Smax;
            const __half* kb = ap.kcache + (size_t)bb * ap.c_bs + hh * DA_DH;
            const __half* vb = ap.vcache + (size_t)bb * ap.c_bs + hh * DA_DH;
            for (int j = epi_tid; j < 2 * kvl; j += MG_EPI_THREADS) {  // one 128-byte line = half a K or V row
              const int row = j >> 1, hf2 = j & 1;
              prefetch_l2(kb + (size_t)row * ap.c_ts + hf2 * 64);
              prefetch_l2(vb + (size_t)row * ap.c_ts + hf2 * 64);
            }
          }
        }
        while (c.valid && c.op == oi) {
          const int tile = c.tile;
          const int n = c.kb_end - c.kb;
          const long long tb = (long long)tile * o.kb_total;
          const int first = (int)(tb / o.per), last = (int)((tb + o.kb_total - 1) / o.per);
          const int n_seg = last - first + 1;
          const int seg = (int)blockIdx.x - first;
          const bool swiglu = o.epi == MG_EPI_SWIGLU;
          float* ws = p.partial + o.partial_off + ((size_t)tile * o.max_seg + seg) * (MG_BN * MG_BM);
          const int f = tile * MG_BM + lrow;

          if (seg == 0 && hsel == 0) {
            // finisher: pull what the hand-off will read (other owners' partials, the residual rows) towards L2 while the
            // MMAs of this segment are still running; at full streaming load a DRAM miss costs several microseconds
            const float* w0p = p.partial + o.partial_off + (size_t)tile * o.max_seg * (MG_BN * MG_BM);
            if (lane == 0) {
              for (int j = 0; j < o.T; ++j) {
                for (int sg = 1; sg < n_seg; ++sg) prefetch_l2(w0p + (size_t)sg * (MG_BN * MG_BM) + (size_t)j * MG_BM + lrow);
                if (o.epi == MG_EPI_RES32 && f < o.F) prefetch_l2(reinterpret_cast<const float*>(o.out) + (size_t)j * o.ldo + f);
              }
            }
          }
          if (o.has_tail && o.norm_rows > 0 && epi_tid < (o.D >> 5)) prefetch_l2(o.gamma + epi_tid * 32);
          mbar_wait(&tfull[as], aphase);
          tc_fence_after();
          if (p.trace && blockIdx.x == 0 && epi_tid == 0) p.trace[3 * oi + 2] = gtime();
          if (p.trace && epi_tid == 0) p.trace[3 * p.n_ops + ((size_t)oi * gridDim.x + blockIdx.x) * 2] = gtime();
          // The CTA that owns the tile's FIRST k-range reaches it at the END of its own range, i.e. last in time: it is the
          // tile's finisher and keeps its partial on chip; the other owners park theirs in the workspace and move on.
          const bool is_fin = seg == 0;
          float acc[MG_TMAX];
          if (hsel == 0) {
            uint32_t r[16];
            tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + as * MG_BN, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < MG_TMAX; ++j) acc[j] = __uint_as_float(r[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
          if (++as == 2) {
            as = 0;
            aphase ^= 1;
          }

          if (!is_fin) {
            if (hsel == 0) {
#pragma unroll
              for (int j = 0; j < MG_TMAX; ++j)
                if (j < o.T) ws[(size_t)j * MG_BM + lrow] = acc[j];
            }
            __threadfence();
            mg_epi_bar();
            if (epi_tid == 0) red_release_add(&p.tile_counters[o.counter_off + tile], 1);
          } else {
            if (n_seg > 1) {
              if (epi_tid == 0) {
                int* cnt = &p.tile_counters[o.counter_off + tile];
                const long long t0 = clock64();
                while (ld_acquire(cnt) != n_seg - 1) {
                  if (clock64() - t0 > 4000000000LL) {
                    printf("[myriad_b200] decode-step watchdog: block %d op %d tile %d waits for %d partials\n", blockIdx.x, oi, tile,
                           n_seg - 1);
                    __trap();
                  }
                }
                *cnt = 0;  // every other owner has arrived: leave the counter clean for the next launch
              }
              mg_epi_bar();
              __threadfence();
            }
            MG_TR(0);  // (last writer wins: the closing CTA's values survive for ops with a tail)
            const float* w0 = p.partial + o.partial_off + (size_t)tile * o.max_seg * (MG_BN * MG_BM);
            if (hsel == 0) {
              // fixed summation order: own k-range first, then the other owners by k. All loads (partials of the other
              // owners, residual) are issued before the first use: one memory round trip for the whole tile.
              const size_t SEG = (size_t)MG_BN * MG_BM;
              const bool res = (o.epi == MG_EPI_RES32) && f < o.F;
              const float* hrow = reinterpret_cast<const float*>(o.out) + f;
              float hres[MG_TMAX];
#pragma unroll
              for (int jb = 0; jb < MG_TMAX; jb += 4) {  // four tokens at a time: everything stays in registers
                if (jb < o.T) {
                  float p1[4], p2[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int j = jb + u;
                    const bool on = j < o.T;
                    p1[u] = (on && n_seg > 1) ? __ldcg(w0 + SEG + (size_t)j * MG_BM + lrow) : 0.f;
                    p2[u] = (on && n_seg > 2) ? __ldcg(w0 + 2 * SEG + (size_t)j * MG_BM + lrow) : 0.f;
                    hres[j] = (on && res) ? __ldcg(hrow + (size_t)j * o.ldo) : 0.f;
                  }
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int j = jb + u;
                    float a = (acc[j] + p1[u]) + p2[u];
                    if (j < o.T)
                      for (int sg = 3; sg < n_seg; ++sg) a += __ldcg(w0 + (size_t)sg * SEG + (size_t)j * MG_BM + lrow);
                    acc[j] = a;
                  }
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u) hres[jb + u] = 0.f;
                }
              }
              if (!swiglu) {
                if (f < o.F) {
#pragma unroll
                  for (int j = 0; j < MG_TMAX; ++j) {
                    if (j < o.T) {
                      if (o.epi == MG_EPI_RES32)
                        reinterpret_cast<float*>(o.out)[(size_t)j * o.ldo + f] = hres[j] + acc[j];
                      else if (o.epi == MG_EPI_F32)
                        reinterpret_cast<float*>(o.out)[(size_t)j * o.ldo + f] = acc[j];
                      else
                        reinterpret_cast<__half*>(o.out)[(size_t)j * o.ldo + f] = __float2half_rn(acc[j]);
                    }
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < MG_TMAX; ++j)
                  if (j < o.T) s_scores[j * MG_BM + lrow] = acc[j];
              }
            }
            if (swiglu) {
              // weight rows interleaved in blocks of 64: lane l = gate row, lane l + 64 = the matching up row
              mg_epi_bar();
              const int i = tile * 64 + (epi_tid & 63);
              if (i < (o.F >> 1)) {
                for (int col = epi_tid >> 6; col < o.T; col += MG_EPI_THREADS / 64) {
                  const float gv = s_scores[col * MG_BM + (epi_tid & 63)], uv = s_scores[col * MG_BM + 64 + (epi_tid & 63)];
                  reinterpret_cast<__half*>(o.out)[(size_t)col * o.ldo + i] = __float2half_rn(mg_swiglu(gv, uv));
                }
              }
            }
            // one more finished tile of this op. Ops without a tail publish the count itself (fire and forget: consumers
            // poll it); ops with a tail (and the step's last op) elect the closing CTA with the returned count.
            MG_TR(1);
            __threadfence();
            mg_epi_bar();
            MG_TR(2);
            if (o.has_tail || oi == p.n_ops - 1) {
              if (epi_tid == 0) {
                const int old = atomicAdd(&p.op_done[oi], 1);
                *s_flag = (old == o.n_tiles - 1);
              }
              mg_epi_bar();
              MG_TR(3);
              if (*s_flag) finished_op = true;
            } else if (epi_tid == 0) {
              red_release_add(&p.op_done[oi], 1);
            }
          }
          cur_advance(c, p, n);
        }
      }

      if (finished_op) {
        // ---------------- op tail, by the CTA that completed it ----------------
        MG_TR(4);
        __threadfence();
        if (o.norm_rows > 0) {
          // RMSNorm (modeling_llama.py:66-74): dst16 = fp16( gamma * (x * rsqrt(mean(x^2) + eps)) ), fp32 statistics.
          // Two warps per row, four rows at a time; a row's values stay in registers between the two passes (D <= 4096),
          // so the tail costs about two memory round trips whatever the number of rows.
          const int grp = (warp - 2) >> 1, hf = (warp - 2) & 1;
          for (int base = 0; base < o.norm_rows; base += 4) {
            const int t = base + grp;
            const bool t_ok = t < o.norm_rows;
            const float* x = o.norm_src + (size_t)(t_ok ? t : 0) * o.D;
            const int nv = o.D >> 2;  // float4 per row
            float4 v[16];
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c4 = hf * 32 + lane + 64 * i;
              v[i] = (t_ok && c4 < nv) ? __ldcg(reinterpret_cast<const float4*>(x) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
            for (int c4 = hf * 32 + lane + 64 * 16; t_ok && c4 < nv; c4 += 64) {  // D > 4096: the rest is re-read in pass 2
              const float4 w = __ldcg(reinterpret_cast<const float4*>(x) + c4);
              ss += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sft);
            MG_TR(5);
            mg_epi_bar();
            if (lane == 0) s_red[warp - 2] = ss;
            mg_epi_bar();
            const float rstd = rsqrtf((s_red[2 * grp] + s_red[2 * grp + 1]) / o.D + o.eps);
            if (t_ok) {
              __half* y = o.norm_dst + (size_t)t * o.D;
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {  // gamma in two batches of 8 independent loads
                float4 gm[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int c4 = hf * 32 + lane + 64 * (h8 * 8 + i);
                  gm[i] = (c4 < nv) ? __ldg(reinterpret_cast<const float4*>(o.gamma) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int c4 = hf * 32 + lane + 64 * (h8 * 8 + i);
                  if (c4 < nv) {
                    const float4 xv = v[h8 * 8 + i];
                    __half2 h0 = __floats2half2_rn(gm[i].x * (xv.x * rstd), gm[i].y * (xv.y * rstd));
                    __half2 h1 = __floats2half2_rn(gm[i].z * (xv.z * rstd), gm[i].w * (xv.w * rstd));
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    reinterpret_cast<uint2*>(y)[c4] = pk;
                  }
                }
              }
              for (int c4 = hf * 32 + lane + 64 * 16; c4 < nv; c4 += 64) {
                const float4 w = __ldcg(reinterpret_cast<const float4*>(x) + c4);
                const float4 gm = __ldg(reinterpret_cast<const float4*>(o.gamma) + c4);
                __half2 h0 = __floats2half2_rn(gm.x * (w.x * rstd), gm.y * (w.y * rstd));
                __half2 h1 = __floats2half2_rn(gm.z * (w.z * rstd), gm.w * (w.w * rstd));
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                reinterpret_cast<uint2*>(y)[c4] = pk;
              }
            }
          }
        }
        MG_TR(6);
        fence_proxy_async_all();
        __threadfence();
        mg_epi_bar();
        MG_TR(7);
        if (epi_tid == 0) {
          if (p.trace) p.trace[3 * oi] = gtime();
          if (oi == p.n_ops - 1) {
            // the step is complete: every CTA is past every wait, leave the hand-off state clean for the next launch
            for (int k = 0; k < p.n_ops; ++k) {
              p.flags[k] = 0;
              p.op_done[k] = 0;
            }
          } else {
            st_release(p.flags + oi, 1);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, MG_TMEM_COLS);
}

static size_t mg_smem_bytes() {
  return (size_t)MG_STAGES * MG_STAGE_BYTES + 1024 + 512 + ((sizeof(DecodeAttnSmem) + 15) & ~size_t(15)) + MG_MAX_SCORES * sizeof(float);
}

}  // namespace myr

using namespace myr;

// blob = [CUtensorMap maps[2 * n_ops]] [MegaOp ops[n_ops]]
extern "C" size_t myr_mega_plan_bytes(int32_t n_ops) {
  return (size_t)n_ops * 2 * sizeof(CUtensorMap) + (size_t)n_ops * sizeof(MegaOp);
}

extern "C" int myr_mega_plan(const myr_mega_op* ops, int32_t n_ops, void* host_blob, size_t blob_bytes, size_t* workspace_bytes,
                             int32_t* n_tile_counters) {
  MYR_CHECK_ARG(ops && host_blob && workspace_bytes && n_tile_counters && n_ops > 0, "mega_plan: bad arguments");
  MYR_CHECK_ARG(blob_bytes >= myr_mega_plan_bytes(n_ops), "mega_plan: blob too small");
  MYR_CHECK_ARG((reinterpret_cast<uintptr_t>(host_blob) & 63) == 0, "mega_plan: blob must be 64-byte aligned");
  const int grid = sm_count();
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(host_blob);
  MegaOp* dops = reinterpret_cast<MegaOp*>(maps + 2 * n_ops);
  memset(host_blob, 0, myr_mega_plan_bytes(n_ops));
  long long partial_floats = 0;
  int counters = 0;
  for (int i = 0; i < n_ops; ++i) {
    const myr_mega_op& s = ops[i];
    MegaOp& d = dops[i];
    d.kind = s.kind;
    d.norm_src = reinterpret_cast<const float*>(s.norm_src);
    d.norm_dst = reinterpret_cast<__half*>(s.norm_dst);
    d.gamma = reinterpret_cast<const float*>(s.gamma);
    d.eps = s.eps;
    d.D = s.D;
    d.norm_rows = (s.norm_src && s.norm_dst && s.gamma) ? s.norm_rows : 0;
    MYR_CHECK_ARG(d.norm_rows == 0 || (s.D % 4 == 0 && s.D > 0), "mega_plan: op %d: norm width %d must be a multiple of 4", i, s.D);
    if (s.kind == MG_EMBED) {
      MYR_CHECK_ARG(s.table && s.ids && s.h32 && s.T > 0 && s.T <= MG_TMAX && s.D % 8 == 0, "mega_plan: op %d: bad embed arguments", i);
      d.table = reinterpret_cast<const __half*>(s.table);
      d.ids = reinterpret_cast<const int*>(s.ids);
      d.h32 = reinterpret_cast<float*>(s.h32);
      d.T = s.T;
    } else if (s.kind == MG_ATTN) {
      const myr_decode_attn_args* a = &s.attn;
      MYR_CHECK_ARG(a->qkv && a->pos && a->cos_table && a->sin_table && a->kcache && a->vcache && a->out && a->kv_len,
                    "mega_plan: op %d: attention null pointer", i);
      MYR_CHECK_ARG(a->dh == DA_DH && a->cache_len > 0 && a->cache_len <= MG_MAX_SCORES, "mega_plan: op %d: dh must be %d and cache_len <= %d",
                    i, DA_DH, MG_MAX_SCORES);
      MYR_CHECK_ARG(a->lora_r == 0 || a->lora_r == 8, "mega_plan: op %d: LoRA rank must be 8", i);
      DecodeAttnParams& q = d.attn;
      q.qkv = reinterpret_cast<const __half*>(a->qkv); q.ldq = a->ldq;
      q.B = a->B; q.H = a->H; q.Smax = a->cache_len;
      q.pos = reinterpret_cast<const int*>(a->pos);
      q.cos_t = reinterpret_cast<const float*>(a->cos_table); q.sin_t = reinterpret_cast<const float*>(a->sin_table);
      q.kcache = reinterpret_cast<__half*>(a->kcache); q.vcache = reinterpret_cast<__half*>(a->vcache);
      q.c_ts = a->cache_token_stride; q.c_bs = a->cache_batch_stride;
      q.cache_off = reinterpret_cast<const int*>(a->cache_off_dev); q.cache_off_host = a->cache_off;
      q.kv_len = reinterpret_cast<const int*>(a->kv_len);
      q.lora_bq = reinterpret_cast<const __half*>(a->lora_bq); q.lora_bv = reinterpret_cast<const __half*>(a->lora_bv);
      q.lora_r = (a->lora_bq && a->lora_bv) ? a->lora_r : 0; q.lora_scale = a->lora_scale;
      q.scale = a->scale;
      q.out = reinterpret_cast<__half*>(a->out); q.ldo = a->ldo;
      q.next_layer_stride = 0;  // the persistent kernel prefetches the slice itself while the qkv weights stream
      q.trace = nullptr;
      d.n_tasks = a->B * a->H;
    } else if (s.kind == MG_GEMM) {
      MYR_CHECK_ARG(s.x && s.w && s.out && s.T > 0 && s.T <= MG_TMAX && s.F > 0 && s.K > 0, "mega_plan: op %d: bad GEMM (T <= %d)", i, MG_TMAX);
      MYR_CHECK_ARG(s.ldx % 8 == 0 && s.ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(s.x) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(s.w) & 15) == 0,
                    "mega_plan: op %d: operands must be 16-byte aligned", i);
      MYR_CHECK_ARG(s.epi >= MG_EPI_F16 && s.epi <= MG_EPI_F32, "mega_plan: op %d: bad epilogue", i);
      MYR_CHECK_ARG(s.epi != MG_EPI_SWIGLU || s.F % 128 == 0, "mega_plan: op %d: SwiGLU needs F %% 128 == 0", i);
      d.epi = s.epi;
      d.T = s.T; d.F = s.F; d.K = s.K;
      d.out = s.out; d.ldo = s.ldo;
      d.kb_total = ceil_div(s.K, MG_BK);
      d.n_tiles = ceil_div(s.F, MG_BM);
      d.total_kb = (long long)d.n_tiles * d.kb_total;
      // equal contiguous k-block ranges over all CTAs; whole tiles per CTA when that is already balanced
      const int tiles_per_cta = ceil_div(d.n_tiles, grid);
      const double dp_eff = (double)d.n_tiles / ((double)tiles_per_cta * grid);
      int per = tiles_per_cta * d.kb_total;
      d.max_seg = 1;
      if (dp_eff < 0.95) {
        const int sk = (int)((d.total_kb + grid - 1) / grid);
        if (sk >= 2) {
          per = sk;
          d.max_seg = (d.kb_total + per - 1) / per + 1;
        }
      }
      d.per = per;
      d.partial_off = partial_floats;
      d.counter_off = counters;
      partial_floats += (long long)d.n_tiles * d.max_seg * MG_BN * MG_BM;
      counters += d.n_tiles;
      uint64_t dims[2], strides[1];
      uint32_t box[2];
      dims[0] = (uint64_t)s.K; dims[1] = (uint64_t)s.F; box[0] = MG_BK; box[1] = MG_BM;
      strides[0] = (uint64_t)s.ldw * 2;
      int rc = make_tmap_f16(&maps[2 * i], s.w, 2, dims, strides, box);
      if (rc) return rc;
      dims[1] = (uint64_t)s.T; box[1] = MG_BN;
      strides[0] = (uint64_t)s.ldx * 2;
      rc = make_tmap_f16(&maps[2 * i + 1], s.x, 2, dims, strides, box);
      if (rc) return rc;
    } else {
      set_error("mega_plan: op %d: unknown kind %d", i, s.kind);
      return MYR_ERR_INVALID;
    }
  }
  // dependencies: op i reads what op i - 1 wrote; completion of an op with a tail is its flag, otherwise its unit count
  for (int i = 0; i < n_ops; ++i) {
    MegaOp& d = dops[i];
    d.has_tail = d.norm_rows > 0 ? 1 : 0;
    d.total_units = d.kind == MG_GEMM ? d.n_tiles : (d.kind == MG_ATTN ? d.n_tasks : 1);
    if (d.kind == MG_EMBED) d.has_tail = 1;  // published through its flag (single CTA)
    if (i == 0) {
      d.dep_idx = -1;
      d.dep_target = 0;
    } else if (dops[i - 1].has_tail) {
      d.dep_idx = i - 1;
      d.dep_target = 1;
    } else {
      d.dep_idx = n_ops + (i - 1);
      d.dep_target = dops[i - 1].total_units;
    }
  }
  // workspace: flags[n_ops] | op_done[n_ops] | tile counters | (64-byte aligned) partials
  size_t ints = (size_t)2 * n_ops + counters;
  ints = (ints + 15) / 16 * 16;
  *workspace_bytes = ints * sizeof(int) + (size_t)partial_floats * sizeof(float);
  *n_tile_counters = counters;
  return MYR_OK;
}

extern "C" int myr_mega_launch(const void* dev_blob, int32_t n_ops, void* workspace, size_t workspace_bytes, int32_t n_tile_counters,
                               void* trace, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(dev_blob && workspace && n_ops > 0, "mega_launch: bad arguments");
  MYR_CHECK_ARG((reinterpret_cast<uintptr_t>(dev_blob) & 63) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "mega_launch: blob must be 64-byte aligned, workspace 16-byte aligned");
  MegaParams p;
  p.maps = reinterpret_cast<const CUtensorMap*>(dev_blob);
  p.ops = reinterpret_cast<const MegaOp*>(p.maps + 2 * n_ops);
  p.n_ops = n_ops;
  int* ints = reinterpret_cast<int*>(workspace);
  p.flags = ints;
  p.op_done = ints + n_ops;
  p.tile_counters = ints + 2 * n_ops;
  size_t n_ints = (size_t)2 * n_ops + n_tile_counters;
  n_ints = (n_ints + 15) / 16 * 16;
  MYR_CHECK_ARG(workspace_bytes >= n_ints * sizeof(int), "mega_launch: workspace too small");
  p.partial = reinterpret_cast<float*>(ints + n_ints);
  p.trace = reinterpret_cast<long long*>(trace);
  {
    const char* e = getenv("MYR_MEGA_INFLIGHT");
    int v = e ? atoi(e) : 4;
    p.inflight = v < 1 ? 1 : (v > MG_STAGES ? MG_STAGES : v);
  }
  static bool attr_set = false;
  if (!attr_set) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mg_smem_bytes()));
    attr_set = true;
  }
  MYR_CHECK_CUDA(launch_kernel(decode_step_kernel, dim3(sm_count()), dim3(MG_THREADS), mg_smem_bytes(), stream, true, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
