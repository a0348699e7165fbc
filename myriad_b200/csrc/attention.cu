// FlashAttention-style forward on tcgen05 for sm_100a.
//
//   O[b, i, h, :] = softmax_j( scale * Q[b, i, h, :] . K[b, j, h, :]  + mask ) V[b, j, h, :]
//
// One CTA = one (128-query tile, head, batch). Per KV tile of BN keys (BN in {32..256}, multiple of 32):
//   TMA   : K, V tiles -> 128B-swizzled smem (4-D tensor maps: OOB rows / head-dim padding are zero-filled)
//   MMA 1 : S = Q K^T          tcgen05.mma  M=128, N=BN, K=dh_pad      (A = Q K-major, B = K K-major) -> TMEM
//   softmax: thread r owns query row r (TMEM lane r): tcgen05.ld, online max/sum in fp32, exp2; P -> smem fp16
//            in the UMMA K-major SWIZZLE_128B layout (aliases the K tile, which MMA 1 has consumed)
//   MMA 2 : O_j = P V          tcgen05.mma  M=128, N=dh_pad, K=BN      (A = P K-major, B = V MN-major) -> TMEM
//   acc   : registers: acc = acc * alpha + O_j
// Two CTAs per SM (<= 256 TMEM columns, <= ~100 KB smem each) overlap one CTA's softmax with the other's MMAs.
//
// Replaces (reference, eager): eva_vit.py:128-144 (ViT, dh=88, N=257, q pre-scaled), Qformer.py:228-265 (self:
// Q=K<=81; cross: K=257; dh=64, scores / sqrt(dh)), modeling_llama.py:197-215 (causal + padding, dh=128, fp32
// softmax) including decode against the pre-allocated KV cache (replaces the torch.cat cache growth :190-195).
// The additive S x S masks of modeling_llama.py:25-54,442-463 are never materialised: a key j is visible to
// query i iff j < kv_len[b] and (not causal or j <= q_off + i).
#include "common.h"
#include "ptx.cuh"

namespace myr {

struct AttnKernelParams {
  int Sq, Skv, dh, dhp, DB;  // DB = number of 64-wide head-dim blocks (1 or 2)
  int BN, KB;                // keys per tile, number of 64-key blocks of P
  int kp_bytes, v_bytes;
  int tmem_cols, ocol;
  float scale_log2;
  int causal, q_off;
  const int* kv_len;         // [B] device, or null -> Skv
  __half* out;
  long long o_ts, o_bs, o_hs;  // element strides of out: token, batch, head
  uint32_t idesc_qk, idesc_pv;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NCH>  // NCH = dh_pad / 32 output chunks per row (2, 3 or 4)
__global__ void __launch_bounds__(128, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKP = sQ + p.DB * 16384;
  uint8_t* sV = sKP + p.kp_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.v_bytes);
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * 128;
  const int kv_len = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
  int kv_end = kv_len;
  if (p.causal) kv_end = min(kv_len, p.q_off + q0 + 128);
  const int n_tiles = (kv_end + p.BN - 1) / p.BN;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // q / k / v (and kv_len) come from the previous kernels
  pdl_launch_dependents();
  const uint32_t t_s = tmem_base + (uint32_t(warp * 32) << 16);
  const uint32_t t_o = t_s + (uint32_t)p.ocol;

  if (tid == 0 && n_tiles > 0) {
    mbar_arrive_expect_tx(bar_q, (uint32_t)p.DB * 16384u);
    for (int db = 0; db < p.DB; ++db) tma_load_4d(sQ + db * 16384, &tmQ, bar_q, db * 64, h, q0, b);
  }

  float acc[NCH * 32];
#pragma unroll
  for (int i = 0; i < NCH * 32; ++i) acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const int row = tid;  // query row within the tile == TMEM lane
  const int q_pos = p.q_off + q0 + row;
  const int nchunk_s = p.BN / 32;

  for (int j = 0; j < n_tiles; ++j) {
    const uint32_t par = j & 1;
    const int k0 = j * p.BN;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_k, (uint32_t)(p.DB * p.BN * 128));
      for (int db = 0; db < p.DB; ++db) tma_load_4d(sKP + db * p.BN * 128, &tmK, bar_k, db * 64, h, k0, b);
      mbar_arrive_expect_tx(bar_v, (uint32_t)(p.DB * p.BN * 128));
      for (int db = 0; db < p.DB; ++db) tma_load_4d(sV + db * p.BN * 128, &tmV, bar_v, db * 64, h, k0, b);
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, par);
      tc_fence_after();
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sKP);
      const int ksteps = p.dhp / 16;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = make_smem_desc(aq + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
        const uint64_t db_ = make_smem_desc(ak + (ks >> 2) * (p.BN * 128) + (ks & 3) * 32, 16, 1024);
        tc_mma_f16(tmem_base, da, db_, p.idesc_qk, ks > 0 ? 1u : 0u);
      }
      tc_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, par);
    tc_fence_after();

    // ---- pass 1: row maximum of the scaled, masked scores
    const bool edge = (k0 + p.BN > kv_len) || (p.causal && (k0 + p.BN - 1 > p.q_off + q0));
    float m_tile = -INFINITY;
    for (int c = 0; c < nchunk_s; ++c) {
      uint32_t r[32];
      tmem_ld32(t_s + c * 32, r);
      tmem_ld_wait();
      if (!edge) {
#pragma unroll
        for (int i = 0; i < 32; ++i) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int key = k0 + c * 32 + i;
          const bool ok = key < kv_len && (!p.causal || key <= q_pos);
          if (ok) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
        }
      }
    }
    const float m_new = fmaxf(m_run, m_tile);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked row so far: keep everything at zero
    const float alpha = fast_exp2((m_run - m_use) * p.scale_log2);  // m_run = -inf -> 0
    const float moff = m_use * p.scale_log2;

    // ---- pass 2: P = exp2(s * scale_log2 - moff) -> fp16 -> swizzled smem; row sum in fp32
    float l_tile = 0.f;
    for (int c = 0; c < nchunk_s; ++c) {
      uint32_t r[32];
      tmem_ld32(t_s + c * 32, r);
      tmem_ld_wait();
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -moff));
        float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -moff));
        if (edge) {
          const int key = k0 + c * 32 + i;
          if (!(key < kv_len && (!p.causal || key <= q_pos))) p0 = 0.f;
          if (!(key + 1 < kv_len && (!p.causal || key + 1 <= q_pos))) p1 = 0.f;
        }
        l_tile += p0 + p1;
        __half2 hp = __floats2half2_rn(p0, p1);
        packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hp);
      }
      const int c0 = c * 32;
      uint8_t* base = sKP + (c0 >> 6) * 16384 + row * 128;
      const int slot0 = (c0 & 63) >> 3;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v = make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
        *reinterpret_cast<uint4*>(base + (((slot0 + g) ^ (row & 7)) << 4)) = v;
      }
    }
    l_run = l_run * alpha + l_tile;
    m_run = m_new;

    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, par);
      const uint32_t ap = smem_u32(sKP), av = smem_u32(sV);
      const int ksteps = p.BN / 16;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = make_smem_desc(ap + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
        const uint64_t db_ = make_smem_desc(av + ks * 2048, (uint32_t)(p.BN * 128), 1024);
        tc_mma_f16(tmem_base + (uint32_t)p.ocol, da, db_, p.idesc_pv, ks > 0 ? 1u : 0u);
      }
      tc_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, par);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t r[32];
      tmem_ld32(t_o + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], alpha, __uint_as_float(r[i]));
    }
    tc_fence_before();  // orders these TMEM reads before the next tile's MMAs (issued after the next __syncthreads)
  }

  // ---- finalise: O = acc / l  -> fp16, 16-byte stores along the head dim
  if (q0 + row < p.Sq) {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __half* o = p.out + (long long)b * p.o_bs + (long long)(q0 + row) * p.o_ts + (long long)h * p.o_hs;
#pragma unroll
    for (int c8 = 0; c8 < NCH * 4; ++c8) {
      if (c8 * 8 < p.dh) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2 hp = __floats2half2_rn(acc[c8 * 8 + 2 * i] * inv, acc[c8 * 8 + 2 * i + 1] * inv);
          w[i] = *reinterpret_cast<uint32_t*>(&hp);
        }
        *reinterpret_cast<uint4*>(o + c8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

static int pick_bn(int skv) {
  // smallest padded key count first, then the larger tile (fewer serial softmax/MMA phases)
  int best = 128, best_pad = 1 << 30;
  const int cands[4] = {128, 96, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int nt = ceil_div(skv, bn);
    const int pad = nt * bn + 64 * nt;  // padded keys + a per-tile serial-phase overhead (in key units)
    if (pad < best_pad) {
      best_pad = pad;
      best = bn;
    }
  }
  return best;
}

static uint32_t pow2_cols(int c) {
  uint32_t v = 32;
  while ((int)v < c) v <<= 1;
  return v;
}

}  // namespace myr

namespace myr {
bool attn2_enabled();  // attention2.cu: the warp-specialised kernel (default); MYR_ATTN2=0 or a bn_hint select this file's kernel
int attn2_launch(const myr_attn_args* a, cudaStream_t stream);
}

using namespace myr;

extern "C" int myr_attention_fwd(const myr_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MYR_CHECK_ARG(a != nullptr, "attention: null args");
  MYR_CHECK_ARG(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Skv > 0, "attention: bad shape B=%d H=%d Sq=%d Skv=%d", a->B,
                a->H, a->Sq, a->Skv);
  MYR_CHECK_ARG(a->dh % 8 == 0 && a->dh >= 16 && a->dh <= 128, "attention: head dim %d unsupported (8 | dh <= 128)", a->dh);
  MYR_CHECK_ARG(a->q && a->k && a->v && a->out, "attention: null pointer");
  MYR_CHECK_ARG(a->bn_hint == 0 || (a->bn_hint % 32 == 0 && a->bn_hint >= 32 && a->bn_hint <= 256),
                "attention: bn_hint %d must be a multiple of 32 in [32,256]", a->bn_hint);
  if (attn2_enabled() && a->bn_hint == 0) return attn2_launch(a, stream);

  AttnKernelParams p;
  p.Sq = a->Sq; p.Skv = a->Skv; p.dh = a->dh;
  p.dhp = ceil_div(a->dh, 32) * 32;
  if (p.dhp < 64) p.dhp = 64;
  p.DB = ceil_div(p.dhp, 64);
  p.BN = a->bn_hint ? a->bn_hint : pick_bn(a->Skv);
  p.KB = ceil_div(p.BN, 64);
  const int k_bytes = p.DB * p.BN * 128, p_bytes = p.KB * 16384;
  p.kp_bytes = ((k_bytes > p_bytes ? k_bytes : p_bytes) + 1023) & ~1023;
  p.v_bytes = (k_bytes + 1023) & ~1023;
  p.ocol = p.BN;
  p.tmem_cols = (int)pow2_cols(p.BN + p.dhp);
  MYR_CHECK_ARG(p.tmem_cols <= 512, "attention: BN=%d + dh=%d exceeds tensor memory", p.BN, p.dhp);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal; p.q_off = a->q_off;
  p.kv_len = reinterpret_cast<const int*>(a->kv_len);
  p.out = reinterpret_cast<__half*>(a->out);
  p.o_ts = a->o_token_stride; p.o_bs = a->o_batch_stride; p.o_hs = a->o_head_stride;
  p.idesc_qk = make_idesc_f16(128, p.BN, 0, 0);
  p.idesc_pv = make_idesc_f16(128, p.dhp, 0, 1);
  MYR_CHECK_ARG(p.o_ts % 8 == 0 && p.o_bs % 8 == 0 && p.o_hs % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                "attention: output must be 16-byte aligned per (token, head)");

  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    const void* ptrs[3] = {a->q, a->k, a->v};
    const int64_t ts[3] = {a->q_token_stride, a->k_token_stride, a->v_token_stride};
    const int64_t bs[3] = {a->q_batch_stride, a->k_batch_stride, a->v_batch_stride};
    const int64_t hs[3] = {a->q_head_stride, a->k_head_stride, a->v_head_stride};
    const int rows[3] = {a->Sq, a->Skv, a->Skv};
    const int boxr[3] = {128, p.BN, p.BN};
    CUtensorMap* maps[3] = {&tmQ, &tmK, &tmV};
    for (int i = 0; i < 3; ++i) {
      MYR_CHECK_ARG(ts[i] % 8 == 0 && bs[i] % 8 == 0 && hs[i] % 8 == 0 && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0,
                    "attention: operand %d strides/pointer must be 16-byte aligned", i);
      dims[0] = (uint64_t)a->dh; dims[1] = (uint64_t)a->H; dims[2] = (uint64_t)rows[i]; dims[3] = (uint64_t)a->B;
      strides[0] = (uint64_t)hs[i] * 2; strides[1] = (uint64_t)ts[i] * 2; strides[2] = (uint64_t)bs[i] * 2;
      box[0] = 64; box[1] = 1; box[2] = (uint32_t)boxr[i]; box[3] = 1;
      int rc = make_tmap_f16(maps[i], ptrs[i], 4, dims, strides, box);
      if (rc) return rc;
    }
  }

  const size_t smem_bytes = (size_t)p.DB * 16384 + p.kp_bytes + p.v_bytes + 1024 + 128;
  const int nch = p.dhp / 32;
  dim3 grid(ceil_div(a->Sq, 128), a->H, a->B);
  static bool attr_set = false;
  if (!attr_set) {
    MYR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MYR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (nch <= 2)
    MYR_CHECK_CUDA(launch_kernel(attn_fwd_kernel<2>, grid, dim3(128), smem_bytes, stream, true, tmQ, tmK, tmV, p));
  else if (nch == 3)
    MYR_CHECK_CUDA(launch_kernel(attn_fwd_kernel<3>, grid, dim3(128), smem_bytes, stream, true, tmQ, tmK, tmV, p));
  else
    MYR_CHECK_CUDA(launch_kernel(attn_fwd_kernel<4>, grid, dim3(128), smem_bytes, stream, true, tmQ, tmK, tmV, p));
  MYR_CHECK_LAUNCH();
  return MYR_OK;
}
