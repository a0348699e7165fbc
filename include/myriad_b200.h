/*
 * myriad_b200 — C ABI of the B200-native Myriad hot path (libmyriad_b200.so).
 *
 * The reference (tzjtatata/Myriad) is 100 % Python and has NO FFI/plugin ABI: its only
 * plug-in point is the Python class registry (minigpt4/common/registry.py:83-109). This header is therefore
 * the *new* boundary: one extern "C" entry point per fused device op that the registry-registered replacement
 * modules (the minigpt4/models package in this repo) call through ctypes. Every entry point cites the reference
 * code whose eager ATen kernels it replaces.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless stated; caller (PyTorch) owns all
 *     memory, including workspaces; the library keeps no persistent device allocations.
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without syncing.
 *   - return value: 0 on success, negative myr_status on failure; message via myr_last_error().
 *   - dtype codes: MYR_F16 = 0, MYR_F32 = 1.
 *   - "fp16 rounding points" are part of the contract: they reproduce where the reference's CUDA path
 *     (torch.cuda.amp.autocast fp16 / fp16 weights) rounds, so greedy token ids can match bit-exactly.
 */
#ifndef MYRIAD_B200_H_
#define MYRIAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum myr_status {
  MYR_OK = 0,
  MYR_ERR_INVALID = -1, /* bad argument (shape/alignment/dtype) */
  MYR_ERR_CUDA = -2,    /* CUDA runtime/driver error */
  MYR_ERR_UNSUPPORTED = -3,
  MYR_ERR_WORKSPACE = -4 /* workspace too small */
};
enum myr_dtype { MYR_F16 = 0, MYR_F32 = 1 };
enum myr_act { MYR_ACT_NONE = 0, MYR_ACT_GELU_ERF = 1, MYR_ACT_RELU = 2, MYR_ACT_SWIGLU = 3 };

/* ---- library ------------------------------------------------------------------------------------- */
int myr_version(void);                            /* ABI version (integer, bumps on breaking change) */
/* sizeof() of an argument struct as compiled into the library (0 myr_gemm_args, 1 myr_attn_args, 2 myr_norm_args,
 * 3 myr_rope_args, 4 myr_decode_attn_args, 5 myr_mega_op; else 0): lets a binding verify its own struct layout. */
size_t myr_abi_sizeof(int32_t which);
int myr_last_error(char* buf, size_t buf_bytes);  /* copies last error message of calling thread */
int myr_device_sm_count(void);
unsigned long long myr_launch_count(void);        /* kernels launched (or captured into a CUDA graph) by this library */

/* ---- GEMM (tcgen05 + TMA) --------------------------------------------------------------------------
 * out[t, f] = epilogue( sum_k X[t, k] * W[f, k] )          t < T tokens, f < F features
 * Replaces every nn.Linear / F.linear / 1x1- and im2col-conv on the path:
 *   eva_vit.py:124,146 (qkv, proj), eva_vit.py:55-59 (fc1, GELU, fc2), Qformer.py:186-198,286,359,372,
 *   myriad.py:263 (llama_proj), modeling_llama.py:140,179-181,226 (gate/up/down, q/k/v/o), :690 (lm_head).
 * Operands are fp16, accumulation fp32 in tensor memory. W rows sit on the 128 TMEM lanes
 * ("features on lanes"), tokens on the UMMA N side, so the same kernel serves decode (T small) and prefill.
 *   x_mn_major / w_mn_major = 0: operand stored [rows, K] (K contiguous, nn.Linear layout);
 *                           = 1: operand stored [K, rows] (rows contiguous) — used by dgrad/wgrad.
 * Epilogue order (each step optional): v = acc (+ bias[f]); if round_acc: v = fp16(v);
 *   if f < scale_cols: v = fp16(v * scale); if act: v = fp16(act(v)); v *= alpha; if res: v += res[t, f]; store as out_dtype.
 * act = MYR_ACT_SWIGLU (modeling_llama.py:139-140 fused into the gate/up projection): W holds gate and up rows
 *   interleaved in blocks of 64 ([gate 0..63 | up 0..63 | gate 64..127 | ...], F = 2 * I rows) and the kernel writes
 *   out[t, i] = silu(fp16(gate_i)) * fp16(up_i) for i < I (fp16, no bias / residual).
 * Alignment: x, w 16-byte aligned; ldx, ldw (and batch strides) multiples of 8 elements; out / res / bias use 16-byte
 *   vector accesses when their pointers and strides allow it, scalar accesses otherwise.
 * Workspace: COUNTERS (first 64 KiB, int32, MUST be zero when first handed to the library; the library leaves them zero)
 *   followed by fp32 partial tiles of split tiles ("stream-K": every CTA gets an equal, contiguous range of k-blocks; a
 *   tile shared by several CTAs is finished, deterministically, by the last one to arrive).
 */
typedef struct {
  const void* x; int64_t ldx;
  const void* w; int64_t ldw;
  int32_t T, F, K;
  int32_t x_mn_major, w_mn_major;
  const void* bias;            /* fp16 [F] or NULL */
  int32_t act;                 /* myr_act */
  int32_t round_acc;
  int32_t scale_cols; float scale;
  const void* res; int32_t res_dtype; int64_t ldr; /* may alias out (accumulate) */
  void* out; int32_t out_dtype; int64_t ldo;
  void* workspace; size_t workspace_bytes;         /* split-K partials; see myr_gemm_workspace_bytes */
  int32_t bn_hint;             /* token tile (multiple of 16, <= 256), 0 = auto */
  int32_t ksplit_hint;         /* 0 = auto */
  int32_t out_group_rows;      /* 0 = plain; else out row t is stored at (t / rows) * out_group_stride + (t % rows) * ldo: */
  int64_t out_group_stride;    /*   writes token groups straight into a concatenated [B, L, F] buffer (myriad.py:249-266) */
  int32_t nb0, nb1;            /* batched GEMM over nb0 x nb1 independent problems (0 = 1): attention backward per (batch, head) */
  int64_t x_bs0, x_bs1, w_bs0, w_bs1, o_bs0, o_bs1; /* element strides of the two batch dims; no residual / split-K when batched */
  int32_t alpha_set; float alpha; /* if alpha_set: v *= alpha (fp32, unrounded) after the activation, before the residual add */
  int32_t pdl;                 /* launch with programmatic stream serialization (overlaps the previous kernel's tail) */
  int32_t w_static;            /* w is not written by any kernel in flight: its tiles may be prefetched before the PDL wait */
  /* Hand-over of an RMSNorm between two small-batch launches (T <= 4), so that the second one does not have to read and
   * normalise the fp32 stream on its critical path. Producer (the projection whose output + residual IS the stream, e.g.
   * o_proj / down_proj): post_out16[t * post_ld + f] = rn_f16(out[t, f] * post_gamma[f]) and post_ss = { n_parts, -, -, -,
   * sums of out[t, f]^2 over each block of 8 features [n_parts][4] } (post_ss must hold 4 + F / 2 floats). Consumer: x = post_out16 and
   * norm_ss = post_ss (+ norm_eps): the result is multiplied by rsqrt(sum / K + eps) per token in the epilogue. */
  const void* post_gamma; void* post_out16; int64_t post_ld; void* post_ss;
  const void* norm_ss; float norm_eps;
} myr_gemm_args;
/* Small-batch path: T <= 4 with K-major fp16 operands, K % 128 == 0, act NONE or SWIGLU and no scale_cols / round_acc / alpha /
 * row groups / batch / hints runs on the weight-streaming kernel of csrc/gemv.cu (row groups of 16 output rows reduced inside
 * one CTA: no split tiles; the last int of the workspace's counter area hands out row groups, zero on entry and on exit). myr_set_gemv(0) (or env MYR_GEMV=0) routes those shapes to the tcgen05 kernel instead;
 * returns the previous setting. */
int32_t myr_set_gemv(int32_t enabled);
/* Work split of a small-batch launch on a device with `sms` SMs (pure host arithmetic, no GPU needed): out8 = { units of 8
 * output rows (SwiGLU: 8 gate/up pairs), units per 16-row MMA group, CTAs, units [0, u_static) split evenly by CTA index
 * (the rest handed out in groups through the atomic counter), counter used (0/1), ring stages, 1024-k stages per group,
 * bytes of shared memory besides the ring }. */
int myr_gemv_plan(int32_t F, int32_t K, int32_t act, int32_t sms, int32_t have_counter, int32_t* out8);
/* The same for the multi-token streaming kernel (csrc/gemv_mt.cu, 5 <= T <= 32: weights AND tokens go through one TMA ring, row
 * groups of up to 32 output rows / 16 SwiGLU pairs): out6 = { 8-token slices NT = ceil(T / 8), units of 8 rows (8 pairs), units per
 * row group, CTAs, units [0, u_static) split evenly by CTA index, ring stages of 512 k }. */
int myr_gemv_mt_plan(int32_t T, int32_t F, int32_t K, int32_t act, int32_t sms, int32_t have_counter, int32_t* out6);
size_t myr_gemm_workspace_bytes(int32_t T, int32_t F, int32_t K);
int myr_gemm_f16(const myr_gemm_args* args, void* stream);
/* Profiling aid: while set, every GEMM launch writes 148 x 6 %globaltimer stamps (CTA start, predecessor released, last MMA
 * issued, last accumulator ready, epilogue done, -) at `buf` and advances `buf` by that much; NULL stops tracing. */
void myr_gemm_set_trace(void* buf);
/* Programmatic dependent launch on/off for the whole library (default on; env MYR_PDL=0 disables). */
void myr_set_pdl(int32_t enabled);

/* ---- attention (tcgen05 flash forward) ---------------------------------------------------------------
 * out[b, i, h, :] = softmax_j(scale * q[b,i,h,:] . k[b,j,h,:] + mask) v[b,j,h,:], fp16 in/out, fp32 softmax.
 * Replaces eva_vit.py:128-144, Qformer.py:228-265, modeling_llama.py:197-215 (+ the additive masks of
 * :25-54,442-463, never materialised: key j visible to query i iff j < kv_len[b] and (!causal || j <= q_off + i)).
 * q/k/v/out are addressed by element strides (token, batch, head), so fused qkv rows and the pre-allocated KV
 * cache are read in place. All strides and pointers must be 16-byte aligned; dh % 8 == 0, dh <= 128.
 */
typedef struct {
  const void* q; int64_t q_token_stride, q_batch_stride, q_head_stride;
  const void* k; int64_t k_token_stride, k_batch_stride, k_head_stride;
  const void* v; int64_t v_token_stride, v_batch_stride, v_head_stride;
  void* out;     int64_t o_token_stride, o_batch_stride, o_head_stride;
  int32_t B, H, Sq, Skv, dh;
  float scale;
  int32_t causal, q_off;
  const void* kv_len;      /* int32 [B] device (valid keys per batch row) or NULL = Skv */
  int32_t bn_hint;         /* keys per tile (multiple of 32), 0 = auto */
} myr_attn_args;
int myr_attention_fwd(const myr_attn_args* args, void* stream);
/* profiling aid: attention launches after this call write %globaltimer stamps of their CTA (0,0,0) at buf (256 x int64:
 * per KV tile scores seen / probabilities handed over / P.V issued / iteration end); NULL switches it off */
void myr_attn_set_trace(void* buf);

/* ---- LayerNorm / RMSNorm (+ fused LoraAdaptorV2) ----------------------------------------------------
 * y = norm(x [+ W2 (W1 x)]) * gamma (+ beta); statistics in fp32. Replaces blip2.py:119-125 (ln_vision),
 * eva_vit.py:173-180 norm1/norm2 (eps 1e-6), Qformer.py LayerNorm (eps 1e-12), modeling_llama.py:66-74 (rms=1),
 * and with adaptor_w1/w2 set networks.py:81-93 fused in front of ln_vision (myriad.py:248).
 * x: fp16 or fp32 rows; outputs (each optional): out16 (fp16, GEMM operand), out32 (fp32 residual stream),
 * pre32 (fp32 pre-norm value), stats [rows, 2] = (mean, rstd). D % 4 == 0, D <= 4096, rank <= 4.
 */
typedef struct {
  const void* x; int32_t x_dtype; int64_t ldx;
  int32_t rows, D;
  const void* gamma; const void* beta;     /* fp32 [D]; beta NULL for RMSNorm */
  float eps; int32_t rms;
  const void* adaptor_w1; const void* adaptor_w2; int32_t rank; /* fp32 [rank, D], [D, rank] or NULL */
  void* out16; int64_t ld16;
  void* out32; int64_t ld32;
  void* pre32; int64_t ldpre;
  void* stats;
} myr_norm_args;
int myr_norm_fwd(const myr_norm_args* args, void* stream);

/* ---- RoPE + KV-cache append (modeling_llama.py:109-123,190-195) -------------------------------------
 * qkv: [B*S, 3*H*dh] fp16 fused rows; q is rotated in place, rotated k and v are written to cache slot
 * cache_off + s of batch row b. cos/sin: fp32 [max_pos, dh/2]. pos: int32 [B*S]. cache_off_dev (device int32
 * scalar) overrides cache_off when non-NULL (CUDA-graph replay of decode steps).
 */
typedef struct {
  void* qkv; int64_t ldq;
  int32_t B, S, H, dh;
  const void* pos;
  const void* cos_table; const void* sin_table;
  void* kcache; void* vcache; int64_t cache_token_stride, cache_batch_stride;
  int32_t cache_off; const void* cache_off_dev;
  /* optional fused peft LoRA on q_proj / v_proj (myriad.py:171-178; third-party peft semantics y += (alpha/r) B (A x)):
   * the fused qkv GEMM carries the 2r rows of A_q, A_v after the 3*H*dh qkv rows, so xa = x A^T arrives in columns
   * [3*H*dh, 3*H*dh + 2r) of each qkv row; this kernel adds lora_scale * B xa to q and v before rotating / caching. r = 8. */
  const void* lora_bq; const void* lora_bv; int32_t lora_r; float lora_scale;
} myr_rope_args;
int myr_rope_cache(const myr_rope_args* args, void* stream);

/* ---- decode-step attention (one new token per sequence): LoRA-B + RoPE + KV-cache append + attention over the cache
 * in ONE launch, one CTA per (head, batch row). Replaces, for Sq = 1, modeling_llama.py:109-123 (rotary), :190-195 (cache
 * growth by torch.cat), :197-215 (scores / fp32 softmax / PV) and the peft LoRA update of q and v (myriad.py:171-178).
 * qkv: [B, ldq] fp16 rows q | k | v | xa_q | xa_v (see myr_rope_args). The new k / v are written to cache slot
 * cache_off (or *cache_off_dev) of each batch row; kv_len[b] = visible keys including the new one. dh must be 128. */
typedef struct {
  const void* qkv; int64_t ldq;
  int32_t B, H, dh;
  const void* pos;                       /* int32 [B] */
  const void* cos_table; const void* sin_table;
  void* kcache; void* vcache; int64_t cache_token_stride, cache_batch_stride; int32_t cache_len;
  int32_t cache_off; const void* cache_off_dev;
  const void* kv_len;                    /* int32 [B] device */
  const void* lora_bq; const void* lora_bv; int32_t lora_r; float lora_scale;
  float scale;
  void* out; int64_t ldo;                /* fp16 [B, H * dh] */
  int64_t next_layer_stride;             /* elements between this layer's and the next layer's cache (0 = last layer): the
                                            kernel prefetches the next layer's slice into L2 */
  int32_t kv_cap;                        /* 0, or an upper bound (<= 256, <= cache_len) on kv_len for this launch (and every replay
                                            of a graph holding it): K / V then arrive by TMA in 512 * kv_cap bytes of shared memory */
  void* split_ws; size_t split_ws_bytes; /* optional workspace for LONG caches (cache_len > 256, kv_cap == 0, B <= 128): persistent CTAs
                                            then stream the flat list of 128-key chunks of all (row, head) pairs (K / V by TMA, online
                                            softmax); (row, head)s cut between CTAs leave partial (max, sum, P.V) results here, combined
                                            in order by the last one to arrive - deterministic. Needs
                                            myr_decode_attention_ws_bytes(B, H, cache_len) bytes, 16-byte aligned, ZEROED once; the
                                            kernel leaves its counters at zero. NULL: one CTA walks the whole cache of a (row, head). */
} myr_decode_attn_args;
int myr_decode_attention(const myr_decode_attn_args* args, void* stream);
int64_t myr_decode_attention_ws_bytes(int32_t B, int32_t H, int32_t cache_len);

/* ---- persistent decode-step kernel ----------------------------------------------------------------------------------
 * One launch = one greedy-decode step of the whole LLaMA stack for T <= 16 sequences (modeling_llama.py:466-716 with a
 * single new token per sequence): the caller describes the step as a list of ops; myr_mega_plan turns it into a device
 * blob (tensor maps + op records; copy it to the device verbatim, 64-byte aligned) and reports the zero-initialised
 * workspace it needs; myr_mega_launch runs it. Ops execute in list order; op i may read what ops < i wrote.
 *   MYR_MEGA_GEMM : out = epi(x[T,K] w[F,K]^T), weights streamed once (stream-K over all SMs). epi: MYR_MEGA_F16 store fp16,
 *                   MYR_MEGA_RES32 out(fp32) += acc (residual stream, in place), MYR_MEGA_SWIGLU (64-row interleaved gate/up,
 *                   fp16 out [T, F/2]), MYR_MEGA_F32 store fp32.
 *   MYR_MEGA_ATTN : myr_decode_attention semantics (LoRA-B + RoPE + cache append + attention), one CTA per (head, row).
 *   MYR_MEGA_EMBED: h32[t,:] = table[ids[t],:] (embedding gather, ids int32 on the device).
 * Any op may carry a tail executed once the op is complete: norm_dst(fp16)[t,:] = RMSNorm(norm_src(fp32)[t,:]) * gamma for
 * t < norm_rows (modeling_llama.py:66-74) — the fp16 operand of the next GEMM. */
enum myr_mega_kind { MYR_MEGA_GEMM = 0, MYR_MEGA_ATTN = 1, MYR_MEGA_EMBED = 2 };
enum myr_mega_epi { MYR_MEGA_F16 = 0, MYR_MEGA_RES32 = 1, MYR_MEGA_SWIGLU = 2, MYR_MEGA_F32 = 3 };
typedef struct {
  int32_t kind, epi;
  const void* x; int64_t ldx; const void* w; int64_t ldw; int32_t T, F, K;
  void* out; int64_t ldo;
  const void* norm_src; void* norm_dst; const void* gamma; float eps; int32_t D; int32_t norm_rows;
  const void* table; const void* ids; void* h32;
  myr_decode_attn_args attn;
} myr_mega_op;
size_t myr_mega_plan_bytes(int32_t n_ops);
int myr_mega_plan(const myr_mega_op* ops, int32_t n_ops, void* host_blob, size_t blob_bytes, size_t* workspace_bytes,
                  int32_t* n_tile_counters);
/* trace: optional device int64 [3 * n_ops] (%globaltimer ns per op: completed | first CTA saw its input | first CTA drained a
 * tile) for profiling the op chain; NULL in production. */
int myr_mega_launch(const void* dev_blob, int32_t n_ops, void* workspace, size_t workspace_bytes, int32_t n_tile_counters,
                    void* trace, void* stream);

/* SwiGLU modeling_llama.py:139-140: out[t, i] = silu(gate_up[t, i]) * gate_up[t, I + i] (fp16). */
int myr_swiglu(const void* gate_up, int64_t ld_gu, void* out, int64_t ld_out, int32_t T, int32_t I, void* stream);
/* Embedding gather myriad.py:308-311: out[r, :] = table[ids[r], :]; table fp16 [V, D]; ids int32 or int64 (device). */
int myr_embed(const void* table, int32_t D, const void* ids, int32_t ids_are_int64, int32_t n, void* out,
              int32_t out_dtype, int64_t ld_out, void* stream);
/* Strided row copy / cast (places token groups into the concatenated buffers of myriad.py:249-266,372). */
int myr_copy_rows(const void* src, int32_t src_dtype, int64_t src_ld, int64_t src_group_stride, void* dst,
                  int32_t dst_dtype, int64_t dst_ld, int64_t dst_group_stride, int32_t groups, int32_t rows_per_group,
                  int32_t D, void* stream);
/* ViT token assembly eva_vit.py:326-331: x[b,0] = cls + pos[0]; x[b,1+p] = patch[b,p] + pos[1+p] (fp32). */
int myr_vit_assemble(const void* patch, const void* cls, const void* pos, void* x, int32_t B, int32_t N, int32_t D,
                     void* stream);
/* ViT patchify eva_vit.py:196-203: image fp32 [B,C,HW,HW] -> fp16 [B*g*g, ldp] columns (c, ky, kx), pad zeroed. */
int myr_patchify(const void* image, void* patches, int32_t B, int32_t C, int32_t HW, int32_t P, int32_t ldp, void* stream);

/* ---- expert-prior conv stacks (networks.py:95-197), NHWC fp16 ---------------------------------------- */
/* conv3x3(pad 1) + bias + ReLU + maxpool2 fused; in fp32 or fp16 [B,H,W,Cin]; w fp32 [Cout,3,3,Cin]; out fp16. */
int myr_conv3x3_relu_pool(const void* in, int32_t in_dtype, const void* w, const void* bias, void* out, int32_t B,
                          int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream);
/* im2col on NHWC fp16: out [(b,oy,ox), (ky,kx,c)], zero padding `pad`, stride 1. */
int myr_im2col(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t KH, int32_t KW, int32_t pad,
               void* stream);
int myr_maxpool2(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

/* ---- greedy decoding step (HF generate greedy search + conversation.py:96-107), all state on device ------
 * state (int32): [0] step [1] done [2] cache_off [3] - | unfinished[B] | cur_tok[B] | kv_len[B] | pos[B] |
 * tokens[B, max_new_tokens]. scratch: int32 [B]. stop_seqs: int32 [n_stops, stop_max_len], -1 padded.
 */
int myr_greedy_step(const void* logits, int64_t ld_logits, int32_t B, int32_t V, void* state, void* scratch,
                    int32_t max_new_tokens, int32_t min_new_tokens, int32_t eos, const void* stop_seqs, int32_t n_stops,
                    int32_t stop_max_len, void* stream);

/* ==== training (backward) entry points =================================================================
 * Gradients flow loss -> lm_head -> 32 LLaMA layers -> inputs_embeds -> {VETokenizer, base_prompts, llama_proj ->
 * Q-Former -> {VEInstructor, ln_vision -> LoraAdaptorV2}} (everything else is frozen: dgrad only). Matrix products
 * reuse myr_gemm_f16 with MN-major operands; attention backward = batched myr_gemm_f16 around the two softmax kernels.
 */
/* clamp_CE_loss modeling_llama.py:718-728 on [R, V] fp32 logits, labels int64 [R] (-100 = ignored).
 * fwd: row_loss [R], stats [R,2] = (max, sumexp), loss_out[2] = (mean loss, #supervised rows).
 * bwd: dlogits fp16 [R, V] = loss_scale / count * (softmax - onehot), zero where p_y is clamped or the row ignored. */
int myr_clamp_ce_fwd(const void* logits, int64_t ld, int32_t R, int32_t V, const void* labels, void* row_loss, void* stats,
                     void* loss_out, void* stream);
int myr_clamp_ce_bwd(const void* logits, int64_t ld, int32_t R, int32_t V, const void* labels, const void* stats,
                     const void* loss_out, float loss_scale, void* dlogits, int64_t ldd, void* stream);
/* LayerNorm / RMSNorm input gradient (frozen gamma/beta): x fp32 pre-norm rows, dy fp16|fp32, out = dx (+ add). */
int myr_norm_bwd(const void* x, int64_t ldx, const void* dy, int32_t dy_dtype, int64_t lddy, const void* gamma, float eps,
                 int32_t rms, int32_t rows, int32_t D, const void* add, int64_t ldadd, void* out32, int64_t ldo, void* out16,
                 int64_t ldo16, void* stream);
int myr_swiglu_bwd(const void* gate_up, int64_t ld_gu, const void* dact, int64_t ld_da, void* dgu, int64_t ld_dgu, int32_t T,
                   int32_t I, void* stream);
/* LoRA dropout (peft lora_dropout = 0.05, reference myriad.py:175; applied to the input of each LoRA branch while training).
 * Counter-based: element i of [rows, D] is kept iff hash(seed, offset + i) >= p * 2^32 and scaled by 1 / (1 - p); the backward
 * (acc += mask * g / (1 - p)) regenerates the mask from the same (seed, offset). myr_dropout_mask exports the 0/1 mask. */
int myr_dropout_fwd(const void* x_f16, int64_t ldx, void* out_f16, int64_t ldo, int32_t rows, int32_t D, float p, uint64_t seed,
                    uint64_t offset, void* stream);
int myr_dropout_bwd_add(const void* g_f32, int64_t ldg, void* acc_f32, int64_t lda, int32_t rows, int32_t D, float p, uint64_t seed,
                        uint64_t offset, void* stream);
int myr_dropout_mask(void* out_u8, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
int myr_gelu_fwd(const void* pre, void* out, int64_t n, void* stream);
int myr_gelu_bwd(const void* pre, const void* dy, void* dpre, int64_t n, void* stream);
int myr_rope_bwd(void* dqkv, int64_t ld, int32_t T, int32_t H, int32_t dh, const void* pos, const void* cos_table,
                 const void* sin_table, void* stream);
/* P = softmax(scale * S + mask) on materialised score rows [B*H*Sq, cols] (fp32 -> fp16); dS = scale * P * (dP - sum dP P). */
int myr_softmax_rows(const void* S, int64_t lds, void* P, int64_t ldp, int32_t B, int32_t H, int32_t Sq, int32_t Skv, int32_t cols,
                     float scale, int32_t causal, const void* kv_len, void* stream);
int myr_softmax_bwd_rows(const void* P, int64_t ldp, const void* dP, int64_t lddp, void* dS, int64_t lds, int64_t n_rows,
                         int32_t cols, float scale, void* stream);
/* Fused attention backward for short sequences (dh 64 / 128, Skv <= 256): dQ, dK, dV from Q, K, V, dO in one launch, the
 * scores never leave the SM. Replaces the autograd backward of the eager softmax(Q K^T * scale + mask) V of
 * modeling_llama.py:196-222 / Qformer.py:180-260. Operands fp16 with head stride dh, token stride *_ts and row (batch) stride *_bs
 * in elements (multiples of 8); mask as myr_softmax_rows (causal needs Sq == Skv; kv_len int32 [B] or NULL). */
int myr_attn_bwd_small_supported(int32_t Sq, int32_t Skv, int32_t dh);
int myr_attn_bwd_small(const void* q, int64_t q_ts, int64_t q_bs, const void* k, int64_t k_ts, int64_t k_bs, const void* v,
                       int64_t v_ts, int64_t v_bs, const void* dO, int64_t do_ts, int64_t do_bs, void* dq, int64_t dq_ts,
                       int64_t dq_bs, void* dk, int64_t dk_ts, int64_t dk_bs, void* dv, int64_t dv_ts, int64_t dv_bs, int32_t B,
                       int32_t H, int32_t Sq, int32_t Skv, int32_t dh, float scale, int32_t causal, const void* kv_len,
                       void* stream);
/* dst[r] = src[idx[r]] (scatter = 0) or dst[idx[r]] = src[r] (scatter = 1), with dtype cast; idx int32 device. */
int myr_index_rows(const void* src, int32_t src_dtype, int64_t src_ld, void* dst, int32_t dst_dtype, int64_t dst_ld,
                   const void* idx, int32_t R, int32_t D, int32_t scatter, void* stream);
/* out[c] (+)= scale * sum over groups x rows of src (bias and base_prompts gradients), deterministic. */
int myr_colsum(const void* src, int32_t src_dtype, int64_t ld, int64_t group_stride, int32_t groups, int32_t rows, int32_t D,
               float scale, void* out, int32_t accumulate, void* stream);
/* LoraAdaptorV2 (networks.py:81-93) weight gradients; scratch fp32 [rows, 2*rank]. */
int myr_adaptor_bwd(const void* x, const void* dy, const void* w1, const void* w2, void* scratch, void* dw1, void* dw2,
                    int32_t rows, int32_t D, int32_t rank, float scale, void* stream);
/* fused AdamW over flat fp32 buffers (runner_base.py:105-139), gradient unscale + inf/nan skip (GradScaler) folded in. */
int myr_adamw_step(void* params, const void* grads, void* exp_avg, void* exp_avg_sq, const void* wd_mask, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float inv_scale, void* found_inf,
                   void* stream);
int myr_memset_zero(void* ptr, size_t bytes, void* stream);
/* conv stacks, training side (networks.py:98-127,159-188): unfused conv+ReLU keeping the pre-pool map, pool+ReLU backward,
 * direct wgrad (+bias) / dgrad for small channel counts, col2im for the im2col/GEMM layers. NHWC, w [Cout,3,3,Cin] fp32. */
int myr_conv3x3_relu(const void* in, int32_t in_dtype, const void* w, const void* bias, void* out, int32_t B, int32_t H, int32_t W,
                     int32_t Cin, int32_t Cout, void* stream);
int myr_pool_relu_bwd(const void* y, const void* dpool, int32_t dpool_dtype, void* dy, int32_t B, int32_t H, int32_t W, int32_t C,
                      void* stream);
int myr_conv3x3_wgrad(const void* in, int32_t in_dtype, const void* dy, void* dw, void* db, int32_t B, int32_t H, int32_t W,
                      int32_t Cin, int32_t Cout, float scale, void* stream);
int myr_conv3x3_dgrad(const void* dy, const void* w, void* din, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                      void* stream);
int myr_col2im(const void* dcols, void* din, int32_t B, int32_t H, int32_t W, int32_t C, int32_t KH, int32_t KW, int32_t pad,
               void* stream);


/* ---- LoRA branches of q_proj / v_proj in training (peft, myriad.py:171-178: y += s * B (A dropout(x)), rank 8), CUDA-core kernels: as
 * tcgen05 GEMMs with F = 8 or K = 8 these cost 15-20 us per launch. Dropout masks are regenerated from (seed, offset + t * D + d) as in
 * myr_dropout_fwd (p = 0: none); off_q / off_v are the mask streams of the two branches. All summation orders are fixed.
 * fwd: xa fp32 [T, 16] = [drop_q(x1) A_q^T | drop_v(x1) A_v^T] (A fp16 [16, D] = A_q rows then A_v rows), then
 *      qkv[t, col_q + d] += s * xa[t, 0:8] . B_q[d, :],  qkv[t, col_v + d] += s * xa[t, 8:16] . B_v[d, :]   (qkv fp16, B fp16 [D, 8]).
 * bwd: d_xa (scratch fp32 [T, 16]) = s * dY B;  dB_j fp32 [D, 8] = s * inv_scale * dY_j^T xa_j;  dA fp32 [16, D] = inv_scale * d_xa^T drop(x1);
 *      dx1 fp32 [T, D] += mask_j * d_xa_j A_j (branch q first, then v). dY_j = dqkv[:, col_j : col_j + D] (fp16). */
int myr_lora_fwd(const void* x1, int64_t ldx, const void* A, const void* bq, const void* bv, void* xa, void* qkv, int64_t ldq, int64_t col_q,
                 int64_t col_v, int32_t T, int32_t D, int32_t r, float s, float p, uint64_t seed, uint64_t off_q, uint64_t off_v, void* stream);
int myr_lora_bwd(const void* dqkv, int64_t ldq, int64_t col_q, int64_t col_v, const void* xa, const void* A, const void* bq, const void* bv,
                 const void* x1, int64_t ldx, void* dxa_scratch, void* dbq, void* dbv, void* dA, void* dx1, int64_t ldd, int32_t T, int32_t D,
                 int32_t r, float s, float inv_scale, float p, uint64_t seed, uint64_t off_q, uint64_t off_v, void* stream);

/* ---- vision expert heads (adrefexpert_v2.py:245-301); the ImageBind-Huge trunk (imagebind_model.py:486-504,
 * transformer.py:104-170) runs on myr_patchify / myr_gemm_f16 / myr_norm / myr_attention_fwd ----------------------------------- */
/* tokens[i].transpose(0,1)[:, 1:, :] (adrefexpert_v2.py:26-27,215-216): x fp32 [B, N, D] -> out fp16 [B, N-1, D] without the class
 * token; normalize != 0: every row divided by max(|row|, 1e-8) (operand of F.cosine_similarity, :270). */
int myr_expert_tap(const void* x, void* out16, int32_t B, int32_t N, int32_t D, int32_t normalize, void* stream);
/* zero-shot logits :285-286: logits[b*P + p][k] = scale * tokens[b*P + p] . text[b][k] / |tokens[b*P + p]|; tokens fp32 (row stride
 * ld), text fp32 [B, 2, C], logits fp32 [B*P, 2]. */
int myr_expert_logits(const void* tokens, int64_t ld, const void* text, void* logits, int32_t B, int32_t P, int32_t C, float scale,
                      void* stream);
/* :287-301: logits fp32 [L, B, G*G, 2] -> maps fp32 [B, OUT*OUT] = mean over L of softmax(bilinear_align_corners(logits))[:, 1],
 * masks fp32 [B, G*G] = mean over L of softmax(logits)[:, 1]. */
int myr_expert_maps(const void* logits, void* maps, void* masks, int32_t L, int32_t B, int32_t G, int32_t OUT, void* stream);
/* k-shot :270-272: acc[row] (+)= weight * max_r S[row][r]; S fp32 [rows, R] (row stride ld) = cosine similarities. */
int myr_expert_rowmax(const void* S, int64_t ld, void* acc, int32_t rows, int32_t R, float weight, int32_t accumulate, void* stream);
/* k-shot :274-278: sim fp32 [B, G*G] -> simmask [B, G*G] = 1 - sim, maps [B, OUT*OUT] = 1 - bilinear_align_corners(sim). */
int myr_expert_sim_maps(const void* sim, void* maps, void* simmask, int32_t B, int32_t G, int32_t OUT, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MYRIAD_B200_H_ */
