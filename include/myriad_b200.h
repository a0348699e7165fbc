/*
 * myriad_b200 — C ABI of the B200-native Myriad hot path (libmyriad_b200.so).
 *
 * The reference (tzjtatata/Myriad) is 100 % Python and has NO FFI/plugin ABI: its only
 * plug-in point is the Python class registry (minigpt4/common/registry.py:83-109). This header is therefore
 * the *new* boundary: one extern "C" entry point per fused device op that the registry-registered replacement
 * modules (minigpt4/models/*.py in this repo) call through ctypes. Every entry point cites the reference
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
enum myr_act { MYR_ACT_NONE = 0, MYR_ACT_GELU_ERF = 1 };

/* ---- library ------------------------------------------------------------------------------------- */
int myr_version(void);                            /* ABI version (integer, bumps on breaking change) */
int myr_last_error(char* buf, size_t buf_bytes);  /* copies last error message of calling thread */
int myr_device_sm_count(void);

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
 *   if f < scale_cols: v = fp16(v * scale); if act: v = fp16(act(v)); if res: v += res[t, f]; store as out_dtype.
 * Alignment: x, w, out 16-byte aligned; ldx, ldw multiples of 8 elements; K multiple of 8.
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
} myr_gemm_args;
size_t myr_gemm_workspace_bytes(int32_t T, int32_t F, int32_t K);
int myr_gemm_f16(const myr_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MYRIAD_B200_H_ */
