"""Small driver for `ncu --set full` captures of the hot kernels in their bench configurations (run under gpurun):
   ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 12 -o gpurun_out/prof_gemm python scripts/profile_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math

import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 4
# decode weight streaming (T = 4, gemv_kernel) in layer order: norm+qkv (+LoRA A rows), o (+res), norm+gate/up (+SwiGLU), down (+res);
# fresh weights per launch (no L2 reuse)
x = torch.randn(B, 4096, device=dev).half()
a = torch.randn(B, 11008, device=dev).half()
res = torch.randn(B, 4096, device=dev)
qkv = torch.empty(B, 12304, device=dev, dtype=torch.float16)
act = torch.empty(B, 11008, device=dev, dtype=torch.float16)
for i in range(2):
    wq = (torch.randn(12304, 4096, device=dev) * 0.02).half()
    wo = (torch.randn(4096, 4096, device=dev) * 0.02).half()
    wgu = (torch.randn(22016, 4096, device=dev) * 0.02).half()
    wd = (torch.randn(4096, 11008, device=dev) * 0.02).half()
    gamma = torch.ones(4096, device=dev)
    ya, yb = torch.zeros(B, 4096, device=dev, dtype=torch.float16), torch.zeros(B, 4096, device=dev, dtype=torch.float16)
    ssa, ssb = torch.zeros(K.NORM_SS_FLOATS, device=dev), torch.zeros(K.NORM_SS_FLOATS, device=dev)
    K.gemm(a, wd, res=res, out=res, w_static=True, post_norm=(gamma, yb, ssb))  # small-batch kernel (gemv.cu), RMSNorm hand-over
    K.gemm(yb, wq, out=qkv, w_static=True, norm_ss=(ssb, 1e-6))
    K.gemm(x, wo, res=res, out=res, w_static=True, post_norm=(gamma, ya, ssa))
    K.gemm(ya, wgu, act=K.ACT_SWIGLU, out=act, w_static=True, norm_ss=(ssa, 1e-6))
    K.gemm(a, wd, res=res, out=res, w_static=True, post_norm=(gamma, yb, ssb))
# ViT GEMMs at the bench batch (T = 4 * 257): qkv, fc1 + GELU, fc2 + residual
T = B * 257
h = torch.randn(T, 1408, device=dev).half()
w1 = (torch.randn(6144, 1408, device=dev) * 0.02).half()
b1 = torch.zeros(6144, device=dev).half()
w2 = (torch.randn(1408, 6144, device=dev) * 0.02).half()
b2 = torch.zeros(1408, device=dev).half()
xr = torch.zeros(T, 1408, device=dev)
m = K.gemm(h, w1, bias=b1, act=K.ACT_GELU)
K.gemm(m, w2, bias=b2, res=xr, out=xr)
# LLaMA prefill GEMMs (T = 524): gate/up + SwiGLU and qkv on the CTA-pair kernel (gemm2.cu), o_proj + residual (few tiles: gemm.cu)
hp = torch.randn(524, 4096, device=dev).half()
K.gemm(hp, wgu, act=K.ACT_SWIGLU)
K.gemm(hp, wq)
rp = torch.zeros(524, 4096, device=dev)
K.gemm(hp, wo, res=rp, out=rp)
# Q-Former cross-attention K/V projection of all six cross layers (T = 1028, F = 9216)
wkv = (torch.randn(9216, 1408, device=dev) * 0.02).half()
K.gemm(h, wkv, bias=torch.zeros(9216, device=dev).half())
# ViT attention (dh = 88, N = 257), LLaMA prefill attention (dh = 128, S = 131, causal)
D = 1408
qkv = torch.randn(T, 3 * D, device=dev).half()
out = torch.empty(T, D, device=dev, dtype=torch.float16)
s = (3 * D, 257 * 3 * D, 88)
for i in range(2):
    K.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], out, B, 16, 257, 257, 88, 1.0, s, s, s, (D, 257 * D, 88))
S, Dl = 131, 4096
ql = torch.randn(B * S, 3 * Dl, device=dev).half()
ol = torch.empty(B * S, Dl, device=dev, dtype=torch.float16)
sl = (3 * Dl, S * 3 * Dl, 128)
K.attention(ql, ql[:, Dl:], ql[:, 2 * Dl:], ol, B, 32, S, S, 128, 1 / math.sqrt(128), sl, sl, sl, (Dl, S * Dl, 128), causal=True)
# long-sequence prefill attention of the sweep (BASELINE configs[4]): S = 2048, causal and full
S2 = 2048
q2 = torch.randn(B * S2, 3 * Dl, device=dev).half()
o2 = torch.empty(B * S2, Dl, device=dev, dtype=torch.float16)
s2 = (3 * Dl, S2 * 3 * Dl, 128)
for causal in (True, False):
    K.attention(q2, q2[:, Dl:], q2[:, 2 * Dl:], o2, B, 32, S2, S2, 128, 1 / math.sqrt(128), s2, s2, s2, (Dl, S2 * Dl, 128), causal=causal)
# decode attention (one launch per layer and step): 4 sequences, 32 heads, 160 cached tokens
kc = torch.randn(B, 256, Dl, device=dev).half()
vc = torch.randn(B, 256, Dl, device=dev).half()
qd = torch.randn(B, 3 * Dl + 16, device=dev).half()
pos = torch.full((B,), 160, dtype=torch.int32, device=dev)
kvl = torch.full((B,), 161, dtype=torch.int32, device=dev)
cos = torch.randn(2048, 64, device=dev)
od = torch.empty(B, Dl, device=dev, dtype=torch.float16)
bq = torch.randn(Dl, 8, device=dev).half()
for i in range(2):
    K.decode_attention(qd, B, 32, 128, pos, cos, cos, kc, vc, kvl, od, 1 / math.sqrt(128), cache_off=160, lora=(bq, bq, 8, 2.0))
    # the decode step's variant: K / V tiles by TMA into shared memory (visible cache bounded by 168 rows)
    K.decode_attention(qd, B, 32, 128, pos, cos, cos, kc, vc, kvl, od, 1 / math.sqrt(128), cache_off=160, lora=(bq, bq, 8, 2.0), kv_cap=168)
# decode projections at the sweep's batch 16 (gemv_mt_kernel): qkv and gate/up + SwiGLU
x16 = torch.randn(16, 4096, device=dev).half()
for i in range(2):
    K.gemm(x16, wq, w_static=True)
    K.gemm(x16, wgu, act=K.ACT_SWIGLU, w_static=True)
# decode attention over a long cache (decode_attn_stream_kernel): 4 sequences, 2064 visible tokens
Sl = 2080
kcl = torch.randn(B, Sl, Dl, device=dev).half()
vcl = torch.randn(B, Sl, Dl, device=dev).half()
posl = torch.full((B,), 2063, dtype=torch.int32, device=dev)
kvll = torch.full((B,), 2064, dtype=torch.int32, device=dev)
cosl = torch.randn(4096, 64, device=dev)
wsl = torch.zeros(K.decode_attn_split_bytes(B, 32, Sl), device=dev, dtype=torch.uint8)
for i in range(2):
    K.decode_attention(qd, B, 32, 128, posl, cosl, cosl, kcl, vcl, kvll, od, 1 / math.sqrt(128), cache_off=2063, lora=(bq, bq, 8, 2.0),
                       split_ws=wsl)
torch.cuda.synchronize()
print("done")
