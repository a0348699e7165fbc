"""Small driver for `ncu --set full` captures of the hot kernels in their bench configurations (run under gpurun):
   ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 8 -o gpurun_out/prof_gemm python scripts/profile_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 4
# decode weight-streaming GEMMs (T = 4): gate_up and down of LLaMA-7B, distinct weights per launch (no L2 reuse)
x = torch.randn(B, 4096, device=dev).half()
a = torch.randn(B, 11008, device=dev).half()
res = torch.zeros(B, 4096, device=dev)
for i in range(3):
    wgu = (torch.randn(22016, 4096, device=dev) * 0.02).half()
    wd = (torch.randn(4096, 11008, device=dev) * 0.02).half()
    K.gemm(x, wgu)
    K.gemm(a, wd, res=res, out=res)
# ViT GEMMs at the bench batch (T = 4 * 257)
T = B * 257
h = torch.randn(T, 1408, device=dev).half()
w1 = (torch.randn(6144, 1408, device=dev) * 0.02).half()
b1 = torch.zeros(6144, device=dev).half()
for i in range(2):
    K.gemm(h, w1, bias=b1, act=K.ACT_GELU)
# ViT attention (dh = 88, N = 257) and LLaMA prefill attention (dh = 128, S = 131, causal)
D = 1408
qkv = torch.randn(T, 3 * D, device=dev).half()
out = torch.empty(T, D, device=dev, dtype=torch.float16)
s = (3 * D, 257 * 3 * D, 88)
for i in range(2):
    K.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], out, B, 16, 257, 257, 88, 1.0, s, s, s, (D, 257 * D, 88))
torch.cuda.synchronize()
print("done")
