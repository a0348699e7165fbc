"""ncu driver: one flash-attention forward launch at LLaMA S = 2048 (B = 4, H = 32, dh = 128), full or causal (argv[1])."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
B, H, S, dh = 4, 32, 2048, 128
causal = len(sys.argv) > 1 and sys.argv[1] == "causal"
q = torch.randn(B, S, H, dh, device=dev).half()
k = torch.randn(B, S, H, dh, device=dev).half()
v = torch.randn(B, S, H, dh, device=dev).half()
out = torch.empty_like(q)
st = lambda t: (t.stride(1), t.stride(0), t.stride(2))
for _ in range(3):
    K.attention(q, k, v, out, B, H, S, S, dh, 1.0 / math.sqrt(dh), st(q), st(k), st(v), st(out), causal=causal)
torch.cuda.synchronize()
