"""ncu driver for decode_attn_tma_kernel at the benchmark's shape (B = 4, H = 32, 163-slot cache):
   ncu --set full --clock-control none --import-source on -k regex:decode_attn_tma -s 2 -c 1 -o /tmp/da python scripts/tma_attn_ncu.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
B, H, dh, Smax, off = 4, 32, 128, 163, 150
Dl = H * dh
kc = torch.randn(B, Smax, Dl, device=dev).half()
vc = torch.randn(B, Smax, Dl, device=dev).half()
qd = torch.randn(B, 3 * Dl + 16, device=dev).half()
pos = torch.full((B,), off, dtype=torch.int32, device=dev)
kvl = torch.full((B,), off + 1, dtype=torch.int32, device=dev)
cos = torch.randn(4096, 64, device=dev)
od = torch.empty(B, Dl, device=dev, dtype=torch.float16)
bq = torch.randn(Dl, 8, device=dev).half()
for i in range(4):
    K.decode_attention(qd, B, H, dh, pos, cos, cos, kc, vc, kvl, od, 1 / math.sqrt(dh), cache_off=off, lora=(bq, bq, 8, 2.0), kv_cap=Smax)
torch.cuda.synchronize()
print("done")
