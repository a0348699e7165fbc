"""ncu driver for decode_attn_stream_kernel: one launch at B x S given on the command line (default 32 x 2064), caches from HBM.
   ncu --set full --clock-control none --import-source on -k regex:decode_attn_stream -s 2 -c 1 -o /tmp/ds python scripts/stream_attn_ncu.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2064
H, dh, Smax = 32, 128, 2080
Dl = H * dh
kcs = [torch.randn(B, Smax, Dl, device=dev).half() for _ in range(3)]
vcs = [torch.randn(B, Smax, Dl, device=dev).half() for _ in range(3)]
qd = torch.randn(B, 3 * Dl + 16, device=dev).half()
pos = torch.full((B,), S - 1, dtype=torch.int32, device=dev)
kvl = torch.full((B,), S, dtype=torch.int32, device=dev)
cos = torch.randn(4096, 64, device=dev)
od = torch.empty(B, Dl, device=dev, dtype=torch.float16)
bq = torch.randn(Dl, 8, device=dev).half()
ws = torch.zeros(K.decode_attn_split_bytes(B, H, Smax), device=dev, dtype=torch.uint8)
for i in range(4):
    K.decode_attention(qd, B, H, dh, pos, cos, cos, kcs[i % 3], vcs[i % 3], kvl, od, 1 / math.sqrt(dh), cache_off=S - 1, lora=(bq, bq, 8, 2.0),
                       split_ws=ws)
torch.cuda.synchronize()
print("done")
