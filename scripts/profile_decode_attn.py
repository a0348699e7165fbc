"""ncu driver: decode attention over a long cache (B = 4, H = 32, 2064 visible tokens), split and single-CTA variants."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
B, H, dh, Smax = 4, 32, 128, 2080
Dl = H * dh
kc = torch.randn(B, Smax, Dl, device=dev).half()
vc = torch.randn(B, Smax, Dl, device=dev).half()
qd = torch.randn(B, 3 * Dl + 16, device=dev).half()
pos = torch.full((B,), 2063, dtype=torch.int32, device=dev)
kvl = torch.full((B,), 2064, dtype=torch.int32, device=dev)
cos = torch.randn(4096, 64, device=dev)
od = torch.empty(B, Dl, device=dev, dtype=torch.float16)
bq = torch.randn(Dl, 8, device=dev).half()
ws = torch.zeros(K.decode_attn_split_bytes(B, H, Smax), device=dev, dtype=torch.uint8)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for i in range(3):
    flush.zero_()
    K.decode_attention(qd, B, H, dh, pos, cos, cos, kc, vc, kvl, od, 1 / math.sqrt(dh), cache_off=2063, lora=(bq, bq, 8, 2.0), split_ws=ws)
    flush.zero_()
    K.decode_attention(qd, B, H, dh, pos, cos, cos, kc, vc, kvl, od, 1 / math.sqrt(dh), cache_off=2063, lora=(bq, bq, 8, 2.0))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, w in (("split", ws), ("single", None)):
    flush.zero_()
    torch.cuda.synchronize()
    e0.record()
    K.decode_attention(qd, B, H, dh, pos, cos, cos, kc, vc, kvl, od, 1 / math.sqrt(dh), cache_off=2063, lora=(bq, bq, 8, 2.0), split_ws=w)
    e1.record()
    torch.cuda.synchronize()
    print(name, "%.1f us cold" % (e0.elapsed_time(e1) * 1e3), "=> %.0f GB/s" % (2 * B * 2064 * Dl * 2 / e0.elapsed_time(e1) / 1e6))
