"""B200: decode attention over long caches (csrc/decode_attn_stream.cu vs the one-CTA-per-(head, row) kernel of decode_attn.cu):
us per launch inside a CUDA graph that rotates over 6 caches (> L2), i.e. every launch reads its K / V from HBM as a decode
step does (32 layers, 32 different caches)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
H, dh = 32, 128
Dl = H * dh
PEAK = 6551.0


def graph_time(fn, n_rot, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(n_rot):
            fn(i)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n_rot):
                fn(i)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n_rot)


def main():
    print("| B | visible keys | stream us | GB/s | frac of %d | one CTA per (head, row) us |" % PEAK)
    print("|---:|---:|---:|---:|---:|---:|")
    for B, S in ((4, 272), (4, 1040), (4, 2064), (16, 272), (16, 2064), (32, 272), (32, 1040), (32, 2064)):
        Smax = 2080
        n_rot = 6 if B <= 16 else 3
        kcs = [torch.randn(B, Smax, Dl, device=dev).half() for _ in range(n_rot)]
        vcs = [torch.randn(B, Smax, Dl, device=dev).half() for _ in range(n_rot)]
        qd = torch.randn(B, 3 * Dl + 16, device=dev).half()
        pos = torch.full((B,), S - 1, dtype=torch.int32, device=dev)
        kvl = torch.full((B,), S, dtype=torch.int32, device=dev)
        cos = torch.randn(4096, 64, device=dev)
        od = torch.empty(B, Dl, device=dev, dtype=torch.float16)
        bq = torch.randn(Dl, 8, device=dev).half()
        ws = torch.zeros(K.decode_attn_split_bytes(B, H, Smax), device=dev, dtype=torch.uint8)

        def run(i, w):
            K.decode_attention(qd, B, H, dh, pos, cos, cos, kcs[i], vcs[i], kvl, od, 1 / math.sqrt(dh), cache_off=S - 1,
                               lora=(bq, bq, 8, 2.0), split_ws=w)

        t1 = graph_time(lambda i: run(i, ws), n_rot)
        t0 = graph_time(lambda i: run(i, None), n_rot)
        gbs = 2 * B * S * Dl * 2 / t1 / 1e3
        print("| %d | %d | %.1f | %.0f | %.3f | %.1f |" % (B, S, t1, gbs, gbs / PEAK, t0))
        del kcs, vcs


if __name__ == "__main__":
    main()
