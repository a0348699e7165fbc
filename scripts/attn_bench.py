"""Flash-attention forward: the warp-specialised kernel (csrc/attention2.cu, default) against the round-1 kernel
(csrc/attention.cu, MYR_ATTN2=0) and torch SDPA on the shapes of the path. CUDA-graph timed (launches back to back on the device).
Run under gpurun: python scripts/attn_bench.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
SHAPES = [  # name, B, H, Sq, Skv, dh, causal
    ("vit B=4", 4, 16, 257, 257, 88, False),
    ("vit B=8", 8, 16, 257, 257, 88, False),
    ("qf cross B=4", 4, 12, 81, 257, 64, False),
    ("llama prefill S=131", 4, 32, 131, 131, 128, True),
    ("llama S=1024", 4, 32, 1024, 1024, 128, True),
    ("llama S=2048", 4, 32, 2048, 2048, 128, True),
    ("llama S=2048 full", 4, 32, 2048, 2048, 128, False),
]


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for name, B, H, Sq, Skv, dh, causal in SHAPES:
    torch.manual_seed(0)
    q = torch.randn(B, Sq, H, dh, device=dev).half()
    k = torch.randn(B, Skv, H, dh, device=dev).half()
    v = torch.randn(B, Skv, H, dh, device=dev).half()
    out = torch.empty_like(q)
    scale = 1.0 / math.sqrt(dh)
    st = lambda t: (t.stride(1), t.stride(0), t.stride(2))

    def ours():
        K.attention(q, k, v, out, B, H, Sq, Skv, dh, scale, st(q), st(k), st(v), st(out), causal=causal)

    if "--trace" in sys.argv and Sq >= 1024:
        import ctypes
        os.environ["MYR_ATTN2"] = "1"
        tb = torch.zeros(256, dtype=torch.int64, device=dev)
        ours()
        torch.cuda.synchronize()
        K.lib().myr_attn_set_trace(ctypes.c_void_p(tb.data_ptr()))
        ours()
        K.lib().myr_attn_set_trace(ctypes.c_void_p(0))
        torch.cuda.synchronize()
        t = tb.cpu()
        t0 = int(t[0])
        f = lambda a: " ".join("%.2f" % ((int(v) - t0) / 1e3) for v in a[:16])
        print("  S0 seen   us:", f(t[0:64]))
        print("  P0 handed us:", f(t[64:128]))
        print("  PV0 issue us:", f(t[128:192]))
        print("  iter end  us:", f(t[192:256]), flush=True)
    res = {}
    outs = {}
    for label, env in (("new", "1"), ("old", "0")):
        os.environ["MYR_ATTN2"] = env
        res[label] = timed(ours)
        outs[label] = out.clone()
    os.environ["MYR_ATTN2"] = "1"
    qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
    ref = torch.nn.functional.scaled_dot_product_attention(qt.float(), kt.float(), vt.float(), is_causal=causal, scale=scale).transpose(1, 2)
    err = {l: ((outs[l].float() - ref).abs().max() / ref.abs().max()).item() for l in outs}
    res["sdpa"] = timed(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal, scale=scale))
    fl = 4.0 * B * H * Sq * Skv * dh * (0.5 if causal and Sq == Skv else 1.0)
    print("%-22s new %8.1f us %6.1f TF/s (err %.1e) | old %8.1f us %6.1f TF/s (err %.1e) | torch sdpa %8.1f us %6.1f TF/s" % (
        name, res["new"], fl / res["new"] / 1e6, err["new"], res["old"], fl / res["old"] / 1e6, err["old"], res["sdpa"],
        fl / res["sdpa"] / 1e6), flush=True)
