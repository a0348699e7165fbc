"""Decode weight-streaming chain (T = 4, LLaMA-7B shapes, a different matrix per launch so nothing is re-served from L2):
small-batch kernel (csrc/gemv.cu) vs the tcgen05 kernel, per shape and as the 4-GEMM layer chain. Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import kernels as K
dev = torch.device("cuda:0")
torch.manual_seed(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NL = 8
def ev():
    return torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
def mk(F, Kd):
    return [(torch.randn(F, Kd, device=dev) * 0.02).half() for _ in range(NL)]
wq, wo, wgu, wd = mk(12304, 4096), mk(4096, 4096), mk(22016, 4096), mk(4096, 11008)
x = torch.randn(T, 4096, device=dev).half()
a = torch.randn(T, 11008, device=dev).half()
h = torch.zeros(T, 4096, device=dev)
gamma = torch.ones(4096, device=dev)
qkv = torch.empty(T, 12304, device=dev, dtype=torch.float16)
act = torch.empty(T, 11008, device=dev, dtype=torch.float16)
for gemv in (True, False):
    K.set_gemv(gemv)
    name = "gemv  " if gemv else "tcgen05"
    for nm, ws, fn in (("qkv 100.8MB", wq, lambda w: K.gemm(x, w, out=qkv, w_static=True)),
                       ("o 33.6MB", wo, lambda w: K.gemm(x, w, res=h, out=h, w_static=True)),
                       ("gate/up 180.4MB", wgu, lambda w: K.gemm(x, w, act=K.ACT_SWIGLU, out=act, w_static=True)),
                       ("down 90.2MB", wd, lambda w: K.gemm(a, w, res=h, out=h, w_static=True))):
        def chain():
            for w in ws:
                fn(w)
        t = timeit(chain) / NL
        mb = ws[0].numel() * 2 / 1e6
        print("%s %-16s %6.2f us/launch  %.2f TB/s" % (name, nm, t, mb / t), flush=True)
    def layer_chain():
        for i in range(NL):
            K.gemm(x, wq[i], out=qkv, w_static=True)
            K.gemm(x, wo[i], res=h, out=h, w_static=True)
            K.gemm(x, wgu[i], act=K.ACT_SWIGLU, out=act, w_static=True)
            K.gemm(a, wd[i], res=h, out=h, w_static=True)
    t = timeit(layer_chain) / NL
    tot = sum(w[0].numel() * 2 for w in (wq, wo, wgu, wd)) / 1e6
    print("%s layer chain (4 GEMMs, %.0f MB): %.2f us/layer  %.2f TB/s" % (name, tot, t, tot / t), flush=True)
    if gemv:
        ya, yb = torch.zeros(T, 4096, device=dev, dtype=torch.float16), torch.zeros(T, 4096, device=dev, dtype=torch.float16)
        ssa, ssb = torch.zeros(K.NORM_SS_FLOATS, device=dev), torch.zeros(K.NORM_SS_FLOATS, device=dev)
        ssb[0] = 1.0
        def layer_chain_norm():
            for i in range(NL):
                K.gemm(yb, wq[i], out=qkv, w_static=True, norm_ss=(ssb, 1e-6))
                K.gemm(x, wo[i], res=h, out=h, w_static=True, post_norm=(gamma, ya, ssa))
                K.gemm(ya, wgu[i], act=K.ACT_SWIGLU, out=act, w_static=True, norm_ss=(ssa, 1e-6))
                K.gemm(a, wd[i], res=h, out=h, w_static=True, post_norm=(gamma, yb, ssb))
        t = timeit(layer_chain_norm) / NL
        print("%s layer chain with the RMSNorm hand-over: %.2f us/layer  %.2f TB/s" % (name, t, tot / t), flush=True)
K.set_gemv(True)
# steady state: one long launch (F = 65536, K = 4096: 537 MB) and (F = 16384, K = 11008: 361 MB)
for F, Kd in ((65536, 4096), (16384, 11008)):
    wbig = (torch.randn(F, Kd, device=dev) * 0.02).half()
    xb = torch.randn(T, Kd, device=dev).half()
    ob = torch.empty(T, F, device=dev, dtype=torch.float16)
    for gemv in (True, False):
        K.set_gemv(gemv)
        t = timeit(lambda: K.gemm(xb, wbig, out=ob, w_static=True), reps=10)
        print("%s single launch F=%d K=%d (%.0f MB): %.1f us  %.2f TB/s" % ("gemv  " if gemv else "tcgen05", F, Kd, F * Kd * 2 / 1e6, t, F * Kd * 2 / 1e6 / t), flush=True)
    del wbig
K.set_gemv(True)
