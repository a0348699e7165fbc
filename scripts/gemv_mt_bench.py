"""B200: the multi-token weight-streaming kernel (csrc/gemv_mt.cu, 5 <= T <= 32) at the LLaMA-7B decode shapes against the
tcgen05 path it replaces and the T <= 4 kernel: us per launch inside a CUDA graph of 8 rotating weight matrices (> L2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
PEAK = 6551.0


def timed(fn, n_rot, reps=5):
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
    shapes = [("qkv", 12304, 4096, K.ACT_NONE), ("o", 4096, 4096, K.ACT_NONE), ("gate/up", 22016, 4096, K.ACT_SWIGLU),
              ("down", 4096, 11008, K.ACT_NONE), ("lm_head", 32000, 4096, K.ACT_NONE)]
    print("| shape | T | gemv_mt us | GB/s | frac | tcgen05 path us |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, F, Kd, act in shapes:
        n_rot = 8
        ws = [(torch.randn(F, Kd, device=dev) * 0.02).half() for _ in range(n_rot)]
        for T in (4, 8, 16, 32):
            x = torch.randn(T, Kd, device=dev).half()
            Fo = F // 2 if act == K.ACT_SWIGLU else F
            out = torch.empty(T, Fo, device=dev, dtype=torch.float16)
            res = None
            f = lambda i: K.gemm(x, ws[i], act=act, out=out, w_static=True)
            t1 = timed(f, n_rot)
            old = K.set_gemv(False)
            try:
                t2 = timed(f, n_rot)
            finally:
                K.set_gemv(old)
            gbs = F * Kd * 2 / t1 / 1e3
            print("| %s %dx%d | %d | %.1f | %.0f | %.3f | %.1f |" % (name, F, Kd, T, t1, gbs, gbs / PEAK, t2))


if __name__ == "__main__":
    main()
