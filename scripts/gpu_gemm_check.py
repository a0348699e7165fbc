"""GPU check of myr_gemm_f16 vs torch (fp32 reference of the same fp16 operands) + timing. Run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import kernels as K

torch.manual_seed(0)
dev = torch.device("cuda:0")

def ref(x, w, bias=None, act=0, res=None):
    y = x.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if act == 1:
        y = torch.nn.functional.gelu(y.half().float()).half().float() if False else torch.nn.functional.gelu(y)
    if res is not None:
        y = y + res.float()
    return y

def run(T, F, Kd, bias=False, act=0, res=None, out_dtype=torch.float16, **kw):
    x = torch.randn(T, Kd, device=dev).half()
    w = (torch.randn(F, Kd, device=dev) / Kd ** 0.5).half()
    b = torch.randn(F, device=dev).half() if bias else None
    r = None
    if res == "f16":
        r = torch.randn(T, F, device=dev).half()
    elif res == "f32":
        r = torch.randn(T, F, device=dev)
    y = K.gemm(x, w, bias=b, act=act, res=r, out_dtype=out_dtype, **kw)
    torch.cuda.synchronize()
    yr = ref(x, w, b, act, r)
    err = (y.float() - yr).abs().max().item()
    tol = 2e-2 if out_dtype == torch.float16 else 2e-3
    ok = err < tol and torch.isfinite(y.float()).all().item()
    print("T=%5d F=%5d K=%5d bias=%d act=%d res=%s out=%s kw=%s  max_err=%.3e %s" % (
        T, F, Kd, bias, act, res, str(out_dtype)[6:], kw, err, "OK" if ok else "FAIL"), flush=True)
    return ok

def bench(T, F, Kd, iters=20, **kw):
    x = torch.randn(T, Kd, device=dev).half()
    w = (torch.randn(F, Kd, device=dev) / Kd ** 0.5).half()
    out = torch.empty(T, F, device=dev, dtype=torch.float16)
    for _ in range(3):
        K.gemm(x, w, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.gemm(x, w, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    for _ in range(3):
        torch.matmul(x, w.t(), out=out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(x, w.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    fl = 2.0 * T * F * Kd
    by = 2.0 * (T * Kd + F * Kd + T * F)
    print("bench T=%5d F=%5d K=%5d %s: ours %.3f ms %.1f TF/s %.0f GB/s | cublas %.3f ms %.1f TF/s" % (
        T, F, Kd, kw, ms, fl / ms / 1e9, by / ms / 1e6, ms_t, fl / ms_t / 1e9), flush=True)

ok = True
if "--stream-only" in sys.argv:
    run = lambda *a, **k: True
print("sm count", K.lib().myr_device_sm_count())
# basic shapes first (single tile, single k-block)
ok &= run(16, 128, 64)
ok &= run(16, 128, 128)
ok &= run(256, 128, 64)
ok &= run(256, 256, 512)
ok &= run(100, 200, 136)
ok &= run(2056, 4224, 1408, bias=True)
ok &= run(2056, 1408, 1408, bias=True, res="f32", out_dtype=torch.float32)
ok &= run(2056, 6144, 1408, bias=True, act=1)
ok &= run(2056, 1408, 6144, bias=True, res="f16")
ok &= run(4, 4096, 4096)
ok &= run(4, 11008, 4096)
ok &= run(4, 4096, 11008, res="f32", out_dtype=torch.float32)
ok &= run(4, 32000, 4096, out_dtype=torch.float32)
ok &= run(524, 12288, 4096)
ok &= run(257, 768, 1408, bias=True)
ok &= run(81, 3072, 768, bias=True, act=1)
ok &= run(300, 256, 512, bn_hint=64)
ok &= run(300, 256, 4096, ksplit_hint=4)
print("ALL OK" if ok else "SOME FAILED", flush=True)
def bench_stream(T, F, Kd, n_w=6, iters=5, act=0):
    """weight streaming: rotate over distinct weights so nothing is served from L2 (6 x 180 MB >> 126 MB)."""
    x = torch.randn(T, Kd, device=dev).half()
    ws = [(torch.randn(F, Kd, device=dev) / Kd ** 0.5).half() for _ in range(n_w)]
    out = torch.empty(T, F // 2 if act == 3 else F, device=dev, dtype=torch.float16)
    for w in ws:
        K.gemm(x, w, out=out, act=act, w_static=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for w in ws:
            K.gemm(x, w, out=out, act=act, w_static=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (iters * n_w)
    # same sequence inside a CUDA graph (no host launch overhead: what the decode step sees)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            K.gemm(x, w, out=out, act=act, w_static=True)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms_g = e0.elapsed_time(e1) / (iters * n_w)
    print("   in a CUDA graph: %.4f ms/launch %.0f GB/s" % (ms_g, 2.0 * (T * Kd + F * Kd + T * F) / ms_g / 1e6), flush=True)
    for w in ws:
        torch.matmul(x, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        for w in ws:
            torch.matmul(x, w.t())
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / (iters * n_w)
    by = 2.0 * (T * Kd + F * Kd + T * F)
    print("stream T=%3d F=%5d K=%5d act=%d: ours %.4f ms %.0f GB/s | cublas %.4f ms %.0f GB/s" % (
        T, F, Kd, act, ms, by / ms / 1e6, ms_t, by / ms_t / 1e6), flush=True)

for shp in [(4, 12304, 4096), (4, 4096, 4096), (4, 22016, 4096), (4, 4096, 11008), (4, 32000, 4096), (16, 22016, 4096)]:
    bench_stream(*shp)
bench_stream(4, 22016, 4096, act=3)
if "--stream-only" in sys.argv:
    sys.exit(0)
for shp in [(2056, 4224, 1408), (2056, 6144, 1408), (2056, 1408, 6144), (2056, 1408, 1408), (8224, 6144, 1408),
            (4, 4096, 4096), (4, 22016, 4096), (4, 4096, 11008), (4, 32000, 4096), (524, 12288, 4096),
            (524, 22016, 4096), (524, 4096, 11008), (8192, 8192, 8192)]:
    bench(*shp)
