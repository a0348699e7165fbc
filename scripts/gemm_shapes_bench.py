"""Tensor-bound GEMM shapes of the benchmark (ViT B=4 -> T=1028, LLaMA prefill T=524, training T=656): myr_gemm_f16 against
cuBLAS (torch.matmul) on the same operands, rotating over several weight matrices so the weights come from HBM like in the
model. Run under gpurun: python scripts/gemm_shapes_bench.py [--check]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import kernels as K

dev = torch.device("cuda:0")
torch.manual_seed(0)

SHAPES = [  # name, T, F, K, kwargs
    ("vit qkv", 1028, 4224, 1408, dict(bias=True)),
    ("vit proj+res", 1028, 1408, 1408, dict(bias=True, res="f32")),
    ("vit fc1+gelu", 1028, 6144, 1408, dict(bias=True, act=K.ACT_GELU)),
    ("vit fc2+res", 1028, 1408, 6144, dict(bias=True, res="f32")),
    ("vit8 fc1+gelu", 2056, 6144, 1408, dict(bias=True, act=K.ACT_GELU)),
    ("vit8 fc2+res", 2056, 1408, 6144, dict(bias=True, res="f32")),
    ("llama qkv", 524, 12304, 4096, dict()),
    ("llama o+res", 524, 4096, 4096, dict(res="f32")),
    ("llama gate/up swiglu", 524, 22016, 4096, dict(act=K.ACT_SWIGLU)),
    ("llama down+res", 524, 4096, 11008, dict(res="f32")),
    ("train qkv", 656, 12304, 4096, dict()),
    ("train o+res", 656, 4096, 4096, dict(res="f32")),
    ("train gate/up", 656, 22016, 4096, dict()),
    ("train down+res", 656, 4096, 11008, dict(res="f32")),
    ("train d_down", 656, 11008, 4096, dict()),
    ("train d_gate/up", 656, 4096, 22016, dict()),
    ("train d_qkv", 656, 4096, 12288, dict()),
    ("s2048 gate/up swiglu", 8192, 22016, 4096, dict(act=K.ACT_SWIGLU)),
    ("s2048 down+res", 8192, 4096, 11008, dict(res="f32")),
]


def one(name, T, F, Kd, kw, check, nw=6, iters=30):
    kw = dict(kw)
    x = torch.randn(T, Kd, device=dev).half()
    ws = [(torch.randn(F, Kd, device=dev) / Kd ** 0.5).half() for _ in range(nw)]
    bias = torch.randn(F, device=dev).half() if kw.pop("bias", False) else None
    res_kind = kw.pop("res", None)
    act = kw.get("act", 0)
    Fo = F // 2 if act == K.ACT_SWIGLU else F
    res = torch.randn(T, Fo, device=dev) if res_kind == "f32" else None
    out = torch.empty(T, Fo, device=dev, dtype=torch.float32 if res is not None else torch.float16)
    ref_out = torch.empty(T, F, device=dev, dtype=torch.float16)

    def ours(i):
        K.gemm(x, ws[i % nw], bias=bias, res=res, out=out, **kw)

    def cublas(i):
        torch.matmul(x, ws[i % nw].t(), out=ref_out)

    res_t = {}
    for label, fn in (("ours", ours), ("cublas", cublas)):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        res_t[label] = e0.elapsed_time(e1) / iters
    fl = 2.0 * T * F * Kd
    err = ""
    if check:
        K.gemm(x, ws[0], bias=bias, res=None, out=None if act != K.ACT_SWIGLU else None, **kw) if False else None
        y = K.gemm(x, ws[0], bias=bias, **kw).float()
        r = x.float() @ ws[0].float().t()
        if bias is not None:
            r = r + bias.float()
        if act == K.ACT_GELU:
            r = torch.nn.functional.gelu(r)
        if act == K.ACT_SWIGLU:
            I = F // 2
            g = r.reshape(T, I // 64, 2, 64)
            r = (torch.nn.functional.silu(g[:, :, 0]) * g[:, :, 1]).reshape(T, I)
        e = (y - r).abs().max().item() / max(1.0, r.abs().max().item())
        err = " relerr %.1e%s" % (e, "" if e < 4e-3 else " FAIL")
    print("%-22s T=%5d F=%5d K=%5d: ours %7.1f us %6.0f TF/s | cublas %7.1f us %6.0f TF/s | ours/cublas %.2f%s" % (
        name, T, F, Kd, res_t["ours"] * 1e3, fl / res_t["ours"] / 1e9, res_t["cublas"] * 1e3, fl / res_t["cublas"] / 1e9,
        res_t["cublas"] / res_t["ours"], err), flush=True)
    return res_t["ours"]


def main():
    check = "--check" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    tot = {}
    for name, T, F, Kd, kw in SHAPES:
        if only and not any(o in name for o in only):
            continue
        tot[name] = one(name, T, F, Kd, kw, check)
    if not only:
        vit = 39 * sum(tot[n] for n in ("vit qkv", "vit proj+res", "vit fc1+gelu", "vit fc2+res"))
        pre = 32 * sum(tot[n] for n in ("llama qkv", "llama o+res", "llama gate/up swiglu", "llama down+res"))
        print("sum of GEMM time: ViT (39 blocks, B=4) %.2f ms, LLaMA prefill (32 layers, T=524) %.2f ms" % (vit, pre))


if __name__ == "__main__":
    main()
