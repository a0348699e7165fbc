"""Correctness + timing sweep of the CTA-pair GEMM (csrc/gemm2.cu) over operand arrangement / tile width / pairs per
cluster, forced through MYR_G2_MODE / MYR_G2_BN / MYR_G2_P. Each group runs in its own process (a protocol bug traps the
context; the other groups still report). Run under gpurun:  python scripts/gemm2_sweep.py [group ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {
    "vit_fc1": (1028, 6144, 1408, dict(bias=True, act="gelu")),
    "vit_fc2": (1028, 1408, 6144, dict(bias=True, res=True)),
    "vit_qkv": (1028, 4224, 1408, dict(bias=True)),
    "vit_proj": (1028, 1408, 1408, dict(bias=True, res=True)),
    "ll_qkv": (524, 12304, 4096, dict()),
    "ll_o": (524, 4096, 4096, dict(res=True)),
    "ll_gu": (524, 22016, 4096, dict(act="swiglu")),
    "ll_down": (524, 4096, 11008, dict(res=True)),
    "big_gu": (8192, 22016, 4096, dict(act="swiglu")),
    "big_down": (8192, 4096, 11008, dict(res=True)),
    "tr_gu": (656, 22016, 4096, dict()),
    "qf_kv": (1028, 9216, 1408, dict(bias=True)),
    "small": (300, 768, 768, dict(bias=True)),
    "tr_o": (656, 4096, 4096, dict(res=True)),
    "tr_down": (656, 4096, 11008, dict(res=True)),
    "tr_dqkv": (656, 4096, 12288, dict(res=True)),
    "tr_dgu": (656, 4096, 22016, dict(res=True)),
}

GROUPS = {
    # name: list of (shape, mode, bn, P, S[, extra env]); mode -1 / bn 0 / P 0 / S 0 = let the plan choose; mode -2 = gemm.cu's kernel
    "old": [(s, -2, 0, 0, 0) for s in SHAPES],
    "auto": [(s, -1, 0, 0, 0) for s in SHAPES],
    "m0": [("vit_fc1", 0, 176, 1, 1), ("vit_fc1", 0, 208, 1, 1), ("vit_qkv", 0, 176, 1, 1), ("vit_qkv", 0, 208, 1, 1), ("vit_qkv", 0, 128, 1, 1),
           ("vit_fc2", 0, 176, 1, 1), ("vit_fc2", 0, 176, 1, 2), ("vit_fc2", 0, 176, 1, 4), ("vit_fc2", 0, 208, 1, 3),
           ("vit_proj", 0, 176, 1, 1), ("vit_proj", 0, 176, 1, 2), ("vit_proj", 0, 128, 1, 1), ("vit_proj", 0, 96, 1, 1),
           ("ll_qkv", 0, 176, 1, 1), ("ll_o", 0, 176, 1, 1), ("ll_o", 0, 176, 1, 2), ("ll_o", 0, 176, 1, 3), ("ll_o", 0, 176, 1, 4),
           ("ll_gu", 0, 176, 1, 1), ("ll_down", 0, 176, 1, 1), ("ll_down", 0, 176, 1, 2), ("ll_down", 0, 176, 1, 3), ("ll_down", 0, 176, 1, 4),
           ("tr_gu", 0, 224, 1, 1), ("big_gu", 0, 256, 1, 1), ("big_down", 0, 256, 1, 1), ("qf_kv", 0, 176, 1, 1), ("small", 0, 160, 1, 1)],
    "pf": [(sh, 0, bn, 1, 1, {"MYR_G2_PF": str(pf)}) for sh, bn in (("ll_gu", 176), ("ll_qkv", 176), ("vit_fc1", 176), ("big_gu", 256))
           for pf in (0, 4, 8, 16)],
    "prof": [("vit_qkv", 0, 208, 1, 1), ("ll_o", 0, 176, 1, 1), ("vit_fc1", 0, 208, 1, 1)],
    "lat": [(sh, 0, bn, 1, 1, {"MYR_G2_STAGES": str(st), "NW": str(nw)}) for sh, bn in (("vit_qkv", 208), ("ll_gu", 176), ("ll_qkv", 176))
            for st, nw in ((6, 8), (4, 8), (3, 8), (2, 8), (6, 1))] +
           [("ll_gu", 0, bn, 1, 1, {"NW": "2"}) for bn in (96, 128, 176, 256)] + [("big_gu", 0, bn, 1, 1, {"NW": "2"}) for bn in (128, 192, 256)],
    "abl": [("ll_gu", 0, bn, 1, 1, {"NW": "2", "MYR_G2_DBG": str(d)}) for bn in (176,) for d in (0, 1, 2, 4, 6, 8, 3, 5)],
    "tl": [("ll_gu", 0, 176, 1, 1, {"NW": "2", "MYR_G2_DBG": str(d)}) for d in (16, 17, 21)],
    "few": [("vit_fc2", 0, 176, 1, 1), ("vit_fc2", 0, 176, 1, 2), ("vit_fc2", 0, 176, 1, 3), ("vit_fc2", 0, 128, 1, 1), ("vit_fc2", 0, 128, 1, 2),
            ("vit_fc2", 0, 208, 1, 2), ("vit_proj", 0, 176, 1, 1), ("vit_proj", 0, 128, 1, 1), ("vit_proj", 0, 176, 1, 2),
            ("ll_o", 0, 176, 1, 1), ("ll_o", 0, 144, 1, 1), ("ll_o", 0, 176, 1, 2), ("ll_down", 0, 176, 1, 1), ("ll_down", 0, 176, 1, 2),
            ("ll_down", 0, 144, 1, 1), ("ll_down", 0, 144, 1, 2), ("vit_qkv", 0, 144, 1, 1), ("vit_qkv", 0, 176, 1, 1), ("vit_qkv", 0, 208, 1, 1),
            ("vit_fc1", 0, 208, 1, 1), ("qf_kv", 0, 208, 1, 1)],
    "mc": [(sh, 0, bn, P, 1) for sh, bn in (("big_gu", 256), ("big_down", 256), ("ll_gu", 176), ("ll_qkv", 176), ("vit_fc1", 176), ("tr_gu", 224), ("qf_kv", 208))
           for P in (1, 2, 4)],
    "tr": [(sh, 0, bn, 1, S) for sh in ("tr_o", "tr_down", "tr_dqkv", "tr_dgu") for bn, S in ((176, 1), (224, 1), (224, 2), (224, 3), (176, 2), (256, 3))] +
          [(sh, 0, bn, 1, S) for sh in ("ll_o", "ll_down") for bn, S in ((144, 1), (176, 1), (176, 2), (176, 3))],
    "m1": [("vit_fc1", 1, 256, 1, 1), ("ll_gu", 1, 256, 1, 1), ("big_gu", 1, 256, 1, 1), ("big_down", 1, 256, 1, 1)],
}


def run_group(name):
    import torch

    from myriad_b200 import kernels as K
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for entry in GROUPS[name]:
        shape, mode, bn, P, S = entry[:5]
        extra = entry[5] if len(entry) > 5 else {}
        os.environ["MYR_G2_PF"] = "0"
        os.environ.update(extra)
        T, F, Kd, kw = SHAPES[shape]
        os.environ["MYR_G2_MIN_T"] = "1000000" if mode == -2 else "128"
        for k, v in (("MYR_G2_MODE", mode), ("MYR_G2_BN", bn), ("MYR_G2_P", P), ("MYR_G2_S", S)):
            if v > 0 or (k == "MYR_G2_MODE" and v >= 0):
                os.environ[k] = str(v)
            else:
                os.environ.pop(k, None)
        nw = 8 if T * F < 5e7 else 2  # rotate over more weight bytes than the 126 MB L2 holds
        nw = int(extra.get("NW", nw))
        os.environ["MYR_G2_STAGES"] = extra.get("MYR_G2_STAGES", "0")
        os.environ["MYR_G2_DBG"] = extra.get("MYR_G2_DBG", "0")
        x = torch.randn(T, Kd, device=dev).half()
        ws = [(torch.randn(F, Kd, device=dev) / Kd ** 0.5).half() for _ in range(nw)]
        bias = torch.randn(F, device=dev).half() if kw.get("bias") else None
        act = {"gelu": K.ACT_GELU, "swiglu": K.ACT_SWIGLU}.get(kw.get("act"), 0)
        Fo = F // 2 if act == K.ACT_SWIGLU else F
        res = torch.randn(T, Fo, device=dev) if kw.get("res") else None
        out = torch.empty(T, Fo, device=dev, dtype=torch.float32 if res is not None else torch.float16)
        y = K.gemm(x, ws[0], bias=bias, act=act, res=res, out=out.clone()).float()
        torch.cuda.synchronize()
        rows = torch.arange(0, T, max(1, T // 256), device=dev)
        rows = torch.cat([rows, torch.tensor([T - 1], device=dev)])
        r = x[rows].float() @ ws[0].float().t()
        if bias is not None:
            r = r + bias.float()
        if act == K.ACT_GELU:
            r = torch.nn.functional.gelu(r)
        if act == K.ACT_SWIGLU:
            g = r.reshape(len(rows), Fo // 64, 2, 64)
            r = (torch.nn.functional.silu(g[:, :, 0].half().float()) * g[:, :, 1].half().float()).reshape(len(rows), Fo)
        if res is not None:
            r = r + res[rows]
        err = (y[rows] - r).abs().max().item() / max(1.0, r.abs().max().item())
        iters = 20

        def chain():
            for i in range(iters):
                K.gemm(x, ws[i % nw], bias=bias, act=act, res=res, out=out)

        # the launches are captured in a CUDA graph: the number is the GPU time of back-to-back (PDL-chained) launches, not the
        # host's launch rate (a 1028 x 6144 x 1408 GEMM runs shorter than torch / ctypes take to issue it)
        chain()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 1e9
        for _ in range(5):  # best of 5 replays (a single replay varies by +-10 % with the clock state)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1) / iters)
        # one traced launch: set-up done / first MMA wave / accumulator ready / epilogue done, relative to the earliest CTA
        import ctypes
        tr = torch.zeros(148 * 6 * 2, dtype=torch.int64, device=dev)
        K.lib().myr_gemm_set_trace(ctypes.c_void_p(tr.data_ptr()))
        K.gemm(x, ws[0], bias=bias, act=act, res=res, out=out)
        K.lib().myr_gemm_set_trace(ctypes.c_void_p(0))
        torch.cuda.synchronize()
        if int(os.environ.get("MYR_G2_DBG", "0")) & 16:
            x = tr[888:888 + 192].cpu()
            t00 = int(x[0])
            fmt = lambda v: " ".join("%.2f" % ((int(a) - t00) / 1e3) for a in v[:40])
            print("TL leader issue  us:", fmt(x[0:64]))
            print("TL leader mmardy us:", fmt(x[64:128]))
            print("TL peer   issue  us:", fmt(x[128:192]), flush=True)
        t = tr[:148 * 6].reshape(148, 6).cpu()
        live = t[:, 0] > 0
        tl = ""
        if live.any():
            t0 = int(t[live, 0].min())
            cols = [("setup", 0), ("mma_last", 2), ("acc_ready", 3), ("epi_done", 4)]
            tl = " | trace us: " + " ".join("%s %.1f..%.1f" % (n, (int(t[live & (t[:, c] > 0), c].min()) - t0) / 1e3,
                                                             (int(t[live & (t[:, c] > 0), c].max()) - t0) / 1e3)
                                           for n, c in cols if (live & (t[:, c] > 0)).any())
        print("RESULT %-9s mode=%2d bn=%3d P=%d S=%d %s T=%5d F=%5d K=%5d  %8.1f us %6.0f TF/s  relerr %.1e %s%s" % (
            shape, mode, bn, P, S, " ".join("%s=%s" % kv for kv in extra.items()), T, F, Kd, ms * 1e3, 2.0 * T * F * Kd / ms / 1e9, err, "OK" if err < 4e-3 else "FAIL", tl), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--group":
        run_group(sys.argv[2])
    else:
        groups = sys.argv[1:] or ["old", "m0", "pf", "auto"]
        for g in groups:
            print("=== group %s" % g, flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], timeout=300, stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True)
                print(r.stdout[-20000:], flush=True)
                print("=== group %s exit %d" % (g, r.returncode), flush=True)
            except subprocess.TimeoutExpired as e:
                print((e.stdout or b"")[-3000:] if isinstance(e.stdout, (bytes, str)) else "", flush=True)
                print("=== group %s TIMEOUT" % g, flush=True)
