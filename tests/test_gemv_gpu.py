"""GPU: the small-batch (T <= 4) weight-streaming kernel (csrc/gemv.cu) behind myr_gemm_f16 against the fp32 definition,
against the tcgen05 kernel it stands in for, and — for the RMSNorm hand-over — against the oracle's LlamaRMSNorm
(modeling_llama.py:66-74). Tolerance: fp16 operands, fp32 accumulation -> 1e-3 of the output scale; outputs of repeated
launches must be bit-identical (fixed-order reductions)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from myriad_b200 import kernels
    assert kernels.lib().myr_version() >= 1
    return kernels


@pytest.fixture(scope="module")
def O():
    from oracle import myriad_oracle
    return myriad_oracle


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, std=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * std


def close(a, b, tol, what=""):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all(), what + " has non-finite values"
    err = (a - b).abs().max().item()
    scale = max(1.0, b.abs().max().item())
    assert err <= tol * scale, "%s: max err %.3e > %.1e * %.2f" % (what, err, tol, scale)


SHAPES = [(4, 4096, 4096), (4, 12304, 4096), (4, 4096, 11008), (3, 32000, 4096), (1, 4096, 4096), (2, 100, 128), (4, 768, 1408),
          (4, 5, 128), (4, 1408, 6144), (4, 40000, 512)]


@pytest.mark.parametrize("T,F,Kd", SHAPES)
def test_gemv_vs_fp32_and_tensor_core_path(K, T, F, Kd):
    x, w = rnd(T, Kd, seed=1).half().to(dev()), (rnd(F, Kd, seed=2) / Kd ** 0.5).half().to(dev())
    b = rnd(F, seed=3).half().to(dev())
    r32 = rnd(T, F, seed=4).to(dev())
    ref = x.float() @ w.float().t() + b.float() + r32
    n0 = K.launch_count()
    y1 = K.gemm(x, w, bias=b, res=r32, out_dtype=torch.float32, w_static=True)
    assert K.launch_count() - n0 == 1, "one launch: no split-tile reduce, no workspace pass"
    y2 = K.gemm(x, w, bias=b, res=r32, out_dtype=torch.float32, w_static=True)
    close(y1, ref, 1e-3, "gemv")
    assert torch.equal(y1, y2), "fixed-order reduction must be reproducible"
    old = K.set_gemv(False)
    try:
        y3 = K.gemm(x, w, bias=b, res=r32, out_dtype=torch.float32)
    finally:
        K.set_gemv(old)
    close(y1, y3, 1e-3, "gemv vs tcgen05")
    # fp16 output, no residual, padded row strides on both operands
    xp = torch.zeros(T, Kd + 64, device=dev(), dtype=torch.float16)
    xp[:, :Kd] = x
    wp = torch.zeros(F, Kd + 8, device=dev(), dtype=torch.float16)
    wp[:, :Kd] = w
    y4 = K.gemm(xp[:, :Kd], wp[:, :Kd], K=Kd)
    close(y4, x.float() @ w.float().t(), 2e-3, "gemv fp16 out, strided operands")


def test_gemv_residual_in_place_and_fp16_residual(K):
    T, F, Kd = 4, 4096, 11008
    x, w = rnd(T, Kd, seed=1).half().to(dev()), (rnd(F, Kd, seed=2) / Kd ** 0.5).half().to(dev())
    r32 = rnd(T, F, seed=4).to(dev())
    ref = x.float() @ w.float().t() + r32
    acc = r32.clone()
    K.gemm(x, w, res=acc, out=acc, w_static=True)  # the fp32 residual stream is updated in place (o_proj / down_proj)
    close(acc, ref, 1e-3, "in-place residual")
    r16 = r32.half()
    y = K.gemm(x, w, res=r16)
    close(y, x.float() @ w.float().t() + r16.float(), 2e-3, "fp16 residual")


def _interleave64(gate, up):
    I, Kd = gate.shape
    return torch.stack([gate.reshape(I // 64, 64, Kd), up.reshape(I // 64, 64, Kd)], 1).reshape(2 * I, Kd)


@pytest.mark.parametrize("T,I,Kd", [(4, 11008, 4096), (1, 64, 256), (3, 1024, 640), (4, 192, 4096)])
def test_gemv_swiglu(K, T, I, Kd):
    x = rnd(T, Kd, seed=1).half().to(dev())
    g, u = (rnd(I, Kd, seed=2) / Kd ** 0.5).half(), (rnd(I, Kd, seed=3) / Kd ** 0.5).half()
    y = K.gemm(x, _interleave64(g, u).to(dev()), act=K.ACT_SWIGLU, w_static=True)
    assert y.shape == (T, I) and y.dtype == torch.float16
    gf, uf = x.float().cpu() @ g.float().t(), x.float().cpu() @ u.float().t()
    close(y, torch.nn.functional.silu(gf.half().float()) * uf.half().float(), 3e-3, "gemv swiglu")
    old = K.set_gemv(False)
    try:
        y_tc = K.gemm(x, _interleave64(g, u).to(dev()), act=K.ACT_SWIGLU)
    finally:
        K.set_gemv(old)
    close(y, y_tc, 2e-3, "gemv swiglu vs tcgen05")
    assert (y != y_tc).float().mean().item() < 0.05  # one fp16 ulp where the fp32 sums round differently


def test_hand_over_rejected_on_large_batch(K):
    """The RMSNorm hand-over exists on the T <= 4 path only; the tensor-core path must refuse it loudly, not ignore it."""
    x = rnd(8, 256, seed=1).half().to(dev())
    w = rnd(64, 256, seed=2).half().to(dev())
    ss = torch.zeros(K.NORM_SS_FLOATS, device=dev())
    with pytest.raises(RuntimeError):
        K.gemm(x, w, norm_ss=(ss, 1e-6))


def test_gemv_dependent_chain_in_cuda_graph(K):
    """o_proj -> (norm) gate/up+SwiGLU -> down_proj launched back to back with programmatic dependent launch and replayed from
    a CUDA graph: each kernel pre-loads its weights before the previous one has finished and must still see its output."""
    T, D, I = 4, 1024, 2816
    torch.manual_seed(0)
    ctx = rnd(T, D, seed=1).half().to(dev())
    wo = (rnd(D, D, seed=2) / D ** 0.5).half().to(dev())
    g, u = (rnd(I, D, seed=3) / D ** 0.5).half(), (rnd(I, D, seed=4) / D ** 0.5).half()
    wgu = _interleave64(g, u).to(dev())
    wd = (rnd(D, I, seed=5) / I ** 0.5).half().to(dev())
    gamma = (1.0 + 0.1 * rnd(D, seed=6)).to(dev())
    h0 = rnd(T, D, seed=7).to(dev())
    h = h0.clone()
    act = torch.empty(T, I, device=dev(), dtype=torch.float16)

    y16 = torch.zeros(T, D, device=dev(), dtype=torch.float16)
    ss = torch.zeros(K.NORM_SS_FLOATS, device=dev(), dtype=torch.float32)

    def chain():
        K.gemm(ctx, wo, res=h, out=h, w_static=True, post_norm=(gamma, y16, ss))
        K.gemm(y16, wgu, act=K.ACT_SWIGLU, out=act, w_static=True, norm_ss=(ss, 1e-6))
        K.gemm(act, wd, res=h, out=h, w_static=True)

    chain()
    torch.cuda.synchronize()
    eager = h.clone()
    # fp32 definition with the kernels' rounding points
    r = h0.cpu() + ctx.float().cpu() @ wo.float().cpu().t()
    xn = (r * torch.rsqrt(r.pow(2).mean(-1, keepdim=True) + 1e-6) * gamma.cpu()).half().float()
    a = (torch.nn.functional.silu((xn @ g.float().t()).half().float()) * (xn @ u.float().t()).half().float()).half().float()  # noqa
    ref = r + a @ wd.float().cpu().t()
    close(eager, ref, 2e-3, "chain vs fp32 definition")
    graph = torch.cuda.CUDAGraph()
    h.copy_(h0)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph):
            chain()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        h.copy_(h0)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(h, eager), "graph replay must equal the eager launches exactly"


def test_gemv_rmsnorm_hand_over(K, O):
    """o_proj-style producer writes rn_f16(h * gamma) + per-CTA sums of squares; the consumer multiplies by the scale in its
    epilogue: == RMSNorm(h) @ W^T of the oracle (modeling_llama.py:66-74), and == the in-kernel fused-norm path."""
    T, D, F = 4, 4096, 12304
    ctx = rnd(T, D, seed=1).half().to(dev())
    wo = (rnd(D, D, seed=2) / D ** 0.5).half().to(dev())
    h0 = rnd(T, D, seed=3, std=2.0).to(dev())
    gamma = (1.0 + 0.1 * rnd(D, seed=4)).to(dev())
    w = (rnd(F, D, seed=5) / D ** 0.5).half().to(dev())
    h = h0.clone()
    y16 = torch.zeros(T, D, device=dev(), dtype=torch.float16)
    ss = torch.zeros(K.NORM_SS_FLOATS, device=dev(), dtype=torch.float32)
    K.gemm(ctx, wo, res=h, out=h, w_static=True, post_norm=(gamma, y16, ss))
    y = K.gemm(y16, w, out_dtype=torch.float32, w_static=True, norm_ss=(ss, 1e-6))
    href = h0.cpu() + ctx.float().cpu() @ wo.float().cpu().t()
    close(h, href, 1e-3, "producer output")
    assert int(ss[0].item()) >= 1
    tot = ss[4:4 + 4 * int(ss[0].item())].reshape(-1, 4).sum(0).cpu()
    close(tot, href.pow(2).sum(-1), 1e-4, "sum of squares partials")
    close(y, O.rms_norm(href, gamma.cpu(), 1e-6) @ w.float().cpu().t(), 2e-3, "hand-over vs oracle")
    x16 = torch.empty(T, D, device=dev(), dtype=torch.float16)
    K.norm(h, gamma, None, 1e-6, rms=True, out16=x16)
    y2 = K.gemm(x16, w, out_dtype=torch.float32, w_static=True)
    close(y, y2, 5e-4, "hand-over vs norm kernel + projection")


# ------------------------------------------------------------------------------------------------------------------
# 5 <= T <= 32: weights and tokens streamed through one ring (csrc/gemv_mt.cu)
MT_SHAPES = [(16, 4096, 4096), (32, 12304, 4096), (5, 4096, 11008), (9, 32000, 4096), (24, 4096, 4096), (17, 1024, 128),
             (32, 4096, 11008), (8, 1100, 640), (31, 12304, 4096)]


@pytest.mark.parametrize("T,F,Kd", MT_SHAPES)
def test_gemv_mt_vs_fp32_and_tensor_core_path(K, T, F, Kd):
    x, w = rnd(T, Kd, seed=1).half().to(dev()), (rnd(F, Kd, seed=2) / Kd ** 0.5).half().to(dev())
    b = rnd(F, seed=3).half().to(dev())
    r32 = rnd(T, F, seed=4).to(dev())
    ref = x.float() @ w.float().t() + b.float() + r32
    n0 = K.launch_count()
    y1 = K.gemm(x, w, bias=b, res=r32, out_dtype=torch.float32, w_static=True)
    assert K.launch_count() - n0 == 1, "one launch: no split-tile reduce, no workspace pass"
    y2 = K.gemm(x, w, bias=b, res=r32, out_dtype=torch.float32, w_static=True)
    close(y1, ref, 1e-3, "gemv_mt")
    assert torch.equal(y1, y2), "fixed-order reduction must be reproducible"
    old = K.set_gemv(False)
    try:
        y3 = K.gemm(x, w, bias=b, res=r32, out_dtype=torch.float32)
    finally:
        K.set_gemv(old)
    close(y1, y3, 1e-3, "gemv_mt vs tcgen05")
    # fp16 output, no residual / bias, padded row strides on both operands, in-place fp32 residual
    xp = torch.zeros(T, Kd + 64, device=dev(), dtype=torch.float16)
    xp[:, :Kd] = x
    wp = torch.zeros(F, Kd + 8, device=dev(), dtype=torch.float16)
    wp[:, :Kd] = w
    y4 = K.gemm(xp[:, :Kd], wp[:, :Kd], K=Kd)
    close(y4, x.float() @ w.float().t(), 2e-3, "gemv_mt fp16 out, strided operands")
    acc = r32.clone()
    K.gemm(x, w, res=acc, out=acc, w_static=True)
    close(acc, x.float() @ w.float().t() + r32, 1e-3, "gemv_mt in-place residual")


@pytest.mark.parametrize("T,I,Kd", [(16, 11008, 4096), (32, 11008, 4096), (5, 1024, 640), (20, 1088, 4096), (12, 1032 * 8 // 8 // 64 * 64, 256)])
def test_gemv_mt_swiglu(K, T, I, Kd):
    x = rnd(T, Kd, seed=1).half().to(dev())
    g, u = (rnd(I, Kd, seed=2) / Kd ** 0.5).half(), (rnd(I, Kd, seed=3) / Kd ** 0.5).half()
    n0 = K.launch_count()
    y = K.gemm(x, _interleave64(g, u).to(dev()), act=K.ACT_SWIGLU, w_static=True)
    assert K.launch_count() - n0 == 1
    assert y.shape == (T, I) and y.dtype == torch.float16
    gf, uf = x.float().cpu() @ g.float().t(), x.float().cpu() @ u.float().t()
    close(y, torch.nn.functional.silu(gf.half().float()) * uf.half().float(), 3e-3, "gemv_mt swiglu")
    old = K.set_gemv(False)
    try:
        y_tc = K.gemm(x, _interleave64(g, u).to(dev()), act=K.ACT_SWIGLU)
    finally:
        K.set_gemv(old)
    close(y, y_tc, 2e-3, "gemv_mt swiglu vs tcgen05")
    assert (y != y_tc).float().mean().item() < 0.05


def test_gemv_mt_dependent_chain_in_cuda_graph(K):
    """o_proj -> gate/up -> down of one decode step at batch 16 as dependent launches (programmatic dependent launch: weights
    are requested before the predecessor finishes, tokens after): eager == CUDA-graph replay, bit for bit, over 3 replays."""
    T, D, I = 16, 4096, 11008
    ctx = rnd(T, D, seed=1).half().to(dev())
    wo = (rnd(D, D, seed=2) / D ** 0.5).half().to(dev())
    g, u = (rnd(I, D, seed=3) / D ** 0.5).half(), (rnd(I, D, seed=4) / D ** 0.5).half()
    wgu = _interleave64(g, u).to(dev())
    wd = (rnd(D, I, seed=5) / I ** 0.5).half().to(dev())
    h0 = rnd(T, D, seed=6).to(dev())
    x16 = torch.empty(T, D, device=dev(), dtype=torch.float16)
    act = torch.empty(T, I, device=dev(), dtype=torch.float16)
    gamma = torch.ones(D, device=dev())

    def chain(h):
        K.gemm(ctx, wo, res=h, out=h, w_static=True)
        K.norm(h, gamma, None, 1e-6, rms=True, out16=x16)
        K.gemm(x16, wgu, act=K.ACT_SWIGLU, out=act, w_static=True)
        K.gemm(act, wd, res=h, out=h, w_static=True)

    h_e = h0.clone()
    chain(h_e)
    torch.cuda.synchronize()
    hf = h0.float().cpu() + ctx.float().cpu() @ wo.float().cpu().t()
    xn = (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-6)).half().float()
    a = torch.nn.functional.silu((xn @ g.float().t()).half().float()) * (xn @ u.float().t()).half().float()
    ref = hf + a.half().float() @ wd.float().cpu().t()
    close(h_e, ref, 2e-3, "decode-step chain at batch 16")
    h_g = h0.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain(h_g)  # warm-up on the side stream
        h_g.copy_(h0)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            chain(h_g)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        h_g.copy_(h0)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(h_g, h_e), "CUDA-graph replay must equal eager launches"
