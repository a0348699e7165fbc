"""GPU: every C-ABI kernel against the CPU oracle (oracle/myriad_oracle.py) or its plain definition, on the same
seeded inputs. Tolerances: fp16 operands / fp32 accumulation vs an fp32 oracle -> |err| <= 2^-9 * scale per
rounding; stated per test."""
import math

import pytest
import torch

from myriad_b200 import synthetic as syn

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


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("T,F,Kd", [(16, 128, 64), (100, 200, 136), (257, 768, 1408), (2056, 1408, 1408), (4, 4096, 4096),
                                    (524, 1024, 4096), (9, 4096, 25600)])
def test_gemm_shapes(K, T, F, Kd):
    x, w, b = rnd(T, Kd, seed=1).half(), (rnd(F, Kd, seed=2) / Kd ** 0.5).half(), rnd(F, seed=3).half()
    y = K.gemm(x.to(dev()), w.to(dev()), bias=b.to(dev()))
    close(y, x.float() @ w.float().t() + b.float(), 2e-3, "gemm")


def test_gemm_epilogues(K, O):
    T, F, Kd = 300, 384, 256
    x, w, b = rnd(T, Kd, seed=1).half(), (rnd(F, Kd, seed=2) / 16).half(), rnd(F, seed=3).half()
    r32 = rnd(T, F, seed=4)
    base = x.float() @ w.float().t() + b.float()
    xd, wd, bd = x.to(dev()), w.to(dev()), b.to(dev())
    close(K.gemm(xd, wd, bias=bd, act=K.ACT_GELU), O.gelu_erf(base), 2e-3, "gelu")
    close(K.gemm(xd, wd, bias=bd, act=K.ACT_RELU), base.clamp_min(0), 2e-3, "relu")
    y = K.gemm(xd, wd, bias=bd, res=r32.to(dev()), out_dtype=torch.float32)
    close(y, base + r32, 1e-4, "res f32")
    # residual aliasing the output (in-place accumulate into the fp32 stream)
    acc = r32.to(dev()).clone()
    K.gemm(xd, wd, bias=bd, res=acc, out=acc)
    close(acc, base + r32, 1e-4, "res aliased")
    # q-scaling of the first columns (ViT q * dh^-0.5 folded into the qkv epilogue)
    y = K.gemm(xd, wd, bias=bd, scale_cols=128, scale=0.25)
    ref = base.clone()
    ref[:, :128] *= 0.25
    close(y, ref, 2e-3, "scale_cols")
    # strided output slice
    big = torch.zeros(T, 2 * F, device=dev(), dtype=torch.float32)
    K.gemm(xd, wd, bias=bd, out=big[:, F:])
    close(big[:, F:], base, 1e-4, "strided out")
    assert float(big[:, :F].abs().max()) == 0.0


def test_gemm_mn_major_operands(K):
    # dgrad / wgrad forms: operands stored [K, rows]
    T, F, Kd = 192, 256, 320
    x, w = rnd(T, Kd, seed=1).half(), (rnd(F, Kd, seed=2) / 16).half()
    ref = x.float() @ w.float().t()
    xd, wd = x.to(dev()), w.to(dev())
    xt, wt = xd.t().contiguous(), wd.t().contiguous()  # [K, T], [K, F]
    close(K.gemm(xd, wt, w_mn_major=True, F=F, K=Kd), ref, 2e-3, "w mn-major")
    close(K.gemm(xt, wd, x_mn_major=True, T=T, K=Kd, bn_hint=64), ref, 2e-3, "x mn-major")
    close(K.gemm(xt, wt, x_mn_major=True, w_mn_major=True, T=T, F=F, K=Kd, bn_hint=128), ref, 2e-3, "both mn-major")


def _interleave64(gate, up):
    """[I, K] gate and up rows -> [2I, K] in blocks of 64: gate 0..63, up 0..63, gate 64..127, ... (MYR_ACT_SWIGLU layout)."""
    I, Kd = gate.shape
    return torch.stack([gate.reshape(I // 64, 64, Kd), up.reshape(I // 64, 64, Kd)], 1).reshape(2 * I, Kd)


@pytest.mark.parametrize("T,I,Kd", [(4, 1024, 512), (4, 11008, 4096), (33, 512, 256), (300, 512, 256), (524, 2048, 1024)])
def test_gemm_swiglu_epilogue(K, T, I, Kd):
    """gate/up projection with the SwiGLU of modeling_llama.py:139-140 fused (both operand arrangements, split and
    unsplit tiles); rounding points identical to the unfused gemm + myr_swiglu path."""
    x = rnd(T, Kd, seed=1).half()
    g, u = (rnd(I, Kd, seed=2) / Kd ** 0.5).half(), (rnd(I, Kd, seed=3) / Kd ** 0.5).half()
    xd = x.to(dev())
    y = K.gemm(xd, _interleave64(g, u).to(dev()), act=K.ACT_SWIGLU)
    assert y.shape == (T, I)
    gu = K.gemm(xd, torch.cat([g, u]).to(dev()))
    ref = torch.empty(T, I, device=dev(), dtype=torch.float16)
    K.swiglu(gu, ref, T, I)
    # same rounding points as the unfused path; the k-ranges of split tiles differ between the two weight layouts, so the
    # fp32 sums may differ in their last bit and the fp16 results by one ulp
    close(y, ref, 1e-3, "fused vs unfused SwiGLU")
    assert (y != ref).float().mean().item() < 0.02
    gf, uf = x.float() @ g.float().t(), x.float() @ u.float().t()
    close(y, torch.nn.functional.silu(gf) * uf, 3e-3, "swiglu vs fp32")


@pytest.mark.parametrize("T,F,Kd", [(4, 22016, 4096), (4, 4096, 11008), (1, 32000, 4096), (16, 12304, 4096), (48, 768, 1408),
                                    (100, 256, 8192), (200, 512, 16384)])
def test_gemm_stream_k(K, T, F, Kd):
    """Shapes whose tiles are shared between CTAs (equal k-block ranges): the finisher's fixed-order sum must make the
    result reproducible run to run, and the arrival counters must be left clean for the next launch."""
    x, w = rnd(T, Kd, seed=1).half().to(dev()), (rnd(F, Kd, seed=2) / Kd ** 0.5).half().to(dev())
    r32 = rnd(T, F, seed=4).to(dev())
    ref = x.float() @ w.float().t() + r32
    y1 = K.gemm(x, w, res=r32, out_dtype=torch.float32)
    y2 = K.gemm(x, w, res=r32, out_dtype=torch.float32)
    close(y1, ref, 1e-3, "stream-K")
    assert torch.equal(y1, y2), "split tiles must reduce deterministically"
    y3 = K.gemm(x, w, res=r32, out_dtype=torch.float32, ksplit_hint=1)
    close(y3, ref, 1e-3, "data-parallel")


def test_gemm_small_t_mn_major_and_pdl_chain(K):
    """T <= 64 with MN-major operands (dgrad / wgrad forms in the lanes = features arrangement) and a chain of dependent
    GEMMs launched back to back with programmatic dependent launch (each consumes the previous one's output)."""
    T, F, Kd = 48, 320, 256
    x, w = rnd(T, Kd, seed=1).half(), (rnd(F, Kd, seed=2) / 16).half()
    ref = x.float() @ w.float().t()
    xd, wd = x.to(dev()), w.to(dev())
    xt, wt = xd.t().contiguous(), wd.t().contiguous()
    close(K.gemm(xd, wt, w_mn_major=True, F=F, K=Kd), ref, 2e-3, "w mn-major, small T")
    close(K.gemm(xt, wd, x_mn_major=True, T=T, K=Kd), ref, 2e-3, "x mn-major, small T")
    close(K.gemm(xt, wt, x_mn_major=True, w_mn_major=True, T=T, F=F, K=Kd), ref, 2e-3, "both mn-major, small T")
    D = 1024
    ws = [(rnd(D, D, seed=10 + i) / D ** 0.5).half().to(dev()) for i in range(6)]
    for T in (4, 200):
        h = rnd(T, D, seed=3).half().to(dev())
        cur, ref = h, h.float()
        for wi in ws:
            cur = K.gemm(cur, wi, w_static=True)
            ref = (ref @ wi.float().t()).half().float()
        close(cur, ref, 5e-3, "pdl chain T=%d" % T)


# ------------------------------------------------------------------------------------------- attention
def ref_attention(q, k, v, scale, causal=False, q_off=0, kv_len=None):
    # q [B,Sq,H,dh], k/v [B,Skv,H,dh] fp32
    s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    B, H, Sq, Skv = s.shape
    j = torch.arange(Skv)[None, None, None, :]
    i = torch.arange(Sq)[None, None, :, None] + q_off
    ok = torch.ones(B, 1, Sq, Skv, dtype=torch.bool)
    if causal:
        ok = ok & (j <= i)
    if kv_len is not None:
        ok = ok & (j < kv_len[:, None, None, None])
    s = s.masked_fill(~ok, -float("inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhij,bjhd->bihd", p, v)


@pytest.mark.parametrize("name,B,H,Sq,Skv,dh,causal,kvl,bn", [
    ("vit", 2, 4, 257, 257, 88, False, None, 0),
    ("qf_self", 2, 12, 81, 81, 64, False, None, 0),
    ("qf_cross", 2, 12, 81, 257, 64, False, None, 0),
    ("llama_prefill", 2, 4, 131, 131, 128, True, [131, 100], 0),
    ("llama_long", 1, 2, 700, 700, 128, True, None, 0),
    ("decode", 3, 4, 1, 256, 128, False, [200, 131, 7], 0),
    ("bn32", 1, 2, 40, 70, 64, False, None, 32),
    ("bn128", 1, 2, 300, 300, 128, True, None, 128),
])
def test_attention(K, name, B, H, Sq, Skv, dh, causal, kvl, bn):
    q = rnd(B, Sq, H, dh, seed=1).half()
    k = rnd(B, Skv, H, dh, seed=2).half()
    v = rnd(B, Skv, H, dh, seed=3).half()
    scale = 1.0 / math.sqrt(dh)
    kv_len = torch.tensor(kvl, dtype=torch.int32) if kvl is not None else None
    ref = ref_attention(q.float(), k.float(), v.float(), scale, causal, 0 if Sq > 1 else 0, kv_len)
    qd, kd, vd = q.to(dev()), k.to(dev()), v.to(dev())
    out = torch.full((B, Sq, H, dh), float("nan"), device=dev(), dtype=torch.float16)
    st = lambda t: (t.stride(1), t.stride(0), t.stride(2))
    K.attention(qd, kd, vd, out, B, H, Sq, Skv, dh, scale, st(qd), st(kd), st(vd), st(out), causal=causal,
                kv_len=kv_len.to(dev()) if kv_len is not None else None, bn_hint=bn)
    close(out, ref, 3e-3, name)


def test_attention_fused_qkv_layout(K):
    # ViT layout: one [B*N, 3*H*dh] buffer, q pre-scaled, output [B*N, H*dh]
    B, N, H, dh = 2, 257, 4, 88
    D = H * dh
    qkv = rnd(B * N, 3 * D, seed=5).half()
    ref = ref_attention(*(qkv.float().reshape(B, N, 3, H, dh)[:, :, i] for i in range(3)), 1.0)
    qd = qkv.to(dev())
    out = torch.empty(B * N, D, device=dev(), dtype=torch.float16)
    s = (3 * D, N * 3 * D, dh)
    K.attention(qd, qd[:, D:], qd[:, 2 * D:], out, B, H, N, N, dh, 1.0, s, s, s, (D, N * D, dh))
    close(out.reshape(B, N, H, dh), ref, 3e-3, "fused qkv")


# ----------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("D,rows", [(1408, 70), (768, 33), (4096, 9), (176, 5)])
def test_layernorm(K, O, D, rows):
    x, g, b = rnd(rows, D, seed=1), 1 + 0.1 * rnd(D, seed=2), 0.1 * rnd(D, seed=3)
    ref = O.layer_norm(x, g, b, 1e-6)
    o16 = torch.empty(rows, D, device=dev(), dtype=torch.float16)
    o32 = torch.empty(rows, D, device=dev(), dtype=torch.float32)
    K.norm(x.to(dev()), g.to(dev()), b.to(dev()), 1e-6, out16=o16, out32=o32)
    close(o32, ref, 1e-5, "ln f32")
    close(o16, ref, 1e-3, "ln f16")
    K.norm(x.half().to(dev()), g.to(dev()), b.to(dev()), 1e-6, out32=o32)
    close(o32, O.layer_norm(x.half().float(), g, b, 1e-6), 1e-5, "ln f16 in")


def test_rmsnorm(K, O):
    x, g = rnd(7, 4096, seed=1) * 3, 1 + 0.1 * rnd(4096, seed=2)
    o16 = torch.empty(7, 4096, device=dev(), dtype=torch.float16)
    K.norm(x.to(dev()), g.to(dev()), None, 1e-6, rms=True, out16=o16)
    close(o16, O.rms_norm(x, g, 1e-6), 1e-3, "rms")


def test_adaptor_layernorm_fused(K, O):
    D, rows = 1408, 40
    sd = {"expert_adaptor.conv1.weight": rnd(4, D, seed=1, std=0.05), "expert_adaptor.conv2.weight": rnd(D, 4, seed=2, std=0.05)}
    x, g, b = rnd(rows, D, seed=3), 1 + 0.1 * rnd(D, seed=4), 0.1 * rnd(D, seed=5)
    pre_ref = O.lora_adaptor(sd, x)
    ref = O.layer_norm(pre_ref, g, b, 1e-5)
    o16 = torch.empty(rows, D, device=dev(), dtype=torch.float16)
    pre = torch.empty(rows, D, device=dev(), dtype=torch.float32)
    stats = torch.empty(rows, 2, device=dev(), dtype=torch.float32)
    K.norm(x.to(dev()), g.to(dev()), b.to(dev()), 1e-5, out16=o16, w1=sd["expert_adaptor.conv1.weight"].to(dev()),
           w2=sd["expert_adaptor.conv2.weight"].to(dev()).contiguous(), pre32=pre, stats=stats)
    close(pre, pre_ref, 1e-5, "adaptor pre")
    close(o16, ref, 1e-3, "adaptor+ln")
    close(stats[:, 0], pre_ref.mean(-1), 1e-5, "mean")


# ------------------------------------------------------------------------------------------ rope/cache
def test_rope_cache(K, O):
    B, S, H, dh, Smax = 2, 5, 4, 128, 16
    qkv = rnd(B * S, 3 * H * dh, seed=1).half()
    pos = torch.tensor([[3, 4, 5, 6, 7], [0, 1, 2, 3, 4]], dtype=torch.int32)
    cos, sin = O.rope_tables(dh, 64)
    x = qkv.float().reshape(B, S, 3, H, dh)
    qr = O.apply_rope(x[:, :, 0].transpose(1, 2), cos, sin, pos.long()).transpose(1, 2)
    kr = O.apply_rope(x[:, :, 1].transpose(1, 2), cos, sin, pos.long()).transpose(1, 2)
    kc = torch.zeros(B, Smax, H * dh, device=dev(), dtype=torch.float16)
    vc = torch.zeros_like(kc)
    qd = qkv.to(dev())
    K.rope_cache(qd, B, S, H, dh, pos.reshape(-1).to(dev()), cos[:, :dh // 2].contiguous().to(dev()),
                 sin[:, :dh // 2].contiguous().to(dev()), kc, vc, cache_off=3)
    close(qd.reshape(B, S, 3, H, dh)[:, :, 0], qr, 2e-3, "q rope")
    close(kc[:, 3:3 + S].reshape(B, S, H, dh), kr, 2e-3, "k cache")
    close(vc[:, 3:3 + S].reshape(B, S, H, dh), x[:, :, 2], 1e-6, "v cache")
    assert float(kc[:, :3].abs().max()) == 0.0 and float(kc[:, 3 + S:].abs().max()) == 0.0
    off = torch.tensor([9], dtype=torch.int32, device=dev())
    K.rope_cache(qkv.to(dev()), B, S, H, dh, pos.reshape(-1).to(dev()), cos[:, :dh // 2].contiguous().to(dev()),
                 sin[:, :dh // 2].contiguous().to(dev()), kc, vc, cache_off_dev=off)
    close(kc[:, 9:9 + S].reshape(B, S, H, dh), kr, 2e-3, "k cache (device offset)")


def test_swiglu_embed_copy(K):
    T, I = 37, 1024
    gu = rnd(T, 2 * I, seed=1).half()
    out = torch.empty(T, I, device=dev(), dtype=torch.float16)
    K.swiglu(gu.to(dev()), out, T, I)
    close(out, torch.nn.functional.silu(gu[:, :I].float()) * gu[:, I:].float(), 2e-3, "swiglu")
    table = rnd(50, 256, seed=2).half()
    ids = torch.tensor([3, 49, 0, 7, 7], dtype=torch.int64)
    o32 = torch.empty(5, 256, device=dev(), dtype=torch.float32)
    K.embed(table.to(dev()), ids.to(dev()), o32)
    close(o32, table[ids].float(), 0.0, "embed f32")
    o16 = torch.empty(5, 256, device=dev(), dtype=torch.float16)
    K.embed(table.to(dev()), ids.int().to(dev()), o16)
    close(o16, table[ids].float(), 0.0, "embed f16")
    # place a [B, 9, D] group into rows 4..12 of a [B, 20, D] buffer with cast
    B, D = 3, 64
    src = rnd(B * 9, D, seed=3).half().to(dev())
    dst = torch.zeros(B, 20, D, device=dev(), dtype=torch.float32)
    K.copy_rows(src, dst[:, 4:], B, 9, D, D, 9 * D, D, 20 * D)
    close(dst[:, 4:13], src.float().reshape(B, 9, D), 0.0, "copy_rows")
    assert float(dst[:, :4].abs().max()) == 0.0 and float(dst[:, 13:].abs().max()) == 0.0


# --------------------------------------------------------------------------------------- ViT embedding
def test_patch_embed_assemble(K, O):
    d = syn.VitDims(img=56, dim=176, depth=1, heads=2, mlp_hidden=768)
    sd = syn.make_state_dict(syn.MyriadDims(vit=d, use_instructor=False, use_tokenizer=False), 0, only_prefix="visual_encoder")
    image, _ = syn.make_inputs(2, seed=3, img=56)
    ref = O.vit_patch_embed(sd, image, d)
    ref = torch.cat([sd["visual_encoder.cls_token"].expand(2, -1, -1), ref], 1) + sd["visual_encoder.pos_embed"]
    Kp = 3 * 14 * 14
    ldp = (Kp + 7) // 8 * 8
    patches = torch.empty(2 * 16, ldp, device=dev(), dtype=torch.float16)
    K.patchify(image.to(dev()), patches, 2, 3, 56, 14)
    w = torch.zeros(d.dim, ldp, dtype=torch.float16)
    w[:, :Kp] = sd["visual_encoder.patch_embed.proj.weight"].reshape(d.dim, Kp).half()
    pe = K.gemm(patches, w.to(dev()), bias=sd["visual_encoder.patch_embed.proj.bias"].half().to(dev()), out_dtype=torch.float32)
    x = torch.empty(2, 17, d.dim, device=dev(), dtype=torch.float32)
    K.vit_assemble(pe, sd["visual_encoder.cls_token"].to(dev()), sd["visual_encoder.pos_embed"].to(dev()), x, 2, 17, d.dim)
    close(x, ref, 2e-3, "patch embed")


# ------------------------------------------------------------------------------------------ conv stack
def test_conv_kernels(K, O):
    import torch.nn.functional as F
    B, H, W, Cin, Cout = 2, 12, 12, 4, 16
    x, w, b = rnd(B, Cin, H, W, seed=1).half().float(), rnd(Cout, Cin, 3, 3, seed=2, std=0.2), rnd(Cout, seed=3, std=0.1)
    ref = F.max_pool2d(F.relu(F.conv2d(x, w, b, padding=1)), 2)
    out = torch.empty(B, H // 2, W // 2, Cout, device=dev(), dtype=torch.float16)
    K.conv3x3_relu_pool(x.permute(0, 2, 3, 1).contiguous().half().to(dev()), w.permute(0, 2, 3, 1).contiguous().to(dev()),
                        b.to(dev()), out, B, H, W, Cin, Cout)
    close(out.permute(0, 3, 1, 2), ref, 1e-3, "direct conv")
    # im2col + GEMM + pool path
    Cin, Cout = 64, 128
    x, w, b = rnd(B, Cin, H, W, seed=4).half().float(), (rnd(Cout, Cin, 3, 3, seed=5) / 24).half().float(), rnd(Cout, seed=6, std=0.1).half().float()
    ref = F.max_pool2d(F.relu(F.conv2d(x, w, b, padding=1)), 2)
    xn = x.permute(0, 2, 3, 1).contiguous().half().to(dev())
    cols = torch.empty(B * H * W, 9 * Cin, device=dev(), dtype=torch.float16)
    K.im2col(xn, cols, B, H, W, Cin, 3, 3, 1)
    y = K.gemm(cols, w.permute(0, 2, 3, 1).reshape(Cout, -1).half().to(dev()), bias=b.half().to(dev()), act=K.ACT_RELU)
    pooled = torch.empty(B, H // 2, W // 2, Cout, device=dev(), dtype=torch.float16)
    K.maxpool2(y, pooled, B, H, W, Cout)
    close(pooled.permute(0, 3, 1, 2), ref, 2e-3, "im2col conv")


# ---------------------------------------------------------------------------------------------- greedy
def test_greedy_step(K):
    B, V, max_new = 3, 1000, 6
    state = torch.zeros(4 + 4 * B + B * max_new, dtype=torch.int32)
    state[4:4 + B] = 1                      # unfinished
    state[4 + 2 * B:4 + 3 * B] = 10         # kv_len
    state[4 + 3 * B:4 + 4 * B] = 9          # pos
    state[2] = 9                            # cache_off
    state = state.to(dev())
    scratch = torch.zeros(B, dtype=torch.int32, device=dev())
    stops = torch.tensor([[835, -1], [77, 88]], dtype=torch.int32, device=dev())
    seq = [[5, 77, 88], [2, 9, 9], [7, 7, 7]]  # row 1 hits eos at step 0 -> suppressed (min_new=1) -> second best
    for step in range(3):
        logits = torch.zeros(B, V)
        for b in range(B):
            logits[b, seq[b][step]] = 5.0
        logits[1, 11] = 4.0
        K.greedy_step(logits.to(dev()), state, scratch, B, V, max_new, 1, 2, stops, 2, 2)
    s = state.cpu()
    toks = s[4 + 4 * B:].reshape(B, max_new)
    assert toks[0, :3].tolist() == [5, 77, 88]
    assert toks[1, :3].tolist() == [11, 9, 9]      # eos suppressed at step 0 only
    assert int(s[0]) == 3 and int(s[1]) == 1        # stopped by row-0 stop sequence (77, 88)
    assert s[4 + 2 * B:4 + 3 * B].tolist() == [13] * B and s[4 + 3 * B:4 + 4 * B].tolist() == [12] * B and int(s[2]) == 12


@pytest.mark.parametrize("kv_cap", [0, 64, 96])  # 0: K / V through registers; else by TMA into shared memory
@pytest.mark.parametrize("lora", [False, True])
def test_decode_attention_matches_rope_plus_flash(K, O, lora, kv_cap):
    _decode_attention_case(K, O, lora, kv_cap, B=3, H=4, Smax=96, past=[40, 57, 1], off=60)


@pytest.mark.parametrize("off", [131, 147, 162])
@pytest.mark.parametrize("kv_cap", [0, 163])
def test_decode_attention_bench_shape(K, O, kv_cap, off):
    """The decode-attention launch of the benchmark itself (bench.py / BASELINE configs[2]): 32 heads, 4 sequences, cache of
    131 prefill + up to 32 new tokens => kv_cap = 163 (engine.py greedy_decode), visible cache 132 ... 163."""
    _decode_attention_case(K, O, True, kv_cap, B=4, H=32, Smax=163, past=[off, off, off - 7, 3], off=off)


@pytest.mark.parametrize("off", [280, 600, 1023, 1024, 2070])
def test_decode_attention_long_cache_split(K, O, off):
    """Caches beyond 256 tokens (the sweep's S = 256 / 1024 / 2048): one CTA per 128-key chunk + ordered combine, against the same
    references; `off` puts the new token's slot inside a chunk, at a chunk's last key, at a chunk's first key, near the end."""
    _decode_attention_case(K, O, True, 0, B=2, H=4, Smax=2080, past=[off, off - 5], off=off, split=True)


@pytest.mark.parametrize("B,H,Smax,off", [(16, 32, 320, 290), (1, 32, 320, 299), (32, 32, 320, 258), (5, 32, 1100, 1050), (3, 8, 4100, 4090)])
def test_decode_attention_stream_many_rows(K, O, B, H, Smax, off):
    """The persistent long-cache kernel at the throughput sweep's shapes: batch 16 / 32 with 32 heads (several (row, head) segments per
    CTA, segments cut between CTAs and combined from partial results), a single row (fewer chunks than SMs: every (row, head) is split
    over three CTAs), five rows over 1050 slots, a 4090-slot cache. Same references as the other decode-attention tests."""
    _decode_attention_case(K, O, True, 0, B=B, H=H, Smax=Smax, past=[off - (i % 7) for i in range(B)], off=off, split=True)


def _decode_attention_case(K, O, lora, kv_cap, B, H, Smax, past, off, split=False):
    """myr_decode_attention (one launch per decode step and layer) against the prefill pair myr_rope_cache + myr_attention_fwd
    on the same cache, and against the oracle's rotary / attention arithmetic."""
    torch.manual_seed(4)
    dh, r = 128, 8
    D = H * dh
    ldq = 3 * D + (2 * r if lora else 0)
    kc = (torch.randn(B, Smax, D) * 0.5).half()
    vc = (torch.randn(B, Smax, D) * 0.5).half()
    qkv = (torch.randn(B, ldq) * 0.5).half()
    bq, bv = (torch.randn(D, r) * 0.05).half(), (torch.randn(D, r) * 0.05).half()
    # the graph-replayed decode writes every row's new token to the same slot `off`; rows are right-aligned by kv_len
    # make "past" keys of each row live in slots [off - past, off)
    pos = torch.tensor([p for p in past], dtype=torch.int32)
    kv_len = torch.tensor([off + 1] * B, dtype=torch.int32)
    cos, sin = O.rope_tables(dh, max(256, Smax + 8))
    half = dh // 2
    cosd, sind = cos[:, :half].contiguous().to(dev()), sin[:, :half].contiguous().to(dev())
    lo = (bq.to(dev()), bv.to(dev()), r, 2.0) if lora else None
    # reference pair on its own copy of the cache
    kc1, vc1, q1 = kc.to(dev()).clone(), vc.to(dev()).clone(), qkv.to(dev()).clone()
    K.rope_cache(q1, B, 1, H, dh, pos.to(dev()), cosd, sind, kc1, vc1, cache_off=off, lora=lo)
    ref = torch.empty(B, D, device=dev(), dtype=torch.float16)
    cs = (kc1.stride(1), kc1.stride(0), dh)
    K.attention(q1, kc1, vc1, ref, B, H, 1, Smax, dh, 1.0 / math.sqrt(dh), (ldq, ldq, dh), cs, cs, (D, D, dh), causal=False,
                kv_len=kv_len.to(dev()))
    kc2, vc2 = kc.to(dev()).clone(), vc.to(dev()).clone()
    out = torch.empty(B, D, device=dev(), dtype=torch.float16)
    ws = torch.zeros(K.decode_attn_split_bytes(B, H, Smax), device=dev(), dtype=torch.uint8) if split else None
    for rep in range(2 if split else 1):  # second launch: the arrival counters must have been left at zero
        K.decode_attention(qkv.to(dev()), B, H, dh, pos.to(dev()), cosd, sind, kc2, vc2, kv_len.to(dev()), out, 1.0 / math.sqrt(dh),
                           cache_off=off, lora=lo, kv_cap=kv_cap, split_ws=ws)
    if split:
        assert int(ws[:B * H * 4].view(torch.int32).abs().sum()) == 0
    close(kc2, kc1, 1e-3, "appended k")
    close(vc2, vc1, 1e-3, "appended v")
    print("cache append: %d k / %d v values differ in the last bit from myr_rope_cache" % (
        int((kc1 != kc2).sum()), int((vc1 != vc2).sum())))
    close(out, ref, 2e-3, "decode attention vs rope+flash")
    # oracle arithmetic (fp32)
    x = qkv.float()
    q, k, v = x[:, :D], x[:, D:2 * D], x[:, 2 * D:3 * D]
    if lora:
        q = q + 2.0 * x[:, 3 * D:3 * D + r] @ bq.float().t()
        v = v + 2.0 * x[:, 3 * D + r:] @ bv.float().t()
    q = O.apply_rope(q.reshape(B, 1, H, dh).transpose(1, 2), cos, sin, pos.long()[:, None])
    k = O.apply_rope(k.reshape(B, 1, H, dh).transpose(1, 2), cos, sin, pos.long()[:, None])
    kk = kc.float().reshape(B, Smax, H, dh).transpose(1, 2).clone()
    vv = vc.float().reshape(B, Smax, H, dh).transpose(1, 2).clone()
    kk[:, :, off] = k[:, :, 0]
    vv[:, :, off] = v.reshape(B, H, dh)
    s = (q @ kk[:, :, :off + 1].transpose(-1, -2)) / math.sqrt(dh)
    o = (torch.softmax(s, -1) @ vv[:, :, :off + 1]).transpose(1, 2).reshape(B, D)
    close(out, o, 3e-3, "decode attention vs oracle arithmetic")
