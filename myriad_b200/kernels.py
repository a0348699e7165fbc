"""Thin Python bindings over the C-ABI (include/myriad_b200.h). PyTorch supplies device memory and the
current CUDA stream only; all arithmetic happens in libmyriad_b200.so.
"""
import ctypes

import torch

from ._lib import check, lib

F16, F32 = 0, 1
ACT_NONE, ACT_GELU = 0, 1


def _dt(t):
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.float32:
        return F32
    raise TypeError("unsupported dtype %s" % t.dtype)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("x", ctypes.c_void_p), ("ldx", ctypes.c_int64),
        ("w", ctypes.c_void_p), ("ldw", ctypes.c_int64),
        ("T", ctypes.c_int32), ("F", ctypes.c_int32), ("K", ctypes.c_int32),
        ("x_mn_major", ctypes.c_int32), ("w_mn_major", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("act", ctypes.c_int32),
        ("round_acc", ctypes.c_int32),
        ("scale_cols", ctypes.c_int32), ("scale", ctypes.c_float),
        ("res", ctypes.c_void_p), ("res_dtype", ctypes.c_int32), ("ldr", ctypes.c_int64),
        ("out", ctypes.c_void_p), ("out_dtype", ctypes.c_int32), ("ldo", ctypes.c_int64),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("bn_hint", ctypes.c_int32), ("ksplit_hint", ctypes.c_int32),
        ("out_group_rows", ctypes.c_int32), ("out_group_stride", ctypes.c_int64),
        ("nb0", ctypes.c_int32), ("nb1", ctypes.c_int32),
        ("x_bs0", ctypes.c_int64), ("x_bs1", ctypes.c_int64), ("w_bs0", ctypes.c_int64), ("w_bs1", ctypes.c_int64),
        ("o_bs0", ctypes.c_int64), ("o_bs1", ctypes.c_int64),
        ("alpha_set", ctypes.c_int32), ("alpha", ctypes.c_float),
        ("pdl", ctypes.c_int32), ("w_static", ctypes.c_int32),
        ("post_gamma", ctypes.c_void_p), ("post_out16", ctypes.c_void_p), ("post_ld", ctypes.c_int64), ("post_ss", ctypes.c_void_p),
        ("norm_ss", ctypes.c_void_p), ("norm_eps", ctypes.c_float),
    ]


NORM_SS_FLOATS = 4 + 4 * 1024  # size of a post_ss / norm_ss buffer: header + per-8-row-item sums of squares (F <= 8192)


_ws_cache = {}


def workspace(nbytes, device):
    """Grow-only per-device scratch buffer (arrival counters + split-tile partials). Owned by PyTorch, lent to the
    library; zero-filled on creation because the counters at its head must start at zero (include/myriad_b200.h)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 96 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def gemm(x, w, bias=None, act=ACT_NONE, res=None, out=None, out_dtype=torch.float16, scale_cols=0, scale=1.0,
         round_acc=False, x_mn_major=False, w_mn_major=False, T=None, F=None, K=None, bn_hint=0, ksplit_hint=0,
         out_group_rows=0, out_group_stride=0, ldo=None, ldx=None, ldw=None, batch=None, alpha=None, pdl=True, w_static=False,
         post_norm=None, norm_ss=None):
    # batch = (nb0, nb1, (x_bs0, x_bs1), (w_bs0, w_bs1), (o_bs0, o_bs1)): independent problems, element strides
    """out[t, f] = epilogue(sum_k x[t, k] * w[f, k]);  x: [T, K] fp16 (row stride may exceed K), w: [F, K] fp16.
    post_norm = (gamma [F] fp32, y16 [T, F] fp16, ss [NORM_SS_FLOATS] fp32): producer side of an RMSNorm hand-over (T <= 4): also
    writes y16 = rn_f16(out * gamma) and the sums of out^2 over each block of 8 features; the consumer passes x = y16, norm_ss = (ss, eps)."""
    assert x.dtype == torch.float16 and w.dtype == torch.float16
    assert x.stride(-1) == 1 and w.stride(-1) == 1
    assert batch is None or (T is not None and F is not None and K is not None and out is not None and ldo is not None)
    if T is None:
        T = x.shape[0] if not x_mn_major else x.shape[1]
    if K is None:
        K = x.shape[1] if not x_mn_major else x.shape[0]
    if F is None:
        F = w.shape[0] if not w_mn_major else w.shape[1]
    if out is None:
        out = torch.empty((T, F // 2 if act == ACT_SWIGLU else F), dtype=out_dtype, device=x.device)
    ws = workspace(96 << 20, x.device)
    a = GemmArgs()
    a.x, a.ldx = x.data_ptr(), (ldx if ldx is not None else x.stride(0))
    a.w, a.ldw = w.data_ptr(), (ldw if ldw is not None else w.stride(0))
    if batch is not None:
        a.nb0, a.nb1 = batch[0], batch[1]
        (a.x_bs0, a.x_bs1), (a.w_bs0, a.w_bs1), (a.o_bs0, a.o_bs1) = batch[2], batch[3], batch[4]
    a.T, a.F, a.K = T, F, K
    a.x_mn_major, a.w_mn_major = int(x_mn_major), int(w_mn_major)
    a.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float16 and bias.numel() >= F
    a.act = act
    a.round_acc = int(round_acc)
    a.scale_cols, a.scale = scale_cols, scale
    if res is not None:
        assert res.stride(-1) == 1
        a.res, a.res_dtype, a.ldr = res.data_ptr(), _dt(res), res.stride(0)
    else:
        a.res, a.res_dtype, a.ldr = None, 0, 0
    assert out.stride(-1) == 1
    a.out, a.out_dtype, a.ldo = out.data_ptr(), _dt(out), (ldo if ldo is not None else out.stride(0))
    assert out.dim() == 2 or ldo is not None, "flat output needs an explicit row stride"
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.bn_hint, a.ksplit_hint = bn_hint, ksplit_hint
    a.out_group_rows, a.out_group_stride = out_group_rows, out_group_stride
    if alpha is not None:
        a.alpha_set, a.alpha = 1, alpha
    a.pdl, a.w_static = int(pdl), int(w_static)
    if post_norm is not None:
        pg, y16, ss = post_norm
        assert pg.dtype == torch.float32 and y16.dtype == torch.float16 and ss.dtype == torch.float32 and ss.numel() >= NORM_SS_FLOATS
        a.post_gamma, a.post_out16, a.post_ld, a.post_ss = pg.data_ptr(), y16.data_ptr(), y16.stride(0), ss.data_ptr()
    if norm_ss is not None:
        assert norm_ss[0].dtype == torch.float32
        a.norm_ss, a.norm_eps = norm_ss[0].data_ptr(), norm_ss[1]
    check(lib().myr_gemm_f16(ctypes.byref(a), _stream()), "myr_gemm_f16")
    return out


ACT_RELU = 2
ACT_SWIGLU = 3


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("q", ctypes.c_void_p), ("q_ts", ctypes.c_int64), ("q_bs", ctypes.c_int64), ("q_hs", ctypes.c_int64),
        ("k", ctypes.c_void_p), ("k_ts", ctypes.c_int64), ("k_bs", ctypes.c_int64), ("k_hs", ctypes.c_int64),
        ("v", ctypes.c_void_p), ("v_ts", ctypes.c_int64), ("v_bs", ctypes.c_int64), ("v_hs", ctypes.c_int64),
        ("out", ctypes.c_void_p), ("o_ts", ctypes.c_int64), ("o_bs", ctypes.c_int64), ("o_hs", ctypes.c_int64),
        ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("Sq", ctypes.c_int32), ("Skv", ctypes.c_int32),
        ("dh", ctypes.c_int32),
        ("scale", ctypes.c_float),
        ("causal", ctypes.c_int32), ("q_off", ctypes.c_int32),
        ("kv_len", ctypes.c_void_p),
        ("bn_hint", ctypes.c_int32),
    ]


def attention(q, k, v, out, B, H, Sq, Skv, dh, scale, q_strides, k_strides, v_strides, o_strides, causal=False, q_off=0,
              kv_len=None, bn_hint=0):
    """Strided flash attention. q/k/v/out: fp16 tensors (any view; only data_ptr is used); *_strides =
    (token, batch, head) element strides."""
    a = AttnArgs()
    a.q, (a.q_ts, a.q_bs, a.q_hs) = q.data_ptr(), q_strides
    a.k, (a.k_ts, a.k_bs, a.k_hs) = k.data_ptr(), k_strides
    a.v, (a.v_ts, a.v_bs, a.v_hs) = v.data_ptr(), v_strides
    a.out, (a.o_ts, a.o_bs, a.o_hs) = out.data_ptr(), o_strides
    a.B, a.H, a.Sq, a.Skv, a.dh = B, H, Sq, Skv, dh
    a.scale = scale
    a.causal, a.q_off = int(causal), q_off
    a.kv_len = kv_len.data_ptr() if kv_len is not None else None
    a.bn_hint = bn_hint
    check(lib().myr_attention_fwd(ctypes.byref(a), _stream()), "myr_attention_fwd")
    return out


class NormArgs(ctypes.Structure):
    _fields_ = [
        ("x", ctypes.c_void_p), ("x_dtype", ctypes.c_int32), ("ldx", ctypes.c_int64),
        ("rows", ctypes.c_int32), ("D", ctypes.c_int32),
        ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("eps", ctypes.c_float), ("rms", ctypes.c_int32),
        ("w1", ctypes.c_void_p), ("w2", ctypes.c_void_p), ("rank", ctypes.c_int32),
        ("out16", ctypes.c_void_p), ("ld16", ctypes.c_int64),
        ("out32", ctypes.c_void_p), ("ld32", ctypes.c_int64),
        ("pre32", ctypes.c_void_p), ("ldpre", ctypes.c_int64),
        ("stats", ctypes.c_void_p),
    ]


def norm(x, gamma, beta, eps, rms=False, out16=None, out32=None, w1=None, w2=None, pre32=None, stats=None):
    """x: [rows, D] fp16/fp32 (row stride free). gamma/beta/w1/w2 fp32."""
    rows, D = x.shape
    a = NormArgs()
    a.x, a.x_dtype, a.ldx = x.data_ptr(), _dt(x), x.stride(0)
    a.rows, a.D = rows, D
    assert gamma.dtype == torch.float32 and (beta is None or beta.dtype == torch.float32)
    a.gamma, a.beta = gamma.data_ptr(), (beta.data_ptr() if beta is not None else None)
    a.eps, a.rms = eps, int(rms)
    if w1 is not None:
        assert w1.dtype == torch.float32 and w2.dtype == torch.float32 and w1.is_contiguous() and w2.is_contiguous()
        a.w1, a.w2, a.rank = w1.data_ptr(), w2.data_ptr(), w1.shape[0]
    for name, t, ld, dt in (("out16", out16, "ld16", torch.float16), ("out32", out32, "ld32", torch.float32),
                            ("pre32", pre32, "ldpre", torch.float32)):
        if t is not None:
            assert t.dtype == dt and t.stride(-1) == 1
            setattr(a, name, t.data_ptr())
            setattr(a, ld, t.stride(0))
    a.stats = stats.data_ptr() if stats is not None else None
    check(lib().myr_norm_fwd(ctypes.byref(a), _stream()), "myr_norm_fwd")


class RopeArgs(ctypes.Structure):
    _fields_ = [
        ("qkv", ctypes.c_void_p), ("ldq", ctypes.c_int64),
        ("B", ctypes.c_int32), ("S", ctypes.c_int32), ("H", ctypes.c_int32), ("dh", ctypes.c_int32),
        ("pos", ctypes.c_void_p),
        ("cos", ctypes.c_void_p), ("sin", ctypes.c_void_p),
        ("kcache", ctypes.c_void_p), ("vcache", ctypes.c_void_p), ("c_ts", ctypes.c_int64), ("c_bs", ctypes.c_int64),
        ("cache_off", ctypes.c_int32), ("cache_off_dev", ctypes.c_void_p),
        ("lora_bq", ctypes.c_void_p), ("lora_bv", ctypes.c_void_p), ("lora_r", ctypes.c_int32), ("lora_scale", ctypes.c_float),
    ]


def rope_cache(qkv, B, S, H, dh, pos, cos, sin, kcache, vcache, cache_off=0, cache_off_dev=None, lora=None):
    """lora = (B_q fp16 [H*dh, r], B_v fp16 [H*dh, r], r, scale): xa = x A^T is read from qkv[:, 3*H*dh:3*H*dh + 2r]."""
    a = RopeArgs()
    a.qkv, a.ldq = qkv.data_ptr(), qkv.stride(0)
    a.B, a.S, a.H, a.dh = B, S, H, dh
    assert pos.dtype == torch.int32
    a.pos, a.cos, a.sin = pos.data_ptr(), cos.data_ptr(), sin.data_ptr()
    a.kcache, a.vcache = kcache.data_ptr(), vcache.data_ptr()
    a.c_ts, a.c_bs = kcache.stride(1), kcache.stride(0)
    a.cache_off = cache_off
    a.cache_off_dev = cache_off_dev.data_ptr() if cache_off_dev is not None else None
    if lora is not None:
        bq, bv, r, scale = lora
        assert bq.dtype == torch.float16 and bv.dtype == torch.float16 and bq.is_contiguous() and bv.is_contiguous()
        a.lora_bq, a.lora_bv, a.lora_r, a.lora_scale = bq.data_ptr(), bv.data_ptr(), r, scale
    check(lib().myr_rope_cache(ctypes.byref(a), _stream()), "myr_rope_cache")


class DecodeAttnArgs(ctypes.Structure):
    _fields_ = [
        ("qkv", ctypes.c_void_p), ("ldq", ctypes.c_int64),
        ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("dh", ctypes.c_int32),
        ("pos", ctypes.c_void_p), ("cos", ctypes.c_void_p), ("sin", ctypes.c_void_p),
        ("kcache", ctypes.c_void_p), ("vcache", ctypes.c_void_p), ("c_ts", ctypes.c_int64), ("c_bs", ctypes.c_int64),
        ("cache_len", ctypes.c_int32),
        ("cache_off", ctypes.c_int32), ("cache_off_dev", ctypes.c_void_p),
        ("kv_len", ctypes.c_void_p),
        ("lora_bq", ctypes.c_void_p), ("lora_bv", ctypes.c_void_p), ("lora_r", ctypes.c_int32), ("lora_scale", ctypes.c_float),
        ("scale", ctypes.c_float),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64),
        ("next_layer_stride", ctypes.c_int64),
        ("kv_cap", ctypes.c_int32),
        ("split_ws", ctypes.c_void_p), ("split_ws_bytes", ctypes.c_size_t),
    ]


def decode_attn_split_bytes(B, H, cache_len):
    """Workspace (zero-initialised by the caller, once) of the long-cache decode attention: counters + partial results."""
    return int(lib().myr_decode_attention_ws_bytes(B, H, cache_len)) + 64


def decode_attention(qkv, B, H, dh, pos, cos, sin, kcache, vcache, kv_len, out, scale, cache_off=0, cache_off_dev=None, lora=None,
                     next_layer_stride=0, kv_cap=0, split_ws=None):
    """One new token per sequence: LoRA-B + RoPE + KV-cache append + attention over the cache in one launch."""
    a = DecodeAttnArgs()
    a.qkv, a.ldq = qkv.data_ptr(), qkv.stride(0)
    a.B, a.H, a.dh = B, H, dh
    assert pos.dtype == torch.int32 and kv_len.dtype == torch.int32
    a.pos, a.cos, a.sin = pos.data_ptr(), cos.data_ptr(), sin.data_ptr()
    a.kcache, a.vcache = kcache.data_ptr(), vcache.data_ptr()
    a.c_ts, a.c_bs, a.cache_len = kcache.stride(1), kcache.stride(0), kcache.shape[1]
    a.cache_off = cache_off
    a.cache_off_dev = cache_off_dev.data_ptr() if cache_off_dev is not None else None
    a.kv_len = kv_len.data_ptr()
    if lora is not None:
        bq, bv, r, s = lora
        a.lora_bq, a.lora_bv, a.lora_r, a.lora_scale = bq.data_ptr(), bv.data_ptr(), r, s
    a.scale = scale
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    a.next_layer_stride = next_layer_stride
    a.kv_cap = kv_cap
    if split_ws is not None:
        a.split_ws, a.split_ws_bytes = split_ws.data_ptr(), split_ws.numel() * split_ws.element_size()
    check(lib().myr_decode_attention(ctypes.byref(a), _stream()), "myr_decode_attention")
    return out


MEGA_GEMM, MEGA_ATTN, MEGA_EMBED = 0, 1, 2
MEGA_F16, MEGA_RES32, MEGA_SWIGLU, MEGA_F32 = 0, 1, 2, 3


class MegaOp(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32), ("epi", ctypes.c_int32),
        ("x", ctypes.c_void_p), ("ldx", ctypes.c_int64), ("w", ctypes.c_void_p), ("ldw", ctypes.c_int64),
        ("T", ctypes.c_int32), ("F", ctypes.c_int32), ("K", ctypes.c_int32),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64),
        ("norm_src", ctypes.c_void_p), ("norm_dst", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("eps", ctypes.c_float),
        ("D", ctypes.c_int32), ("norm_rows", ctypes.c_int32),
        ("table", ctypes.c_void_p), ("ids", ctypes.c_void_p), ("h32", ctypes.c_void_p),
        ("attn", DecodeAttnArgs),
    ]


def _mega_tail(op, norm):
    if norm is not None:
        src, dst, gamma, eps = norm
        assert src.dtype == torch.float32 and dst.dtype == torch.float16 and gamma.dtype == torch.float32
        assert src.is_contiguous() and dst.is_contiguous()
        op.norm_src, op.norm_dst, op.gamma, op.eps = src.data_ptr(), dst.data_ptr(), gamma.data_ptr(), eps
        op.D, op.norm_rows = src.shape[1], src.shape[0]


def mega_gemm(x, w, out, epi, norm=None):
    """x fp16 [T, K], w fp16 [F, K], out [T, F] (or [T, F/2] for SwiGLU). norm = (src32, dst16, gamma, eps) op tail."""
    op = MegaOp()
    op.kind, op.epi = MEGA_GEMM, epi
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.stride(1) == 1 and w.stride(1) == 1
    op.x, op.ldx, op.w, op.ldw = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0)
    op.T, op.F, op.K = x.shape[0], w.shape[0], w.shape[1]
    assert x.shape[1] == w.shape[1]
    assert out.dtype == (torch.float16 if epi in (MEGA_F16, MEGA_SWIGLU) else torch.float32)
    op.out, op.ldo = out.data_ptr(), out.stride(0)
    _mega_tail(op, norm)
    return op


def mega_embed(table, ids, h32, norm=None):
    op = MegaOp()
    op.kind = MEGA_EMBED
    assert table.dtype == torch.float16 and ids.dtype == torch.int32 and h32.dtype == torch.float32 and h32.is_contiguous()
    op.table, op.ids, op.h32, op.T, op.D = table.data_ptr(), ids.data_ptr(), h32.data_ptr(), ids.numel(), table.shape[1]
    _mega_tail(op, norm)
    return op


def mega_attn(qkv, B, H, dh, pos, cos, sin, kcache, vcache, kv_len, out, scale, cache_off_dev, lora=None):
    op = MegaOp()
    op.kind = MEGA_ATTN
    a = op.attn
    a.qkv, a.ldq = qkv.data_ptr(), qkv.stride(0)
    a.B, a.H, a.dh = B, H, dh
    a.pos, a.cos, a.sin = pos.data_ptr(), cos.data_ptr(), sin.data_ptr()
    a.kcache, a.vcache = kcache.data_ptr(), vcache.data_ptr()
    a.c_ts, a.c_bs, a.cache_len = kcache.stride(1), kcache.stride(0), kcache.shape[1]
    a.cache_off_dev = cache_off_dev.data_ptr()
    a.kv_len = kv_len.data_ptr()
    if lora is not None:
        bq, bv, r, s = lora
        a.lora_bq, a.lora_bv, a.lora_r, a.lora_scale = bq.data_ptr(), bv.data_ptr(), r, s
    a.scale = scale
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    return op


class MegaPlan:
    """Device blob (tensor maps + op records) and zeroed workspace of one persistent decode-step launch. Holds references
    to nothing: the caller keeps every tensor named by the ops alive."""

    def __init__(self, ops, device):
        n = len(ops)
        arr = (MegaOp * n)(*ops)
        nbytes = lib().myr_mega_plan_bytes(n)
        host = torch.zeros(nbytes + 64, dtype=torch.uint8)
        off = (-host.data_ptr()) % 64
        ws_bytes, n_cnt = ctypes.c_size_t(0), ctypes.c_int32(0)
        check(lib().myr_mega_plan(arr, n, ctypes.c_void_p(host.data_ptr() + off), ctypes.c_size_t(nbytes), ctypes.byref(ws_bytes),
                                  ctypes.byref(n_cnt)), "myr_mega_plan")
        self.n_ops, self.n_counters = n, n_cnt.value
        self.blob = host[off:off + nbytes].to(device)
        assert self.blob.data_ptr() % 64 == 0
        self.workspace = torch.zeros(max(ws_bytes.value, 64), dtype=torch.uint8, device=device)

    def launch(self, trace=None):
        check(lib().myr_mega_launch(_p(self.blob), self.n_ops, _p(self.workspace), ctypes.c_size_t(self.workspace.numel()),
                                    self.n_counters, _p(trace), _stream()), "myr_mega_launch")


def _i64(v):
    return ctypes.c_int64(v)


def swiglu(gate_up, out, T, I):
    check(lib().myr_swiglu(_p(gate_up), _i64(gate_up.stride(0)), _p(out), _i64(out.stride(0)), T, I, _stream()), "myr_swiglu")


def embed(table, ids, out):
    assert table.dtype == torch.float16 and ids.dtype in (torch.int32, torch.int64)
    check(lib().myr_embed(_p(table), table.shape[1], _p(ids), int(ids.dtype == torch.int64), ids.numel(), _p(out), _dt(out),
                          _i64(out.stride(0)), _stream()), "myr_embed")


def copy_rows(src, dst, groups, rows_per_group, D, src_ld, src_gs, dst_ld, dst_gs):
    check(lib().myr_copy_rows(_p(src), _dt(src), _i64(src_ld), _i64(src_gs), _p(dst), _dt(dst), _i64(dst_ld), _i64(dst_gs),
                              groups, rows_per_group, D, _stream()), "myr_copy_rows")


def vit_assemble(patch, cls, pos, x, B, N, D):
    check(lib().myr_vit_assemble(_p(patch), _p(cls), _p(pos), _p(x), B, N, D, _stream()), "myr_vit_assemble")


def patchify(image, patches, B, C, HW, P):
    check(lib().myr_patchify(_p(image), _p(patches), B, C, HW, P, patches.stride(0), _stream()), "myr_patchify")


def conv3x3_relu_pool(x, w, bias, out, B, H, W, Cin, Cout):
    check(lib().myr_conv3x3_relu_pool(_p(x), _dt(x), _p(w), _p(bias), _p(out), B, H, W, Cin, Cout, _stream()),
          "myr_conv3x3_relu_pool")


def im2col(x, out, B, H, W, C, KH, KW, pad):
    check(lib().myr_im2col(_p(x), _p(out), B, H, W, C, KH, KW, pad, _stream()), "myr_im2col")


def maxpool2(x, out, B, H, W, C):
    check(lib().myr_maxpool2(_p(x), _p(out), B, H, W, C, _stream()), "myr_maxpool2")


def greedy_step(logits, state, scratch, B, V, max_new, min_new, eos, stops, n_stops, stop_max_len):
    check(lib().myr_greedy_step(_p(logits), _i64(logits.stride(0)), B, V, _p(state), _p(scratch), max_new, min_new, eos,
                                _p(stops), n_stops, stop_max_len, _stream()), "myr_greedy_step")


def set_gemv(enabled):
    """Route T <= 4 GEMMs to the small-batch CUDA-core kernel (True, default) or to the tcgen05 kernel; returns the old setting."""
    return bool(lib().myr_set_gemv(int(bool(enabled))))


_graph_replay_launches = 0


def launch_count():
    """Kernels of this library launched so far in this process: eager launches + (nodes captured) x (graph replays)."""
    return int(lib().myr_launch_count()) + _graph_replay_launches


def note_graph_replay(n_nodes):
    global _graph_replay_launches
    _graph_replay_launches += n_nodes


# ----------------------------------------------------------------------------------------------- training kernels
def _f32(v):
    return ctypes.c_float(v)


def clamp_ce_fwd(logits, labels, row_loss, stats, loss_out):
    R, V = logits.shape
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64
    check(lib().myr_clamp_ce_fwd(_p(logits), _i64(logits.stride(0)), R, V, _p(labels), _p(row_loss), _p(stats), _p(loss_out), _stream()),
          "myr_clamp_ce_fwd")


def clamp_ce_bwd(logits, labels, stats, loss_out, loss_scale, dlogits):
    R, V = logits.shape
    check(lib().myr_clamp_ce_bwd(_p(logits), _i64(logits.stride(0)), R, V, _p(labels), _p(stats), _p(loss_out), _f32(loss_scale),
                                 _p(dlogits), _i64(dlogits.stride(0)), _stream()), "myr_clamp_ce_bwd")


def norm_bwd(x, dy, gamma, eps, rms=False, add=None, out32=None, out16=None):
    rows, D = x.shape
    assert x.dtype == torch.float32
    check(lib().myr_norm_bwd(_p(x), _i64(x.stride(0)), _p(dy), _dt(dy), _i64(dy.stride(0)), _p(gamma), _f32(eps), int(rms), rows, D,
                             _p(add), _i64(add.stride(0) if add is not None else 0), _p(out32),
                             _i64(out32.stride(0) if out32 is not None else 0), _p(out16),
                             _i64(out16.stride(0) if out16 is not None else 0), _stream()), "myr_norm_bwd")


def swiglu_bwd(gu, dact, dgu, T, I):
    check(lib().myr_swiglu_bwd(_p(gu), _i64(gu.stride(0)), _p(dact), _i64(dact.stride(0)), _p(dgu), _i64(dgu.stride(0)), T, I, _stream()),
          "myr_swiglu_bwd")


def _u64(v):
    return ctypes.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)


def dropout_fwd(x, out, p, seed, offset):
    rows, D = x.shape
    check(lib().myr_dropout_fwd(_p(x), _i64(x.stride(0)), _p(out), _i64(out.stride(0)), rows, D, _f32(p), _u64(seed), _u64(offset), _stream()),
          "myr_dropout_fwd")
    return out


def dropout_bwd_add(g, acc, p, seed, offset):
    rows, D = g.shape
    assert g.dtype == torch.float32 and acc.dtype == torch.float32
    check(lib().myr_dropout_bwd_add(_p(g), _i64(g.stride(0)), _p(acc), _i64(acc.stride(0)), rows, D, _f32(p), _u64(seed), _u64(offset),
                                    _stream()), "myr_dropout_bwd_add")


def dropout_mask(out_u8, p, seed, offset):
    check(lib().myr_dropout_mask(_p(out_u8), _i64(out_u8.numel()), _f32(p), _u64(seed), _u64(offset), _stream()), "myr_dropout_mask")
    return out_u8


def gelu_fwd(pre, out):
    check(lib().myr_gelu_fwd(_p(pre), _p(out), _i64(pre.numel()), _stream()), "myr_gelu_fwd")


def gelu_bwd(pre, dy, dpre):
    check(lib().myr_gelu_bwd(_p(pre), _p(dy), _p(dpre), _i64(pre.numel()), _stream()), "myr_gelu_bwd")


def rope_bwd(dqkv, T, H, dh, pos, cos, sin):
    check(lib().myr_rope_bwd(_p(dqkv), _i64(dqkv.stride(0)), T, H, dh, _p(pos), _p(cos), _p(sin), _stream()), "myr_rope_bwd")


def softmax_rows(S, P, B, H, Sq, Skv, cols, scale, causal=False, kv_len=None):
    check(lib().myr_softmax_rows(_p(S), _i64(cols), _p(P), _i64(cols), B, H, Sq, Skv, cols, _f32(scale), int(causal), _p(kv_len),
                                 _stream()), "myr_softmax_rows")


def softmax_bwd_rows(P, dP, dS, n_rows, cols, scale):
    check(lib().myr_softmax_bwd_rows(_p(P), _i64(cols), _p(dP), _i64(cols), _p(dS), _i64(cols), _i64(n_rows), cols, _f32(scale),
                                     _stream()), "myr_softmax_bwd_rows")


def attn_bwd_small_supported(Sq, Skv, dh):
    return bool(lib().myr_attn_bwd_small_supported(Sq, Skv, dh))


def attn_bwd_small(q, k, v, dctx, dq, dk, dv, B, H, Sq, Skv, dh, scale, causal=False, kv_len=None):
    """q / k / v / dq / dk / dv: (tensor view at head 0, token stride, row stride); dctx fp16 [B * Sq, H * dh]."""
    a = []
    for t, ts, bs in (q, k, v, (dctx, H * dh, Sq * H * dh), dq, dk, dv):
        a += [_p(t), _i64(ts), _i64(bs)]
    check(lib().myr_attn_bwd_small(*a, B, H, Sq, Skv, dh, _f32(scale), int(causal), _p(kv_len), _stream()), "myr_attn_bwd_small")


def index_rows(src, dst, idx, D, scatter=False):
    assert idx.dtype == torch.int32
    check(lib().myr_index_rows(_p(src), _dt(src), _i64(src.stride(0)), _p(dst), _dt(dst), _i64(dst.stride(0)), _p(idx), idx.numel(), D,
                               int(scatter), _stream()), "myr_index_rows")


def colsum(src, ld, group_stride, groups, rows, D, out, scale=1.0, accumulate=False):
    check(lib().myr_colsum(_p(src), _dt(src), _i64(ld), _i64(group_stride), groups, rows, D, _f32(scale), _p(out), int(accumulate),
                           _stream()), "myr_colsum")


def adaptor_bwd(x, dy, w1, w2, scratch, dw1, dw2, rows, D, rank, scale):
    check(lib().myr_adaptor_bwd(_p(x), _p(dy), _p(w1), _p(w2), _p(scratch), _p(dw1), _p(dw2), rows, D, rank, _f32(scale), _stream()),
          "myr_adaptor_bwd")


def adamw_step(params, grads, m, v, wd_mask, lr, beta1, beta2, eps, wd, step, inv_scale=1.0, found_inf=None):
    check(lib().myr_adamw_step(_p(params), _p(grads), _p(m), _p(v), _p(wd_mask), _i64(params.numel()), _f32(lr), _f32(beta1),
                               _f32(beta2), _f32(eps), _f32(wd), step, _f32(inv_scale), _p(found_inf), _stream()), "myr_adamw_step")


def memset_zero(t):
    check(lib().myr_memset_zero(_p(t), ctypes.c_size_t(t.numel() * t.element_size()), _stream()), "myr_memset_zero")


def conv3x3_relu(x, w, bias, out, B, H, W, Cin, Cout):
    check(lib().myr_conv3x3_relu(_p(x), _dt(x), _p(w), _p(bias), _p(out), B, H, W, Cin, Cout, _stream()), "myr_conv3x3_relu")


def pool_relu_bwd(y, dpool, dy, B, H, W, C):
    check(lib().myr_pool_relu_bwd(_p(y), _p(dpool), _dt(dpool), _p(dy), B, H, W, C, _stream()), "myr_pool_relu_bwd")


def conv3x3_wgrad(x, dy, dw, db, B, H, W, Cin, Cout, scale):
    check(lib().myr_conv3x3_wgrad(_p(x), _dt(x), _p(dy), _p(dw), _p(db), B, H, W, Cin, Cout, _f32(scale), _stream()), "myr_conv3x3_wgrad")


def conv3x3_dgrad(dy, w, din, B, H, W, Cin, Cout):
    check(lib().myr_conv3x3_dgrad(_p(dy), _p(w), _p(din), B, H, W, Cin, Cout, _stream()), "myr_conv3x3_dgrad")


def col2im(dcols, din, B, H, W, C, KH, KW, pad):
    check(lib().myr_col2im(_p(dcols), _p(din), B, H, W, C, KH, KW, pad, _stream()), "myr_col2im")


# ---- vision expert heads (csrc/expert.cu; adrefexpert_v2.py:245-301)
def expert_tap(x, out16, B, N, D, normalize=False):
    check(lib().myr_expert_tap(_p(x), _p(out16), B, N, D, int(normalize), _stream()), "myr_expert_tap")


def expert_logits(tokens, text, logits, B, P, C, scale=100.0):
    assert tokens.dtype == torch.float32 and text.dtype == torch.float32 and logits.dtype == torch.float32
    check(lib().myr_expert_logits(_p(tokens), _i64(tokens.stride(0)), _p(text), _p(logits), B, P, C, _f32(scale), _stream()),
          "myr_expert_logits")


def expert_maps(logits, maps, masks, L, B, G, OUT):
    check(lib().myr_expert_maps(_p(logits), _p(maps), _p(masks), L, B, G, OUT, _stream()), "myr_expert_maps")


def expert_rowmax(S, ld, acc, rows, R, weight, accumulate):
    check(lib().myr_expert_rowmax(_p(S), _i64(ld), _p(acc), rows, R, _f32(weight), int(accumulate), _stream()), "myr_expert_rowmax")


def expert_sim_maps(sim, maps, simmask, B, G, OUT):
    check(lib().myr_expert_sim_maps(_p(sim), _p(maps), _p(simmask), B, G, OUT, _stream()), "myr_expert_sim_maps")


# ---- LoRA branches in training (csrc/train.cu; peft semantics of myriad.py:171-178, rank 8)
def lora_fwd(x1, A, bq, bv, xa, qkv, col_q, col_v, s, p=0.0, seed=0, off_q=0, off_v=0):
    T, D = x1.shape
    assert x1.dtype == A.dtype == bq.dtype == bv.dtype == qkv.dtype == torch.float16 and xa.dtype == torch.float32
    assert A.shape == (16, D) and A.is_contiguous() and bq.is_contiguous() and bv.is_contiguous() and xa.is_contiguous()
    check(lib().myr_lora_fwd(_p(x1), _i64(x1.stride(0)), _p(A), _p(bq), _p(bv), _p(xa), _p(qkv), _i64(qkv.stride(0)), _i64(col_q), _i64(col_v),
                             T, D, 8, _f32(s), _f32(p), _u64(seed), _u64(off_q), _u64(off_v), _stream()), "myr_lora_fwd")


def lora_bwd(dqkv, col_q, col_v, xa, A, bq, bv, x1, dxa, dbq, dbv, dA, dx1, s, inv_scale, p=0.0, seed=0, off_q=0, off_v=0):
    T, D = x1.shape
    assert dqkv.dtype == torch.float16 and xa.dtype == dxa.dtype == dbq.dtype == dbv.dtype == dA.dtype == dx1.dtype == torch.float32
    assert dbq.is_contiguous() and dbv.is_contiguous() and dA.is_contiguous() and dxa.is_contiguous()
    check(lib().myr_lora_bwd(_p(dqkv), _i64(dqkv.stride(0)), _i64(col_q), _i64(col_v), _p(xa), _p(A), _p(bq), _p(bv), _p(x1), _i64(x1.stride(0)),
                             _p(dxa), _p(dbq), _p(dbv), _p(dA), _p(dx1), _i64(dx1.stride(0)), T, D, 8, _f32(s), _f32(inv_scale), _f32(p), _u64(seed),
                             _u64(off_q), _u64(off_v), _stream()), "myr_lora_bwd")
