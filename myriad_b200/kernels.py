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
    ]


_ws_cache = {}


def workspace(nbytes, device):
    """Grow-only per-device scratch buffer (split-K partials etc.). Owned by PyTorch, lent to the library."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 64 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def gemm(x, w, bias=None, act=ACT_NONE, res=None, out=None, out_dtype=torch.float16, scale_cols=0, scale=1.0,
         round_acc=False, x_mn_major=False, w_mn_major=False, T=None, F=None, K=None, bn_hint=0, ksplit_hint=0):
    """out[t, f] = epilogue(sum_k x[t, k] * w[f, k]);  x: [T, K] fp16 (row stride may exceed K), w: [F, K] fp16."""
    assert x.dtype == torch.float16 and w.dtype == torch.float16
    assert x.stride(-1) == 1 and w.stride(-1) == 1
    if T is None:
        T = x.shape[0] if not x_mn_major else x.shape[1]
    if K is None:
        K = x.shape[1] if not x_mn_major else x.shape[0]
    if F is None:
        F = w.shape[0] if not w_mn_major else w.shape[1]
    if out is None:
        out = torch.empty((T, F), dtype=out_dtype, device=x.device)
    ws = workspace(64 << 20, x.device)
    a = GemmArgs()
    a.x, a.ldx = x.data_ptr(), x.stride(0)
    a.w, a.ldw = w.data_ptr(), w.stride(0)
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
    a.out, a.out_dtype, a.ldo = out.data_ptr(), _dt(out), out.stride(0)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.bn_hint, a.ksplit_hint = bn_hint, ksplit_hint
    check(lib().myr_gemm_f16(ctypes.byref(a), _stream()), "myr_gemm_f16")
    return out
