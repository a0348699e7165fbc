"""CPU: the C-ABI library builds, loads without a GPU, and exports every symbol include/myriad_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "myriad_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(myr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from myriad_b200._build import build_library
    lib = ctypes.CDLL(build_library())
    names = _declared()
    assert len(names) >= 15, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/myriad_b200.h but not exported: %s" % missing
    lib.myr_version.restype = ctypes.c_int
    assert lib.myr_version() >= 1


def test_header_is_plain_c():
    """include/myriad_b200.h is the boundary a non-C++ host binds to: it must compile as C99 on its own."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    r = subprocess.run([gcc, "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", os.path.join(ROOT, "include", "myriad_b200.h")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_ctypes_struct_layouts_match_the_library():
    """The ctypes mirrors in myriad_b200/kernels.py have the size the C compiler gave the structs of include/myriad_b200.h
    (a drifted binding would pass garbage to the kernels without any error)."""
    from myriad_b200 import kernels as K
    from myriad_b200._lib import lib
    lib().myr_abi_sizeof.restype = ctypes.c_size_t
    for which, cls in enumerate((K.GemmArgs, K.AttnArgs, K.NormArgs, K.RopeArgs, K.DecodeAttnArgs, K.MegaOp)):
        assert lib().myr_abi_sizeof(which) == ctypes.sizeof(cls), cls.__name__
    assert lib().myr_abi_sizeof(99) == 0


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected with a status code and a message before any CUDA call (errors never throw)."""
    from myriad_b200 import kernels as K
    from myriad_b200._lib import last_error, lib
    a = K.GemmArgs()
    assert lib().myr_gemm_f16(ctypes.byref(a), None) == -1
    assert "gemm" in last_error()
    assert lib().myr_gemm_f16(None, None) == -1
    b = K.AttnArgs()
    assert lib().myr_attention_fwd(ctypes.byref(b), None) == -1
    assert "attention" in last_error()
    # the fused training kernels state their limits instead of falling back silently
    import torch
    L = lib()
    assert L.myr_attn_bwd_small_supported(164, 164, 128) == 1 and L.myr_attn_bwd_small_supported(81, 81, 64) == 1
    assert L.myr_attn_bwd_small_supported(33, 256, 64) == 1
    assert L.myr_attn_bwd_small_supported(81, 257, 64) == 0 and L.myr_attn_bwd_small_supported(16, 16, 96) == 0
    t = torch.zeros(4, 96, dtype=torch.float16)
    ptr, i64 = ctypes.c_void_p(t.data_ptr()), ctypes.c_int64
    triple = [ptr, i64(96), i64(384)]
    rc = L.myr_attn_bwd_small(*(triple * 7), 1, 1, 4, 4, 96, ctypes.c_float(0.1), 0, None, None)
    assert rc == -1 and "dh must be 64 or 128" in last_error()
    x = torch.zeros(4, 100, dtype=torch.float16)
    px = ctypes.c_void_p(x.data_ptr())
    rc = L.myr_lora_fwd(px, i64(100), px, px, px, px, px, i64(100), i64(0), i64(0), 4, 100, 8, ctypes.c_float(1.0), ctypes.c_float(0.0),
                        ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0), None)
    assert rc == -1 and "lora_fwd" in last_error()

def test_no_cpu_fallback_in_product():
    """The product path never imports the oracle (the oracle is the checker, not a fallback)."""
    for pkg in ("myriad_b200", "minigpt4"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
