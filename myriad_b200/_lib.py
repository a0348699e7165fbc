"""ctypes loader for libmyriad_b200.so — the C-ABI boundary (include/myriad_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute wrapper raises.
"""
import ctypes
import os

from ._build import LIB_PATH

_lib = None


class MyriadLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MyriadLibraryError(
                "libmyriad_b200.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.myr_version.restype = ctypes.c_int
        _lib.myr_last_error.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        _lib.myr_device_sm_count.restype = ctypes.c_int
        _lib.myr_launch_count.restype = ctypes.c_ulonglong
        _lib.myr_mega_plan_bytes.restype = ctypes.c_size_t
        _lib.myr_gemm_workspace_bytes.restype = ctypes.c_size_t
        _lib.myr_decode_attention_ws_bytes.restype = ctypes.c_int64
    return _lib


def last_error():
    buf = ctypes.create_string_buffer(1024)
    lib().myr_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc, what):
    if rc != 0:
        raise MyriadLibraryError("%s failed (status %d): %s" % (what, rc, last_error()))
