"""myriad_b200 — B200-native (sm_100a) implementation of the Myriad data-parallel hot path.

csrc/      hand-written CUDA kernels + the C-ABI (include/myriad_b200.h)
kernels.py ctypes bindings (PyTorch lends device memory and streams)
"""
__version__ = "0.1.0"
