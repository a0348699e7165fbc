"""Timeline of a decode layer's kernels inside a CUDA graph (globaltimer stamps written by the GEMM kernel itself)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import kernels as K, synthetic as syn
from myriad_b200.engine import MyriadEngine
dev = torch.device("cuda:0")
dims = syn.full_dims(lora_r=8)
dims.llama.layers = 6
eng = MyriadEngine(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=4, max_seq=256)
torch.manual_seed(0)
x = (torch.randn(4, 131, 4096, device=dev) * 0.5)
eng.greedy_decode(x.clone(), 8, ((100000,),))
st = list(eng._decode_graphs.values())[0]
n_gemm = 4 * dims.llama.layers + 1
buf = torch.zeros(n_gemm * 148 * 6, dtype=torch.int64, device=dev)
K.lib().myr_gemm_set_trace(ctypes.c_void_p(buf.data_ptr()))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng._decode_step(st)
K.lib().myr_gemm_set_trace(ctypes.c_void_p(0))
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
t = buf.cpu().reshape(n_gemm, 148, 6)
names = ["qkv", "o", "gu", "down"] * dims.llama.layers + ["lm_head"]
t0 = int(t[4][:, 0][t[4][:, 0] > 0].min())
for i in range(4, 13):
    a = t[i]
    ok = a[:, 0] > 0
    f = lambda col, fn: (int(fn(a[ok][:, col])) - t0) / 1e3
    print("%-7s start min %7.2f max %7.2f | released min %7.2f max %7.2f | lastMMA max %7.2f | acc ready max %7.2f | done min %7.2f max %7.2f" % (
        names[i], f(0, torch.min), f(0, torch.max), f(1, torch.min), f(1, torch.max), f(2, torch.max), f(3, torch.max), f(4, torch.min), f(4, torch.max)))
