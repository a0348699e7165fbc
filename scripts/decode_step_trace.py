"""Timeline of a decode layer's kernels inside a CUDA graph (globaltimer stamps written by the kernels themselves).
Small-batch path: per layer qkv (fused RMSNorm) -> decode attention -> o_proj -> gate/up (fused RMSNorm, SwiGLU) -> down_proj."""
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
n_k = 5 * dims.llama.layers + 2
buf = torch.zeros(n_k * 148 * 6, dtype=torch.int64, device=dev)
K.lib().myr_gemm_set_trace(ctypes.c_void_p(buf.data_ptr()))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng._decode_step(st)
K.lib().myr_gemm_set_trace(ctypes.c_void_p(0))
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
t = buf.cpu().reshape(n_k, 148, 6)
names = ["qkv", "attn", "o", "gu", "down"] * dims.llama.layers + ["lm_head"]
legend = {"gemv": "start | released | first x chunk at the consumers | first weight stage landed (fused norm: first x chunk staged) | last stage consumed (fused norm: all x staged) | done",
          "attn": "start | released | rope done | scores | softmax | done"}
print("gemv stamps:", legend["gemv"]); print("attn stamps:", legend["attn"])
t0 = int(t[5][:, 0][t[5][:, 0] > 0].min())
for i in range(5, 16):
    a = t[i]
    ok = a[:, 0] > 0
    a = a[ok]
    cols = []
    for c in range(6):
        v = a[:, c]
        v = v[v > 0]
        cols.append("%7.2f..%7.2f" % ((int(v.min()) - t0) / 1e3, (int(v.max()) - t0) / 1e3) if len(v) else "      -")
    print("%-7s %s" % (names[i], " | ".join(cols)))
