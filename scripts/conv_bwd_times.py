"""Per-call CUDA-event times of the conv-trunk backward (networks.py:98-127 stacks) inside one training step. Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import synthetic as syn
from myriad_b200.training import MyriadTrainer
import myriad_b200.training as T_
dev = torch.device("cuda:0")
dims = syn.full_dims(lora_r=8)
dims.llama.layers = 2
dims.vit.depth = 2
tr = MyriadTrainer(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=4, max_seq=256)
image, maps = syn.make_inputs(4, seed=4321, device="cpu")
image, maps = image.to(dev), maps.to(dev)
ids_b, ids_a = syn.make_prompt_ids(dims.llama.vocab)
text = torch.randint(3, dims.llama.vocab, (4, 32)); tmask = torch.ones(4, 32, dtype=torch.long)
text[:, 16:] = dims.llama.eos; tmask[:, 16:] = 0
marks, active = [], [False]
def wrap(name):
    fn = getattr(T_.K, name)
    def w(*a, **k):
        if not active[0]:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        shape = ""
        if name == "gemm":
            shape = " T=%s F=%s K=%s" % (k.get("T", a[0].shape[0]), k.get("F", a[1].shape[0]), k.get("K", a[0].shape[1]))
        elif name.startswith("conv3x3") or name == "pool_relu_bwd":
            shape = " " + str([x for x in a if isinstance(x, int)])
        marks.append((name + shape, e0, e1)); return r
    setattr(T_.K, name, w)
for n in ("pool_relu_bwd", "conv3x3_wgrad", "conv3x3_dgrad", "gemm", "colsum", "col2im"):
    wrap(n)
orig = tr._conv_trunk_bwd
def timed_bwd(*a, **k):
    active[0] = True
    try:
        return orig(*a, **k)
    finally:
        active[0] = False
tr._conv_trunk_bwd = timed_bwd
for _ in range(2):
    tr.train_step(image, maps, 1, ids_b, ids_a, text, tmask)
torch.cuda.synchronize(); marks.clear()
tr.train_step(image, maps, 1, ids_b, ids_a, text, tmask)
torch.cuda.synchronize()
tot = 0.0
for name, e0, e1 in marks[:len(marks) // 2]:
    t = e0.elapsed_time(e1) * 1e3; tot += t
    print("%-60s %9.1f us" % (name, t))
print("one stack: %.1f us" % tot)
