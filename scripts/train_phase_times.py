"""Phase breakdown of one training step (myriad_stage2_lora_finetune_b4: stage 1, LoRA r = 8, batch 4, L = 164) with CUDA events
around the trainer's own phases. Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import synthetic as syn, kernels as K
from myriad_b200.training import MyriadTrainer
dev = torch.device("cuda:0")
dims = syn.full_dims(lora_r=8)
tr = MyriadTrainer(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=4, max_seq=256)
image, maps = syn.make_inputs(4, seed=4321, device="cpu")
image, maps = image.to(dev), maps.to(dev)
ids_b, ids_a = syn.make_prompt_ids(dims.llama.vocab)
g = torch.Generator().manual_seed(99)
text = torch.randint(3, dims.llama.vocab, (4, 32), generator=g)
tmask = torch.ones(4, 32, dtype=torch.long)
text[:, 16:] = dims.llama.eos
tmask[:, 16:] = 0
marks = []
def ev(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        ev(label + ":begin"); r = fn(*a, **k); ev(label + ":end"); return r
    setattr(obj, name, w)
for name, label in (("vit_forward", "vit fwd"), ("build_inputs_embeds", "encode fwd (vit + adaptor + q-former + experts + embeds)"),
                    ("_llama_train_fwd", "llama fwd (+ lm_head)"), ("_llama_train_bwd", "llama bwd"), ("_qformer_bwd", "q-former bwd"),
                    ("optimizer_step", "all-reduce + AdamW + refresh")):
    wrap(tr, name, label)
FINE = (("_ve_head_bwd", "expert head bwd (1x1 / 5x5 conv wgrad + trunk bwd)"), ("_conv_trunk_bwd", "  conv trunk bwd"),
        ("_conv_trunk_train", "  conv trunk fwd (inside encode fwd)"))
for name, label in FINE:
    wrap(tr, name, label)
import myriad_b200.training as T_
for name, label in (("clamp_ce_fwd", "clamp-CE fwd"), ("clamp_ce_bwd", "clamp-CE bwd"), ("adaptor_bwd", "adaptor bwd"), ("memset_zero", "zero flat grads")):
    wrap(T_.K, name, label)
for _ in range(3):
    tr.train_step(image, maps, 1, ids_b, ids_a, text, tmask)
torch.cuda.synchronize()
for rep in range(2):
    marks.clear()
    ev("step:begin")
    tr.train_step(image, maps, 1, ids_b, ids_a, text, tmask)
    ev("step:end")
    torch.cuda.synchronize()
    d = dict(marks)
    n0 = K.launch_count()
    print("step %.2f ms" % d["step:begin"].elapsed_time(d["step:end"]))
    for label in ("vit fwd", "encode fwd (vit + adaptor + q-former + experts + embeds)", "llama fwd (+ lm_head)", "llama bwd", "q-former bwd",
                  "all-reduce + AdamW + refresh"):
        print("  %-62s %.2f ms" % (label, d[label + ":begin"].elapsed_time(d[label + ":end"])))
    # phases that run several times per step: summed over their calls
    for label in [l for _, l in FINE] + ["clamp-CE fwd", "clamp-CE bwd", "adaptor bwd", "zero flat grads"]:
        b = [e for n, e in marks if n == label + ":begin"]
        e_ = [e for n, e in marks if n == label + ":end"]
        if b:
            print("  %-62s %.2f ms (%d calls)" % (label, sum(x.elapsed_time(y) for x, y in zip(b, e_)), len(b)))
