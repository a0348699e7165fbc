"""ncu driver: ONE training step (myriad_stage2_lora_finetune_b4) between cudaProfilerStart / Stop, after three warm-up steps:
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv \
       python scripts/train_launch_list.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import synthetic as syn
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
for _ in range(3):
    tr.train_step(image, maps, 1, ids_b, ids_a, text, tmask)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.train_step(image, maps, 1, ids_b, ids_a, text, tmask)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
