import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import synthetic as syn, kernels as K
from myriad_b200.engine import MyriadEngine
d = syn.tiny_dims(lora_r=0)
sd = syn.make_state_dict(d, 0)
eng = MyriadEngine(sd, d, device="cuda:0", max_batch=4, max_seq=256)
x = syn.synth("llama_in", (2, 12, d.llama.hidden), 0.5, 0, round_fp16=False)
mode = sys.argv[1] if len(sys.argv) > 1 else "eager"
toks = eng.greedy_decode(x[:, :7].contiguous().cuda(), 12, ((100,), (101, 102)), use_graph=(mode == "graph"))
torch.cuda.synchronize()
print(mode, "tokens", toks.tolist())
