"""Phase breakdown of one bench step (myriad_generate_b4) with CUDA events. Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myriad_b200 import synthetic as syn, kernels as K
from myriad_b200.engine import MyriadEngine
dev = torch.device("cuda:0")
dims = syn.full_dims(lora_r=8)
eng = MyriadEngine(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=4, max_seq=256)
image, maps = syn.make_inputs(4, seed=1234, device="cpu")
image, maps = image.to(dev), maps.to(dev)
ids_b, ids_a = syn.make_prompt_ids(dims.llama.vocab)
stops = ((835,), (2277, 29937))
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for _ in range(3):
    eng.generate(image, maps, ids_b, ids_a, max_new_tokens=32, stop_seqs=stops)
torch.cuda.synchronize()
for rep in range(2):
    e0 = ev()
    x = eng.vit_forward(image)
    e1 = ev()
    emb = eng.build_inputs_embeds(image, maps, 1, ids_b, ids_a)
    e2 = ev()
    logits = eng.llama_prefill(emb.clone())
    e3 = ev()
    toks = eng.greedy_decode(emb, 32, stops)
    e4 = ev()
    torch.cuda.synchronize()
    print("vit %.2f ms | build_inputs_embeds (vit+qformer+experts+embeds) %.2f ms | prefill %.2f ms | greedy_decode (prefill + 32 steps) %.2f ms -> %.3f ms/decode step" % (
        e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4), (e3.elapsed_time(e4) - e2.elapsed_time(e3)) / 32), flush=True)
# decode step alone: replay the captured graph 32x without host syncs
st = list(eng._decode_graphs.values())[0]
torch.cuda.synchronize()
e0 = ev()
for _ in range(32):
    st.graph.replay()
e1 = ev()
torch.cuda.synchronize()
print("graph replay only: %.3f ms/step (%d kernel nodes)" % (e0.elapsed_time(e1) / 32, st.graph_nodes))
