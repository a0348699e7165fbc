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
# op-chain trace of the persistent decode kernel (globaltimer, ns)
if st.mega is not None:
    n_ops, G = st.mega.n_ops, 148
    tr = torch.zeros(3 * n_ops + n_ops * G * 2 + n_ops * 8, dtype=torch.int64, device=dev)
    st.mega.launch(trace=tr)
    torch.cuda.synchronize()
    tc = tr.cpu()
    t = tc[:3 * n_ops].reshape(-1, 3)
    per = tc[3 * n_ops:3 * n_ops + n_ops * G * 2].reshape(n_ops, G, 2)
    fine = tc[3 * n_ops + n_ops * G * 2:].reshape(n_ops, 8)
    t0 = int(t[0, 0])
    names = ["embed"] + ["qkv", "attn", "o", "gu", "down"] * dims.llama.layers + ["lm_head"]
    prev = t0
    agg = {}
    for i in range(n_ops):
        done = int(t[i, 0])
        if 6 <= i < 16:
            d = per[i, :, 0]; s_ = per[i, :, 1]
            d = d[d > 0]; s_ = s_[s_ > 0]
            line = "op %3d %-7s done +%7.2f (dur %6.2f)" % (i, names[i], (done - t0) / 1e3, (done - prev) / 1e3)
            if len(d):
                line += " | last drain per CTA: min +%7.2f med +%7.2f max +%7.2f" % ((int(d.min()) - t0) / 1e3, (int(d.median()) - t0) / 1e3, (int(d.max()) - t0) / 1e3)
            if len(s_):
                line += " | saw input: min +%7.2f max +%7.2f" % ((int(s_.min()) - t0) / 1e3, (int(s_.max()) - t0) / 1e3)
            print(line)
            if names[i] in ('o', 'down'):
                print('      closer: ' + ' '.join('%s+%.2f' % (nm, (int(fine[i, k]) - t0) / 1e3) for k, nm in enumerate(['sum', 'stored', 'fenced', 'elected', 'tail', 'pass1', 'normed', 'flag'])))
        agg.setdefault(names[i], []).append((done - prev) / 1e3)
        prev = done
    print({k: round(sum(v) / len(v), 2) for k, v in agg.items()}, "total %.1f us" % ((prev - t0) / 1e3))
