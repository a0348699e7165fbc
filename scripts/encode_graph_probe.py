"""Is the encoder / prefill part of a benchmark step bound by the host's launch rate? Time vit_forward, encode_img and llama_prefill at the
benchmark batch as eager launches (what MyriadEngine.generate does) and replayed from a CUDA graph (device time only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from myriad_b200 import synthetic as syn
from myriad_b200.engine import MyriadEngine

dev = torch.device("cuda:0")
dims = syn.full_dims(lora_r=8)
eng = MyriadEngine(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=4, max_seq=256)
image, maps = syn.make_inputs(4, seed=1, device="cpu")
image, maps = image.to(dev), maps.to(dev)
ids_b, ids_a = syn.make_prompt_ids(dims.llama.vocab)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    return g.replay


emb = eng.build_inputs_embeds(image, maps, 1, ids_b, ids_a)
for name, fn in (("vit_forward", lambda: eng.vit_forward(image)), ("encode_img (stage 1)", lambda: eng.encode_img(image, maps, 1)),
                 ("llama_prefill (S = 131)", lambda: eng.llama_prefill(emb.clone()))):
    t_e = timed(fn)
    t_g = timed(graphed(fn))
    print("%-26s eager %7.3f ms   CUDA graph %7.3f ms" % (name, t_e, t_g), flush=True)
