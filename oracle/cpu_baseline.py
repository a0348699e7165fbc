"""TEST/BENCH INFRASTRUCTURE ONLY — times the CPU oracle (a port of the reference's fp32 CPU path) on a bounded sample
of the benchmark workload. Imported only by bench.py's `cpu_baseline` leg and `--impl reference`.

Workload = Myriad.generate on one synthetic 224x224 image: ViT-g (39 blocks) -> adaptor+ln_vision -> VEInstructor ->
Q-Former (12 layers, 81 queries) -> llama_proj -> VETokenizer -> prompt wrap (32-token prompt, S = 131) -> Vicuna-7B
prefill -> `new_tokens` greedy steps. The encoder side is run in full. The LLaMA body is run on `llama_layers_sample`
of its 32 identical layers at full width and scaled by 32 / sample (27 GB of fp32 weights would take minutes just
to draw); lm_head and the per-step decode are measured on the same sample. All of this is stated in `sample`.
"""
import time

import torch

from myriad_b200 import synthetic as syn
from oracle import myriad_oracle as O


class CpuSample:
    def __init__(self, llama_layers_sample=4, lora_r=8, seed=0):
        self.full = syn.full_dims(lora_r=lora_r)
        self.k = llama_layers_sample
        self.d = syn.MyriadDims(llama=syn.LlamaDims(layers=llama_layers_sample), lora_r=lora_r)
        t0 = time.perf_counter()
        self.sd = syn.make_state_dict(self.d, seed)
        self.t_weights = time.perf_counter() - t0
        self.image, self.maps = syn.make_inputs(1, seed=1234)
        self.ids_b, self.ids_a = syn.make_prompt_ids(self.d.llama.vocab)

    @torch.no_grad()
    def run(self, new_tokens=32, decode_steps_sample=2):
        d, sd = self.d, self.sd
        scale = self.full.llama.layers / self.k
        t = {}
        t0 = time.perf_counter()
        img = O.encode_img(sd, self.image, self.maps, 1, d)
        emb = O.prompt_wrap(sd, img, self.ids_b, self.ids_a)
        t["encode"] = time.perf_counter() - t0
        B, S, _ = emb.shape
        mask = torch.ones(B, S, dtype=torch.long)
        pos = torch.arange(S)[None]
        t0 = time.perf_counter()
        h, past = O.llama_layers(sd, emb, O.causal_bias(mask, S, 0), pos, d)
        t["prefill_layers"] = (time.perf_counter() - t0) * scale
        t0 = time.perf_counter()
        logits = O.linear(h[:, -1:], sd["llama_model.lm_head.weight"])
        t["lm_head"] = time.perf_counter() - t0
        nxt = logits[:, -1].argmax(-1)
        dec = 0.0
        for _ in range(decode_steps_sample):
            mask = torch.cat([mask, torch.ones(B, 1, dtype=torch.long)], 1)
            t0 = time.perf_counter()
            x = O.embed_tokens(sd, nxt[:, None])
            h, past = O.llama_layers(sd, x, O.causal_bias(mask, 1, mask.shape[1] - 1), (mask.cumsum(-1) - 1)[:, -1:], d, past)
            dec += (time.perf_counter() - t0) * scale
            t0 = time.perf_counter()
            nxt = O.linear(h[:, -1:], sd["llama_model.lm_head.weight"])[:, -1].argmax(-1)
            dec += time.perf_counter() - t0
        t["decode_step"] = dec / decode_steps_sample
        total = t["encode"] + t["prefill_layers"] + t["lm_head"] + (new_tokens - 1) * t["decode_step"]
        t["total_per_image"] = total
        return 1.0 / total, t

    def describe(self, new_tokens):
        return ("oracle port (fp32 torch-CPU restatement of the reference path), 1 image: full ViT-g + Q-Former + expert "
                "tokens, LLaMA-7B body on %d of 32 layers at full width scaled x%d, lm_head full; prefill S=131, "
                "%d new tokens with the per-step cost measured on 2 steps" % (self.k, self.full.llama.layers // self.k, new_tokens))
