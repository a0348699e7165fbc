"""TEST/BENCH INFRASTRUCTURE ONLY — times the oracle (a port of the reference's own PyTorch path) on a bounded sample of the
benchmark workload. Imported only by bench.py's `cpu_baseline` / `gpu_eager_baseline` legs and `--impl reference`.

Workload = Myriad.generate on a batch of 4 synthetic 224x224 images (BASELINE.json configs[2]): ViT-g (39 blocks) ->
adaptor + ln_vision -> VEInstructor -> Q-Former (12 layers, 81 queries) -> llama_proj -> VETokenizer -> prompt wrap
(32-token prompt, S = 131) -> Vicuna-7B prefill -> `new_tokens` greedy steps.

CpuSample (fp32, host cores = the reference's CPU path, blip2.py:39-47): the encoder side and the prefill / decode run at
the REAL batch of 4; the LLaMA body runs `llama_layers_sample` of its 32 identical layers at full width and is scaled by
32 / sample (27 GB of fp32 weights would take minutes just to draw), lm_head is run in full, the per-step decode cost is
measured on `decode_steps_sample` steps and multiplied out. What was run and every factor is returned in `extrapolation`.

EagerGpuSample (fp16 weights / activations on the same B200 = the reference's CUDA path: PyTorch eager, cuBLAS / ATen kernels,
KV cache by torch.cat): the whole model, the whole workload, nothing extrapolated.
"""
import time

import torch

from myriad_b200 import synthetic as syn
from oracle import myriad_oracle as O

BATCH = 4


class CpuSample:
    def __init__(self, llama_layers_sample=2, lora_r=8, seed=0, batch=BATCH):
        self.full = syn.full_dims(lora_r=lora_r)
        self.k, self.batch = llama_layers_sample, batch
        self.d = syn.MyriadDims(llama=syn.LlamaDims(layers=llama_layers_sample), lora_r=lora_r)
        t0 = time.perf_counter()
        self.sd = syn.make_state_dict(self.d, seed)
        self.t_weights = time.perf_counter() - t0
        self.image, self.maps = syn.make_inputs(batch, seed=1234)
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
        pos = torch.arange(S)[None].expand(B, -1)
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
        t["total_per_batch"] = total
        return self.batch / total, t

    def extrapolation(self, new_tokens, decode_steps_sample=2):
        return {"extrapolated": True, "batch_run": self.batch, "llama_layers_run": self.k, "llama_layers_total": self.full.llama.layers,
                "llama_layer_factor": self.full.llama.layers / self.k, "decode_steps_run": decode_steps_sample,
                "decode_steps_total": new_tokens - 1, "run_in_full": ["ViT-g 39 blocks", "adaptor+ln_vision", "VEInstructor", "Q-Former 12 layers",
                                                                       "llama_proj", "VETokenizer", "prompt wrap", "lm_head"]}

    def describe(self, new_tokens):
        return ("oracle port (fp32 torch-CPU restatement of the reference path) at the real batch of %d images: full ViT-g + Q-Former + "
                "expert tokens + lm_head; LLaMA-7B body on %d of 32 layers at full width scaled x%d; prefill S=131; %d new tokens with "
                "the per-step cost measured on 2 steps" % (self.batch, self.k, self.full.llama.layers // self.k, new_tokens))


class EagerGpuSample:
    """The path this repository replaces (SURVEY.md headline 1): PyTorch eager in fp16 on the same GPU — here the oracle's
    torch ops on .half().cuda() tensors (cuBLAS GEMMs, ATen softmax / LayerNorm, torch.cat KV cache, HF-style greedy loop)."""

    def __init__(self, device, lora_r=8, seed=0, batch=BATCH):
        self.d = syn.full_dims(lora_r=lora_r)
        self.dev, self.batch = device, batch
        lazy = syn.LazyStateDict(self.d, seed=seed, device=device)
        self.sd = {k: lazy[k].half() for k in lazy.keys()}
        image, maps = syn.make_inputs(batch, seed=1234)
        self.image_h, self.maps_h = image.pin_memory(), maps.pin_memory()
        ids_b, ids_a = syn.make_prompt_ids(self.d.llama.vocab)
        self.ids_b, self.ids_a = ids_b.to(device), ids_a.to(device)

    @torch.no_grad()
    def step(self, new_tokens=32, stops=((835,), (2277, 29937))):
        image = self.image_h.to(self.dev, non_blocking=True).half()
        maps = self.maps_h.to(self.dev, non_blocking=True).half()
        img = O.encode_img(self.sd, image, maps, 1, self.d)
        emb = O.prompt_wrap(self.sd, img, self.ids_b, self.ids_a)
        return O.greedy_generate(self.sd, emb, self.d, new_tokens, stops).cpu()

    def time(self, steps=3, warmup=1, new_tokens=32):
        for _ in range(warmup):
            self.step(new_tokens)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            toks = self.step(new_tokens)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return self.batch / (ms / 1e3), ms, int(toks.shape[1])
