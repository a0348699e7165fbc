"""`arch: myriad` — the registry-registered drop-in for the reference's Myriad(Blip2Base) (minigpt4/models/myriad.py:62-517).

Same constructor keys (`from_config` reads the same yaml fields), same methods (`encode_img`, `encode_img_oneshot`,
`prepare_sample`, `prompt_wrap`, `embed_tokens`, `forward(samples) -> {"loss"}`, `generate(samples, **kw) ->
{"token_ids", "ve_anomaly_maps"}`), same externally-used attributes (`llama_tokenizer`, `llama_model.model.embed_tokens`,
`llama_model.generate`, `device`, `maybe_autocast`) and the same state_dict keys for the trainable modules
(`expert_adaptor.conv{1,2}.weight`, `VETokenizer.meta_net.{0,3,6,9,12,15}.{weight,bias}`, `VETokenizer.base_prompts`,
`VEInstructor.meta_net.*`, peft-style LoRA keys) so checkpoints written by the reference runner load unchanged.

All device arithmetic goes through myriad_b200.engine.MyriadEngine -> libmyriad_b200.so (sm_100a). There is no
PyTorch/CPU fallback: without the library or a CUDA device the compute methods raise.

Differences from the reference that a maintainer should know (each is a place where the reference cannot run
offline or is not on the hot path — SURVEY.md headline findings 2-5):
  * vision experts (adrefexpert_v2) are inputs: pass `samples["anomaly_maps"]` / `samples["oneshot_anomaly_maps"]`
    (fp32 [B,1,224,224]) or set `model.vision_expert = callable(images, scenes, querypath=None, testphase=False)`.
  * frozen weights come from the usual checkpoint files when they exist, else (MYRIAD_SYNTHETIC_WEIGHTS=1 or
    `weights=`) from a mapping with the reference key names; nothing is downloaded.
  * rejected with an error (no multi-backend / CPU paths): low_resource, bliva_like, vit_model != eva_clip_g,
    use_grad_checkpoint, unfrozen ViT / Q-Former / LLaMA base weights.
"""
import logging
import os
import random

import torch
import torch.nn as nn

from minigpt4.common.registry import registry
from minigpt4.models.blip2 import Blip2Base
from minigpt4.models.tokenizer import load_llama_tokenizer
from myriad_b200 import synthetic as syn


def _conv_stack_params(head_out, head_k):
    """Parameter container with the reference layer indices (networks.py:98-127 / 159-188); never called."""
    layers, c = [], syn.CONV_CHANNELS
    for j in range(5):
        layers += [nn.Conv2d(c[j], c[j + 1], kernel_size=3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(2)]
    layers.append(nn.Conv2d(1024, head_out, kernel_size=head_k, padding=0))
    return nn.Sequential(*layers)


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation runs in myriad_b200.engine (CUDA)")


class _EmbedTokens:
    def __init__(self, owner):
        self._o = owner

    def __call__(self, ids):
        return self._o._embed(ids)


class _LlamaShim:
    """`model.llama_model` as the callers use it: `.model.embed_tokens(ids)`, `.generate(inputs_embeds=, **kw)`,
    `.config.hidden_size` (conversation.py:157,226; evaluation_aqa_dataset.py:83)."""

    def __init__(self, owner):
        self._o = owner
        self.model = type("_M", (), {})()
        self.model.embed_tokens = _EmbedTokens(owner)
        self.model.model = self.model  # peft nesting: llama_model.model.model.embed_tokens (myriad.py:308-311)
        self.config = type("_C", (), {"hidden_size": owner.dims.llama.hidden, "vocab_size": owner.dims.llama.vocab})()

    def generate(self, inputs_embeds=None, **kw):
        return self._o._generate_from_embeds(inputs_embeds, **kw)


@registry.register_model("myriad")
class Myriad(Blip2Base):
    PRETRAINED_MODEL_CONFIG_DICT = {"pretrain_vicuna": "configs/models/minigpt4.yaml"}
    GENERATE_STAGE = 1  # Myriad.generate always encodes with stage 1 (myriad.py:435): VEInstructor queries + VETokenizer tokens

    def __init__(self, vit_model="eva_clip_g", q_former_model="", img_size=224, drop_path_rate=0, use_grad_checkpoint=False,
                 vit_precision="fp16", freeze_vit=True, freeze_qformer=True, freeze_llama=True, use_lora=False,
                 bliva_like=False, round_index=0, k_shot=0, use_ve=False, use_ref=False, do_random=False,
                 adaptor_type="none", num_query_token=32, llama_model="", prompt_path="", prompt_template="", max_txt_len=32,
                 end_sym="\n", low_resource=False, device_8bit=0, weights=None, dims=None, vision_expert=None):
        super().__init__()
        for flag, name in ((low_resource, "low_resource"), (bliva_like, "bliva_like"), (use_grad_checkpoint, "use_grad_checkpoint")):
            if flag:
                raise NotImplementedError("%s=True is not supported by the B200-native Myriad (no 8-bit/CPU/alt paths)" % name)
        if vit_model != "eva_clip_g" or not (freeze_vit and freeze_qformer and freeze_llama) or drop_path_rate:
            raise NotImplementedError("only vit_model=eva_clip_g with frozen ViT / Q-Former / LLaMA and drop_path_rate=0 is supported")
        if dims is None and os.environ.get("MYRIAD_SYNTHETIC_WEIGHTS", "0") == "1" and os.environ.get("MYRIAD_SYNTHETIC_DIMS", "full") == "mid":
            dims = syn.mid_dims(lora_r=8 if use_lora else 0)  # reduced depth / width for tests of the config-driven entry points
        self.dims = dims if dims is not None else syn.full_dims(lora_r=8 if use_lora else 0)
        if dims is None:
            self.dims.vit.img = img_size or 224
            self.dims.qf.num_query = num_query_token or 32
        self.tokenizer = self.init_tokenizer()
        self.low_resource, self.do_random = False, do_random
        self.use_ve, self.use_ref, self.round_index, self.k_shot = use_ve, use_ref, round_index, k_shot
        self.bliva_like = False
        self.lora_config = {"r": 8, "lora_alpha": 16, "lora_dropout": 0.05, "target_modules": ["q_proj", "v_proj"]} if self.dims.lora_r else None
        self.freeze_llama = True
        self.max_txt_len, self.end_sym = max_txt_len, end_sym
        self.vision_expert = vision_expert
        synthetic = weights is not None or os.environ.get("MYRIAD_SYNTHETIC_WEIGHTS", "0") == "1"
        self.llama_tokenizer = load_llama_tokenizer(llama_model, self.dims.llama.vocab, allow_synthetic=synthetic)
        self._frozen = self._resolve_weights(weights, llama_model, q_former_model)
        self._build_trainables()
        self.llama_model = _LlamaShim(self)
        self._engine = None
        if prompt_path:
            with open(prompt_path) as f:
                raw = f.read().splitlines()
            self.prompt_list = [prompt_template.format(p) for p in raw if "<ImageHere>" in p]
        else:
            self.prompt_list = []

    # ------------------------------------------------------------------------------------------ weights
    def _resolve_weights(self, weights, llama_model, q_former_model):
        if weights is not None:
            return weights
        if os.environ.get("MYRIAD_SYNTHETIC_WEIGHTS", "0") == "1":
            logging.warning("Myriad: using seeded synthetic weights (MYRIAD_SYNTHETIC_WEIGHTS=1)")
            dev = "cuda" if torch.cuda.is_available() else "cpu"
            return syn.LazyStateDict(self.dims, seed=int(os.environ.get("MYRIAD_SYNTHETIC_SEED", "0")), device=dev)
        from minigpt4.models.checkpoints import load_reference_checkpoints
        return load_reference_checkpoints(self.dims, llama_model, q_former_model)

    def _build_trainables(self):
        """fp32 master copies of the parameters the reference trains (runner_base.py:111-119 filters requires_grad)."""
        d, w = self.dims, self._frozen
        self.expert_adaptor = _Holder()
        self.expert_adaptor.conv1 = nn.Linear(d.vit.dim, d.adaptor_rank, bias=False)
        self.expert_adaptor.conv2 = nn.Linear(d.adaptor_rank, d.vit.dim, bias=False)
        if d.use_tokenizer:
            self.VETokenizer = _Holder()
            self.VETokenizer.meta_net = _conv_stack_params(4096, 5)
            self.VETokenizer.base_prompts = nn.Parameter(torch.zeros(9, 4096))
        if d.use_instructor:
            self.VEInstructor = _Holder()
            self.VEInstructor.meta_net = _conv_stack_params(768, 1)
        if d.lora_r:
            root = _Holder()
            cur = root
            for name in ("base_model", "model", "model"):
                nxt = _Holder()
                setattr(cur, name, nxt)
                cur = nxt
            cur.layers = nn.ModuleList()
            for _ in range(d.llama.layers):
                layer = _Holder()
                layer.self_attn = _Holder()
                for proj in ("q_proj", "v_proj"):
                    pm = _Holder()
                    pm.lora_A = nn.ModuleDict({"default": nn.Linear(d.llama.hidden, d.lora_r, bias=False)})
                    pm.lora_B = nn.ModuleDict({"default": nn.Linear(d.lora_r, d.llama.hidden, bias=False)})
                    setattr(layer.self_attn, proj, pm)
                cur.layers.append(layer)
            self._lora_params = root
        with torch.no_grad():
            for k, p in self.trainable_state().items():
                if k in w:
                    p.copy_(w[k].to(p.device, p.dtype))

    def trainable_state(self):
        """{reference state_dict key: nn.Parameter} for every parameter with requires_grad."""
        out = {}
        for k, p in self.named_parameters():
            out[k.replace("_lora_params.", "llama_model.")] = p
        return out

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return {k.replace("_lora_params.", "llama_model."): v for k, v in sd.items()}

    def load_state_dict(self, state_dict, strict=False):
        own = self.trainable_state()
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        with torch.no_grad():
            for k, p in own.items():
                if k in state_dict:
                    p.copy_(state_dict[k].to(p.device, p.dtype))
        self._engine = None  # re-prepare device weights on next use
        if strict and (missing or unexpected):
            raise RuntimeError("missing %s unexpected %s" % (missing, unexpected))
        return type("_IncompatibleKeys", (), {"missing_keys": missing, "unexpected_keys": unexpected})()

    class _Merged:
        def __init__(self, frozen, trainable):
            self.f, self.t = frozen, trainable

        def __getitem__(self, k):
            return self.t[k].detach() if k in self.t else self.f[k]

        def __contains__(self, k):
            return k in self.t or k in self.f

    def _trainable_version(self):
        """Changes whenever any trainable nn.Parameter was written in place (optimizer.step, load_state_dict, .copy_):
        torch bumps a tensor's version counter on every in-place update."""
        return tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    @property
    def engine(self):
        """The inference engine over the CURRENT trainable weights: its prepared device copies (NHWC fp16 conv filters, LoRA
        rows inside the fused qkv weight, ...) are refreshed in place whenever a trainable parameter changed since the last
        use, so generate() / encode_img() after optimizer steps — e.g. the per-epoch validation of the runner — see the
        trained weights."""
        if self._engine is None:
            if not torch.cuda.is_available():
                raise RuntimeError("Myriad (B200-native) needs a CUDA device; there is no CPU path")
            from myriad_b200.engine import MyriadEngine
            dev = torch.device("cuda", torch.cuda.current_device())
            self._engine = MyriadEngine(self._Merged(self._frozen, self.trainable_state()), self.dims, device=dev)
            self._engine_version = self._trainable_version()
        elif self._engine_version != self._trainable_version():
            self._engine.refresh_trainables(self._Merged(self._frozen, self.trainable_state()))
            self._engine_version = self._trainable_version()
        return self._engine

    # -------------------------------------------------------------------------------------- public API
    def vit_to_cpu(self):
        raise NotImplementedError("low_resource (ViT on CPU) is not supported")

    def encode_img(self, image, maps, stage):
        """myriad.py:241-272 -> (inputs_llama fp32 [B, n, 4096], atts_llama int64 [B, n])."""
        out = self.engine.encode_img(image.float().contiguous(), maps.float().contiguous(), stage)
        return out, torch.ones(out.shape[:-1], dtype=torch.long, device=out.device)

    def encode_img_oneshot(self, image, oneshotmaps, stage):
        """myriad.py:274-306: the same computation fed the one-shot expert maps."""
        return self.encode_img(image, oneshotmaps, stage)

    def _embed(self, ids):
        from myriad_b200 import kernels as K
        eng = self.engine
        flat = ids.reshape(-1).to(eng.dev)
        out = torch.empty(flat.numel(), self.dims.llama.hidden, device=eng.dev, dtype=torch.float32)
        K.embed(eng.llw.embed, flat.long(), out)
        return out.reshape(*ids.shape, -1)

    def embed_tokens(self, *args, **kwargs):
        return self.llama_model.model.embed_tokens(*args, **kwargs)

    def prepare_sample(self, samples, stage):
        """myriad.py:313-352 with the experts' outputs taken from the batch (or from `self.vision_expert`)."""
        image = samples["image"]
        if "aug_image" in samples and self.training:
            image = torch.cat([image, samples["aug_image"]])
        questions = samples.get({0: "question", 1: "question2", 2: "question3"}[stage], None)
        text_inputs = None
        if self.training:
            text_inputs = samples["text_input"] + samples["aug_text_input"] if "aug_text_input" in samples else samples["text_input"]
        if "anomaly_maps" in samples:
            maps = samples["anomaly_maps"]
            onemaps = samples.get("oneshot_anomaly_maps", maps)
            if self.training and "aug_image" in samples and maps.shape[0] * 2 == image.shape[0]:
                maps, onemaps = torch.cat([maps, samples.get("aug_anomaly_maps", maps)]), torch.cat([onemaps, samples.get("aug_oneshot_anomaly_maps", onemaps)])
        elif self.vision_expert is not None:
            with torch.no_grad():
                scenes, paths = samples["scene"], samples["img_path"]
                if self.training and "aug_image" in samples:
                    scenes, paths = scenes + scenes, paths + paths
                maps, _ = self.vision_expert(image, scenes)
                onemaps, _ = self.vision_expert(image, scenes, querypath=paths, testphase=not self.training)
        else:
            raise KeyError("samples must carry 'anomaly_maps' (and optionally 'oneshot_anomaly_maps'), or set model.vision_expert; "
                           "the ImageBind expert is outside the hot path (SURVEY.md §2 row 10)")
        return image, questions, text_inputs, maps, onemaps

    def _split_prompts(self, prompts, device):
        before, after = [], []
        for p in prompts:
            b, a = p.split("<ImageHere>")
            before.append(self.llama_tokenizer(b, return_tensors="pt", add_special_tokens=False).input_ids[0])
            after.append(self.llama_tokenizer(a, return_tensors="pt", add_special_tokens=False).input_ids[0])
        return torch.stack(before), torch.stack(after)  # equal lengths required, as myriad.py:371 does

    def prompt_wrap(self, img_embeds, atts_img, prompt):
        """myriad.py:354-375 on already-encoded image tokens (kept for API parity; generate/forward fuse this)."""
        if not prompt:
            return img_embeds, atts_img
        ib, ia = self._split_prompts(prompt, img_embeds.device)
        wrapped = torch.cat([self._embed(ib), img_embeds, self._embed(ia)], dim=1)
        return wrapped, atts_img[:, :1].expand(-1, wrapped.shape[1])

    def _generate_from_embeds(self, inputs_embeds, max_new_tokens=20, stopping_criteria=None, do_sample=False, top_p=1.0,
                              temperature=1.0, min_length=0, num_beams=1, **_):
        if num_beams != 1 or temperature != 1.0 or (do_sample and top_p > 0.01):
            raise NotImplementedError("only greedy search is implemented (the eval's do_sample=True, top_p=0.01 nucleus "
                                      "collapses to arg-max; Readme.md:42)")
        stops = []
        for crit in (stopping_criteria or []):
            for s in getattr(crit, "stops", []):
                stops.append(tuple(int(t) for t in torch.as_tensor(s).reshape(-1).tolist()))
        toks = self.engine.greedy_decode(inputs_embeds.float().contiguous().clone(), max_new_tokens, tuple(stops),
                                         min_new_tokens=1 if min_length >= 1 else 0)
        return toks.to(inputs_embeds.device)

    @torch.no_grad()
    def generate(self, samples, **generate_kwargs):
        """myriad.py:433-454: stage 1, no bos, greedy search; returns the NEW token ids and the expert maps."""
        stage = self.GENERATE_STAGE
        image, questions, _, maps, refs = self.prepare_sample(samples, stage)
        use = refs if self.k_shot > 0 else maps
        prompts = ["###Human: " + q + " ###Assistant: " for q in questions]
        ib, ia = self._split_prompts(prompts, image.device)
        dev = self.engine.dev
        emb = self.engine.build_inputs_embeds(image.to(dev).float().contiguous(), use.to(dev).float().contiguous(), stage, ib, ia)
        return {"token_ids": self.llama_model.generate(inputs_embeds=emb, **generate_kwargs), "ve_anomaly_maps": use}

    def forward(self, samples):
        """myriad.py:377-431: stage/task drawn with host RNG per call, targets masked with -100, clamp-CE loss."""
        from minigpt4.models.train_step import myriad_training_forward
        stage = random.choice([0, 1, 2])
        image, questions, text_inputs, maps, onemaps = self.prepare_sample(samples, stage)
        task = random.choice([0, 1])
        return {"loss": myriad_training_forward(self, image, maps if task == 0 else onemaps, stage, questions, text_inputs,
                                                double_prompts=self.training and "aug_image" in samples)}

    @classmethod
    def from_config(cls, cfg):
        g = cfg.get
        model = cls(vit_model=g("vit_model", "eva_clip_g"), q_former_model=g("q_former_model", ""), img_size=g("image_size"),
                    drop_path_rate=g("drop_path_rate", 0), use_grad_checkpoint=g("use_grad_checkpoint", False),
                    vit_precision=g("vit_precision", "fp16"), freeze_vit=g("freeze_vit", True), freeze_qformer=g("freeze_qformer", True),
                    freeze_llama=g("freeze_llama", True), use_lora=g("use_lora", False), bliva_like=g("bliva_like", False),
                    round_index=g("round_index", 0), k_shot=g("k_shot", 0), use_ve=g("use_ve", False), use_ref=g("use_ref", False),
                    num_query_token=g("num_query_token"), llama_model=g("llama_model"), prompt_path=g("prompt_path", ""),
                    prompt_template=g("prompt_template", ""), max_txt_len=g("max_txt_len", 32), end_sym=g("end_sym", "\n"),
                    low_resource=g("low_resource", False), device_8bit=g("device_8bit", 0))
        ckpt_path = g("ckpt", "")
        if ckpt_path:
            if os.path.isfile(ckpt_path):
                print("Load BLIP2-LLM Checkpoint: {}".format(ckpt_path))
                model.load_state_dict(torch.load(ckpt_path, map_location="cpu")["model"], strict=False)
            elif os.environ.get("MYRIAD_SYNTHETIC_WEIGHTS", "0") != "1":
                raise FileNotFoundError(ckpt_path)
        return model
