"""Model dimensions and seeded synthetic weights/inputs with the reference's state_dict key names.

There is no network in this environment, so neither the EVA-ViT-g / BLIP-2 / Vicuna checkpoints nor the
tokenizers exist; benchmarks, tests and the oracle all use weights drawn here. Key names follow
the reference modules so a real checkpoint loads through the same path:
  visual_encoder.*  eva_vit.py:226-340      ln_vision.*      blip2.py:119-125
  expert_adaptor.*  networks.py:71-93       VEInstructor.* / VETokenizer.*  networks.py:95-197
  Qformer.bert.*    Qformer.py:51-560       query_tokens     blip2.py:58-61
  llama_proj.*      myriad.py:207           llama_model.*    modeling_llama.py:401-640
  LoRA (peft naming) llama_model.base_model.model.model.layers.{i}.self_attn.{q,v}_proj.lora_{A,B}.default.weight
"""
import zlib
from dataclasses import dataclass, field

import torch


@dataclass
class VitDims:
    img: int = 224
    patch: int = 14
    dim: int = 1408
    depth: int = 39
    heads: int = 16
    mlp_hidden: int = 6144  # int(1408 * 4.3637), eva_vit.py:422
    ln_eps: float = 1e-6     # eva_vit.py:426

    @property
    def tokens(self):
        return (self.img // self.patch) ** 2 + 1

    @property
    def head_dim(self):
        return self.dim // self.heads


@dataclass
class QformerDims:
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    inter: int = 3072
    cross_freq: int = 2
    num_query: int = 32
    ln_eps: float = 1e-12


@dataclass
class LlamaDims:
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    inter: int = 11008
    vocab: int = 32000
    eps: float = 1e-6
    max_pos: int = 2048
    bos: int = 1
    eos: int = 2

    @property
    def head_dim(self):
        return self.hidden // self.heads


@dataclass
class MyriadDims:
    vit: VitDims = field(default_factory=VitDims)
    qf: QformerDims = field(default_factory=QformerDims)
    llama: LlamaDims = field(default_factory=LlamaDims)
    adaptor_rank: int = 4      # LoraAdaptorV2(dims=1408, input_dim=4), myriad.py:117
    lora_r: int = 0            # 0 = use_lora False (shipped yamls); 8 = peft config of myriad.py:171-178
    lora_alpha: float = 16.0
    use_instructor: bool = True   # needs qf.hidden == 768 (conv stack output is hard-wired, networks.py:128)
    use_tokenizer: bool = True    # needs llama.hidden == 4096 (networks.py:184)


def full_dims(lora_r=0):
    return MyriadDims(lora_r=lora_r)


def tiny_dims(lora_r=0):
    """Smallest config that keeps every structural feature (dh=88 ViT heads, dh=64 Q-Former, dh=128 LLaMA)."""
    return MyriadDims(vit=VitDims(img=56, dim=176, depth=2, heads=2, mlp_hidden=768),
                      qf=QformerDims(hidden=128, layers=2, heads=2, inter=256, num_query=8),
                      llama=LlamaDims(hidden=256, layers=2, heads=2, inter=512, vocab=320),
                      lora_r=lora_r, use_instructor=False, use_tokenizer=False)


def mid_dims(lora_r=0, vit_depth=2, qf_layers=2, llama_layers=1):
    """Real widths where the expert-token conv stacks force them (768 / 4096), reduced depth: the composite
    encode_img / forward / generate golden config."""
    return MyriadDims(vit=VitDims(dim=176, depth=vit_depth, heads=2, mlp_hidden=768),
                      qf=QformerDims(layers=qf_layers),
                      llama=LlamaDims(layers=llama_layers, inter=1024, vocab=1000),
                      lora_r=lora_r)


def _seed_for(name, seed):
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


def synth(name, shape, std, seed, device="cpu", mean=0.0, round_fp16=True):
    """Deterministic N(mean, std) tensor keyed by (name, seed); generated on `device`'s own generator.

    CPU and CUDA generators differ, so parity tests always generate on CPU and copy; benchmarks generate on
    the GPU (7 B parameters in well under a second). Values are rounded to fp16-representable numbers so the
    fp32 oracle and the fp16 device path hold identical weights."""
    g = torch.Generator(device=device)
    g.manual_seed(_seed_for(name, seed))
    t = torch.randn(tuple(shape), generator=g, device=device, dtype=torch.float32)
    if std != 1.0:
        t.mul_(std)
    if mean != 0.0:
        t.add_(mean)
    if round_fp16:
        t = t.half().float()
    return t


CONV_CHANNELS = [1, 4, 16, 64, 256, 1024]  # networks.py:98-122 / 159-182 with dim_in = 1
CONV_IDX = [0, 3, 6, 9, 12]


def state_dict_spec(d: MyriadDims):
    """-> list of (key, shape, std, mean). std follows the reference initialisers with small
    non-zero biases / LN offsets so every fused-epilogue path is exercised."""
    S = []
    v, q, l = d.vit, d.qf, d.llama
    D = v.dim
    p = "visual_encoder."
    S += [(p + "cls_token", (1, 1, D), 0.02, 0), (p + "pos_embed", (1, v.tokens, D), 0.02, 0),
          (p + "patch_embed.proj.weight", (D, 3, v.patch, v.patch), 0.02, 0),
          (p + "patch_embed.proj.bias", (D,), 0.02, 0)]
    for i in range(v.depth):
        b = p + "blocks.%d." % i
        resc = (2.0 * (i + 1)) ** -0.5  # fix_init_weight, eva_vit.py:300-306
        S += [(b + "norm1.weight", (D,), 0.1, 1.0), (b + "norm1.bias", (D,), 0.05, 0),
              (b + "attn.q_bias", (D,), 0.02, 0), (b + "attn.v_bias", (D,), 0.02, 0),
              (b + "attn.qkv.weight", (3 * D, D), 0.02, 0),
              (b + "attn.proj.weight", (D, D), 0.02 * resc, 0), (b + "attn.proj.bias", (D,), 0.02, 0),
              (b + "norm2.weight", (D,), 0.1, 1.0), (b + "norm2.bias", (D,), 0.05, 0),
              (b + "mlp.fc1.weight", (v.mlp_hidden, D), 0.02, 0), (b + "mlp.fc1.bias", (v.mlp_hidden,), 0.02, 0),
              (b + "mlp.fc2.weight", (D, v.mlp_hidden), 0.02 * resc, 0), (b + "mlp.fc2.bias", (D,), 0.02, 0)]
    S += [("ln_vision.weight", (D,), 0.1, 1.0), ("ln_vision.bias", (D,), 0.05, 0),
          ("expert_adaptor.conv1.weight", (d.adaptor_rank, D), 0.02, 0),
          ("expert_adaptor.conv2.weight", (D, d.adaptor_rank), 0.02, 0)]
    for mod, last_out, last_k, on in (("VEInstructor", 768, 1, d.use_instructor), ("VETokenizer", 4096, 5, d.use_tokenizer)):
        if not on:
            continue
        for j, idx in enumerate(CONV_IDX):
            cin, cout = CONV_CHANNELS[j], CONV_CHANNELS[j + 1]
            bound = (1.0 / (cin * 9)) ** 0.5  # kaiming_uniform(a=sqrt(5)) has std = bound / sqrt(3)
            S += [("%s.meta_net.%d.weight" % (mod, idx), (cout, cin, 3, 3), bound / 3 ** 0.5, 0),
                  ("%s.meta_net.%d.bias" % (mod, idx), (cout,), bound / 3 ** 0.5, 0)]
        bound = (1.0 / (1024 * last_k * last_k)) ** 0.5
        S += [("%s.meta_net.15.weight" % mod, (last_out, 1024, last_k, last_k), bound / 3 ** 0.5, 0),
              ("%s.meta_net.15.bias" % mod, (last_out,), bound / 3 ** 0.5, 0)]
    if d.use_tokenizer:
        S += [("VETokenizer.base_prompts", (9, 4096), 1.0, 0)]  # torch.randn, networks.py:189
    H = q.hidden
    S += [("query_tokens", (1, q.num_query, H), 0.02, 0)]
    b = "Qformer.bert."
    S += [(b + "embeddings.LayerNorm.weight", (H,), 0.1, 1.0), (b + "embeddings.LayerNorm.bias", (H,), 0.05, 0)]
    for i in range(q.layers):
        lb = b + "encoder.layer.%d." % i
        atts = [("attention", H)] + ([("crossattention", D)] if i % q.cross_freq == 0 else [])
        for an, kvw in atts:
            S += [(lb + an + ".self.query.weight", (H, H), 0.02, 0), (lb + an + ".self.query.bias", (H,), 0.02, 0),
                  (lb + an + ".self.key.weight", (H, kvw), 0.02, 0), (lb + an + ".self.key.bias", (H,), 0.02, 0),
                  (lb + an + ".self.value.weight", (H, kvw), 0.02, 0), (lb + an + ".self.value.bias", (H,), 0.02, 0),
                  (lb + an + ".output.dense.weight", (H, H), 0.02, 0), (lb + an + ".output.dense.bias", (H,), 0.02, 0),
                  (lb + an + ".output.LayerNorm.weight", (H,), 0.1, 1.0),
                  (lb + an + ".output.LayerNorm.bias", (H,), 0.05, 0)]
        S += [(lb + "intermediate_query.dense.weight", (q.inter, H), 0.02, 0),
              (lb + "intermediate_query.dense.bias", (q.inter,), 0.02, 0),
              (lb + "output_query.dense.weight", (H, q.inter), 0.02, 0),
              (lb + "output_query.dense.bias", (H,), 0.02, 0),
              (lb + "output_query.LayerNorm.weight", (H,), 0.1, 1.0),
              (lb + "output_query.LayerNorm.bias", (H,), 0.05, 0)]
    L = l.hidden
    S += [("llama_proj.weight", (L, H), 0.02, 0), ("llama_proj.bias", (L,), 0.02, 0)]
    b = "llama_model.model."
    # embedding rows ~ unit RMS like a trained model's normalised stream; lm_head sharpened (std 0.08) so greedy
    # arg-max margins sit well above fp16 noise (SURVEY.md §7 "token-exact greedy on random-init weights").
    S += [(b + "embed_tokens.weight", (l.vocab, L), 0.5, 0)]
    for i in range(l.layers):
        lb = b + "layers.%d." % i
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            S += [(lb + "self_attn.%s.weight" % n, (L, L), 0.02, 0)]
        S += [(lb + "mlp.gate_proj.weight", (l.inter, L), 0.02, 0), (lb + "mlp.up_proj.weight", (l.inter, L), 0.02, 0),
              (lb + "mlp.down_proj.weight", (L, l.inter), 0.02, 0),
              (lb + "input_layernorm.weight", (L,), 0.1, 1.0),
              (lb + "post_attention_layernorm.weight", (L,), 0.1, 1.0)]
        if d.lora_r > 0:
            pl = "llama_model.base_model.model.model.layers.%d.self_attn." % i
            for n in ("q_proj", "v_proj"):
                S += [(pl + n + ".lora_A.default.weight", (d.lora_r, L), 0.02, 0),
                      (pl + n + ".lora_B.default.weight", (L, d.lora_r), 0.01, 0)]  # B perturbed off zero
    S += [(b + "norm.weight", (L,), 0.1, 1.0), ("llama_model.lm_head.weight", (l.vocab, L), 0.08, 0)]
    return S


def make_state_dict(d: MyriadDims, seed=0, device="cpu", only_prefix=None):
    sd = {}
    for key, shape, std, mean in state_dict_spec(d):
        if only_prefix is not None and not key.startswith(only_prefix):
            continue
        sd[key] = synth(key, shape, std, seed, device=device, mean=mean)
    return sd


def make_inputs(batch, seed=1234, device="cpu", img=224):
    """image ~ N(0,1) (CLIP-normalised pixels), anomaly maps ~ U(0,1) [B,1,224,224]."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    image = torch.randn(batch, 3, img, img, generator=g, device=device)
    maps = torch.rand(batch, 1, 224, 224, generator=g, device=device)
    return image, maps


def make_prompt_ids(vocab, n_before=6, n_after=26, seed=7):
    """Synthetic token ids standing for '###Human: <Img>' and '</Img> ... ###Assistant: ' (32-token prompt)."""
    g = torch.Generator()
    g.manual_seed(seed)
    ids = torch.randint(3, vocab, (n_before + n_after,), generator=g)
    return ids[:n_before].clone(), ids[n_before:].clone()


class LazyStateDict:
    """Mapping view of make_state_dict that materialises one tensor per lookup (7 B parameters never sit in memory
    twice). Used by the benchmarks at full size; tests use the eager dict so the oracle sees the same tensors."""

    def __init__(self, d: MyriadDims, seed=0, device="cpu"):
        self._spec = {k: (shape, std, mean) for k, shape, std, mean in state_dict_spec(d)}
        self._seed, self._device = seed, device

    def __getitem__(self, key):
        shape, std, mean = self._spec[key]
        return synth(key, shape, std, self._seed, device=self._device, mean=mean)

    def __contains__(self, key):
        return key in self._spec

    def keys(self):
        return self._spec.keys()

    def num_parameters(self):
        n = 0
        for shape, _, _ in self._spec.values():
            m = 1
            for s in shape:
                m *= s
            n += m
        return n
