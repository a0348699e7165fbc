"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference module files from /root/reference by path.

Used only by oracle/gen_golden.py (in the build container, where /root/reference exists) to pin
oracle/myriad_oracle.py against the reference's own code. Nothing here travels to the GPU box and nothing
in the product (myriad_b200/, minigpt4/) imports it.

The reference package does not import as shipped (minigpt4/models/__init__.py:18-27 imports missing modules)
and targets transformers 4.28 / timm; the shims below are the minimal symbol moves needed to execute
eva_vit.py, Qformer.py, networks.py and modeling_llama.py under transformers 5.5.
"""
import importlib.machinery
import importlib.util
import math
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MYRIAD_REFERENCE", "/root/reference")
_MODELS = os.path.join(REF_ROOT, "minigpt4", "models")


def available():
    return os.path.isdir(_MODELS)


def _load(name, fname):
    spec = importlib.util.spec_from_file_location(name, os.path.join(_MODELS, fname))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _install_timm_stub():
    if "timm.models.layers" in sys.modules:
        return
    import transformers  # noqa: F401  (must resolve its optional-dependency probes before the stub exists)
    from transformers import BertConfig, LlamaConfig  # noqa: F401
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")

    def drop_path(x, drop_prob=0.0, training=False):
        if not drop_prob or not training:
            return x
        raise NotImplementedError("drop_path > 0 is not on the hot path (configs use 0)")

    layers.drop_path = drop_path
    layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    layers.trunc_normal_ = _trunc_normal_
    registry.register_model = lambda fn: fn
    timm.models = models
    models.layers = layers
    models.registry = registry
    for m in (timm, models, layers, registry):
        m.__spec__ = importlib.machinery.ModuleSpec(m.__name__, None)
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                        "timm.models.registry": registry})


def _install_minigpt4_stub():
    if "minigpt4.common.dist_utils" in sys.modules:
        return
    pk = types.ModuleType("minigpt4")
    common = types.ModuleType("minigpt4.common")
    du = types.ModuleType("minigpt4.common.dist_utils")

    def download_cached_file(*a, **k):
        raise RuntimeError("no network in the oracle")

    du.download_cached_file = download_cached_file
    pk.common = common
    common.dist_utils = du
    sys.modules.update({"minigpt4": pk, "minigpt4.common": common, "minigpt4.common.dist_utils": du})


def _install_transformers_shims():
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for n in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = getattr(pu, "find_pruneable_heads_and_indices", lambda *a, **k: None)


_cache = {}


def eva_vit():
    if "eva_vit" not in _cache:
        _install_timm_stub()
        saved = {k: sys.modules.get(k) for k in ("minigpt4", "minigpt4.common", "minigpt4.common.dist_utils")}
        _install_minigpt4_stub()
        _cache["eva_vit"] = _load("_ref_eva_vit", "eva_vit.py")
        for k, v in saved.items():  # do not shadow the repo's own minigpt4 package
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return _cache["eva_vit"]


def networks():
    if "networks" not in _cache:
        _cache["networks"] = _load("_ref_networks", "networks.py")
    return _cache["networks"]


def modeling_llama():
    if "llama" not in _cache:
        _cache["llama"] = _load("_ref_modeling_llama", "modeling_llama.py")
    return _cache["llama"]


def qformer():
    if "qformer" not in _cache:
        _install_transformers_shims()
        mod = _load("_ref_qformer", "Qformer.py")
        # transformers 5.x: init_weights() requires tied-weight bookkeeping the 4.x-era fork never set up, and
        # get_head_mask was removed. The reference passes head_mask=None, i.e. [None]*num_layers.
        mod.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
        mod.BertModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
        _cache["qformer"] = mod
    return _cache["qformer"]


def build_vit(img_size, patch_size, embed_dim, depth, num_heads, mlp_ratio):
    """VisionTransformer exactly as create_eva_vit_g constructs it (eva_vit.py:416-428), minus the download."""
    from functools import partial
    m = eva_vit()
    return m.VisionTransformer(img_size=img_size, patch_size=patch_size, use_mean_pooling=False, embed_dim=embed_dim,
                               depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=True,
                               drop_path_rate=0.0, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                               use_checkpoint=False).eval()


def build_qformer(hidden, layers, heads, inter, encoder_width, num_query, cross_freq=2):
    """Blip2Base.init_Qformer (blip2.py:49-63) + the trimming of myriad.py:151-156."""
    from transformers import BertConfig
    m = qformer()
    cfg = BertConfig(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                     intermediate_size=inter)
    cfg.encoder_width = encoder_width
    cfg.add_cross_attention = True
    cfg.cross_attention_freq = cross_freq
    cfg.query_length = num_query
    q = m.BertLMHeadModel(config=cfg)
    q.cls = None
    q.bert.embeddings.word_embeddings = None
    q.bert.embeddings.position_embeddings = None
    for layer in q.bert.encoder.layer:
        layer.output = None
        layer.intermediate = None
    return q.eval()


def build_llama(hidden, layers, heads, inter, vocab, max_pos=2048, eps=1e-6):
    from transformers import LlamaConfig
    m = modeling_llama()
    cfg = LlamaConfig(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                      num_key_value_heads=heads, intermediate_size=inter, vocab_size=vocab,
                      max_position_embeddings=max_pos, rms_norm_eps=eps, hidden_act="silu",
                      pad_token_id=0, bos_token_id=1, eos_token_id=2, tie_word_embeddings=False)
    cfg.use_cache = True
    return m.LlamaForCausalLM(cfg).eval()
