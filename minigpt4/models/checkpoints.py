"""Frozen-weight loading from the same files the reference reads (nothing is downloaded; there is no network):
  EVA-ViT-g   eva_vit_g.pth            (eva_vit.py:429-436; LAVIS cache or ./pretrained_models/)
  Q-Former    BLIP-2 checkpoint        (blip2.py:91-110: checkpoint["model"] with Qformer.*, query_tokens, ln_vision.*)
  llama_proj  ./pretrained_models/pretrained_minigpt4_7b.pth   (myriad.py:210-217, hard-coded path in the reference)
  Vicuna-7B   HF directory `llama_model` (pytorch_model-*.bin or *.safetensors)          (myriad.py:193-196)
Returned mapping uses the reference state_dict key names (see myriad_b200/synthetic.py header)."""
import glob
import os

import torch


def _need(path, what):
    if not path or not os.path.exists(path):
        raise FileNotFoundError(
            "%s not found at %r. Place the reference's checkpoint files on disk, or set MYRIAD_SYNTHETIC_WEIGHTS=1 to run the "
            "hot path on seeded synthetic weights (benchmarks/tests)." % (what, path))
    return path


def load_reference_checkpoints(dims, llama_model, q_former_model):
    sd = {}
    vit_path = None
    for cand in (os.path.join("pretrained_models", "eva_vit_g.pth"),
                 os.path.join(torch.hub.get_dir(), "checkpoints", "eva_vit_g.pth")):
        if os.path.exists(cand):
            vit_path = cand
            break
    vit = torch.load(_need(vit_path, "EVA-ViT-g weights (eva_vit_g.pth)"), map_location="cpu")
    if isinstance(vit, dict) and "model" in vit and "cls_token" not in vit:
        vit = vit["model"]
    for k, v in vit.items():
        sd["visual_encoder." + k] = v
    qf = torch.load(_need(q_former_model, "BLIP-2 Q-Former checkpoint (q_former_model)"), map_location="cpu")["model"]
    for k, v in qf.items():
        if k.startswith(("Qformer.bert.", "ln_vision.")) or k == "query_tokens":
            sd[k] = v
    proj = torch.load(_need(os.path.join("pretrained_models", "pretrained_minigpt4_7b.pth"), "MiniGPT-4 llama_proj checkpoint"),
                      map_location="cpu")["model"]
    sd["llama_proj.weight"], sd["llama_proj.bias"] = proj["llama_proj.weight"], proj["llama_proj.bias"]
    _need(llama_model, "Vicuna-7B directory (llama_model)")
    shards = sorted(glob.glob(os.path.join(llama_model, "pytorch_model*.bin")))
    if shards:
        for s in shards:
            for k, v in torch.load(s, map_location="cpu").items():
                sd["llama_model." + k] = v
    else:
        from safetensors.torch import load_file
        st = sorted(glob.glob(os.path.join(llama_model, "*.safetensors")))
        if not st:
            raise FileNotFoundError("no pytorch_model*.bin / *.safetensors under %r" % llama_model)
        for s in st:
            for k, v in load_file(s).items():
                sd["llama_model." + k] = v
    return sd
