"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the vision expert: the ImageBind vision trunk and the two
map heads of adrefexpert.forward. Only tests/ may import this module; the product (myriad_b200/, minigpt4/) never does.

Follows (paths relative to /root/reference/minigpt4/models/):
  trunk   model/ImageBind/models/imagebind_model.py:142-167,447-470,486-504 (ImageBindModel.forward, vision modality only),
          multimodal_preprocessors.py:121-157,255-271,423-442 (PadIm2Video repeat -> Conv3d stem -> cls + pos_embed),
          transformer.py:94-96,104-170,236-283 (pre-LN, BlockWithMasking over nn.MultiheadAttention, tapped blocks)
  heads   adrefexpert_v2.py:15-28 (LinearLayer), :245-301 (forward: k-shot cosine branch, zero-shot text branch)

Parity pin: oracle/gen_golden_expert.py builds the UNMODIFIED reference ImageBindModel (module files loaded by path) with the
seeded weights of myriad_b200/expert.py and stores its tapped tokens in tests/golden/imagebind_tiny.npz; tests/test_expert_host.py
checks `vision_taps` against them. adrefexpert_v2.py itself cannot be imported (kornia, jsonlines, CUDA at import, checkpoints),
so the heads are restated from its lines with the same torch calls it makes (F.interpolate(mode='bilinear', align_corners=True),
torch.softmax, F.cosine_similarity) — "trunk pinned, heads restated".
"""
import math

import torch
import torch.nn.functional as F

from myriad_b200.expert import _PRE, _TRUNK, ExpertDims

from .myriad_oracle import gelu_erf, layer_norm, linear, softmax_lastdim


def vision_tokens(sd, image, d: ExpertDims):
    """RGBDTPreprocessor.forward (multimodal_preprocessors.py:255-271,273-290): image [B,3,224,224] -> [B, 257, D]."""
    B = image.shape[0]
    video = image.unsqueeze(2).repeat(1, 1, 2, 1, 1)                       # PadIm2Video(pad_type="repeat", ntimes=2), :423-442
    x = F.conv3d(video, sd[_PRE + "rgbt_stem.proj.1.weight"], stride=(2, d.patch, d.patch))  # imagebind_model.py:152-158, bias=False
    x = x.flatten(2).transpose(1, 2)                                        # PatchEmbedGeneric.forward :151-157
    cls = sd[_PRE + "cls_token"].expand(B, -1, -1)
    x = torch.cat([cls, x], dim=1)                                          # :261-265
    return x + sd[_PRE + "pos_embedding_helper.pos_embed"]                  # :266-268 (224 x 224: no interpolation)


def _mha(sd, b, x, heads):
    """nn.MultiheadAttention(bias=True, add_bias_kv=False) self-attention, need_weights=False (transformer.py:94-96)."""
    B, N, D = x.shape
    dh = D // heads
    qkv = linear(x, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, N, heads, dh).transpose(1, 2) * (dh ** -0.5)
    k = k.reshape(B, N, heads, dh).transpose(1, 2)
    v = v.reshape(B, N, heads, dh).transpose(1, 2)
    p = softmax_lastdim(q @ k.transpose(-1, -2))
    ctx = (p @ v).transpose(1, 2).reshape(B, N, D)
    return linear(ctx, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"])


def vision_taps(sd, image, d: ExpertDims):
    """SimpleTransformer.forward with out_layers (transformer.py:236-283): -> list of [B, 257, D], one per tapped block."""
    x = vision_tokens(sd, image, d)
    x = layer_norm(x, sd[_TRUNK + "pre_transformer_layer.0.weight"], sd[_TRUNK + "pre_transformer_layer.0.bias"], d.ln_eps)
    taps = []
    for i in range(max(d.out_layers) + 1):
        b = _TRUNK + "blocks.%d." % i
        # BlockWithMasking.forward, layer_scale_type None (transformer.py:157-159)
        x = x + _mha(sd, b, layer_norm(x, sd[b + "norm_1.weight"], sd[b + "norm_1.bias"], d.ln_eps), d.heads)
        h = layer_norm(x, sd[b + "norm_2.weight"], sd[b + "norm_2.bias"], d.ln_eps)
        x = x + linear(gelu_erf(linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"])), sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        if i in d.out_layers:
            taps.append(x)
    return taps


def zero_shot(sd, taps, text, d: ExpertDims):
    """adrefexpert.forward, querypath None (adrefexpert_v2.py:279-301). taps: list of [B, 257, D]; text [B, 2, dec_dim]."""
    maps, masks = [], []
    for l, t in enumerate(taps):
        tok = linear(t[:, 1:, :], sd["image_decoder.fc.%d.weight" % l], sd["image_decoder.fc.%d.bias" % l])  # :26-27
        tok = tok / tok.norm(dim=-1, keepdim=True)                                                           # :285
        am = 100.0 * tok @ text.transpose(-2, -1)                                                            # :286
        B, L_, C = am.shape
        H = int(math.sqrt(L_))
        grid = am.permute(0, 2, 1).reshape(B, 2, H, H)
        masks.append(torch.softmax(grid, dim=1)[:, 1:, :, :])                                                # :289-291
        up = F.interpolate(grid, size=d.out_size, mode="bilinear", align_corners=True)                       # :292-293
        maps.append(torch.softmax(up, dim=1)[:, 1:, :, :])                                                   # :294-295
    return torch.mean(torch.stack(maps), 0), torch.mean(torch.stack(masks), 0)                               # :298-300


def k_shot(taps_q, taps_ref, d: ExpertDims):
    """adrefexpert.forward with querypath (adrefexpert_v2.py:264-278). taps_q: list of [B, 257, D]; taps_ref: list of [B * k, 257, D]
    (the k references of sample b at rows [b k, (b + 1) k))."""
    B = taps_q[0].shape[0]
    G = d.grid
    sims = []
    for q, r in zip(taps_q, taps_ref):
        qt = q[:, 1:, :].reshape(B, G * G, 1, -1)                     # :266 (after the [:, 1:, :] of :239)
        rt = r[:, 1:, :].reshape(B, 1, -1, q.shape[-1])               # :267
        sims.append(F.cosine_similarity(qt, rt, dim=-1).max(dim=-1).values)  # :268-270
    sim = torch.mean(torch.stack(sims, dim=0), dim=0).reshape(B, 1, G, G)    # :272
    simmask = 1 - sim
    up = F.interpolate(sim, size=d.out_size, mode="bilinear", align_corners=True)
    return 1 - up, simmask
