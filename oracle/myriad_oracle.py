"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the Myriad hot path (the parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (myriad_b200/, minigpt4/) never does and fails loudly without its CUDA library.

Every function restates one piece of the reference in plain torch-CPU fp32 arithmetic on a flat state_dict
(reference key names) and cites the reference file:line it follows (paths relative to /root/reference/).
fp32 with no autocast IS the reference's CPU behaviour (blip2.py:39-47 disables autocast on CPU;
models/__init__.py:76-77 calls .float()).

Parity pin: oracle/gen_golden.py runs the UNMODIFIED reference modules (eva_vit.py, Qformer.py, networks.py,
modeling_llama.py, loaded by path) on seeded weights/inputs and stores their outputs under tests/golden/;
tests/test_oracle_golden.py checks this restatement against those files. The reference itself ships no tests
or golden vectors, and myriad.py / peft / HF generate cannot be imported here, so the glue
(encode_img, prompt_wrap, forward, generate), LoRA and the greedy loop are pinned only against compositions of
the reference's sub-modules written in gen_golden.py — "parity pinned at module level, glue restated".
"""
import math

import torch
import torch.nn.functional as F

from myriad_b200.synthetic import CONV_IDX, MyriadDims


# ----------------------------------------------------------------------------------------------------
# shared primitives
# ----------------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    """nn.LayerNorm in fp32 (blip2.py:119-125 casts to fp32 first; eva_vit.py:426 eps=1e-6; BERT eps=1e-12)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x):
    """nn.GELU() default / ACT2FN['gelu'] — exact erf form (eva_vit.py:45, Qformer.py:353-356)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def softmax_lastdim(s):
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    return e / e.sum(-1, keepdim=True)


# ----------------------------------------------------------------------------------------------------
# EVA ViT  (eva_vit.py)
# ----------------------------------------------------------------------------------------------------
def vit_patch_embed(sd, image, d, p="visual_encoder."):
    """PatchEmbed.forward eva_vit.py:198-204: stride-14 14x14 conv == per-patch GEMM over (c, kh, kw)."""
    B = image.shape[0]
    P = d.patch
    g = d.img // P
    patches = image.reshape(B, 3, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * P * P)
    w = sd[p + "patch_embed.proj.weight"].reshape(d.dim, 3 * P * P)
    return linear(patches, w, sd[p + "patch_embed.proj.bias"])


def vit_attention(sd, x, d, b):
    """Attention.forward eva_vit.py:118-148: fused qkv with bias cat(q_bias, 0, v_bias); q scaled BEFORE QK^T."""
    B, N, C = x.shape
    H, dh = d.heads, d.head_dim
    bias = torch.cat([sd[b + "attn.q_bias"], torch.zeros_like(sd[b + "attn.v_bias"]), sd[b + "attn.v_bias"]])
    qkv = linear(x, sd[b + "attn.qkv.weight"], bias).reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * dh ** -0.5, qkv[1], qkv[2]
    attn = softmax_lastdim(q @ k.transpose(-2, -1))
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return linear(y, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])


def vit_block(sd, x, d, b):
    """Block.forward eva_vit.py:173-180, gamma_1 is None branch (init_values unset at :416-428)."""
    h = layer_norm(x, sd[b + "norm1.weight"], sd[b + "norm1.bias"], d.ln_eps)
    x = x + vit_attention(sd, h, d, b)
    h = layer_norm(x, sd[b + "norm2.weight"], sd[b + "norm2.bias"], d.ln_eps)
    h = gelu_erf(linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))  # Mlp.forward :54-61
    return x + linear(h, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])


def vit_forward(sd, image, d, p="visual_encoder."):
    """VisionTransformer.forward_features eva_vit.py:324-340 (no final norm, no rel-pos bias)."""
    x = vit_patch_embed(sd, image, d, p)
    x = torch.cat([sd[p + "cls_token"].expand(x.shape[0], -1, -1), x], 1) + sd[p + "pos_embed"]
    for i in range(d.depth):
        x = vit_block(sd, x, d, p + "blocks.%d." % i)
    return x


# ----------------------------------------------------------------------------------------------------
# expert-prior modules (networks.py)
# ----------------------------------------------------------------------------------------------------
def lora_adaptor(sd, x, p="expert_adaptor."):
    """LoraAdaptorV2.forward networks.py:81-93 (out_dim == dims branch): x + W2 (W1 x)."""
    return x + linear(linear(x, sd[p + "conv1.weight"]), sd[p + "conv2.weight"])


class _RoundFp16(torch.autograd.Function):
    """fp16 storage of an activation (forward) and of its gradient (backward), everything else fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g.half().float()


# When True, conv_stack stores each post-ReLU map in fp16 — the device path's (and the reference CUDA autocast path's)
# storage precision. The only place on the hot path where that matters qualitatively: max-pool arg-max / ReLU gates are
# discrete, so fp16 storage flips ~1e-3 of them and the weight gradients (random-sign sums over positions) move by 3-10 %
# against the pure-fp32 restatement. tests/test_training_gpu.py compares the device conv gradients with this emulation
# (tight) and with the pure fp32 oracle (loose); every other comparison uses the pure fp32 oracle.
CONV_FP16_ACTS = False


def conv_stack(sd, p, maps):
    """The shared 5 x [conv3x3 pad1, ReLU, maxpool2] trunk, networks.py:98-122 / 159-182. -> [B,1024,7,7]"""
    x = maps
    for idx in CONV_IDX:
        x = F.relu(F.conv2d(x, sd["%smeta_net.%d.weight" % (p, idx)], sd["%smeta_net.%d.bias" % (p, idx)], padding=1))
        if CONV_FP16_ACTS:
            x = _RoundFp16.apply(x)
        x = F.max_pool2d(x, 2)
    return x


def ve_instructor(sd, maps, p="VEInstructor."):
    """VEInstructorV2.forward networks.py:149-153 (version 0: conv1x1 -> 768, 49 tokens)."""
    x = F.conv2d(conv_stack(sd, p, maps), sd[p + "meta_net.15.weight"], sd[p + "meta_net.15.bias"])
    return x.reshape(maps.shape[0], 768, 49).transpose(-2, -1)


def ve_tokenizer(sd, maps, p="VETokenizer."):
    """VETokenizer.forward networks.py:191-197: conv5x5 no pad -> [B,4096,3,3]; 9 learned prompts prepended."""
    x = F.conv2d(conv_stack(sd, p, maps), sd[p + "meta_net.15.weight"], sd[p + "meta_net.15.bias"])
    x = x.reshape(maps.shape[0], 4096, 9).transpose(-2, -1)
    return torch.cat([sd[p + "base_prompts"].expand(maps.shape[0], -1, -1), x], 1)


# ----------------------------------------------------------------------------------------------------
# Q-Former (Qformer.py) — query-only path, all-ones masks (additive mask == 0, Qformer.py:713-802)
# ----------------------------------------------------------------------------------------------------
def bert_attention(sd, p, hidden, kv_src, heads, eps):
    """BertSelfAttention.forward :169-275 + BertSelfOutput :278-289. Scores divided by sqrt(dh) AFTER QK^T."""
    B, Q, H = hidden.shape
    dh = H // heads

    def split(t):
        return t.reshape(B, -1, heads, dh).permute(0, 2, 1, 3)

    q = split(linear(hidden, sd[p + "self.query.weight"], sd[p + "self.query.bias"]))
    k = split(linear(kv_src, sd[p + "self.key.weight"], sd[p + "self.key.bias"]))
    v = split(linear(kv_src, sd[p + "self.value.weight"], sd[p + "self.value.bias"]))
    probs = softmax_lastdim((q @ k.transpose(-1, -2)) / math.sqrt(dh))
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, Q, H)
    out = linear(ctx, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return layer_norm(out + hidden, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)


def qformer_forward(sd, query_embeds, enc, d, p="Qformer.bert."):
    """BertModel.forward :804-965 with query_embeds only -> BertEmbeddings :103-108 (LayerNorm, dropout=id in
    eval) -> BertEncoder -> BertLayer.forward :402-474 (self-attn; cross-attn if layer % cross_freq == 0;
    feed_forward_chunk_query :481-484)."""
    h = layer_norm(query_embeds, sd[p + "embeddings.LayerNorm.weight"], sd[p + "embeddings.LayerNorm.bias"], d.ln_eps)
    for i in range(d.layers):
        lp = p + "encoder.layer.%d." % i
        h = bert_attention(sd, lp + "attention.", h, h, d.heads, d.ln_eps)
        if i % d.cross_freq == 0:
            h = bert_attention(sd, lp + "crossattention.", h, enc, d.heads, d.ln_eps)
        t = gelu_erf(linear(h, sd[lp + "intermediate_query.dense.weight"], sd[lp + "intermediate_query.dense.bias"]))
        t = linear(t, sd[lp + "output_query.dense.weight"], sd[lp + "output_query.dense.bias"])
        h = layer_norm(t + h, sd[lp + "output_query.LayerNorm.weight"], sd[lp + "output_query.LayerNorm.bias"], d.ln_eps)
    return h


# ----------------------------------------------------------------------------------------------------
# encode_img glue (myriad.py:241-272; encode_img_oneshot :274-306 is the same function fed the one-shot maps)
# ----------------------------------------------------------------------------------------------------
def encode_img(sd, image, maps, stage, d: MyriadDims):
    x = vit_forward(sd, image, d.vit)
    x = layer_norm(lora_adaptor(sd, x), sd["ln_vision.weight"], sd["ln_vision.bias"], 1e-5)
    q = sd["query_tokens"].expand(image.shape[0], -1, -1)
    if stage in (1, 2):
        q = torch.cat([q, ve_instructor(sd, maps)], 1)
    h = qformer_forward(sd, q, x, d.qf)
    t = linear(h, sd["llama_proj.weight"], sd["llama_proj.bias"])
    if stage in (0, 1):
        t = torch.cat([t, ve_tokenizer(sd, maps)], 1)
    return t


# ----------------------------------------------------------------------------------------------------
# LLaMA (modeling_llama.py)
# ----------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    """LlamaRMSNorm.forward :66-74."""
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def rope_tables(dh, max_pos, base=10000.0, device="cpu", dtype=torch.float32):
    """LlamaRotaryEmbedding.__init__ :78-91 (freqs duplicated, not interleaved); cached tables are cast to x.dtype (:103-106)."""
    inv = 1.0 / (base ** (torch.arange(0, dh, 2, device=device).float() / dh))
    fr = torch.outer(torch.arange(max_pos, device=device).float(), inv)
    emb = torch.cat([fr, fr], -1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(x, cos, sin, position_ids):
    """rotate_half + apply_rotary_pos_emb :109-123. x: [B,H,S,dh], position_ids: [B,S]."""
    c = cos[position_ids][:, None]
    s = sin[position_ids][:, None]
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], -1)
    return x * c + rot * s


LORA_DROP = None  # tests only: {(layer, "q_proj" | "v_proj"): keep-mask [T, D] already scaled by 1 / (1 - p)} shared with the device path


def _lora(sd, i, name, x, d):
    """peft LoRA (third-party, unpinned; restated from its published definition with the config at
    myriad.py:171-178): y += (alpha / r) * B(A(dropout(x))); dropout is identity in eval / parity runs unless a shared mask is
    supplied through LORA_DROP (peft: every LoRA module owns its nn.Dropout(lora_dropout) applied to its input)."""
    if d.lora_r <= 0:
        return 0.0
    if LORA_DROP is not None:
        x = x * LORA_DROP[(i, name)].reshape(x.shape)
    p = "llama_model.base_model.model.model.layers.%d.self_attn.%s." % (i, name)
    return (d.lora_alpha / d.lora_r) * linear(linear(x, sd[p + "lora_A.default.weight"]), sd[p + "lora_B.default.weight"])


def llama_layers(sd, h, attn_bias, position_ids, d: MyriadDims, past=None, p="llama_model.model."):
    """LlamaModel.forward :466-596 body: 32 x LlamaDecoderLayer.forward :247-299 (LlamaAttention :168-231,
    LlamaMLP :139-140) then final RMSNorm. attn_bias: additive [B,1,Sq,Skv] (0 / finfo.min). past: list of (k,v)."""
    l = d.llama
    B, S, _ = h.shape
    H, dh = l.heads, l.head_dim
    cos, sin = rope_tables(dh, l.max_pos, device=h.device, dtype=h.dtype)
    new_past = []
    for i in range(l.layers):
        lp = p + "layers.%d." % i
        x = rms_norm(h, sd[lp + "input_layernorm.weight"], l.eps)
        q = linear(x, sd[lp + "self_attn.q_proj.weight"]) + _lora(sd, i, "q_proj", x, d)
        k = linear(x, sd[lp + "self_attn.k_proj.weight"])
        v = linear(x, sd[lp + "self_attn.v_proj.weight"]) + _lora(sd, i, "v_proj", x, d)
        q, k, v = (t.reshape(B, S, H, dh).transpose(1, 2) for t in (q, k, v))
        q, k = apply_rope(q, cos, sin, position_ids), apply_rope(k, cos, sin, position_ids)
        if past is not None:
            k = torch.cat([past[i][0], k], 2)
            v = torch.cat([past[i][1], v], 2)
        new_past.append((k, v))
        s = (q @ k.transpose(2, 3)) / math.sqrt(dh) + attn_bias
        s = torch.max(s, torch.tensor(torch.finfo(s.dtype).min, device=s.device, dtype=s.dtype))
        a = (softmax_lastdim(s) @ v).transpose(1, 2).reshape(B, S, l.hidden)
        h = h + linear(a, sd[lp + "self_attn.o_proj.weight"])
        x = rms_norm(h, sd[lp + "post_attention_layernorm.weight"], l.eps)
        x = F.silu(linear(x, sd[lp + "mlp.gate_proj.weight"])) * linear(x, sd[lp + "mlp.up_proj.weight"])
        h = h + linear(x, sd[lp + "mlp.down_proj.weight"])
    return rms_norm(h, sd[p + "norm.weight"], l.eps), new_past


def causal_bias(attention_mask, q_len, past_len=0, dtype=torch.float32):
    """_make_causal_mask + _expand_mask + _prepare_decoder_attention_mask :25-54,442-463.
    attention_mask: [B, past_len + q_len] of 0/1."""
    neg = torch.finfo(dtype).min
    dev = attention_mask.device
    B, kv = attention_mask.shape
    bias = torch.zeros(B, 1, q_len, kv, device=dev, dtype=dtype)
    if q_len > 1:
        qi = torch.arange(q_len, device=dev)[:, None] + past_len
        ki = torch.arange(kv, device=dev)[None, :]
        bias = bias.masked_fill((ki > qi)[None, None], neg)
    pad = (attention_mask == 0)[:, None, None, :].expand(B, 1, q_len, kv)
    bias = bias + torch.zeros_like(bias).masked_fill(pad, neg)
    return bias.clamp_min(neg)


def llama_logits(sd, inputs_embeds, attention_mask, d, position_ids=None, past=None):
    """LlamaForCausalLM.forward :629-716 without labels. position_ids default = arange (:510-517)."""
    B, S, _ = inputs_embeds.shape
    past_len = 0 if past is None else past[0][0].shape[2]
    if position_ids is None:
        position_ids = torch.arange(past_len, past_len + S, device=inputs_embeds.device)[None].expand(B, -1)
    h, new_past = llama_layers(sd, inputs_embeds, causal_bias(attention_mask, S, past_len, dtype=inputs_embeds.dtype), position_ids, d, past)
    return linear(h, sd["llama_model.lm_head.weight"]), new_past


def clamp_ce_loss(logits, labels):
    """clamp_CE_loss :718-728 on shifted logits/labels (:690-703): softmax -> clamp[1e-7, 1-1e-7] -> log -> NLL
    mean over labels != -100."""
    lg = logits[:, :-1].reshape(-1, logits.shape[-1])
    lb = labels[:, 1:].reshape(-1)
    logp = torch.log(softmax_lastdim(lg).clamp(1e-7, 1 - 1e-7))
    keep = lb != -100
    return -(logp[keep, lb[keep]]).mean()


def embed_tokens(sd, ids):
    return sd["llama_model.model.embed_tokens.weight"][ids]


# ----------------------------------------------------------------------------------------------------
# Myriad.forward / Myriad.generate glue (myriad.py:354-454)
# ----------------------------------------------------------------------------------------------------
def prompt_wrap(sd, img_embeds, ids_before, ids_after):
    """prompt_wrap myriad.py:354-375 with pre-tokenised prompt halves (same ids for every sample, as in the
    shipped datasets, anomaly_detection.py:345-347)."""
    B = img_embeds.shape[0]
    pb = embed_tokens(sd, ids_before)[None].expand(B, -1, -1)
    pa = embed_tokens(sd, ids_after)[None].expand(B, -1, -1)
    return torch.cat([pb, img_embeds, pa], 1)


def myriad_loss(sd, image, maps, stage, ids_before, ids_after, text_ids, text_mask, d: MyriadDims):
    """Myriad.forward myriad.py:377-431 with the host RNG choices fixed (stage given; maps already chosen).
    text_ids: [B, Lt] right-padded with eos (pad_token = eos, :182), text_mask: [B, Lt] 0/1."""
    img = prompt_wrap(sd, encode_img(sd, image, maps, stage, d), ids_before, ids_after)
    B, Lw, _ = img.shape
    targets = torch.cat([torch.full((B, Lw + 1), -100, dtype=torch.long),
                         text_ids.masked_fill(text_ids == d.llama.eos, -100)], 1)
    bos = embed_tokens(sd, torch.full((B, 1), d.llama.bos, dtype=torch.long))
    x = torch.cat([bos, img, embed_tokens(sd, text_ids)], 1)
    mask = torch.cat([torch.ones(B, 1 + Lw, dtype=torch.long), text_mask], 1)
    logits, _ = llama_logits(sd, x, mask, d)
    return clamp_ce_loss(logits, targets), logits


def train_grads(sd, d, image, maps, stage, ids_b, ids_a, text, tmask, conv_fp16=False):
    """Loss + gradients of every trainable tensor (runner_base.py:111-119) by autograd over the oracle restatement.
    conv_fp16: emulate fp16 storage of the conv-stack activations (see CONV_FP16_ACTS)."""
    global CONV_FP16_ACTS
    CONV_FP16_ACTS = bool(conv_fp16)
    try:
        return _train_grads(sd, d, image, maps, stage, ids_b, ids_a, text, tmask)
    finally:
        CONV_FP16_ACTS = False


def _train_grads(sd, d, image, maps, stage, ids_b, ids_a, text, tmask):
    keys = [k for k in sd if k.startswith(("expert_adaptor.", "VEInstructor.", "VETokenizer.")) or ".lora_" in k]
    sd2 = dict(sd)
    for k in keys:
        sd2[k] = sd[k].clone().requires_grad_(True)
    loss, _ = myriad_loss(sd2, image, maps, stage, ids_b, ids_a, text, tmask, d)
    loss.backward()
    return loss.detach(), {k: (sd2[k].grad if sd2[k].grad is not None else torch.zeros_like(sd[k])) for k in keys}


def greedy_generate(sd, inputs_embeds, d: MyriadDims, max_new_tokens=90, stop_seqs=((835,), (2277, 29937)),
                    min_new_tokens=1, return_margins=False):
    """Myriad.generate myriad.py:433-454 -> HF generate (third-party, unpinned; restated): greedy search over
    prepare_inputs_for_generation modeling_llama.py:730-760 (prefill with inputs_embeds, then one token per step
    with the KV cache; position_ids = cumsum(mask) - 1), eos suppressed while fewer than min_new_tokens were
    produced (eval passes min_length=1, evaluation_aqa_dataset.py:289-301), finished rows padded with
    pad = eos, StoppingCriteriaSub conversation.py:96-107 (row 0 only). Returns NEW tokens only."""
    l = d.llama
    B, S, _ = inputs_embeds.shape
    dev = inputs_embeds.device
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    logits, past = llama_logits(sd, inputs_embeds, mask, d)
    unfinished = torch.ones(B, dtype=torch.long, device=dev)
    out, margins = [], []
    for step in range(max_new_tokens):
        nl = logits[:, -1].clone()
        if step < min_new_tokens:
            nl[:, l.eos] = -float("inf")
        top2 = nl.topk(2, -1).values
        margins.append(top2[:, 0] - top2[:, 1])
        nxt = nl.argmax(-1)
        nxt = nxt * unfinished + l.eos * (1 - unfinished)
        out.append(nxt)
        unfinished = unfinished * (nxt != l.eos).long()
        row0 = [int(t[0]) for t in out]
        stop = any(len(row0) >= len(s) and tuple(row0[-len(s):]) == tuple(s) for s in stop_seqs)
        if stop or int(unfinished.max()) == 0 or step == max_new_tokens - 1:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=torch.long, device=dev)], 1)
        pos = (mask.cumsum(-1) - 1)[:, -1:]
        logits, past = llama_logits(sd, embed_tokens(sd, nxt[:, None]), mask, d, position_ids=pos, past=past)
    toks = torch.stack(out, 1)
    if return_margins:
        return toks, torch.stack(margins, 1)
    return toks
