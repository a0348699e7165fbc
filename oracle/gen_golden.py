"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The fixtures are committed; the GPU box never needs the reference. Each fixture records the seeds/dims that
regenerate its weights and inputs (myriad_b200/synthetic.py) plus the reference outputs, and this script
asserts on the spot that oracle/myriad_oracle.py reproduces them (the pin).

What is reference code and what is restated here:
  reference modules (by path)  eva_vit.VisionTransformer, networks.{LoraAdaptorV2,VEInstructorV2,VETokenizer},
                               Qformer.BertLMHeadModel(.bert), modeling_llama.LlamaForCausalLM (+ its
                               prepare_inputs_for_generation and tuple KV cache)
  restated glue (not importable here: myriad.py needs peft, CUDA-at-import experts, checkpoints)
                               encode_img composition (myriad.py:241-272), forward targets/concat (:377-431),
                               greedy loop (HF generate is gone from this class in transformers 5.x),
                               peft LoRA (third-party, absent): emulated by wrapping q_proj / v_proj.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from myriad_b200 import synthetic as syn  # noqa: E402
from oracle import myriad_oracle as O  # noqa: E402
from oracle import ref_shims as R  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 0


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _load(module, sd, allowed_missing=()):
    res = module.load_state_dict(sd, strict=False)
    bad = [k for k in res.missing_keys if not any(a in k for a in allowed_missing)]
    assert not bad and not res.unexpected_keys, (bad, res.unexpected_keys)
    return module


def _close(name, a, b, tol=2e-5):
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    print("  pin %-28s max|oracle-ref| = %.3e (ref max %.3e)" % (name, err, scale))
    assert err <= tol * max(1.0, scale), name


def _close_rel(name, a, b, tol=1e-3):
    """Gradients can be small: error relative to the reference tensor's own max (no floor at 1)."""
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    print("  pin %-44s max|oracle-ref| = %.3e (ref max %.3e)" % (name, err, scale))
    assert err <= tol * scale + 1e-12, name


def _save(name, **arrs):
    os.makedirs(GOLDEN, exist_ok=True)
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.0f KB)" % (path, os.path.getsize(path) / 1024))


def build_ref_vit(sd, v):
    vit = R.build_vit(v.img, v.patch, v.dim, v.depth, v.heads, v.mlp_hidden / v.dim + 1e-9)
    assert vit.blocks[0].mlp.fc1.out_features == v.mlp_hidden
    return _load(vit, _sub(sd, "visual_encoder."))


def build_ref_qformer(sd, q, enc_width):
    qf = R.build_qformer(q.hidden, q.layers, q.heads, q.inter, enc_width, q.num_query, q.cross_freq)
    return _load(qf, _sub(sd, "Qformer."), allowed_missing=("position_ids",))


class _PeftLikeLinear(torch.nn.Module):
    """Restatement of peft.tuners.lora.Linear.forward (third-party, absent here): base(x) + B(A(x)) * alpha/r."""

    def __init__(self, base, A, B, scaling):
        super().__init__()
        self.base, self.A, self.B, self.scaling = base, A, B, scaling

    def forward(self, x):
        return self.base(x) + (x @ self.A.t()) @ self.B.t() * self.scaling


def build_ref_llama(sd, d):
    l = d.llama
    m = R.build_llama(l.hidden, l.layers, l.heads, l.inter, l.vocab, l.max_pos, l.eps)
    base = {k: v for k, v in _sub(sd, "llama_model.").items() if not k.startswith("base_model.")}
    _load(m, base, allowed_missing=("inv_freq",))
    if d.lora_r > 0:
        for i, layer in enumerate(m.model.layers):
            for n in ("q_proj", "v_proj"):
                p = "llama_model.base_model.model.model.layers.%d.self_attn.%s." % (i, n)
                setattr(layer.self_attn, n, _PeftLikeLinear(getattr(layer.self_attn, n), sd[p + "lora_A.default.weight"],
                                                            sd[p + "lora_B.default.weight"], d.lora_alpha / d.lora_r))
    return m


def ref_greedy(m, inputs_embeds, l, max_new_tokens, stop_seqs, min_new_tokens=1):
    """Hand-rolled greedy search over the reference forward + prepare_inputs_for_generation + tuple cache."""
    B, S, _ = inputs_embeds.shape
    mask = torch.ones(B, S, dtype=torch.long)
    input_ids = torch.zeros(B, 0, dtype=torch.long)
    past = None
    unfinished = torch.ones(B, dtype=torch.long)
    for step in range(max_new_tokens):
        mi = m.prepare_inputs_for_generation(input_ids, past_key_values=past, attention_mask=mask,
                                             inputs_embeds=inputs_embeds, use_cache=True)
        out = m(**mi, return_dict=True)
        past = out.past_key_values
        nl = out.logits[:, -1].clone()
        if step < min_new_tokens:
            nl[:, l.eos] = -float("inf")
        nxt = nl.argmax(-1) * unfinished + l.eos * (1 - unfinished)
        input_ids = torch.cat([input_ids, nxt[:, None]], 1)
        unfinished = unfinished * (nxt != l.eos).long()
        row0 = input_ids[0].tolist()
        if any(len(row0) >= len(s) and tuple(row0[-len(s):]) == tuple(s) for s in stop_seqs):
            break
        if int(unfinished.max()) == 0:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=torch.long)], 1)
    return input_ids


@torch.no_grad()
def gen_tiny():
    d = syn.tiny_dims()
    sd = syn.make_state_dict(d, SEED)
    image, _ = syn.make_inputs(2, seed=11, img=d.vit.img)
    # --- ViT
    vit = build_ref_vit(sd, d.vit)
    ref = vit.forward_features(image)
    _close("vit_tiny", O.vit_forward(sd, image, d.vit), ref)
    # per-block intermediates pin the block restatement too
    x0 = vit.patch_embed(image)
    _close("vit_patch_embed", O.vit_patch_embed(sd, image, d.vit), x0)
    _save("vit_tiny", seed=SEED, input_seed=11, out=ref, patch_embed=x0)
    # --- adaptor + ln_vision (LoraAdaptorV2 is the reference class; ln_vision = blip2.LayerNorm == fp32 nn.LayerNorm)
    ad = R.networks().LoraAdaptorV2(dims=d.vit.dim, input_dim=d.adaptor_rank)
    _load(ad, _sub(sd, "expert_adaptor."))
    ln = torch.nn.LayerNorm(d.vit.dim)
    _load(ln, _sub(sd, "ln_vision."))
    enc = ln(ad(ref))
    _close("adaptor_ln", O.layer_norm(O.lora_adaptor(sd, ref), sd["ln_vision.weight"], sd["ln_vision.bias"], 1e-5), enc)
    # --- Q-Former (query tokens + 5 extra "instructor" tokens so Q != num_query is covered)
    qf = build_ref_qformer(sd, d.qf, d.vit.dim)
    extra = syn.synth("extra_queries", (2, 5, d.qf.hidden), 1.0, SEED, round_fp16=False)
    qe = torch.cat([sd["query_tokens"].expand(2, -1, -1), extra], 1)
    qout = qf.bert(query_embeds=qe, encoder_hidden_states=enc,
                   encoder_attention_mask=torch.ones(enc.shape[:-1], dtype=torch.long), return_dict=True).last_hidden_state
    _close("qformer_tiny", O.qformer_forward(sd, qe, enc, d.qf), qout)
    _save("qformer_tiny", seed=SEED, enc=enc, query_embeds=qe, out=qout)


@torch.no_grad()
def gen_networks():
    d = syn.mid_dims()
    sd = syn.make_state_dict(d, SEED, only_prefix="VE")
    _, maps = syn.make_inputs(2, seed=12)
    N = R.networks()
    inst = _load(N.VEInstructorV2(), _sub(sd, "VEInstructor."))
    tok = _load(N.VETokenizer(), _sub(sd, "VETokenizer."))
    ri, rt = inst(maps), tok(maps)
    _close("ve_instructor", O.ve_instructor(sd, maps), ri)
    _close("ve_tokenizer", O.ve_tokenizer(sd, maps), rt)
    trunk = tok.meta_net[:15](maps)
    _close("conv_trunk", O.conv_stack(sd, "VETokenizer.", maps), trunk)
    _save("networks", seed=SEED, input_seed=12, instructor=ri, tokenizer=rt[:, :, ::4], trunk_tok=trunk[:, ::8])


@torch.no_grad()
def gen_llama_tiny():
    for lora_r in (0, 8):
        d = syn.tiny_dims(lora_r=lora_r)
        l = d.llama
        sd = syn.make_state_dict(d, SEED, only_prefix="llama_model")
        m = build_ref_llama(sd, d)
        B, S = 2, 12
        x = syn.synth("llama_in", (B, S, l.hidden), 0.5, SEED, round_fp16=False)
        mask = torch.ones(B, S, dtype=torch.long)
        mask[1, 9:] = 0  # right padding, as in training (myriad.py:395-404)
        labels = torch.randint(3, l.vocab, (B, S), generator=torch.Generator().manual_seed(5))
        labels[:, :4] = -100
        labels[1, 9:] = -100
        out = m(inputs_embeds=x, attention_mask=mask, labels=labels, return_dict=True, use_cache=True)
        ol, _ = O.llama_logits(sd, x, mask, d)
        tag = "llama_tiny" + ("_lora" if lora_r else "")
        _close(tag + ".logits", ol, out.logits, 5e-5)
        _close(tag + ".loss", O.clamp_ce_loss(ol, labels), out.loss, 5e-5)
        # greedy decode (all-ones mask; stop sequences chosen inside the tiny vocab)
        xg = x[:, :7].contiguous()
        stops = ((100,), (101, 102))
        toks = ref_greedy(m, xg, l, 12, stops)
        ot, margins = O.greedy_generate(sd, xg, d, 12, stops, return_margins=True)
        assert toks.shape == ot.shape and bool((toks == ot).all()), (toks, ot)
        print("  pin %-28s greedy tokens identical %s, min margin %.3f" % (tag, tuple(toks.shape), margins.min().item()))
        _save(tag, seed=SEED, lora_r=lora_r, x=x, mask=mask, labels=labels, logits=out.logits, loss=out.loss,
              greedy_tokens=toks, greedy_margins=margins)


@torch.no_grad()
def gen_mid():
    """Composite: the reference sub-modules wired as myriad.py:241-272 / :377-431 / :433-454 wires them."""
    d = syn.mid_dims()
    sd = syn.make_state_dict(d, SEED)
    B = 2
    image, maps = syn.make_inputs(B, seed=13)
    vit = build_ref_vit(sd, d.vit)
    N = R.networks()
    ad = _load(N.LoraAdaptorV2(dims=d.vit.dim, input_dim=d.adaptor_rank), _sub(sd, "expert_adaptor."))
    ln = _load(torch.nn.LayerNorm(d.vit.dim), _sub(sd, "ln_vision."))
    inst = _load(N.VEInstructorV2(), _sub(sd, "VEInstructor."))
    tok = _load(N.VETokenizer(), _sub(sd, "VETokenizer."))
    qf = build_ref_qformer(sd, d.qf, d.vit.dim)
    proj = _load(torch.nn.Linear(d.qf.hidden, d.llama.hidden), _sub(sd, "llama_proj."))
    llama = build_ref_llama(sd, d)

    def ref_encode(stage):
        emb = ln(ad(vit.forward_features(image)))
        q = sd["query_tokens"].expand(B, -1, -1)
        if stage in (1, 2):
            q = torch.cat([q, inst(maps)], 1)
        h = qf.bert(query_embeds=q, encoder_hidden_states=emb,
                    encoder_attention_mask=torch.ones(emb.shape[:-1], dtype=torch.long), return_dict=True).last_hidden_state
        t = proj(h)
        if stage in (0, 1):
            t = torch.cat([t, tok(maps)], 1)
        return t

    arrs = {}
    for stage in (0, 1, 2):
        ref = ref_encode(stage)
        _close("encode_img stage %d" % stage, O.encode_img(sd, image, maps, stage, d), ref, 5e-5)
        arrs["encode_stage%d" % stage] = ref[:, :, ::8]
    # training forward (stage 1): prompt wrap, bos, targets, loss
    l = d.llama
    ids_b, ids_a = syn.make_prompt_ids(l.vocab)
    g = torch.Generator().manual_seed(21)
    Lt = 10
    text = torch.randint(3, l.vocab, (B, Lt), generator=g)
    tmask = torch.ones(B, Lt, dtype=torch.long)
    text[1, 7:] = l.eos
    tmask[1, 7:] = 0
    emb = llama.model.embed_tokens
    img = ref_encode(1)
    wrapped = torch.cat([emb(ids_b)[None].expand(B, -1, -1), img, emb(ids_a)[None].expand(B, -1, -1)], 1)
    targets = torch.cat([torch.full((B, wrapped.shape[1] + 1), -100, dtype=torch.long),
                         text.masked_fill(text == l.eos, -100)], 1)
    x = torch.cat([emb(torch.full((B, 1), l.bos)), wrapped, emb(text)], 1)
    am = torch.cat([torch.ones(B, 1 + wrapped.shape[1], dtype=torch.long), tmask], 1)
    out = llama(inputs_embeds=x, attention_mask=am, labels=targets, return_dict=True)
    oloss, ologits = O.myriad_loss(sd, image, maps, 1, ids_b, ids_a, text, tmask, d)
    _close("forward.logits", ologits, out.logits, 1e-4)
    _close("forward.loss", oloss, out.loss, 1e-4)
    # generate (stage 1, no bos, all-ones mask)
    stops = ((835,), (2277, 29937))
    toks = ref_greedy(llama, wrapped, l, 8, stops)
    ot, margins = O.greedy_generate(sd, O.prompt_wrap(sd, O.encode_img(sd, image, maps, 1, d), ids_b, ids_a), d, 8, stops,
                                    return_margins=True)
    assert bool((toks == ot).all()), (toks, ot)
    print("  pin %-28s greedy tokens identical %s, min margin %.3f" % ("generate", tuple(toks.shape), margins.min().item()))
    _save("myriad_mid", seed=SEED, input_seed=13, text=text, text_mask=tmask, loss=out.loss,
          logits_sub=out.logits[:, ::4, ::5], greedy_tokens=toks, greedy_margins=margins, **arrs)


def _sample(t, n=4096):
    """Deterministic subsample of a gradient tensor (flat stride) so fixtures stay small."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n) | 1  # odd stride: does not alias with the power-of-two tensor dims
    return f[::step][:n].clone()


oracle_train_grads = O.train_grads


def gen_mid_train():
    """Training step (Myriad.forward + backward, stage 1). lora_r = 0: the UNMODIFIED reference modules under autograd pin
    the oracle's gradients; lora_r = 8: peft is absent, so the LoRA gradients come from the restated oracle only."""
    B = 2
    image, maps = syn.make_inputs(B, seed=13)
    g = torch.Generator().manual_seed(21)
    Lt = 10
    for lora_r in (0, 8):
        d = syn.mid_dims(lora_r=lora_r)
        l = d.llama
        sd = syn.make_state_dict(d, SEED)
        ids_b, ids_a = syn.make_prompt_ids(l.vocab)
        text = torch.randint(3, l.vocab, (B, Lt), generator=torch.Generator().manual_seed(21))
        tmask = torch.ones(B, Lt, dtype=torch.long)
        text[1, 7:] = l.eos
        tmask[1, 7:] = 0
        arrs = {}
        for stage in (0, 1, 2):
            oloss, ograds = oracle_train_grads(sd, d, image, maps, stage, ids_b, ids_a, text, tmask)
            if lora_r == 0 and stage == 1:
                with torch.no_grad():
                    vit = build_ref_vit(sd, d.vit)
                    N = R.networks()
                    ad = _load(N.LoraAdaptorV2(dims=d.vit.dim, input_dim=d.adaptor_rank), _sub(sd, "expert_adaptor."))
                    ln = _load(torch.nn.LayerNorm(d.vit.dim), _sub(sd, "ln_vision."))
                    inst = _load(N.VEInstructorV2(), _sub(sd, "VEInstructor."))
                    tok = _load(N.VETokenizer(), _sub(sd, "VETokenizer."))
                    qf = build_ref_qformer(sd, d.qf, d.vit.dim)
                    proj = _load(torch.nn.Linear(d.qf.hidden, d.llama.hidden), _sub(sd, "llama_proj."))
                    llama = build_ref_llama(sd, d)
                with torch.enable_grad():
                    for m in (vit, ln, qf, proj, llama):
                        for p_ in m.parameters():
                            p_.requires_grad_(False)
                    emb_img = ln(ad(vit.forward_features(image)))
                    q = torch.cat([sd["query_tokens"].expand(B, -1, -1), inst(maps)], 1)
                    h = qf.bert(query_embeds=q, encoder_hidden_states=emb_img,
                                encoder_attention_mask=torch.ones(emb_img.shape[:-1], dtype=torch.long), return_dict=True).last_hidden_state
                    img = torch.cat([proj(h), tok(maps)], 1)
                    emb = llama.model.embed_tokens
                    wrapped = torch.cat([emb(ids_b)[None].expand(B, -1, -1), img, emb(ids_a)[None].expand(B, -1, -1)], 1)
                    targets = torch.cat([torch.full((B, wrapped.shape[1] + 1), -100, dtype=torch.long),
                                         text.masked_fill(text == l.eos, -100)], 1)
                    x = torch.cat([emb(torch.full((B, 1), l.bos)), wrapped, emb(text)], 1)
                    am = torch.cat([torch.ones(B, 1 + wrapped.shape[1], dtype=torch.long), tmask], 1)
                    out = llama(inputs_embeds=x, attention_mask=am, labels=targets, return_dict=True)
                    out.loss.backward()
                _close("train.loss", oloss, out.loss.detach(), 1e-4)
                for mod, pre in ((ad, "expert_adaptor."), (inst, "VEInstructor."), (tok, "VETokenizer.")):
                    for n_, p_ in mod.named_parameters():
                        _close_rel("train.grad " + pre + n_, ograds[pre + n_], p_.grad, 1e-3)
            arrs["loss_stage%d" % stage] = oloss
            for k, v in ograds.items():
                arrs["s%d:%s" % (stage, k)] = _sample(v)
            # conv-stack gradients under fp16 activation storage (restated emulation; see myriad_oracle.CONV_FP16_ACTS)
            _, egrads = O.train_grads(sd, d, image, maps, stage, ids_b, ids_a, text, tmask, conv_fp16=True)
            for k, v in egrads.items():
                if ".meta_net." in k:
                    arrs["e%d:%s" % (stage, k)] = _sample(v)
        _save("myriad_mid_train" + ("_lora" if lora_r else ""), seed=SEED, input_seed=13, lora_r=lora_r, text=text, text_mask=tmask, **arrs)


def main():
    assert R.available(), "reference tree not found at %s" % R.REF_ROOT
    torch.manual_seed(0)
    if "--train-only" in sys.argv:
        gen_mid_train()
        return
    gen_tiny()
    gen_networks()
    gen_llama_tiny()
    gen_mid()
    gen_mid_train()


if __name__ == "__main__":
    main()
