"""GPU: the registry-registered drop-in class (`arch: myriad`, minigpt4/models/myriad.py) driven the way the reference's
callers drive it — `model.generate(samples, **kw)` as evaluation_aqa_dataset.py:289-340 does and
`model(samples)["loss"].backward()` as base_task.py:233-271 does — against the CPU oracle on the same weights."""
import random

import pytest
import torch

from myriad_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

QUESTION = "<Img><ImageHere></Img> Is there any anomaly in the image?"


def _model(lora_r):
    import minigpt4.models  # noqa: F401  (registers the classes, as the reference's `from minigpt4.models import *` does)
    from minigpt4.common.registry import registry
    d = syn.mid_dims(lora_r=lora_r)
    sd = syn.make_state_dict(d, 0)
    cls = registry.get_model_class("myriad")
    m = cls(use_lora=bool(lora_r), weights=sd, dims=d, llama_model="")
    return m.to("cuda:0"), d, sd


def _samples(B, train):
    image, maps = syn.make_inputs(B, seed=13)
    s = {"image": image.cuda(), "anomaly_maps": maps.cuda(), "oneshot_anomaly_maps": maps.cuda(), "scene": ["bottle"] * B,
         "img_path": ["x.png"] * B, "question": [QUESTION] * B, "question2": [QUESTION] * B, "question3": [QUESTION] * B}
    if train:
        s["text_input"] = ["Yes, there is a scratch on the top.", "No."][:B]
    return s, image, maps


def test_generate_through_registry_class():
    from oracle import myriad_oracle as O
    m, d, sd = _model(0)
    m.eval()
    s, image, maps = _samples(2, False)
    out = m.generate(s, max_new_tokens=8, min_length=1, do_sample=False)
    assert set(out) == {"token_ids", "ve_anomaly_maps"} and out["token_ids"].shape == (2, 8)
    ib, ia = m._split_prompts(["###Human: " + QUESTION + " ###Assistant: "], "cpu")
    emb = O.prompt_wrap(sd, O.encode_img(sd, image, maps, 1, d), ib[0], ia[0])
    toks, margins = O.greedy_generate(sd, emb, d, 8, (), return_margins=True)
    got = out["token_ids"].cpu().tolist()
    for b in range(2):
        for i, (a, c) in enumerate(zip(got[b], toks[b].tolist())):
            if a != c:
                assert float(margins[b, i]) < 0.05, (b, i, a, c)
                break
    print("drop-in generate tokens", got, "oracle", toks.tolist(), "min margin %.3f" % float(margins.min()))


@pytest.mark.parametrize("lora_r", [0, 8])
def test_forward_backward_through_registry_class(lora_r, monkeypatch):
    from oracle import myriad_oracle as O
    m, d, sd = _model(lora_r)
    m.train()
    s, image, maps = _samples(2, True)
    picks = iter([1, 0])  # stage 1, task 0 (zero-shot maps), as random.choice would draw them (myriad.py:378,381)
    monkeypatch.setattr(random, "choice", lambda seq: next(picks))
    loss = m(s)["loss"]
    assert loss.requires_grad and loss.dim() == 0
    (loss * 4.0).backward()  # a GradScaler-style factor must reach the parameter gradients
    ib, ia = m._split_prompts(["###Human: " + QUESTION + " ###Assistant: "], "cpu")
    enc = m.llama_tokenizer([t + m.end_sym for t in s["text_input"]], return_tensors="pt", padding="longest", truncation=True,
                            max_length=m.max_txt_len, add_special_tokens=False)
    oloss, ograds = O.train_grads(sd, d, image, maps, 1, ib[0], ia[0], enc.input_ids, enc.attention_mask, conv_fp16=True)
    assert abs(loss.item() - oloss.item()) < 2e-2
    worst, wk = 0.0, None
    state = m.trainable_state()
    for k, g in ograds.items():
        if g.abs().max() == 0:
            continue
        got = state[k].grad
        assert got is not None, k
        e = ((got.cpu() / 4.0 - g).abs().max() / g.abs().max()).item()
        # conv-stack weights: max-pool arg-max ties resolve differently between the tensor-core summation order and torch's
        # conv (tests/test_training_gpu.py explains); with 2 samples and ~10 supervised tokens they are not averaged out
        tol = 8e-2 if ".meta_net." in k else 3e-2
        if e > 0.3 * tol:
            print("    %-70s rel err %.2e" % (k, e))
        if e / tol > worst:
            worst, wk = e / tol, k
    print("drop-in forward/backward lora_r=%d: loss %.5f (oracle %.5f), worst grad err / tolerance %.2f (%s)" % (
        lora_r, loss.item(), oloss.item(), worst, wk))
    assert worst < 1.0, wk
    # a torch optimizer over the module's parameters (what runner_base.py:105-139 builds) must be able to step
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    before = state["VETokenizer.base_prompts"].detach().clone()
    opt.step()
    assert not torch.equal(before, state["VETokenizer.base_prompts"].detach())
