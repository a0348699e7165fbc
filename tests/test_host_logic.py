"""CPU: host-side mirror of the reference interface (registry, yaml config merge, tokenizer stand-in, conversation,
Myriad construction / state_dict contract / rejected options). No compute calls."""
import os
import types

import pytest
import torch

from myriad_b200 import synthetic as syn


def test_registry_and_exports():
    from minigpt4.common.registry import registry
    from minigpt4.models import BaseModel, Blip2Base, MiniGPT4, Myriad, load_model  # noqa: F401
    assert registry.get_model_class("myriad") is Myriad and registry.get_model_class("mini_gpt4") is MiniGPT4
    assert issubclass(Myriad, Blip2Base) and issubclass(Blip2Base, BaseModel)
    assert Myriad.PRETRAINED_MODEL_CONFIG_DICT["pretrain_vicuna"] == "configs/models/minigpt4.yaml"
    assert os.path.isfile(Myriad.default_config_path("pretrain_vicuna"))
    registry.register("a.b.c", 3)
    assert registry.get("a.b.c") == 3 and registry.get("a.x", default=7, no_warning=True) == 7
    with pytest.raises(KeyError):
        registry.register_model("myriad")(Myriad)


def test_config_merge_order(tmp_path):
    from minigpt4.common.config import Config
    y = tmp_path / "eval.yaml"
    y.write_text("model:\n  arch: myriad\n  model_type: pretrain_vicuna\n  max_txt_len: 160\n  end_sym: \"###\"\n  k_shot: 0\n"
                 "datasets:\n  anomaly_detection:\n    vis_processor:\n      train:\n        name: loc_image_train\nrun:\n  task: image_text_pretrain\n")
    cfg = Config(types.SimpleNamespace(cfg_path=str(y), options=["model.k_shot=4", "run.seed=42"]))
    m = cfg.model_cfg
    assert m.arch == "myriad" and m.max_txt_len == 160 and m.k_shot == 4            # user yaml < --options
    assert m.num_query_token == 32 and m.freeze_vit is True and m.image_size == 224  # class default yaml underneath
    assert cfg.run_cfg.task == "image_text_pretrain" and cfg.run_cfg.seed == 42
    assert cfg.datasets_cfg.anomaly_detection.vis_processor.train.name == "loc_image_train"
    assert m.get("use_lora", False) is False


def test_tokenizer_standin_and_stop_criterion():
    from minigpt4.conversation.conversation import CONV_VISION, Chat, StoppingCriteriaSub  # noqa: F401
    from minigpt4.models.tokenizer import SyntheticLlamaTokenizer
    tok = SyntheticLlamaTokenizer(32000)
    a = tok("###Human: <Img>", return_tensors="pt", add_special_tokens=False).input_ids
    assert a[0, 0].item() == 835 and a.shape == (1, 6)  # 6 tokens, the shape SURVEY §8d assumes for the prompt head
    b = tok(["short", "a longer text here"], return_tensors="pt", padding="longest", truncation=True, max_length=3, add_special_tokens=False)
    assert b.input_ids.shape == (2, 3) and b.attention_mask[0].tolist() == [1, 0, 0] and b.input_ids[0, 1].item() == tok.pad_token_id
    crit = StoppingCriteriaSub(stops=[torch.tensor([835]), torch.tensor([2277, 29937])])
    assert crit(torch.tensor([[5, 2277, 29937], [1, 1, 1]])) and crit(torch.tensor([[5, 835]]))
    assert not crit(torch.tensor([[5, 6, 7], [0, 0, 835]]))  # only row 0 counts (conversation.py:102-107)
    conv = CONV_VISION.copy()
    conv.append_message(conv.roles[0], "<Img><ImageHere></Img>")
    conv.append_message(conv.roles[1], None)
    assert conv.get_prompt().endswith("###Human: <Img><ImageHere></Img>###Assistant:")


def test_myriad_state_dict_contract_and_rejections():
    from minigpt4.models import Myriad
    d = syn.mid_dims(lora_r=8)
    m = Myriad(dims=d, weights=syn.LazyStateDict(d, 0), use_lora=True)
    keys = set(m.state_dict().keys())
    want = {"expert_adaptor.conv1.weight", "expert_adaptor.conv2.weight", "VETokenizer.base_prompts"}
    want |= {"%s.meta_net.%d.%s" % (mod, i, p) for mod in ("VETokenizer", "VEInstructor") for i in (0, 3, 6, 9, 12, 15) for p in ("weight", "bias")}
    want |= {"llama_model.base_model.model.model.layers.0.self_attn.%s.lora_%s.default.weight" % (p, ab) for p in ("q_proj", "v_proj") for ab in "AB"}
    assert keys == want
    assert all(p.requires_grad for p in m.parameters())
    assert sum(p.numel() for p in m.parameters()) == 110_854_128  # 110.73 M + one layer of LoRA r=8
    # the trainables were initialised from the weight mapping
    assert torch.equal(m.expert_adaptor.conv1.weight.detach(), syn.LazyStateDict(d, 0)["expert_adaptor.conv1.weight"])
    sd = {k: torch.full_like(v, 0.5) for k, v in m.state_dict().items()}
    msg = m.load_state_dict(sd, strict=False)
    assert not msg.missing_keys and not msg.unexpected_keys and float(m.VETokenizer.base_prompts.detach().mean()) == 0.5
    for bad in (dict(low_resource=True), dict(bliva_like=True), dict(vit_model="clip_vit_l"), dict(freeze_vit=False), dict(use_grad_checkpoint=True)):
        with pytest.raises(NotImplementedError):
            Myriad(dims=d, weights=syn.LazyStateDict(d, 0), **bad)
    with pytest.raises(KeyError):
        m.prepare_sample({"image": torch.zeros(1, 3, 224, 224), "question2": ["q <ImageHere>"], "scene": ["x"], "img_path": ["p"]}, 1)


def test_missing_checkpoints_fail_loudly(monkeypatch):
    from minigpt4.models import Myriad
    monkeypatch.delenv("MYRIAD_SYNTHETIC_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError):
        Myriad(llama_model="/nonexistent/vicuna", q_former_model="/nonexistent/blip2.pth")


def test_synthetic_weight_spec_matches_reference_sizes():
    d = syn.full_dims(lora_r=8)
    spec = {k: shape for k, shape, _, _ in syn.state_dict_spec(d)}
    import math
    n = lambda pre: sum(math.prod(s) for k, s in spec.items() if k.startswith(pre))
    assert n("visual_encoder.") == 985_894_528          # EVA-ViT-g: 985.9 M params
    assert n("VETokenizer.") == 107_416_504 and n("VEInstructor.") == 3_305_144   # 107.4 M / 3.31 M
    assert n("llama_model.model.layers.") == 6_476_267_520 and n("llama_model.base_model") == 4_194_304  # 6.476 B + 4.19 M LoRA
    assert n("Qformer.") + n("query_tokens") == 105_162_240  # trimmed Q-Former: 105.1 M (SURVEY §8c'')
    assert spec["visual_encoder.blocks.0.mlp.fc1.weight"] == (6144, 1408) and d.vit.head_dim == 88 and d.vit.tokens == 257


# ------------------------------------------------------------------------------------------------ small-batch kernel work split
@pytest.mark.parametrize("F,K,act", [(4096, 4096, 0), (12304, 4096, 0), (4096, 11008, 0), (22016, 4096, 3), (32000, 4096, 0),
                                     (5, 128, 0), (100, 128, 0), (128, 256, 3), (40000, 512, 0), (1408, 6144, 0)])
@pytest.mark.parametrize("have_counter", [0, 1])
def test_gemv_work_split_covers_every_unit_once(F, K, act, have_counter):
    """csrc/gemv.cu: units of 8 output rows are either split evenly by CTA index or handed out in groups through the atomic
    counter. Re-walk the producer's work list on the host from the plan the library computes (myr_gemv_plan, no GPU): every
    unit is requested exactly once, groups never exceed the MMA tile, and the ring fits the shared-memory budget."""
    import ctypes

    from myriad_b200._lib import lib
    out = (ctypes.c_int32 * 8)()
    assert lib().myr_gemv_plan(F, K, act, 148, have_counter, out) == 0
    n_units, gsz, grid, u_static, use_counter, stages, n_kc, fixed = list(out)
    swiglu = act == 3
    assert n_units == (F // 16 if swiglu else -(-F // 8)) and gsz == (1 if swiglu else 2)
    assert 1 <= grid <= 148 and 0 <= u_static <= n_units and (not use_counter or u_static % gsz == 0)
    assert use_counter in (0, 1) and (use_counter or u_static == n_units) and (have_counter or not use_counter)
    assert n_kc == -(-K // 1024) and 2 <= stages <= 6 and stages * 32768 + fixed <= 227 * 1024
    seen = [0] * n_units
    pool = -(-(n_units - u_static) // gsz)
    grabs = 0
    for cta in range(grid):
        u, s1 = cta * u_static // grid, (cta + 1) * u_static // grid
        while u < s1:  # the CTA's slice, in row groups of gsz units; the last one may be a single unit
            nu = min(gsz, s1 - u)
            for i in range(nu):
                seen[u + i] += 1
            u += nu
    for v in range(pool):  # groups handed out by the counter, whoever grabs them
        u0 = u_static + v * gsz
        nu = min(gsz, n_units - u0)
        assert nu >= 1
        for i in range(nu):
            seen[u0 + i] += 1
        grabs += 1
    assert seen == [1] * n_units
    if use_counter:
        # every CTA ends on exactly one failing grab, the last of them (value pool + grid - 1) puts the counter back to zero
        assert pool > 0 and grabs + grid == pool + grid
    if (F, K) in ((12304, 4096), (22016, 4096), (32000, 4096)) and have_counter:
        assert use_counter == 1 and n_units - u_static >= n_units // 4  # the large decode projections use the pool
    if (F, K) in ((4096, 4096), (4096, 11008)):
        assert use_counter == 0  # o_proj / down_proj: fewer than two row groups per CTA, even split only


@pytest.mark.parametrize("T", [5, 8, 16, 17, 32])
@pytest.mark.parametrize("F,K,act", [(4096, 4096, 0), (12304, 4096, 0), (4096, 11008, 0), (22016, 4096, 3), (32000, 4096, 0), (1100, 640, 0)])
def test_gemv_mt_work_split_covers_every_unit_once(T, F, K, act):
    """csrc/gemv_mt.cu (5 <= T <= 32): the producer's work list re-walked on the host from myr_gemv_mt_plan: every 8-row unit (SwiGLU:
    8 gate / up pairs) is requested exactly once, a SwiGLU group of two units starts at an even unit (its 16 pairs stay inside one
    64-row block of the interleaved weight), and ring + reduction buffer fit the shared-memory budget."""
    import ctypes

    from myriad_b200._lib import lib
    out = (ctypes.c_int32 * 6)()
    assert lib().myr_gemv_mt_plan(T, F, K, act, 148, 1, out) == 0
    nt, n_units, gsz, grid, u_static, stages = list(out)
    swiglu = act == 3
    assert nt == -(-T // 8) and n_units == (F // 16 if swiglu else -(-F // 8)) and gsz == (2 if swiglu else 4)
    assert 1 <= grid <= 148 and 0 <= u_static <= n_units and u_static % gsz == 0 or u_static == n_units
    stage_bytes = 32768 + nt * 8192
    assert 2 <= stages <= 6 and stages * stage_bytes + nt * 8192 + 2048 <= 227 * 1024
    seen = [0] * n_units
    for cta in range(grid):
        u, s1 = cta * u_static // grid, (cta + 1) * u_static // grid
        while u < s1:
            nu = 1 if (swiglu and u % 2) else min(gsz, s1 - u)
            assert not (swiglu and nu == 2 and u % 2)
            for i in range(nu):
                seen[u + i] += 1
            u += nu
    u = u_static
    while u < n_units:  # pool groups start at multiples of gsz from u_static
        nu = min(gsz, n_units - u)
        for i in range(nu):
            seen[u + i] += 1
        u += gsz
    assert seen == [1] * n_units
