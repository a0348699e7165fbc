"""CPU: LR schedules / loss-scale schedule / weight-decay split restated from the reference (optims.py, runner_base.py).
When /root/reference is present (build container) the schedules are also checked against the reference's own functions."""
import importlib.util
import math
import os
import sys
import types

import pytest

from myriad_b200 import optim


def test_cosine_warmup_schedule_values():
    s = optim.LinearWarmupCosineLR(max_epoch=5, iters_per_epoch=200, min_lr=1e-6, init_lr=3e-5, warmup_steps=200, warmup_start_lr=1e-6)
    assert s.lr(0, 0) == pytest.approx(1e-6)
    assert s.lr(0, 100) == pytest.approx(1e-6 + (3e-5 - 1e-6) * 0.5)
    assert s.lr(1, 0) == pytest.approx((3e-5 - 1e-6) * 0.5 * (1 + math.cos(math.pi * 200 / 1000)) + 1e-6)
    assert s.lr(4, 199) > 1e-6 and s.lr(4, 199) < 2e-6
    st = optim.LinearWarmupStepLR(max_epoch=5, min_lr=1e-6, init_lr=1e-4, decay_rate=0.5, warmup_steps=10, warmup_start_lr=0)
    assert st.lr(0, 5) == pytest.approx(5e-5) and st.lr(2, 0) == pytest.approx(2.5e-5) and st.lr(40, 0) == 1e-6


@pytest.mark.skipif(not os.path.exists("/root/reference/minigpt4/common/optims.py"), reason="reference tree not mounted")
def test_schedules_match_reference_functions():
    reg = types.ModuleType("minigpt4.common.registry")
    reg.registry = types.SimpleNamespace(register_lr_scheduler=lambda name: (lambda cls: cls))
    saved = {k: sys.modules.get(k) for k in ("minigpt4", "minigpt4.common", "minigpt4.common.registry")}
    try:
        sys.modules["minigpt4.common.registry"] = reg
        spec = importlib.util.spec_from_file_location("_ref_optims", "/root/reference/minigpt4/common/optims.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    class Opt:
        def __init__(self):
            self.param_groups = [{"lr": 0.0}]

    o = Opt()
    r = ref.LinearWarmupCosineLRScheduler(o, max_epoch=4, iters_per_epoch=50, min_lr=8e-5, init_lr=1e-3, warmup_steps=30, warmup_start_lr=1e-6)
    m = optim.LinearWarmupCosineLR(4, 50, 8e-5, 1e-3, 30, 1e-6)
    for ep in range(4):
        for st in (0, 7, 29, 30, 49):
            r.step(ep, st)
            assert o.param_groups[0]["lr"] == pytest.approx(m.lr(ep, st), rel=1e-12), (ep, st)
    r2 = ref.LinearWarmupStepLRScheduler(o, max_epoch=4, min_lr=1e-6, init_lr=1e-3, decay_rate=0.3, warmup_start_lr=1e-5, warmup_steps=20)
    m2 = optim.LinearWarmupStepLR(4, 1e-6, 1e-3, 0.3, 1e-5, 20)
    for ep in range(4):
        for st in (0, 10, 25):
            r2.step(ep, st)
            assert o.param_groups[0]["lr"] == pytest.approx(m2.lr(ep, st), rel=1e-12)


def test_loss_scale_schedule_and_wd_split():
    sc = optim.DynamicLossScale(init_scale=1024.0, growth_interval=3)
    assert sc.update(False) == 1024 and sc.update(False) == 1024 and sc.update(False) == 2048
    assert sc.update(True) == 1024 and sc.update(False) == 1024
    sd = sc.state_dict()
    sc2 = optim.DynamicLossScale()
    sc2.load_state_dict(sd)
    assert sc2.scale == 1024 and sc2._good == 1
    assert optim.no_weight_decay("VETokenizer.meta_net.0.bias", 1) and not optim.no_weight_decay("VETokenizer.meta_net.0.weight", 4)
    assert not optim.no_weight_decay("expert_adaptor.conv1.weight", 2) and not optim.no_weight_decay("VETokenizer.base_prompts", 2)
    assert not optim.no_weight_decay("llama_model.base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight", 2)
