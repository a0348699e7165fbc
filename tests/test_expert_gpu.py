"""GPU: the vision expert (myriad_b200/expert.py + csrc/expert.cu) through the C-ABI against the golden vectors of the UNMODIFIED
reference ImageBind trunk and against the CPU oracle of the heads. Tolerances: activations max|dev - ref| <= 2e-3 * max(1, max|ref|)
(fp16 operands, fp32 accumulation; the k-shot path compares unit vectors rounded to fp16); maps are probabilities in [0, 1]."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imagebind_tiny.npz")
TOL = 2e-3


def dev():
    return torch.device("cuda:0")


def close(a, b, tol, what):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all(), what
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), "%s: max err %.3e" % (what, err)
    return err


@pytest.fixture(scope="module")
def tiny():
    from myriad_b200 import expert as X
    from myriad_b200 import synthetic as syn
    gold = np.load(GOLDEN)
    d = X.tiny_expert_dims()
    sd = X.make_expert_state_dict(d, seed=int(gold["seed"]))
    image, _ = syn.make_inputs(int(gold["B"]), seed=int(gold["image_seed"]))
    refs, _ = syn.make_inputs(int(gold["B"]) * int(gold["k_ref"]), seed=int(gold["ref_seed"]))
    eng = X.VisionExpertEngine(sd, d, device="cuda:0")
    return X, gold, d, sd, image, refs, eng


def test_trunk_taps_vs_reference_imagebind(tiny):
    X, gold, d, sd, image, refs, eng = tiny
    taps, taps_n = eng.trunk(image.to(dev()), raw=True, unit=True)
    for l in range(len(d.out_layers)):
        ref = torch.from_numpy(gold["ref_tap%d" % l])[:, 1:, :].reshape(-1, d.dim)
        e = close(taps[l], ref, TOL, "tap %d" % l)
        close(taps_n[l], ref / ref.norm(dim=-1, keepdim=True), TOL, "unit tap %d" % l)
        print("tap %d: max err %.2e" % (l, e))


def test_zero_shot_and_k_shot_vs_oracle(tiny):
    from oracle import expert_oracle as EO
    X, gold, d, sd, image, refs, eng = tiny
    text = X.make_text_features(2, d, seed=0)
    maps, masks = eng.zero_shot(image.to(dev()), text)
    assert maps.shape == (2, 1, 224, 224) and masks.shape == (2, 1, 16, 16)
    close(maps, torch.from_numpy(gold["oracle_zs_maps"]).float(), 3e-3, "zero-shot maps")
    close(masks, torch.from_numpy(gold["oracle_zs_masks"]), 3e-3, "zero-shot masks")
    km, ks = eng.k_shot(image.to(dev()), refs.to(dev()))
    close(km, torch.from_numpy(gold["oracle_ks_maps"]).float(), 3e-3, "k-shot maps")
    close(ks, torch.from_numpy(gold["oracle_ks_simmask"]), 3e-3, "k-shot simmask")
    # one trunk pass for both heads == the two separate calls, bit for bit
    (m2, k2), (km2, ks2) = eng.both(image.to(dev()), text, ref_images=refs.to(dev()))
    assert torch.equal(m2, maps) and torch.equal(k2, masks) and torch.equal(km2, km) and torch.equal(ks2, ks)


def test_heads_alone_on_reference_taps(tiny):
    """The head kernels fed the REFERENCE trunk's tokens (rounded to fp16): isolates csrc/expert.cu from the trunk's error."""
    from oracle import expert_oracle as EO
    X, gold, d, sd, image, refs, eng = tiny
    taps32 = [torch.from_numpy(gold["ref_tap%d" % l]) for l in range(len(d.out_layers))]
    B, N, D = taps32[0].shape
    t16, t16n = [], []
    for t in taps32:
        a = torch.empty(B * (N - 1), D, device=dev(), dtype=torch.float16)
        b = torch.empty_like(a)
        from myriad_b200 import kernels as K
        K.expert_tap(t.to(dev()).contiguous(), a, B, N, D, normalize=False)
        K.expert_tap(t.to(dev()).contiguous(), b, B, N, D, normalize=True)
        assert torch.equal(a.cpu(), t[:, 1:, :].reshape(-1, D).half())
        t16.append(a)
        t16n.append(b)
    text = X.make_text_features(2, d, seed=0)
    maps, masks = eng.zero_shot_from_taps(t16, text)
    close(maps, torch.from_numpy(gold["oracle_zs_maps"]).float(), 1.5e-3, "zero-shot maps (reference taps)")
    close(masks, torch.from_numpy(gold["oracle_zs_masks"]), 1e-3, "zero-shot masks (reference taps)")
    # query against itself: sim = 1 up to the fp16 rounding of the unit vectors
    km, ks = eng.k_shot_from_taps(t16n, t16n)
    assert km.abs().max().item() < 2e-3 and ks.abs().max().item() < 2e-3


def test_full_width_two_blocks_vs_oracle():
    """ImageBind-Huge widths (1280, 16 heads of 80, MLP 5120), first two blocks tapped, B = 2."""
    from myriad_b200 import expert as X
    from myriad_b200 import synthetic as syn
    from oracle import expert_oracle as EO
    d = X.ExpertDims(depth=2, out_layers=(0, 1))
    sd = X.make_expert_state_dict(d, seed=3)
    image, _ = syn.make_inputs(2, seed=5)
    refs, _ = syn.make_inputs(2, seed=6)
    eng = X.VisionExpertEngine(sd, d, device="cuda:0")
    text = X.make_text_features(2, d, seed=1)
    with torch.no_grad():
        taps = EO.vision_taps(sd, image, d)
        taps_r = EO.vision_taps(sd, refs, d)
        zm, zk = EO.zero_shot(sd, taps, text, d)
        km, ks = EO.k_shot(taps, taps_r, d)
    t16, _ = eng.trunk(image.to(dev()))
    for l in range(2):
        print("full-width tap %d: %.2e" % (l, close(t16[l], taps[l][:, 1:, :].reshape(-1, d.dim), TOL, "tap %d" % l)))
    (m, k), (m2, k2) = eng.both(image.to(dev()), text, ref_images=refs.to(dev()))
    close(m, zm, 3e-3, "zero-shot maps")
    close(k, zk, 3e-3, "zero-shot masks")
    close(m2, km, 3e-3, "k-shot maps")
    close(k2, ks, 3e-3, "k-shot simmask")


def test_drop_in_class_both_branches(monkeypatch):
    """minigpt4.models.adrefexpert_v2.adrefexpert with the reference's call signatures (myriad.py:342-348)."""
    monkeypatch.setenv("MYRIAD_SYNTHETIC_WEIGHTS", "1")
    from minigpt4.models.adrefexpert_v2 import adrefexpert
    from myriad_b200 import expert as X
    from myriad_b200 import synthetic as syn
    from oracle import expert_oracle as EO
    d = X.tiny_expert_dims()
    ex = adrefexpert(round_index=0, k_shot=2, dims=d)
    sd = X.make_expert_state_dict(d, seed=0)
    image, _ = syn.make_inputs(2, seed=21)
    scenes = ["bottle", "metal_nut"]
    for i, c in enumerate(scenes):
        r, _ = syn.make_inputs(2, seed=30 + i)
        ex.register_references(c, r)
    maps, masks = ex(image.cuda(), scenes)
    one, simmask = ex(image.cuda(), scenes, querypath=["unused"] * 2, testphase=True)
    with torch.no_grad():
        taps = EO.vision_taps(sd, image, d)
        text = torch.stack([ex.text_features[c.replace("_", " ")] for c in scenes])
        zm, zk = EO.zero_shot(sd, taps, text, d)
        refs = torch.cat([syn.make_inputs(2, seed=30 + i)[0] for i in range(2)])
        km, ks = EO.k_shot(taps, EO.vision_taps(sd, refs, d), d)
    close(maps, zm, 3e-3, "plugin zero-shot maps")
    close(masks, zk, 3e-3, "plugin zero-shot masks")
    close(one, km, 3e-3, "plugin k-shot maps")
    close(simmask, ks, 3e-3, "plugin k-shot simmask")
