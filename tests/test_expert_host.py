"""CPU: the vision-expert oracle (oracle/expert_oracle.py) against the golden vectors the UNMODIFIED reference ImageBind model produced
(tests/golden/imagebind_tiny.npz, written by oracle/gen_golden_expert.py), properties of the restated map heads, and the drop-in
class's host logic. No CUDA work."""
import os

import numpy as np
import pytest
import torch

from myriad_b200 import expert as X
from myriad_b200 import synthetic as syn
from oracle import expert_oracle as EO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imagebind_tiny.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def setup(gold):
    d = X.tiny_expert_dims()
    sd = X.make_expert_state_dict(d, seed=int(gold["seed"]))
    image, _ = syn.make_inputs(int(gold["B"]), seed=int(gold["image_seed"]))
    refs, _ = syn.make_inputs(int(gold["B"]) * int(gold["k_ref"]), seed=int(gold["ref_seed"]))
    return d, sd, image, refs


def test_oracle_trunk_matches_reference_imagebind(gold, setup):
    d, sd, image, refs = setup
    with torch.no_grad():
        taps = EO.vision_taps(sd, image, d)
        taps_ref = EO.vision_taps(sd, refs, d)
    assert len(taps) == len(d.out_layers)
    for l, t in enumerate(taps):
        ref = torch.from_numpy(gold["ref_tap%d" % l])
        assert t.shape == ref.shape == (2, 257, d.dim)
        assert (t - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())
        m = torch.from_numpy(gold["ref_reftap%d_mean" % l])
        assert (taps_ref[l].mean(dim=(0, 1)) - m).abs().max().item() <= 5e-5


def test_oracle_heads_match_stored_outputs_and_properties(gold, setup):
    d, sd, image, refs = setup
    taps = [torch.from_numpy(gold["ref_tap%d" % l]) for l in range(len(d.out_layers))]
    text = X.make_text_features(2, d, seed=0)
    maps, masks = EO.zero_shot(sd, taps, text, d)
    assert maps.shape == (2, 1, 224, 224) and masks.shape == (2, 1, 16, 16)
    assert (maps - torch.from_numpy(gold["oracle_zs_maps"]).float()).abs().max().item() < 2e-3  # stored as fp16
    assert (masks - torch.from_numpy(gold["oracle_zs_masks"])).abs().max().item() < 1e-5
    assert 0.0 < maps.min().item() and maps.max().item() < 1.0
    # align_corners=True: the four corners of the up-sampled map are the corner cells of the 16 x 16 grid of each layer's softmax
    # only when softmax and interpolation commute, which they do not; but a constant logit grid must give a constant map
    flat = [torch.ones_like(t) for t in taps]
    m2, k2 = EO.zero_shot(sd, flat, text, d)
    assert (m2 - m2[:, :, :1, :1]).abs().max().item() < 1e-6 and (k2 - k2[:, :, :1, :1]).abs().max().item() < 1e-6
    # k-shot with the query itself among the references: every patch finds itself, sim = 1, anomaly map = 0
    am, sm = EO.k_shot(taps, taps, d)
    assert am.abs().max().item() < 1e-5 and sm.abs().max().item() < 1e-5


def test_expert_dims_and_state_dict_names():
    d = X.ExpertDims()
    assert (d.tokens, d.head_dim, d.mlp_hidden, d.grid) == (257, 80, 5120, 16)
    keys = {k for k, _, _, _ in X.expert_state_dict_spec(X.tiny_expert_dims())}
    # names of adrefexpert.state_dict() (imagebind_model.py modality_* dicts + image_decoder, adrefexpert_v2.py:104-108)
    for k in ("visual_encoder.modality_preprocessors.vision.cls_token", "visual_encoder.modality_preprocessors.vision.rgbt_stem.proj.1.weight",
              "visual_encoder.modality_preprocessors.vision.pos_embedding_helper.pos_embed",
              "visual_encoder.modality_trunks.vision.pre_transformer_layer.0.weight",
              "visual_encoder.modality_trunks.vision.blocks.3.attn.in_proj_weight",
              "visual_encoder.modality_trunks.vision.blocks.0.mlp.fc2.bias", "image_decoder.fc.3.weight"):
        assert k in keys, k


def test_drop_in_class_host_logic(monkeypatch, tmp_path):
    monkeypatch.setenv("MYRIAD_SYNTHETIC_WEIGHTS", "1")
    from minigpt4.models.adrefexpert_v2 import MVTEC_CLASS_NAMES, adrefexpert, load_and_transform_vision_data
    ex = adrefexpert(round_index=1, k_shot=2, dims=X.tiny_expert_dims(), data_root=str(tmp_path))
    assert ex.mvtec_references["bottle"] == [os.path.join(str(tmp_path), "mvtec", "bottle", "train", "good", n) for n in ("004.png", "005.png")]
    assert set(ex.mvtec_references) == set(MVTEC_CLASS_NAMES)
    assert ex.text_features["metal nut"].shape == (2, ex.dims.dec_dim)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ex.engine  # no CPU path
    monkeypatch.delenv("MYRIAD_SYNTHETIC_WEIGHTS")
    with pytest.raises(FileNotFoundError):
        adrefexpert()
    # ImageBind's image transform on a file
    from PIL import Image
    arr = (np.random.RandomState(0).rand(300, 260, 3) * 255).astype(np.uint8)
    Image.fromarray(arr).save(tmp_path / "a.png")
    x = load_and_transform_vision_data([str(tmp_path / "a.png")])
    assert x.shape == (1, 3, 224, 224) and torch.isfinite(x).all() and -2.5 < x.min().item() < x.max().item() < 2.5
