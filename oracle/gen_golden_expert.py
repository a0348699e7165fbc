"""TEST INFRASTRUCTURE ONLY — generates tests/golden/imagebind_tiny.npz by running the UNMODIFIED reference ImageBind vision model.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden_expert
Loads /root/reference/minigpt4/models/model/ImageBind/models/{helpers,multimodal_preprocessors,transformer,imagebind_model}.py by
path (stubs for the absent timm / ftfy / iopath imports, none of which the vision modality executes), builds ImageBindModel at the
dims of myriad_b200.expert.tiny_expert_dims() with the seeded weights of make_expert_state_dict, runs the reference forward on seeded
224 x 224 images and stores the tapped tokens; asserts on the spot that oracle/expert_oracle.vision_taps reproduces them (the pin).
The map heads (adrefexpert_v2.py:245-301) are not importable (kornia / jsonlines / CUDA at import): their oracle outputs are stored
next to the taps, marked `oracle_`, so the GPU tests have fixed targets.
"""
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from myriad_b200 import expert as X  # noqa: E402
from myriad_b200 import synthetic as syn  # noqa: E402
from oracle import expert_oracle as EO  # noqa: E402
from oracle import ref_shims as R  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
IB_DIR = os.path.join(R.REF_ROOT, "minigpt4", "models", "model", "ImageBind", "models")


def load_reference_imagebind():
    R._install_timm_stub()
    layers = sys.modules["timm.models.layers"]

    class DropPath(torch.nn.Module):  # drop_path = 0.0 for the vision trunk (imagebind_model.py:329)
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    layers.DropPath = DropPath
    for name in ("ftfy", "iopath", "iopath.common", "iopath.common.file_io"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    sys.modules["iopath.common.file_io"].g_pathmgr = None  # only the text tokenizer opens files
    pkg = types.ModuleType("ref_imagebind_models")
    pkg.__path__ = [IB_DIR]
    pkg.__spec__ = importlib.machinery.ModuleSpec("ref_imagebind_models", None, is_package=True)
    sys.modules["ref_imagebind_models"] = pkg
    return importlib.import_module("ref_imagebind_models.imagebind_model")


def main():
    d = X.tiny_expert_dims()
    sd = X.make_expert_state_dict(d, seed=0)
    ib = load_reference_imagebind()
    small = dict(text_embed_dim=64, text_num_blocks=1, text_num_heads=2, audio_embed_dim=64, audio_num_blocks=1, audio_num_heads=2,
                 depth_embed_dim=64, depth_num_blocks=1, depth_num_heads=2, thermal_embed_dim=64, thermal_num_blocks=1,
                 thermal_num_heads=2, imu_embed_dim=64, imu_num_blocks=1, imu_num_heads=2)
    model = ib.ImageBindModel(vision_embed_dim=d.dim, vision_num_blocks=d.depth, vision_num_heads=d.heads, out_embed_dim=d.dec_dim,
                              layers=list(d.out_layers), **small).eval()
    ref_sd = {k[len(X.VE):]: v for k, v in sd.items() if k.startswith(X.VE)}
    res = model.load_state_dict(ref_sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert not [k for k in res.missing_keys if ".vision." in k and "modality_heads" not in k], res.missing_keys
    B, k_ref = 2, 2
    image, _ = syn.make_inputs(B, seed=77)
    refs, _ = syn.make_inputs(B * k_ref, seed=78)
    with torch.no_grad():
        # adrefexpert_v2.py:201-205: inputs = {VISION: torch.stack([images], dim=0)} -> visual_encoder(inputs)['vision'][1]
        taps = [t.transpose(0, 1).contiguous() for t in model({"vision": torch.stack([image], dim=0)})["vision"][1]]
        taps_ref = [t.transpose(0, 1).contiguous() for t in model({"vision": torch.stack([refs], dim=0)})["vision"][1]]
        mine = EO.vision_taps(sd, image, d)
        mine_ref = EO.vision_taps(sd, refs, d)
    for l, (a, b) in enumerate(zip(mine + mine_ref, taps + taps_ref)):
        err = (a - b).abs().max().item()
        print("  pin tap %d  max|oracle-ref| = %.3e (ref max %.3e)" % (l, err, b.abs().max().item()))
        assert err <= 3e-5 * max(1.0, b.abs().max().item())
    text = X.make_text_features(B, d, seed=0)
    with torch.no_grad():
        zs_maps, zs_masks = EO.zero_shot(sd, taps, text, d)
        ks_maps, ks_simmask = EO.k_shot(taps, taps_ref, d)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "imagebind_tiny.npz"), seed=0, image_seed=77, ref_seed=78, B=B, k_ref=k_ref,
                        **{"ref_tap%d" % l: t.numpy().astype(np.float32) for l, t in enumerate(taps)},
                        **{"ref_reftap%d_mean" % l: t.mean(dim=(0, 1)).numpy().astype(np.float32) for l, t in enumerate(taps_ref)},
                        oracle_zs_maps=zs_maps.numpy().astype(np.float16), oracle_zs_masks=zs_masks.numpy(),
                        oracle_ks_maps=ks_maps.numpy().astype(np.float16), oracle_ks_simmask=ks_simmask.numpy())
    print("wrote imagebind_tiny.npz: taps", [tuple(t.shape) for t in taps], "zs", tuple(zs_maps.shape), "ks", tuple(ks_maps.shape))


if __name__ == "__main__":
    main()
