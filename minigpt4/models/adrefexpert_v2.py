"""Drop-in for the reference's vision expert `adrefexpert` (minigpt4/models/adrefexpert_v2.py:99-301) on the sm_100a kernels.

Same call contract as the reference class: `expert(images, cls_names)` -> zero-shot (anomaly_maps [B,1,224,224], masks [B,1,16,16]);
`expert(images, cls_names, querypath=paths, testphase=...)` -> k-shot (1 - sim maps, simmask). `Myriad` calls both per batch
(myriad.py:342-348); assign an instance to `model.vision_expert`.

What differs from the reference, on purpose:
  * the ImageBind-Huge vision trunk and both heads run in `myriad_b200.expert.VisionExpertEngine` (no torch arithmetic);
  * the TEXT side of the zero-shot branch (prompt ensemble through ImageBind's text tower, :41-96) is a table lookup:
    `text_features[class] = [2, 1024]` (normal, abnormal), computed once offline by the reference and passed in / loaded from
    `text_features_path` (torch.save of {class: tensor}); seeded synthetic rows under MYRIAD_SYNTHETIC_WEIGHTS=1;
  * reference ("normal") images of the k-shot branch: the reference re-reads and re-encodes the same files on every call
    (:256-261); here every class's reference tokens are encoded once and cached (`_ref_bank`, the name the reference reserves for it);
    files are decoded with PIL using ImageBind's transform (data.py: resize 224 bicubic, centre crop, CLIP mean / std), or register
    tensors with `register_references`.
"""
import logging
import os

import torch
import torch.nn as nn

from myriad_b200 import expert as X

CLASS_NAMES = ['bottle', 'cable', 'capsule', 'carpet', 'grid', 'hazelnut', 'leather', 'metal nut', 'pill', 'screw', 'tile',
               'toothbrush', 'transistor', 'wood', 'zipper', 'object', 'candle', 'cashew', 'chewinggum', 'fryum', 'macaroni', 'pcb',
               'pipe fryum', 'macaroni1', 'macaroni2', 'pcb1', 'pcb2', 'pcb3', 'pcb4', 'capsules']  # adrefexpert_v2.py:38-39
MVTEC_CLASS_NAMES = ['bottle', 'cable', 'capsule', 'carpet', 'grid', 'hazelnut', 'leather', 'metal_nut', 'pill', 'screw', 'tile',
                     'toothbrush', 'transistor', 'wood', 'zipper']  # :150
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def load_and_transform_vision_data(paths):
    """ImageBind data.load_and_transform_vision_data: RGB, resize 224 (bicubic), centre crop 224, CLIP normalisation -> [n,3,224,224]."""
    from PIL import Image
    out = []
    for p in paths:
        with open(p, "rb") as fh:
            img = Image.open(fh).convert("RGB")
        w, h = img.size
        s = 224.0 / min(w, h)
        img = img.resize((max(224, round(w * s)), max(224, round(h * s))), Image.BICUBIC)
        w, h = img.size
        l, t = (w - 224) // 2, (h - 224) // 2
        img = img.crop((l, t, l + 224, t + 224))
        x = torch.frombuffer(bytearray(img.tobytes()), dtype=torch.uint8).reshape(224, 224, 3).permute(2, 0, 1).float() / 255.0
        out.append((x - torch.tensor(CLIP_MEAN).view(3, 1, 1)) / torch.tensor(CLIP_STD).view(3, 1, 1))
    return torch.stack(out)


class adrefexpert(nn.Module):
    def __init__(self, round_index=0, k_shot=1, pt="mvtec", weights=None, text_features=None, text_features_path=None, dims=None,
                 data_root="./data/TrainADDataset", device="cuda:0"):
        super().__init__()
        self.dims = dims if dims is not None else X.ExpertDims()
        self.round_index, self.k_shot = round_index, max(1, k_shot)  # :130-131
        synthetic = os.environ.get("MYRIAD_SYNTHETIC_WEIGHTS", "0") == "1"
        if weights is None:
            if not synthetic:
                raise FileNotFoundError("adrefexpert: pass weights= (imagebind_huge.pth + image_decoder.* state dict with the "
                                        "reference's key names) or set MYRIAD_SYNTHETIC_WEIGHTS=1")
            logging.warning("adrefexpert: using seeded synthetic weights (MYRIAD_SYNTHETIC_WEIGHTS=1)")
            weights = X.make_expert_state_dict(self.dims, seed=int(os.environ.get("MYRIAD_SYNTHETIC_SEED", "0")))
        if text_features is None and text_features_path is not None:
            text_features = torch.load(text_features_path, map_location="cpu")
        if text_features is None:
            if not synthetic:
                raise FileNotFoundError("adrefexpert: the zero-shot branch needs text_features= {class: [2, %d] tensor}" % self.dims.dec_dim)
            t = X.make_text_features(len(CLASS_NAMES), self.dims)
            text_features = {c: t[i] for i, c in enumerate(CLASS_NAMES)}
        self.text_features = {k.replace("_", " "): v.float() for k, v in text_features.items()}
        self._weights, self._device, self._engine = weights, device, None
        # reference-image paths as the reference lays them out (:150-160)
        self.mvtec_references = {}
        for c in MVTEC_CLASS_NAMES:
            names = [str(round_index * 4 + i).zfill(3) + ".png" for i in range(4)][:self.k_shot]
            self.mvtec_references[c] = [os.path.join(data_root, "mvtec", c, "train", "good", n) for n in names]
        self.visa_references = {}
        self._ref_bank = {}  # class -> per tapped layer unit-norm fp16 tokens [k * 256, D] (device)

    @property
    def engine(self):
        if self._engine is None:
            self._engine = X.VisionExpertEngine(self._weights, self.dims, device=self._device)
        return self._engine

    def register_references(self, cls_name, images):
        """images fp32 [k, 3, 224, 224] (already normalised): the normal references of `cls_name`."""
        refs = self.engine.encode_refs(images.to(self.engine.dev, torch.float32).contiguous(), 1)
        self._ref_bank[cls_name] = refs

    def _refs_for(self, cls_names):
        per_layer = None
        for c in cls_names:
            if c not in self._ref_bank:
                paths = self.visa_references.get(c) or self.mvtec_references.get(c)
                if not paths:
                    raise KeyError("adrefexpert: no reference images known for class %r" % c)
                self.register_references(c, load_and_transform_vision_data(paths))
            bank = self._ref_bank[c]
            per_layer = [[t] for t in bank] if per_layer is None else [a + [t] for a, t in zip(per_layer, bank)]
        rows = {t.shape[0] for t in per_layer[0]}
        assert len(rows) == 1, "every sample of a batch needs the same number of reference images"
        return [torch.cat(ts) for ts in per_layer]

    @torch.no_grad()
    def forward(self, images, cls_names, return_masks=True, querypath=None, testphase=False):
        eng = self.engine
        images = images.to(eng.dev, torch.float32).contiguous()
        if querypath:  # adrefexpert_v2.py:247-278
            _, tn = eng.trunk(images, raw=False, unit=True)
            return eng.k_shot_from_taps(tn, self._refs_for(list(cls_names)))
        text = torch.stack([self.text_features[c.replace("_", " ")] for c in cls_names])  # :72-96 as a table
        return eng.zero_shot(images, text)
