"""Image / caption processors registered under the reference's names (minigpt4/processors/blip_processors.py):
`blip_caption` :32-72, `blip2_image_train` :75-117, `loc_image_train` :120-191, `blip2_image_eval` :194-221.

Images leave a processor as CLIP-normalised fp32 CHW tensors (mean / std of :24-26) — the `image` the hot path's
PatchEmbed consumes. Implemented with torch ops on the decoded array; `loc_image_train` keeps the
shipped configuration (`identity: True`, loraadapter_simple_myriad_finetune.yaml) and the resize-shortest-edge + random
crop geometry, written here directly because the reference's mmdet transforms are not installable offline;
`strong_aug` is rejected."""
import re

import numpy as np
import torch
import torch.nn.functional as F

from minigpt4.common.registry import registry
from minigpt4.processors.base_processor import BaseProcessor

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _to_chw_float(img):
    """PIL image / HWC uint8 array / CHW float tensor -> CHW fp32 in [0, 1]."""
    if isinstance(img, torch.Tensor):
        return img.float()
    a = np.asarray(img)
    if a.ndim == 2:
        a = a[:, :, None].repeat(3, 2)
    t = torch.from_numpy(np.array(a)).permute(2, 0, 1).float()
    return t / 255.0 if a.dtype == np.uint8 else t


class BlipImageBaseProcessor(BaseProcessor):
    def __init__(self, mean=None, std=None):
        super().__init__()
        m = torch.tensor(CLIP_MEAN if mean is None else tuple(mean), dtype=torch.float32)
        s = torch.tensor(CLIP_STD if std is None else tuple(std), dtype=torch.float32)
        self._mean, self._std = m.view(3, 1, 1), s.view(3, 1, 1)

    def normalize(self, chw):
        return (chw - self._mean) / self._std

    def denormalize(self, chw):
        return chw * self._std + self._mean

    def _resize_normalize(self, item, size):
        x = _to_chw_float(item)
        if x.shape[-2:] != (size, size):
            x = F.interpolate(x[None], size=(size, size), mode="bicubic", align_corners=False, antialias=True)[0].clamp(0, 1)
        return self.normalize(x)


@registry.register_processor("blip_caption")
class BlipCaptionProcessor(BaseProcessor):
    _PUNCT, _SPACES = re.compile(r"([.!\"()*#:;~])"), re.compile(r"\s{2,}")

    def __init__(self, prompt="", max_words=50):
        super().__init__()
        self.prompt, self.max_words = prompt, max_words

    def pre_caption(self, caption):
        words = self._SPACES.sub(" ", self._PUNCT.sub(" ", caption.lower())).rstrip("\n").strip(" ").split(" ")
        return " ".join(words[:self.max_words])

    def __call__(self, caption):
        return self.prompt + self.pre_caption(caption)

    @classmethod
    def from_config(cls, cfg=None):
        cfg = cfg or {}
        return cls(prompt=cfg.get("prompt", ""), max_words=cfg.get("max_words", 50))


class _ResizeProcessor(BlipImageBaseProcessor):
    def __init__(self, image_size=224, mean=None, std=None, **_):
        super().__init__(mean, std)
        self.image_size = image_size
        self.transform = lambda item: self._resize_normalize(item, self.image_size)

    @classmethod
    def from_config(cls, cfg=None):
        cfg = cfg or {}
        return cls(image_size=cfg.get("image_size", 224), mean=cfg.get("mean", None), std=cfg.get("std", None),
                   min_scale=cfg.get("min_scale", 0.5), max_scale=cfg.get("max_scale", 1.0))


@registry.register_processor("blip2_image_train")
class Blip2ImageTrainProcessor(_ResizeProcessor):
    pass


@registry.register_processor("blip2_image_eval")
class Blip2ImageEvalProcessor(_ResizeProcessor):
    pass


@registry.register_processor("loc_image_train")
class LocImageTrainProcessor(BlipImageBaseProcessor):
    """dict in ({'img': HWC uint8, optional 'gt_seg_map': HW}) -> dict out with 'img' a normalised CHW tensor."""

    def __init__(self, image_size=224, mean=None, std=None, min_scale=0.5, max_scale=1.0, strong_aug=False, identity=False,
                 debug_mode=False):
        super().__init__(mean, std)
        if strong_aug:
            raise NotImplementedError("loc_image_train: strong_aug needs mmdet transforms (not on the hot path)")
        self.image_size, self.identity, self.debug_mode = image_size, identity, debug_mode

    def _geometry(self, sample):
        """resize the shortest edge to image_size, then a random image_size crop; the mask follows the image."""
        img = np.asarray(sample["img"])
        h, w = img.shape[:2]
        s = self.image_size
        scale = s / min(h, w)
        nh, nw = max(s, round(h * scale)), max(s, round(w * scale))
        t = torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1)[None].float()
        t = F.interpolate(t, size=(nh, nw), mode="bilinear", align_corners=False)[0]
        y0, x0 = np.random.randint(0, nh - s + 1), np.random.randint(0, nw - s + 1)
        out = dict(sample)
        out["img"] = t[:, y0:y0 + s, x0:x0 + s].round().clamp(0, 255).byte().permute(1, 2, 0).numpy()
        if "gt_seg_map" in sample:
            m = torch.from_numpy(np.ascontiguousarray(sample["gt_seg_map"]))[None, None].float()
            out["gt_seg_map"] = F.interpolate(m, size=(nh, nw), mode="nearest")[0, 0, y0:y0 + s, x0:x0 + s].numpy()
        return out

    def __call__(self, data_sample):
        ret = dict(data_sample) if self.identity else self._geometry(data_sample)
        if not self.debug_mode:
            ret["img"] = self.normalize(_to_chw_float(ret["img"]))
        if "gt_bboxes" in ret:
            ret["gt_bboxes"] = np.asarray(ret["gt_bboxes"]).tolist()
        return ret

    @classmethod
    def from_config(cls, cfg=None):
        cfg = cfg or {}
        return cls(image_size=cfg.get("image_size", 224), mean=cfg.get("mean", None), std=cfg.get("std", None),
                   min_scale=cfg.get("min_scale", 0.5), max_scale=cfg.get("max_scale", 1.0), strong_aug=cfg.get("strong_aug", False),
                   identity=cfg.get("identity", False))
