"""`AnomalyDetectionDataset`: the training set of the Myriad finetune (reference
minigpt4/datasets/datasets/anomaly_detection.py:97-363; evaluation_aqa_dataset.py:22 imports the class too).

Every item is one NORMAL industrial image plus, in the train stage, a synthetic-anomaly copy of it (NSA cut-paste from
another random image, datasets/self_sup_tasks.py) — so a loader batch of b items becomes 2b images inside
`Myriad.prepare_sample` (runner_base.py:546-549 halves the configured batch size for this dataset). Keys of an item
(reference :342-362): image, scene, question / question2 / question3 (the same constant prompt containing `<ImageHere>`),
text_input, image_id, is_anomaly, img_path and, for training, aug_image and aug_text_input.

Annotation files are jsonl (`img_path`, `ve_path`, `caption`, `is_anomaly`), read with the json module.

Synthetic mode (`vis_root` == "synthetic", or MYRIAD_SYNTHETIC_DATA=1): no files are touched; images are seeded random
textures in MVTec shape and each item also carries the vision-expert outputs (`anomaly_maps`, `oneshot_anomaly_maps`,
`aug_anomaly_maps`, `aug_oneshot_anomaly_maps`, fp32 [1, 224, 224]) which `Myriad.prepare_sample` consumes when no expert
module is attached (the ImageBind expert is outside the built path)."""
import json
import os

import cv2
import numpy as np
import torch
from PIL import Image

from minigpt4.datasets.datasets.base_dataset import BaseDataset, imreader as _image_reader
from minigpt4.datasets.self_sup_tasks import patch_ex

QuestionPrompts = [
    "This image may be simulated by photo editing. According on IAD expert opinions, find out if there are defects in this image.",
    "This image may be simulated by photo editing. According to IAD expert opinions and corresponding visual descriptions, find out if there are defects in this image.",
    "This image may be simulated by photo editing. According to IAD expert visual descriptions, find out if there are defects in this image.",
]

_W = (0.03, 0.4)
# per-class NSA bounds (patch half-widths as fractions of the image side), logistic label parameters (k, x0) and
# background levels (grey level, tolerance) of the MVTec classes — data of the reference (:44-62), not code
MVTEC_WIDTH_BOUNDS_PCT = {
    "bottle": (_W, _W), "cable": ((0.05, 0.4), (0.05, 0.4)), "capsule": ((0.03, 0.15), _W), "hazelnut": ((0.03, 0.35), (0.03, 0.35)),
    "metal_nut": (_W, _W), "pill": ((0.03, 0.2), _W), "screw": ((0.03, 0.12), (0.03, 0.12)), "toothbrush": (_W, (0.03, 0.2)),
    "transistor": (_W, _W), "zipper": (_W, (0.03, 0.2)), "carpet": (_W, _W), "grid": (_W, _W), "leather": (_W, _W), "tile": (_W, _W),
    "wood": (_W, _W)}
MVTEC_INTENSITY_LOGISTIC_PARAMS = {
    "bottle": (1 / 12, 24), "cable": (1 / 12, 24), "capsule": (1 / 2, 4), "hazelnut": (1 / 12, 24), "metal_nut": (1 / 3, 7),
    "pill": (1 / 3, 7), "screw": (1, 3), "toothbrush": (1 / 6, 15), "transistor": (1 / 6, 15), "zipper": (1 / 6, 15), "carpet": (1 / 3, 7),
    "grid": (1 / 3, 7), "leather": (1 / 3, 7), "tile": (1 / 3, 7), "wood": (1 / 6, 15)}
MVTEC_BACKGROUND = {"bottle": (200, 60), "screw": (200, 60), "capsule": (200, 60), "zipper": (200, 60), "hazelnut": (20, 20),
                    "pill": (20, 20), "toothbrush": (20, 20), "metal_nut": (20, 20)}

_GRID_NAMES = (("top left", "top", "top right"), ("left", "center", "right"), ("bottom left", "bottom", "bottom right"))


def get_position(centers):
    """3 x 3 grid names of (x, y) points on the 224 canvas; first coordinate picks the grid row (reference :65-92)."""
    cell = lambda v: 0 if v <= 1 / 3 else (1 if v <= 2 / 3 else 2)
    return list({_GRID_NAMES[cell(c[0] / 224)][cell(c[1] / 224)] for c in centers})


def imreader(args):
    """preload worker that also decodes the stored vision-expert map (reference :38-43)"""
    _image_reader(args)
    i, records = args
    records[i]["ve"] = cv2.imread(records[i]["ve_path"])


def _resize_center_crop(img, size, crop):
    w, h = img.size
    s = size / min(w, h)
    img = img.resize((max(size, round(w * s)), max(size, round(h * s))), Image.BICUBIC)
    w, h = img.size
    x0, y0 = (w - crop) // 2, (h - crop) // 2
    return img.crop((x0, y0, x0 + crop, y0 + crop))


class AnomalyDetectionDataset(BaseDataset):
    DatasetName = "AnomalyDetection"

    def __init__(self, vis_processor, text_processor, vis_root, ve_root, ann_paths, img_size=224, crop_size=224, version=0,
                 with_mask=False, with_ref=False, with_pos=False, is_preload=False, stage="train", nsa_max_width=0.4,
                 synthetic_len=64, synthetic_seed=0):
        self.version, self.with_mask, self.with_ref, self.with_pos = version, with_mask, with_ref, with_pos
        self.ve_root, self.stage = ve_root, stage
        self.img_size, self.crop_size = img_size, crop_size
        self.synthetic = vis_root == "synthetic" or os.environ.get("MYRIAD_SYNTHETIC_DATA", "0") == "1"
        self.synthetic_len, self.synthetic_seed = synthetic_len, synthetic_seed
        nsa = {"num_patches": 2, "min_object_pct": 0, "min_overlap_pct": 0.25, "gamma_params": (2, 0.05, 0.03), "resize": True,
               "shift": True, "same": False, "mode": cv2.NORMAL_CLONE, "label_mode": "logistic-intensity"}
        if "VISA" in ann_paths[0]:  # VisA has no per-class table: one setting for every class (reference :110-124)
            nsa.update({"width_bounds_pct": ((0.03, nsa_max_width), (0.03, nsa_max_width)), "intensity_logistic_params": (1 / 12, 24),
                        "skip_background": None, "resize_bounds": (0.5, 2)})
        self.self_sup_args = nsa
        self.transform = lambda im: _resize_center_crop(im, self.img_size, self.crop_size)
        super().__init__(vis_processor, text_processor, vis_root, ann_paths, is_preload and not self.synthetic,
                         preload_fn=imreader if with_mask else _image_reader)

    # ------------------------------------------------------------------------------------------ storage
    def load_annotations(self):
        if self.synthetic:
            classes = sorted(MVTEC_WIDTH_BOUNDS_PCT)
            self.annotation = [{"img_path": "mvtec/%s/train/good/%03d.png" % (classes[i % len(classes)], i), "ve_path": "", "caption": "",
                                "is_anomaly": "0"} for i in range(self.synthetic_len)]
            return
        self.annotation = []
        for rel in self.ann_paths:
            with open(os.path.join(self.vis_root, rel)) as fh:
                self.annotation.extend(json.loads(line) for line in fh if line.strip())
        print(f"In {self.DatasetName} Dataset, Has Samples: {len(self.annotation)}")

    def construct_preload_maps(self):
        return [{"path": self.get_image_path(a["img_path"]), "rel_path": a["img_path"], "ve_path": self.get_ve_path(a["ve_path"])}
                for a in self.annotation]

    def post_preload(self, results):
        super().post_preload(results)
        if self.with_mask:
            self._ve_cache = {rec["rel_path"]: rec["ve"] for rec in results}

    def get_ve_path(self, ve_path):
        return os.path.join(self.ve_root, ve_path)

    def get_class_name(self, index):
        return ("mvtec" if "MVTEC" in self.ann_paths[0] or self.synthetic else "visa"), self.annotation[index]["img_path"].split("/")[1]

    def _synthetic_image(self, index):
        """smooth seeded texture on a flat background: enough structure for the NSA blend to change pixels"""
        g = np.random.RandomState(self.synthetic_seed * 100003 + index)
        low = g.randint(0, 255, (8, 8, 3)).astype(np.uint8)
        img = cv2.resize(low, (self.crop_size, self.crop_size), interpolation=cv2.INTER_CUBIC)
        noise = g.randint(0, 24, img.shape).astype(np.int16)
        return Image.fromarray(np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8))

    def prepare_img(self, index):
        if self.synthetic:
            return self._synthetic_image(index)
        rel = self.annotation[index]["img_path"]
        if self.is_preload:
            return self._cache[rel].copy()
        return Image.open(self.get_image_path(rel)).convert("RGB")

    def prepare_ve(self, index):
        ann = self.annotation[index]
        return self._ve_cache[ann["img_path"]].copy() if self.is_preload else cv2.imread(self.get_ve_path(ann["ve_path"]))

    # -------------------------------------------------------------------------------------------- items
    def _simulate_anomaly(self, index, image):
        other = np.random.randint(len(self))
        while other == index and len(self) > 1:
            other = np.random.randint(len(self))
        src = np.asarray(self.transform(self.prepare_img(other)))
        ds, cls = self.get_class_name(index)
        args = dict(self.self_sup_args)
        if ds == "mvtec":
            args.update({"width_bounds_pct": MVTEC_WIDTH_BOUNDS_PCT.get(cls), "intensity_logistic_params": MVTEC_INTENSITY_LOGISTIC_PARAMS.get(cls),
                         "skip_background": MVTEC_BACKGROUND.get(cls)})
            if self.synthetic:
                args["skip_background"] = None
        dest = np.asarray(image)
        for _ in range(50):
            aug, mask, boxes = patch_ex(dest, src, verbose=False, **args)
            if mask.sum() > 0:
                break
        return aug, mask, boxes

    def _expert_maps(self, label=None):
        """stand-in for the vision experts' outputs in synthetic mode: low-amplitude noise, plus the NSA label where pasted"""
        base = torch.rand(1, 224, 224) * 0.2
        if label is not None:
            lab = torch.from_numpy(np.ascontiguousarray(label[..., 0])).float()
            if lab.shape != (224, 224):
                lab = torch.nn.functional.interpolate(lab[None, None], size=(224, 224), mode="bilinear", align_corners=False)[0, 0]
            base = (base + 0.8 * lab[None]).clamp(0, 1)
        return base

    def __getitem__(self, index):
        ann = self.annotation[index]
        image = self.transform(self.prepare_img(index))
        normal_describe = "No, there exists no anomalies in the image."
        if self.version == 0:
            abnormal_describe = "Yes, there exists anomalies in the image."
        elif self.version == 1:
            abnormal_describe = "Yes, there exists anomalies in the image. These anomalies are simulated by photo editing."
        else:
            raise ValueError("Not Support Version.%s" % self.version)
        item = {
            "image": self.vis_processor({"img": np.asarray(image)})["img"],
            "scene": ann["img_path"].split("/")[1],
            "question": "<Img><ImageHere></Img>" + QuestionPrompts[1],
            "question2": "<Img><ImageHere></Img>" + QuestionPrompts[1],
            "question3": "<Img><ImageHere></Img>" + QuestionPrompts[1],
            "text_input": normal_describe,
            "image_id": index,
            "is_anomaly": ann["is_anomaly"] == "1",
            "img_path": os.path.join(self.vis_root or "", ann["img_path"]),
        }
        label = None
        if self.stage == "train":
            aug, label, _ = self._simulate_anomaly(index, image)
            aug_sample = self.vis_processor({"img": aug, "gt_seg_map": label})
            item["aug_image"] = aug_sample["img"]
            item["aug_text_input"] = normal_describe if np.sum(aug_sample["gt_seg_map"]) == 0.0 else abnormal_describe
            label = aug_sample["gt_seg_map"]
        if self.synthetic:
            item["anomaly_maps"], item["oneshot_anomaly_maps"] = self._expert_maps(), self._expert_maps()
            if self.stage == "train":
                item["aug_anomaly_maps"], item["aug_oneshot_anomaly_maps"] = self._expert_maps(label), self._expert_maps(label)
        return item

    def __repr__(self):
        return "%s(n=%d, root=%s, synthetic=%s, with_mask=%s)" % (self.DatasetName, len(self), self.vis_root, self.synthetic, self.with_mask)
