"""`datasets: anomaly_detection:` of the finetune yamls (reference
minigpt4/datasets/builders/anomaly_detection_builder.py:11-52). `build_info.storage: synthetic` (or
MYRIAD_SYNTHETIC_DATA=1) builds the file-free MVTec-shaped set (`build_info.synthetic_len` items)."""
import os
import warnings

from minigpt4.common.registry import registry
from minigpt4.datasets.builders.base_dataset_builder import BaseDatasetBuilder
from minigpt4.datasets.datasets.anomaly_detection import AnomalyDetectionDataset


@registry.register_builder("anomaly_detection")
class AnomalyDetectionBuilder(BaseDatasetBuilder):
    train_dataset_cls = AnomalyDetectionDataset
    DATASET_CONFIG_DICT = {"default": "configs/datasets/anomaly_detection/base.yaml"}

    def build(self):
        self.build_processors()
        cfg, info = self.config, self.config.build_info
        storage = "synthetic" if os.environ.get("MYRIAD_SYNTHETIC_DATA", "0") == "1" else info.storage
        if storage != "synthetic" and not os.path.exists(storage):
            warnings.warn("storage path {} does not exist.".format(storage))
        train = self.train_dataset_cls(
            vis_processor=self.vis_processors["train"], text_processor=self.text_processors["train"],
            ann_paths=list(info.get("ann_paths", ["DC_VISA_train_normal.jsonl"])), img_size=cfg.get("img_size", 224),
            crop_size=cfg.get("crop_size", 224), vis_root=storage, ve_root=info.get("ve_storage", ""), version=cfg.get("version", 0),
            with_mask=cfg.get("with_mask", False), with_pos=cfg.get("with_pos", False), with_ref=cfg.get("with_ref", False),
            is_preload=cfg.get("is_preload", False), nsa_max_width=(cfg.get("augment") or {}).get("nsa_max_width", 0.4),
            synthetic_len=info.get("synthetic_len", 64))
        return {"train": train}
