"""`from minigpt4.datasets.builders import *` (train.py:28): importing registers the builders. Only the anomaly-detection
builder is on the Myriad path; the CC-SBU / LAION / PandaGPT builders of the reference belong to MiniGPT-4 pretraining."""
from minigpt4.common.registry import registry
from minigpt4.datasets.builders.anomaly_detection_builder import AnomalyDetectionBuilder
from minigpt4.datasets.builders.base_dataset_builder import BaseDatasetBuilder, load_dataset_config

__all__ = ["AnomalyDetectionBuilder"]


def load_dataset(name, cfg_path=None, vis_path=None, data_type=None):
    builder_cls = registry.get_builder_class(name)
    if builder_cls is None:
        raise KeyError("Dataset %s not found. Available datasets: %s" % (name, ", ".join(registry.list_datasets())))
    builder = builder_cls(load_dataset_config(cfg_path) if cfg_path is not None else None)
    if vis_path is not None:
        builder.config.build_info.storage = vis_path
    return builder.build_datasets()
