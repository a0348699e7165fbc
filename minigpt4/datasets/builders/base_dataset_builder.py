"""Dataset builder base with the reference's interface (minigpt4/datasets/builders/base_dataset_builder.py:24-236):
config resolution (`DATASET_CONFIG_DICT` default yaml, a path, or the merged node from the task), processor construction
through the registry, `build_datasets()`. The download steps of the reference (:93-166) are dropped: no network."""
import logging

from minigpt4.common import utils
from minigpt4.common.config import load_yaml
from minigpt4.common.registry import registry
from minigpt4.processors.base_processor import BaseProcessor


def load_dataset_config(cfg_path):
    datasets = load_yaml(cfg_path).datasets
    return datasets[next(iter(datasets))]


class BaseDatasetBuilder:
    train_dataset_cls, eval_dataset_cls = None, None
    DATASET_CONFIG_DICT = {}

    def __init__(self, cfg=None):
        if cfg is None:
            self.config = load_dataset_config(self.default_config_path())
        elif isinstance(cfg, str):
            self.config = load_dataset_config(cfg)
        else:
            self.config = cfg
        self.data_type = self.config.get("data_type", "images")
        self.vis_processors = {"train": BaseProcessor(), "eval": BaseProcessor()}
        self.text_processors = {"train": BaseProcessor(), "eval": BaseProcessor()}

    @classmethod
    def default_config_path(cls, type="default"):
        return utils.get_abs_path(cls.DATASET_CONFIG_DICT[type])

    @staticmethod
    def _build_proc_from_cfg(cfg):
        return registry.get_processor_class(cfg.name).from_config(cfg) if cfg is not None else None

    def build_processors(self):
        for kind, table in (("vis_processor", self.vis_processors), ("text_processor", self.text_processors)):
            cfg = self.config.get(kind)
            if cfg is not None:
                table["train"] = self._build_proc_from_cfg(cfg.get("train"))
                table["eval"] = self._build_proc_from_cfg(cfg.get("eval"))

    def build_datasets(self):
        logging.info("Building datasets...")
        return self.build()

    def build(self):
        raise NotImplementedError("%s.build" % type(self).__name__)
