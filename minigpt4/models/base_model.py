"""BaseModel with the reference's interface (minigpt4/models/base_model.py:19-118): `.device`, `from_config`,
`default_config_path`, `load_checkpoint`, `show_n_params`. The contrastive-learning helpers of the reference file
(GatherLayer, concat_all_gather, ...) are never called by Myriad and are out of scope."""
import logging
import os

import torch
import torch.nn as nn

from minigpt4.common.registry import registry


class BaseModel(nn.Module):
    PRETRAINED_MODEL_CONFIG_DICT = {}

    def __init__(self):
        super().__init__()
        self._device_hint = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    @property
    def device(self):
        params = list(self.parameters())
        return params[0].device if params else self._device_hint

    def load_checkpoint(self, url_or_filename):
        if not os.path.isfile(url_or_filename):
            raise RuntimeError("checkpoint url or path is invalid")
        checkpoint = torch.load(url_or_filename, map_location="cpu")
        state_dict = checkpoint["model"] if "model" in checkpoint else checkpoint
        msg = self.load_state_dict(state_dict, strict=False)
        logging.info("load checkpoint from %s", url_or_filename)
        return msg

    @classmethod
    def default_config_path(cls, model_type):
        assert model_type in cls.PRETRAINED_MODEL_CONFIG_DICT, "Unknown model type {}".format(model_type)
        return os.path.join(registry.get_path("library_root"), cls.PRETRAINED_MODEL_CONFIG_DICT[model_type])

    @classmethod
    def from_pretrained(cls, model_type):
        from minigpt4.common.config import load_yaml
        return cls.from_config(load_yaml(cls.default_config_path(model_type)).model)

    def before_evaluation(self, **kwargs):
        pass

    def show_n_params(self, return_str=True):
        tot = sum(p.numel() for p in self.parameters())
        if not return_str:
            return tot
        return "{:.1f}M".format(tot / 1e6) if tot >= 1e6 else "{:.1f}K".format(tot / 1e3)
