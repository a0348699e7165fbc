"""`from minigpt4.models import *` as train.py:29 / evaluation_aqa_dataset.py:31 do. The reference's version of this
file imports seven modules that are not in its repository (models/__init__.py:18-27); this one exports the classes
that exist on the hot path plus `load_model`."""
import torch

from minigpt4.common.registry import registry
from minigpt4.models.base_model import BaseModel
from minigpt4.models.blip2 import Blip2Base
from minigpt4.models.mini_gpt4 import MiniGPT4
from minigpt4.models.myriad import Myriad

__all__ = ["load_model", "BaseModel", "Blip2Base", "MiniGPT4", "Myriad"]


def load_model(name, model_type, is_eval=False, device="cpu", checkpoint=None):
    """reference models/__init__.py:45-80 (the `.float()` for CPU is dropped: there is no CPU path)."""
    model = registry.get_model_class(name).from_pretrained(model_type=model_type)
    if checkpoint is not None:
        model.load_checkpoint(checkpoint)
    if is_eval:
        model.eval()
    return model.to(device)
