"""`from minigpt4.processors import *` (train.py:30, evaluation_aqa_dataset.py:32): importing registers the processors."""
from minigpt4.common.registry import registry
from minigpt4.processors.base_processor import BaseProcessor
from minigpt4.processors.blip_processors import (Blip2ImageEvalProcessor, Blip2ImageTrainProcessor, BlipCaptionProcessor,
                                                  LocImageTrainProcessor)

__all__ = ["BaseProcessor", "Blip2ImageTrainProcessor", "Blip2ImageEvalProcessor", "BlipCaptionProcessor", "LocImageTrainProcessor"]


def load_processor(name, cfg=None):
    return registry.get_processor_class(name).from_config(cfg)
