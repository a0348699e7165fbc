"""Blip2Base with the part of the reference's interface (minigpt4/models/blip2.py:32-125) that the Myriad callers use:
`init_tokenizer`, `maybe_autocast`, plus the module-level `LayerNorm` and `disabled_train`. The reference's
`init_vision_encoder` / `init_Qformer` / `load_from_pretrained` build nn.Modules; here the vision encoder and the Q-Former
are weight sets executed by myriad_b200.engine.MyriadEngine and are loaded by minigpt4/models/checkpoints.py, so those three
methods do not exist."""
import contextlib
import os

import torch
import torch.nn as nn

from minigpt4.models.base_model import BaseModel


def disabled_train(self, mode=True):
    """frozen sub-modules never leave eval mode (blip2.py:113-116)"""
    return self


class LayerNorm(nn.LayerNorm):
    """fp32 LayerNorm regardless of input dtype (blip2.py:119-125); the device path is myr_norm_fwd."""

    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class Blip2Base(BaseModel):
    @classmethod
    def init_tokenizer(cls):
        root = "./pretrained_models/huggingface/models--bert-base-uncased"
        if os.path.isdir(root):
            from transformers import BertTokenizer
            tok = BertTokenizer.from_pretrained(root)
            tok.add_special_tokens({"bos_token": "[DEC]"})
            return tok
        return None  # only used by the (deleted) Q-Former text branch; Myriad never tokenises with it

    def maybe_autocast(self, dtype=torch.float16):
        # the kernels fix their own precision contract (fp16 operands, fp32 accumulate); kept for API compatibility
        if self.device != torch.device("cpu"):
            return torch.autocast("cuda", dtype=dtype)
        return contextlib.nullcontext()
