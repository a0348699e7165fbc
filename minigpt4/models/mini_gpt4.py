"""`arch: mini_gpt4` — plain MiniGPT-4 (reference minigpt4/models/mini_gpt4.py:153-257): the Myriad pipeline without
the expert-prior tokens (no adaptor residual, no VEInstructor / VETokenizer). It falls out of the same kernels:
encode_img runs stage 3 (no expert tokens) with a zero adaptor."""
import torch

from minigpt4.common.registry import registry
from minigpt4.models.myriad import Myriad


@registry.register_model("mini_gpt4")
class MiniGPT4(Myriad):
    PRETRAINED_MODEL_CONFIG_DICT = {"pretrain_vicuna": "configs/models/minigpt4.yaml"}
    GENERATE_STAGE = 3  # no expert tokens: [32 Q-Former queries] only

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        with torch.no_grad():
            self.expert_adaptor.conv2.weight.zero_()  # x + W2(W1 x) == x
        for p in self.parameters():
            p.requires_grad = False

    def encode_img(self, image, maps=None, stage=3):
        """mini_gpt4.py:153-178: (inputs_llama, atts_llama) from the 32 Q-Former queries only."""
        if maps is None:
            maps = torch.zeros(image.shape[0], 1, 224, 224, device=image.device)
        return super().encode_img(image, maps, 3)

    def prepare_sample(self, samples, stage):
        image = samples["image"]
        z = torch.zeros(image.shape[0], 1, 224, 224, device=image.device)
        q = samples.get("question", None)
        return image, q, samples.get("text_input", None), z, z

    def forward(self, samples):
        raise NotImplementedError("plain MiniGPT-4 trains llama_proj (mini_gpt4.py:180-257), which the Myriad finetune keeps frozen; "
                                  "only its inference path (encode_img / generate / Chat) is built on this hot path")
