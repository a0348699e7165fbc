"""Drop-in `minigpt4` package for the Myriad hot path (B200-native).

The reference package does not import as shipped (its models/__init__.py:18-27 imports modules that are not in
the repository); this package exposes the same names — `minigpt4.common.registry.registry`,
`minigpt4.common.config.Config`, `minigpt4.models.{BaseModel, Blip2Base, Myriad, MiniGPT4, load_model}`,
`minigpt4.conversation.conversation.{Conversation, CONV_VISION, StoppingCriteriaSub, Chat}` — so `train.py`,
`evaluation_aqa_dataset.py` and `eval_configs/myriad.yaml` resolve `arch: myriad` to the CUDA implementation in
`myriad_b200`. Only the hot path is implemented; datasets / processors / runners are out of scope.
"""
import os

from minigpt4.common.registry import registry

root_dir = os.path.dirname(os.path.abspath(__file__))
repo_root = os.path.join(root_dir, "..")
if registry.get_path("library_root") is None:
    registry.register_path("library_root", root_dir)
    registry.register_path("repo_root", repo_root)
    registry.register_path("cache_root", os.path.join(repo_root, ".cache"))
registry.register("MAX_INT", 2 ** 31 - 1) if registry.get("MAX_INT", no_warning=True) is None else None
registry.register("SPLIT_NAMES", ["train", "val", "test"]) if registry.get("SPLIT_NAMES", no_warning=True) is None else None
