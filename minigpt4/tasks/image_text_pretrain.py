"""`run.task: image_text_pretrain` (reference minigpt4/tasks/image_text_pretrain.py:12-18): training only, no evaluation."""
from minigpt4.common.registry import registry
from minigpt4.tasks.base_task import BaseTask


@registry.register_task("image_text_pretrain")
class ImageTextPretrainTask(BaseTask):
    def evaluation(self, model, data_loader, cuda_enabled=True):
        return None
