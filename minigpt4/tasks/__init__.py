"""`tasks.setup_task(cfg)` as train.py:98 calls it (reference minigpt4/tasks/__init__.py:13-20)."""
from minigpt4.common.registry import registry
from minigpt4.tasks.base_task import BaseTask
from minigpt4.tasks.image_text_pretrain import ImageTextPretrainTask

__all__ = ["BaseTask", "ImageTextPretrainTask"]


def setup_task(cfg):
    assert "task" in cfg.run_cfg, "Task name must be provided."
    name = cfg.run_cfg.task
    task_cls = registry.get_task_class(name)
    assert task_cls is not None, "Task {} not properly registered.".format(name)
    return task_cls.setup_task(cfg=cfg)
