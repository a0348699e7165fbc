"""BaseProcessor with the reference's interface (minigpt4/processors/base_processor.py:11-26): identity transform,
`from_config(cfg)`, `build(**kwargs)`."""
from minigpt4.common.config import Node


class BaseProcessor:
    def __init__(self):
        self.transform = lambda x: x

    def __call__(self, item):
        return self.transform(item)

    @classmethod
    def from_config(cls, cfg=None):
        return cls()

    def build(self, **kwargs):
        return self.from_config(Node.wrap(kwargs))
