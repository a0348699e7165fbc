"""`from minigpt4.runners import *` (train.py:31): importing registers `runner_base`."""
from minigpt4.runners.runner_base import RunnerBase

__all__ = ["RunnerBase"]
