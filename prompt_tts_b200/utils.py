"""`BaseOutput` contract of the reference (tts/utils.py:15-83): an ordered dict whose keys are also attributes,
so `model(...).sample` (train.py:105) and tuple-style indexing both work."""
from collections import OrderedDict
from typing import Any, Tuple


class BaseOutput(OrderedDict):
    def __init__(self, **kwargs):
        super().__init__()
        for k, v in kwargs.items():
            self[k] = v

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)

    def __setattr__(self, name, value):
        if name in self.keys() and value is not None:
            super().__setitem__(name, value)
        super().__setattr__(name, value)

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self) -> Tuple[Any]:
        return tuple(self[k] for k in self.keys())


class Config(dict):
    """attribute-accessible config (diffusers `register_to_config` contract: `self.config.<ctor arg>`)."""
    __getattr__ = dict.__getitem__
