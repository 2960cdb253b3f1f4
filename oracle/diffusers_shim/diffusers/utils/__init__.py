"""BaseOutput / deprecate / logging (import sites: unet_1d_condition.py:19, transformer_1d.py:8)."""
import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass


class BaseOutput(OrderedDict):
    """Ordered dict whose keys are mirrored to attributes (contract documented in the reference's
    own copy, tts/utils.py:15-83).  `UNet1DConditionOutput(sample=...)` is NOT a dataclass in the
    reference (unet_1d_condition.py:28) so construction goes through OrderedDict(**kw) +
    __setitem__; `Transformer1DModelOutput` is a dataclass and goes through __post_init__."""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        if name in self.keys() and value is not None:
            super().__setitem__(name, value)
        super().__setattr__(name, value)

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


def deprecate(*args, **kwargs):
    return None


class logging:  # noqa: N801  (used as a module-like namespace: `logging.get_logger(__name__)`)
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)
