"""ConfigMixin / register_to_config (import sites: unet_1d_condition.py:18, transformer_1d.py:9).

Semantics kept: every ctor argument (defaults + passed) is recorded on `self.config`, readable as
attributes (`self.config.center_input_sample`, unet_1d_condition.py:602).
"""
import functools
import inspect


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kw):
        cfg = getattr(self, "_cfg", None)
        if cfg is None:
            cfg = _Config()
            object.__setattr__(self, "_cfg", cfg)
        cfg.update(kw)

    @property
    def config(self):
        return self._cfg


def register_to_config(init):
    sig = inspect.signature(init)

    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        vals = {k: v for k, v in bound.arguments.items() if k != "self" and not k.startswith("_")}
        ConfigMixin.register_to_config(self, **vals)
        init(self, *args, **kwargs)

    return wrapped
