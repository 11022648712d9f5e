"""ConfigMixin / register_to_config / FrozenDict restated (diffusers 0.23.0 configuration_utils)."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(name) from e


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kwargs):
        self._internal_dict = FrozenDict(kwargs)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        merged = {k: v for k, v in dict(config).items() if k in sig}
        merged.update(kwargs)
        return cls(**merged)


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for p in sig.parameters.values() if p.name != "self"]
        cfg = {p.name: p.default for p in params}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update(kwargs)
        init(self, *args, **kwargs)
        self.register_to_config(**cfg)

    return inner_init
