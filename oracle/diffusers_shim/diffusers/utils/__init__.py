"""diffusers.utils symbols the reference imports (BaseOutput, logging, deprecate, USE_PEFT_BACKEND)."""
import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields

USE_PEFT_BACKEND = False


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _pylogging.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    return None
