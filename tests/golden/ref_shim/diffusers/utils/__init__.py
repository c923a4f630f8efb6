import logging as _pylog
from collections import OrderedDict
from dataclasses import fields, is_dataclass


class BaseOutput(OrderedDict):
    """dataclass-style output that also indexes like a tuple / dict (diffusers.utils.BaseOutput)."""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    super().__setitem__(f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylog.getLogger(name)


logging = _Logging()
