"""ConfigMixin / register_to_config with the 0.27.2 semantics the reference relies on: the decorator records the
constructor arguments in ``self.config`` BEFORE running ``__init__``; unknown attributes fall through to the config
(quirk D4: the scheduler reads ``self.use_karras_sigmas`` before assigning it)."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = None
    ignore_for_config = []

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        cur = dict(getattr(self, "_internal_dict", {}))
        cur.update(kwargs)
        object.__setattr__(self, "_internal_dict", FrozenDict(cur))

    @property
    def config(self):
        return self._internal_dict

    def __getattr__(self, name):
        d = self.__dict__.get("_internal_dict")
        if d is not None and name in d and name not in self.__dict__:
            return d[name]
        sup = super()
        if hasattr(sup, "__getattr__"):          # nn.Module parameters / buffers / submodules
            return sup.__getattr__(name)
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        ignore = getattr(self, "ignore_for_config", [])
        params = {n: p.default for i, (n, p) in enumerate(inspect.signature(init).parameters.items())
                  if i > 0 and n not in ignore}
        new = {}
        for arg, name in zip(args, params.keys()):
            new[name] = arg
        new.update({k: init_kwargs.get(k, d) for k, d in params.items() if k not in ignore and k not in new})
        self.register_to_config(**new)
        init(self, *args, **init_kwargs)
    return inner_init
