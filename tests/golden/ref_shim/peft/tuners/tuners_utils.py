import warnings


class BaseTunerLayer:
    active_adapter = None
    adapter_layer_names = ()
    other_param_names = ()
    _disable_adapters = False
    _active_adapter = "default"
    merged_adapters = []

    def get_base_layer(self):
        base = self
        while hasattr(base, "base_layer"):
            base = base.base_layer
        return base

    @property
    def weight(self):
        return self.get_base_layer().weight

    @property
    def bias(self):
        return self.get_base_layer().bias

    @property
    def merged(self):
        return bool(self.merged_adapters)

    @property
    def disable_adapters(self):
        return self._disable_adapters

    @property
    def active_adapter(self):
        return self._active_adapter

    @property
    def active_adapters(self):
        a = self._active_adapter
        return [a] if isinstance(a, str) else a

    def enable_adapters(self, enabled):
        self._disable_adapters = not enabled

    def set_adapter(self, adapter_names):
        if isinstance(adapter_names, str):
            adapter_names = [adapter_names]
        for layer_name in self.adapter_layer_names:
            for key, layer in getattr(self, layer_name).items():
                layer.requires_grad_(key in adapter_names)
        self._active_adapter = adapter_names


def check_adapters_to_merge(module, adapter_names=None):
    if adapter_names is None:
        adapter_names = module.active_adapters
    if module.merged:
        merged = set(module.merged_adapters)
        adapter_names = [n for n in adapter_names if n not in merged]
        if not adapter_names:
            warnings.warn("All adapters are already merged, nothing to do.")
    return adapter_names
