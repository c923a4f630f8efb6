"""`from peft.tuners.lora.layer import Linear, BaseTunerLayer` (patch/patch.py:24): the reference vendors the same
layer as models/lora_layer.py, so the shim hands that file back."""
from peft.tuners.tuners_utils import BaseTunerLayer  # noqa: F401
from models.lora_layer import Linear  # noqa: F401
