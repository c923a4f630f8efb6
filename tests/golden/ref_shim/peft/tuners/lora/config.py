from dataclasses import dataclass


@dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: int = 8
    lora_dropout: float = 0.0
    init_lora_weights: object = True
    use_rslora: bool = False
    use_dora: bool = False
