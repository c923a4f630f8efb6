"""Stand-in for peft==0.10.0 plumbing used by the reference's in-tree copy models/lora_layer.py."""
