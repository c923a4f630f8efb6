import torch


def transpose(weight, fan_in_fan_out):
    if not fan_in_fan_out:
        return weight
    if isinstance(weight, torch.nn.Parameter):
        return torch.nn.Parameter(weight.T)
    return weight.T


# names utils/peft_utils.py imports (values as in peft 0.10.0)
EMBEDDING_LAYER_NAMES = ["embed_tokens", "lm_head"]
SAFETENSORS_WEIGHTS_NAME = "adapter_model.safetensors"
WEIGHTS_NAME = "adapter_model.bin"


def check_file_exists_on_hf_hub(repo_id, filename, **kwargs):
    return None        # no network in this environment


def infer_device():
    return "cuda" if torch.cuda.is_available() else "cpu"
