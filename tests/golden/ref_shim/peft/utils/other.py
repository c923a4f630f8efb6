import torch


def transpose(weight, fan_in_fan_out):
    if not fan_in_fan_out:
        return weight
    if isinstance(weight, torch.nn.Parameter):
        return torch.nn.Parameter(weight.T)
    return weight.T
