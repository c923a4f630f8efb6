import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    device = device or torch.device("cpu")
    if isinstance(generator, list):
        shape = (1,) + tuple(shape[1:])
        return torch.cat([torch.randn(shape, generator=g, device=g.device, dtype=dtype).to(device) for g in generator])
    gdev = generator.device if generator is not None else device
    return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)
