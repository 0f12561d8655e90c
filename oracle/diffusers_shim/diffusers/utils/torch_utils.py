import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on ``device`` (on CPU then move when the generator is a CPU one)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None and generator.device.type != device.type and generator.device.type == "cpu":
        rand_device = torch.device("cpu")
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)
