import torch


class EnvBase:
    def __init__(self, device="cpu", batch_size=None, **kwargs):
        self.device = torch.device(device) if isinstance(device, str) else device

    def to(self, device):
        return self
