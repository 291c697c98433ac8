"""`Data` stand-in: an attribute bag with `.to()`; enough for the reference's hot path."""
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    @property
    def num_edges(self):
        return self.edge_index.size(1)
