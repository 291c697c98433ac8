"""`Data`: the attribute bag the reference takes from torch_geometric (not installed here, and not needed: the hot
path only uses it as a namespace for tensors — src/transportation_simulator.py:213-224)."""
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def __contains__(self, key):
        return key in self.__dict__

    def to(self, device, non_blocking=False):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    @property
    def num_nodes(self):
        return self.x.size(0) if hasattr(self, "x") else None

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}")
        return f"Data({', '.join(parts)})"
