"""`Data`: the attribute bag the reference takes from torch_geometric (not installed here, and not needed: the hot
path only uses it as a namespace for tensors — src/transportation_simulator.py:213-224).

One attribute is more than a slot: `x`. `SimulationCoreModel.forward(graph)` may keep the road state of a graph in a
resident link store between calls (core.py, `_ResidentRows`) and bring `graph.x` up to date only when somebody reads
it — reading `graph.x` is that moment. Code that takes the raw device pointer of the rows and writes through it
(agents.rows_state) says so with `rows_written()`; writes through torch ops are seen by the tensor's version counter.
"""
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    # ---- x: the node table, possibly shadowed by a resident link store --------------------------------------------
    @property
    def x(self):
        r = self.__dict__.get("_resident")
        if r is not None and r.dirty:
            r.sync_rows()                   # export: graph.x becomes exactly what the reference would hold
        try:
            return self.__dict__["_x"]
        except KeyError:
            raise AttributeError("x") from None

    @x.setter
    def x(self, value):
        if value is not self.__dict__.get("_x"):
            self.__dict__.pop("_resident", None)        # a new tensor: whatever shadowed the old one is void
        self.__dict__["_x"] = value

    def rows_written(self):
        """Called by code that wrote graph.x through its raw device pointer (no version bump happens there)."""
        self.__dict__["_rows_epoch"] = self.__dict__.get("_rows_epoch", 0) + 1

    def keys(self):
        out = [k for k in self.__dict__ if not k.startswith("_")]
        return (["x"] if "_x" in self.__dict__ else []) + out

    def __contains__(self, key):
        return key in self.keys()

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        if "_x" in self.__dict__:
            state["_x"] = self.x            # synchronised
        return state

    def __setstate__(self, state):
        if "x" in state:                    # files written before x became a property
            state = dict(state)
            state["_x"] = state.pop("x")
        self.__dict__.update(state)

    def to(self, device, non_blocking=False):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    @property
    def num_nodes(self):
        return self.x.size(0) if "_x" in self.__dict__ else None

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}")
        return f"Data({', '.join(parts)})"
