"""MessagePassing stand-in restating PyG 2.5 `propagate` for the call patterns the reference uses."""
import inspect

import torch


def _scatter_builtin(src, index, dim_size, reduce):
    # torch_geometric.utils.scatter (2.5): zeros + scatter_reduce_(include_self=False); mean = sum / count.clamp(1)
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("add", "sum"):
        return src.new_zeros(shape).scatter_add_(0, idx, src)
    if reduce == "mean":
        out = src.new_zeros(shape).scatter_add_(0, idx, src)
        cnt = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(index.numel())).clamp_(min=1)
        return out / cnt.view(-1, *([1] * (src.dim() - 1)))
    if reduce == "max":
        return src.new_zeros(shape).scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    raise NotImplementedError(reduce)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def _params(self, fn):
        return [p for p in inspect.signature(fn).parameters if p not in ("self",)]

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        x = kwargs.get("x")
        dim_size = x.size(0) if x is not None else int(edge_index.max()) + 1

        def collect(names):
            out = {}
            for name in names:
                if name.endswith("_i") or name.endswith("_j"):
                    base = kwargs[name[:-2]]
                    sel = edge_index[i] if name.endswith("_i") else edge_index[j]
                    out[name] = base.index_select(0, sel)
                elif name in kwargs:
                    out[name] = kwargs[name]
            return out

        msg = self.message(**collect(self._params(self.message)))
        agg_fn = type(self).aggregate
        if agg_fn is MessagePassing.aggregate:
            aggr_out = _scatter_builtin(msg, edge_index[i], dim_size, self.aggr)
        else:
            aggr_out = self.aggregate(msg, index=edge_index[i], ptr=None, dim_size=dim_size)
        upd = collect([p for p in self._params(self.update) if p not in ("aggr_out", "inputs")])
        return self.update(aggr_out, **upd)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return _scatter_builtin(inputs, index, dim_size, self.aggr)

    def update(self, aggr_out):
        return aggr_out
