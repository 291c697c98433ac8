"""Stand-in for torch-scatter 2.1.2, CPU semantics (see oracle/shims/README.md). Test infrastructure only.

scatter_max restates csrc/cpu/scatter_cpu.cpp of torch-scatter 2.1.2: `out` is filled with
numeric_limits::lowest(), `arg` with src.size(dim); a sequential sweep over e applies
`if (src[e] > out[idx]) { out[idx] = src[e]; arg[idx] = e; }` (strict → lowest e wins ties; -inf and NaN
never win); finally entries still equal to lowest() are set to 0.
"""
import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    return src.new_zeros(shape).scatter_add_(dim, index, src)


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    if src.dim() == 2 and index.dim() == 1:  # batched [B, E] with a shared [E] index
        outs, args = zip(*[scatter_max(s, index, 0, None, dim_size) for s in src])
        return torch.stack(outs), torch.stack(args)
    assert src.dim() == 1 and index.dim() == 1
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    E = src.numel()
    lowest = torch.finfo(src.dtype).min
    seg_max = src.new_full((dim_size,), lowest).scatter_reduce_(0, index, src, reduce="amax", include_self=True)
    # first (lowest-e) position attaining a max that is strictly above lowest()
    hit = (src == seg_max[index]) & (src > lowest)
    pos = torch.where(hit, torch.arange(E), torch.full((E,), E))
    arg = torch.full((dim_size,), E, dtype=torch.long).scatter_reduce_(0, index, pos, reduce="amin", include_self=True)
    out_v = torch.where(seg_max == lowest, torch.zeros_like(seg_max), seg_max)
    return out_v, arg


def scatter_softmax(src, index, dim=-1, eps=1e-12, dim_size=None):
    # torch_scatter.composite.softmax: subtract per-group max, exp, divide by per-group sum
    n = int(index.max()) + 1 if dim_size is None else dim_size
    idx = index.expand_as(src)
    shape = list(src.shape)
    shape[dim] = n
    mx = src.new_full(shape, torch.finfo(src.dtype).min).scatter_reduce_(dim, idx, src, reduce="amax", include_self=True)
    rec = (src - mx.gather(dim, idx)).exp()
    s = src.new_zeros(shape).scatter_add_(dim, idx, rec)
    return rec / s.gather(dim, idx)
