"""Static topology preprocessing: int32 CSR of the dual graph in both orientations, original edge ids kept.

Built once per `edge_index_routes` tensor with torch ops on its device (plumbing, not on the timed path) and cached
by tensor identity + version, so the drop-in `forward(graph)` can be called every step without re-sorting.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _cabi


class DualTopology:
    """CSR-by-target and CSR-by-source of a [2, E] int64 edge index over `n_links` nodes (row 0 = upstream / source,
    row 1 = downstream / target). Segments keep ascending original edge id (stable sort): the Gumbel arg-max and the
    reference's scatter_add_ both walk edges in that order."""

    def __init__(self, edge_index: torch.Tensor, n_links: int):
        assert edge_index.dim() == 2 and edge_index.size(0) == 2
        E = edge_index.size(1)
        if n_links >= 2 ** 31 or E >= 2 ** 31:
            raise ValueError("topology exceeds int32 indexing")
        dev = edge_index.device
        src, dst = edge_index[0].long(), edge_index[1].long()
        if E and (int(src.min()) < 0 or int(dst.min()) < 0 or int(src.max()) >= n_links or int(dst.max()) >= n_links):
            raise ValueError("edge index out of range for n_links")
        self.n_links, self.n_edges, self.device = n_links, E, dev
        i32 = dict(dtype=torch.int32, device=dev)
        order_in = torch.argsort(dst, stable=True)
        self.in_eid = order_in.to(torch.int32)
        self.in_src = src[order_in].to(torch.int32)
        self.in_ptr = torch.zeros(n_links + 1, **i32)
        self.in_ptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_links), 0)
        order_out = torch.argsort(src, stable=True)
        self.out_eid = order_out.to(torch.int32)
        self.out_dst = dst[order_out].to(torch.int32)
        self.out_ptr = torch.zeros(n_links + 1, **i32)
        self.out_ptr[1:] = torch.cumsum(torch.bincount(src, minlength=n_links), 0)
        self.src32 = src.to(torch.int32)
        self.dst32 = dst.to(torch.int32)
        self.source_sorted = bool(E == 0 or torch.equal(order_out, torch.arange(E, device=dev)))
        self.struct = _cabi.DualCSR(n_links, E, self.in_ptr.data_ptr(), self.in_src.data_ptr(), self.in_eid.data_ptr(),
                                    self.out_ptr.data_ptr(), self.out_dst.data_ptr(),
                                    None if self.source_sorted else self.out_eid.data_ptr())

    def ref(self):
        return C.byref(self.struct)

    def ell(self, edge_attr: torch.Tensor):
        """ELLPACK copy of the first W edges of every link (struct tarl_dual_ell), W = 4 or 8 by maximum degree.
        Cached per edge_attr tensor. Column j of link n sits at [j*pitch + n]; links with more than W edges carry -2
        in the last column and are served from the CSR — and so are, on the in-edge side, links with an in-edge whose
        weight is not >= 1e-3 (there an ineligible edge can win the Gumbel arg-max: they take the literal scan over
        every in-edge)."""
        key = (edge_attr.data_ptr(), edge_attr._version)
        hit = getattr(self, "_ell", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        N, E, dev = self.n_links, self.n_edges, self.device
        in_deg = (self.in_ptr[1:] - self.in_ptr[:-1]).long()
        out_deg = (self.out_ptr[1:] - self.out_ptr[:-1]).long()
        max_deg = int(max(in_deg.max().item() if N else 0, out_deg.max().item() if N else 0))
        W = 4 if max_deg <= 4 else 8
        pitch = max((N + 31) // 32 * 32, 32)
        attr = edge_attr.reshape(-1).to(torch.float32)

        def fill(ptr, deg, values):
            cols = torch.full((W, pitch), -1, dtype=values.dtype, device=dev)
            if E:
                owner = torch.repeat_interleave(torch.arange(N, device=dev), deg)
                rank = torch.arange(E, device=dev) - ptr[:-1].long()[owner]
                keep = rank < W
                cols[rank[keep], owner[keep]] = values[keep]
            return cols

        in_src = fill(self.in_ptr, in_deg, self.in_src)
        in_attr = fill(self.in_ptr, in_deg, attr[self.in_eid.long()])
        out_dst = fill(self.out_ptr, out_deg, self.out_dst)
        if N:
            general = in_deg > W
            if E:
                unsafe = ~(attr >= 1e-3)                       # also catches NaN
                general = general | (torch.zeros(N, dtype=torch.long, device=dev).index_add_(
                    0, self.dst32.long(), unsafe.long()) > 0)
            in_src[W - 1, :N][general] = -2
            out_dst[W - 1, :N][out_deg > W] = -2
        # the weight shared by ALL in-edges of a link (bitwise equal), NaN where they differ: LinkStore parks it in
        # stat_a.w (TARL_STORE_UNIFORM_WEIGHTS) and the direction kernel skips the link's edge-weight column. Links
        # that walk their CSR segment never look at it; a link without in-edges never uses a weight (0).
        uniform = torch.zeros(max(N, 1), dtype=torch.float32, device=dev)
        if N and E:
            uniform[:N] = uniform_in_weights(in_attr[:, :N], in_deg, general)
        struct = _cabi.DualELL(W, pitch, in_src.data_ptr(), in_attr.data_ptr(), out_dst.data_ptr())
        pack = (struct, in_src, in_attr, out_dst, uniform)
        self._ell = (key, pack)
        return pack


def uniform_in_weights(in_attr: torch.Tensor, in_deg: torch.Tensor, general: torch.Tensor) -> torch.Tensor:
    """in_attr [W, N]: the ELL edge-weight columns (column j = weight of the link's j-th in-edge in ascending edge id),
    in_deg [N], general [N] bool (links that walk their CSR segment). Returns fp32 [N]: the weight when every in-edge of
    the link carries bitwise the same one, NaN when they differ or the link is general, 0 for a link without in-edges
    (which never uses a weight). Plain torch: runs wherever the tensors live."""
    W, N = in_attr.shape
    first = in_attr[0].contiguous()
    same = torch.ones(N, dtype=torch.bool, device=in_attr.device)
    for j in range(1, W):
        same &= (in_deg <= j) | (in_attr[j].contiguous().view(torch.int32) == first.view(torch.int32))
    nan = torch.full_like(first, float("nan"))
    return torch.where(in_deg == 0, torch.zeros_like(first), torch.where(same & ~general, first, nan))


_CACHE: dict = {}


def topology_for(edge_index: torch.Tensor, n_links: int) -> DualTopology:
    """Cached DualTopology for this exact tensor (identity, version counter and shape must all match)."""
    key = (id(edge_index), n_links)
    hit = _CACHE.get(key)
    if hit is not None:
        ref, version, ptr, shape, topo = hit
        if ref() is edge_index and version == edge_index._version and ptr == edge_index.data_ptr() and shape == tuple(edge_index.shape):
            return topo
    topo = DualTopology(edge_index, n_links)
    if len(_CACHE) > 64:
        for k in [k for k, v in _CACHE.items() if v[0]() is None]:
            del _CACHE[k]
    _CACHE[key] = (weakref.ref(edge_index), edge_index._version, edge_index.data_ptr(), tuple(edge_index.shape), topo)
    return topo


class GroupCSR:
    """One CSR orientation of an edge list for the MPNN kernels (struct tarl_csr).

    by="source_rank": rows are the RANKS of the distinct source ids (GraphDistribution groups, declared divergence D1:
    the reference indexes by raw source id and only works when sources are exactly 0..K-1).
    by="target": rows are target node ids 0..n_nodes-1 (policy backward, value-net backward).
    by="source": rows are source node ids 0..n_nodes-1 (value-net aggregation)."""

    def __init__(self, edge_index: torch.Tensor, by: str, n_nodes: int | None = None):
        E = edge_index.size(1)
        dev = edge_index.device
        if by == "source_rank":
            self.nodes, key = torch.unique(edge_index[0], return_inverse=True)
            rows = int(self.nodes.numel())
            other = edge_index[1]
        elif by == "target":
            key, rows, other = edge_index[1].long(), int(n_nodes), edge_index[0]
            self.nodes = None
        elif by == "source":
            key, rows, other = edge_index[0].long(), int(n_nodes), edge_index[1]
            self.nodes = None
        else:
            raise ValueError(by)
        order = torch.argsort(key, stable=True)
        self.eid = order.to(torch.int32)
        self.idx = other[order].to(torch.int32)
        self.ptr = torch.zeros(rows + 1, dtype=torch.int32, device=dev)
        if E:
            self.ptr[1:] = torch.cumsum(torch.bincount(key, minlength=rows), 0)
        self.n_rows, self.n_edges = rows, E
        self.struct = _cabi.CSR(rows, E, self.ptr.data_ptr(), self.idx.data_ptr(), self.eid.data_ptr())

    def ref(self):
        return C.byref(self.struct)


_GROUP_CACHE: dict = {}


def group_csr_for(edge_index: torch.Tensor, by: str, n_nodes: int | None = None) -> GroupCSR:
    key = (id(edge_index), by, n_nodes)
    hit = _GROUP_CACHE.get(key)
    if hit is not None:
        ref, version, ptr, shape, csr = hit
        if ref() is edge_index and version == edge_index._version and ptr == edge_index.data_ptr() and shape == tuple(edge_index.shape):
            return csr
    csr = GroupCSR(edge_index, by, n_nodes)
    if len(_GROUP_CACHE) > 64:
        for k in [k for k, v in _GROUP_CACHE.items() if v[0]() is None]:
            del _GROUP_CACHE[k]
    _GROUP_CACHE[key] = (weakref.ref(edge_index), edge_index._version, edge_index.data_ptr(), tuple(edge_index.shape), csr)
    return csr
