"""Seeded synthetic inputs shared by the golden generator, the oracle pin tests and the GPU parity tests.

TEST INFRASTRUCTURE ONLY (lives under oracle/). Pure torch-CPU tensor construction, no reference code involved.
"""
from __future__ import annotations

import torch

from core_port import Cols


def random_dual_graph(g: torch.Generator, N: int, max_out: int = 4, sort_by_source: bool = True):
    """Random dual graph over N links: every link gets 0..max_out distinct downstream links, edge_attr = 1/outdeg
    (what config_network produces, src/transportation_simulator.py:150-168). Returns (edge_index [2,E] int64,
    edge_attr [E,1] fp32)."""
    ups, dns, ws = [], [], []
    for u in range(N):
        k = int(torch.randint(0, max_out + 1, (1,), generator=g))
        if k == 0:
            continue
        d = torch.randperm(N, generator=g)[:k]
        ups += [u] * k
        dns += d.tolist()
        ws += [1.0 / k] * k
    ei = torch.tensor([ups, dns], dtype=torch.long).view(2, -1)
    w = torch.tensor(ws, dtype=torch.float32).view(-1, 1)
    if not sort_by_source and ei.size(1) > 0:
        perm = torch.randperm(ei.size(1), generator=g)
        ei, w = ei[:, perm], w[perm]
    return ei, w


def random_road_state(g: torch.Generator, N: int, Nmax: int, t: float, ei: torch.Tensor, *, garbage: bool = True,
                      jam_fraction: float = 0.15, integral_times: bool = False):
    """Random `x[:N]` rows in the reference layout. Unique agent ids 1..A across all queues, head exit times
    around t, a share of links filled into the jam buffer so that the gridlock branch
    (src/direction_mpnn.py:87-89) is exercised, optional non-zero garbage past the tails."""
    c = Cols(Nmax)
    x = torch.zeros(N, c.F)
    maxn = torch.randint(4, Nmax, (N,), generator=g).float()          # MAXN in [4, Nmax-1]
    fftt = torch.randint(1, 20, (N,), generator=g).float()
    if not integral_times:
        fftt = fftt + torch.rand(N, generator=g)
    x[:, c.MAXN] = maxn
    x[:, c.FFTT] = fftt
    x[:, c.LENGTH] = 100.0
    x[:, c.MAX_FLOW] = torch.randint(600, 2400, (N,), generator=g).float()
    x[:, c.RIDX] = torch.arange(N).float()
    if garbage:
        x[:, c.ARR0:c.ARR0 + Nmax] = torch.randint(0, 50, (N, Nmax), generator=g).float()
        x[:, c.DEP0:c.DEP0 + Nmax] = torch.randint(0, 80, (N, Nmax), generator=g).float()
    num = torch.floor(torch.rand(N, generator=g) * (maxn - 2)).clamp(min=0)           # 0 .. MAXN-3
    jam = torch.rand(N, generator=g) < jam_fraction
    num = torch.where(jam, (maxn - torch.randint(0, 4, (N,), generator=g).float()).clamp(min=0), num)
    num = num.clamp(max=Nmax - 2)
    x[:, c.NUM] = num
    next_id = 1
    for n in range(N):
        k = int(num[n])
        if k == 0:
            continue
        x[n, c.ID0:c.ID0 + k] = torch.arange(next_id, next_id + k).float()
        next_id += k
        arr = t - torch.randint(5, 40, (k,), generator=g).float()
        dep = (t + torch.randint(-25, 6, (1,), generator=g).float()) + torch.cumsum(
            torch.randint(0, 3, (k,), generator=g).float(), 0)
        if not integral_times:
            dep = dep + torch.rand(k, generator=g)
        x[n, c.ARR0:c.ARR0 + k] = arr
        x[n, c.DEP0:c.DEP0 + k] = dep
    x[:, c.SEL] = random_selection(g, N, ei)
    return x, next_id - 1


def random_selection(g: torch.Generator, N: int, ei: torch.Tensor):
    """SELECTED_ROAD per link: mostly one of its downstream links (what Agents.choice / the RL action write),
    sometimes an unrelated link."""
    sel = torch.randint(0, max(N, 1), (N,), generator=g).float()
    if ei.size(1) > 0:
        order = torch.randperm(ei.size(1), generator=g)
        up, dn = ei[0][order], ei[1][order]
        keep = torch.rand(ei.size(1), generator=g) < 0.9
        sel[up[keep]] = dn[keep].float()     # last write wins: a random downstream link for ~all links that have one
    return sel


def uniforms(g: torch.Generator, E: int):
    """E uniforms in (0,1): exactly what torch.rand_like would hand to the Gumbel trick, minus the measure-zero 0."""
    return torch.rand(E, generator=g).clamp_(min=1e-7)


def grid_dual_graph(n: int):
    """n x n grid, one directed link per ordered 4-neighbour pair (N = 4n(n-1)); dual edge a->b iff to(a)==from(b),
    U-turns included, edge_attr = 1/outdeg(to(a)); edges in source order then out-link order, exactly what
    config_network emits for links listed node by node (src/transportation_simulator.py:150-168)."""
    idx = torch.arange(n * n).view(n, n)
    frm = torch.cat([idx[:, :-1].reshape(-1), idx[:, 1:].reshape(-1), idx[:-1, :].reshape(-1), idx[1:, :].reshape(-1)])
    to = torch.cat([idx[:, 1:].reshape(-1), idx[:, :-1].reshape(-1), idx[1:, :].reshape(-1), idx[:-1, :].reshape(-1)])
    return dual_from_links(frm, to, n * n)


def dual_from_links(frm: torch.Tensor, to: torch.Tensor, n_nodes: int):
    N = frm.numel()
    order = torch.argsort(frm, stable=True)                      # links leaving each node, in link order
    counts = torch.bincount(frm, minlength=n_nodes)
    ptr = torch.zeros(n_nodes + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(counts, 0)
    deg = counts[to]                                             # out-degree of each link's to-node
    up = torch.repeat_interleave(torch.arange(N), deg)
    base = torch.repeat_interleave(ptr[to], deg)
    within = torch.arange(up.numel()) - torch.repeat_interleave(torch.cumsum(deg, 0) - deg, deg)
    dn = order[base + within]
    w = (1.0 / deg[up].float()).view(-1, 1)
    return torch.stack([up, dn]), w, frm, to
