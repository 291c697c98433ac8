"""Synthetic networks and warm states of the sizes BASELINE.json names, built directly as tensors.

The reference can only build a network from MATSim XML through a dense [N_tot, N_tot] adjacency
(src/transportation_simulator.py:196-198), which is impossible at 40k / 1M links; these generators emit exactly the
tensors `config_network` would emit for the same link list (same row layout, same dual-edge order: for link j in
file order, one edge to every link leaving to(j) in file order, U-turns included, edge_attr = 1/outdeg —
src/transportation_simulator.py:150-168), without the dense matrices. Pure torch tensor construction (plumbing).
"""
from __future__ import annotations

import torch

from .data import Data
from .feature_helpers import FeatureHelpers

LINK_LENGTH, LINK_FREESPEED, LINK_CAPACITY, LINK_LANES, CELL = 100.0, 13.89, 1800.0, 1.0, 7.5


def grid_links(n: int, device="cpu"):
    """n x n intersections, one directed link per ordered 4-neighbour pair: N = 4 n (n-1)."""
    idx = torch.arange(n * n, device=device).view(n, n)
    frm = torch.cat([idx[:, :-1].reshape(-1), idx[:, 1:].reshape(-1), idx[:-1, :].reshape(-1), idx[1:, :].reshape(-1)])
    to = torch.cat([idx[:, 1:].reshape(-1), idx[:, :-1].reshape(-1), idx[1:, :].reshape(-1), idx[:-1, :].reshape(-1)])
    return frm, to, n * n


def ring_radial_links(rings: int, spokes: int, device="cpu"):
    """`rings` concentric rings of `spokes` intersections; ring links both ways (2 R S) and radial links both ways
    between consecutive rings (2 (R-1) S): N = 999 000 for 500 x 500."""
    idx = torch.arange(rings * spokes, device=device).view(rings, spokes)
    nxt = torch.roll(idx, shifts=-1, dims=1)
    frm = torch.cat([idx.reshape(-1), nxt.reshape(-1), idx[:-1].reshape(-1), idx[1:].reshape(-1)])
    to = torch.cat([nxt.reshape(-1), idx.reshape(-1), idx[1:].reshape(-1), idx[:-1].reshape(-1)])
    return frm, to, rings * spokes


def dual_edges(frm: torch.Tensor, to: torch.Tensor, n_nodes: int):
    """edge_index_routes [2,E] and edge_attr_routes [E,1] in config_network's order."""
    N = frm.numel()
    dev = frm.device
    order = torch.argsort(frm, stable=True)
    counts = torch.bincount(frm, minlength=n_nodes)
    ptr = torch.zeros(n_nodes + 1, dtype=torch.long, device=dev)
    ptr[1:] = torch.cumsum(counts, 0)
    deg = counts[to]
    up = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    start = torch.cumsum(deg, 0) - deg
    within = torch.arange(up.numel(), device=dev) - torch.repeat_interleave(start, deg)
    dn = order[torch.repeat_interleave(ptr[to], deg) + within]
    w = (1.0 / deg[up].to(torch.float32)).view(-1, 1)
    return torch.stack([up, dn]), w


def full_edge_index(frm, to, n_nodes, routes, routes_attr):
    """edge_index / edge_attr of the full graph: routes ++ SRC(v)->link for links leaving v ++ link->DEST(v) for links
    entering v, attr 0 (src/transportation_simulator.py:173-193). SRC(v) = N+2v, DEST(v) = N+2v+1."""
    N = frm.numel()
    dev = frm.device
    links = torch.arange(N, device=dev)
    o_out = torch.argsort(frm, stable=True)          # intersections ascending, their leaving links in file order
    o_in = torch.argsort(to, stable=True)
    src_part = torch.stack([N + 2 * frm[o_out], links[o_out]])
    dst_part = torch.stack([links[o_in], N + 2 * to[o_in] + 1])
    ei = torch.cat([routes, src_part, dst_part], dim=1)
    ea = torch.cat([routes_attr, torch.zeros(2 * N, 1, device=dev)], dim=0)
    return ei, ea


def reorder_links(frm, to, order: str, seed: int = 0):
    """Link numbering = order of the <link> elements in the network file. "direction": as generated (all links of one
    compass direction together); "node": grouped by from-node, the way a network written intersection by intersection
    comes out (neighbouring links get neighbouring ids); "shuffled": a random permutation (worst case for locality)."""
    if order == "direction":
        return frm, to
    if order == "node":
        perm = torch.argsort(frm, stable=True)
    elif order == "shuffled":
        perm = torch.randperm(frm.numel(), generator=torch.Generator().manual_seed(seed)).to(frm.device)
    else:
        raise ValueError(order)
    return frm[perm], to[perm]


def build_graph(frm, to, n_nodes, *, with_full_edges=True) -> tuple[Data, int]:
    """Data with the attribute names the reference's code relies on (src/transportation_simulator.py:213-224), minus
    the dense adj_matrix / src_adj."""
    dev = frm.device
    N = frm.numel()
    maxn = int(LINK_LENGTH * LINK_LANES / CELL) + 1
    Nmax = maxn + 1
    h = FeatureHelpers(Nmax)
    x = torch.zeros(N + 2 * n_nodes, h.num_features, dtype=torch.float32, device=dev)
    x[:N, h.MAX_NUMBER_OF_AGENT] = float(maxn)
    x[:N, h.LENGHT_OF_ROAD] = LINK_LENGTH
    x[:N, h.MAX_FLOW] = LINK_CAPACITY
    x[:N, h.FREE_FLOW_TIME_TRAVEL] = torch.tensor(LINK_LENGTH, dtype=torch.float32) / LINK_FREESPEED
    x[:N, h.ROAD_INDEX] = torch.arange(N, device=dev, dtype=torch.float32)
    x[N:, h.ROAD_INDEX] = -1.0
    routes, routes_attr = dual_edges(frm, to, n_nodes)
    crit = x[:, h.MAX_FLOW] * x[:, h.FREE_FLOW_TIME_TRAVEL] / 3600
    cc = x[:, h.FREE_FLOW_TIME_TRAVEL] * (x[:, h.MAX_NUMBER_OF_AGENT] + 10 - crit)
    g = Data(x=x, edge_index_routes=routes, edge_attr_routes=routes_attr, num_roads=N, critical_number=crit,
             congestion_constant=cc, link_from=frm, link_to=to, num_intersections=n_nodes)
    if with_full_edges:
        g.edge_index, g.edge_attr = full_edge_index(frm, to, n_nodes, routes, routes_attr)
    return g, Nmax


def warm_state(g: Data, Nmax: int, n_agents: int, t: float, seed: int = 0):
    """Steady-state-like queues for timing: queue lengths Poisson(n_agents/N) clipped to Nmax-5, unique agent ids
    1..A in queue order, arrival times before t, head exit times uniform integers in [t-5, t+5] and non-decreasing
    along the queue, SELECTED_ROAD = a uniformly random downstream link. Returns the number of agents placed."""
    h = FeatureHelpers(Nmax)
    dev = g.x.device
    N = int(g.num_roads)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    lam = torch.full((N,), n_agents / N)
    num = torch.poisson(lam, generator=gen).clamp_(max=Nmax - 5).to(dev)
    slots = torch.arange(Nmax, device=dev).view(1, -1)
    live = slots < num.view(-1, 1)
    ids = torch.cumsum(live.reshape(-1).to(torch.int64), 0).view(N, Nmax)
    x = g.x
    x[:N, h.AGENT_POSITION] = torch.where(live, ids.to(torch.float32), torch.zeros((), device=dev))
    head_dep = t + torch.randint(-5, 6, (N, 1), generator=gen).to(dev).to(torch.float32)
    gaps = torch.randint(0, 3, (N, Nmax), generator=gen).to(dev).to(torch.float32)
    gaps[:, 0] = 0
    dep = head_dep + torch.cumsum(gaps, 1)
    x[:N, h.AGENT_TIME_DEPARTURE] = torch.where(live, dep, torch.zeros((), device=dev))
    fftt = x[:N, h.FREE_FLOW_TIME_TRAVEL].view(-1, 1)
    x[:N, h.AGENT_TIME_ARRIVAL] = torch.where(live, dep - fftt.ceil(), torch.zeros((), device=dev))
    x[:N, h.NUMBER_OF_AGENT] = num
    x[:N, h.SELECTED_ROAD] = random_out_neighbour(g, seed + 1)
    return int(num.sum().item())


def random_out_neighbour(g: Data, seed: int):
    """One uniformly random downstream link per link (what Agents.choice draws, src/agents/base.py:446-494);
    links without a downstream link keep 0."""
    ei = g.edge_index_routes
    N = int(g.num_roads)
    dev = ei.device
    gen = torch.Generator(device="cpu").manual_seed(seed)
    deg = torch.bincount(ei[0], minlength=N)
    start = torch.cumsum(deg, 0) - deg
    pick = (torch.rand(N, generator=gen).to(dev) * deg).long().clamp_(max=(deg - 1).clamp(min=0))
    sel = torch.zeros(N, dtype=torch.float32, device=dev)
    has = deg > 0
    # edge_index_routes is source-sorted by construction, so out-edges of link u are [start[u], start[u]+deg[u])
    sel[has] = ei[1][(start + pick)[has]].to(torch.float32)
    return sel


WORKLOADS = {
    # name: (kind, args, agents)
    "grid100": ("grid", (100,), 100_000),
    "ring_radial_1m": ("ring_radial", (500, 500), 2_000_000),
    "grid16": ("grid", (16,), 2_000),
    # size sweep of the bench network (tuning: the state of the 250k form lives in the L2, that of the 4m form cannot)
    "ring_radial_250k": ("ring_radial", (250, 250), 500_000),
    "ring_radial_4m": ("ring_radial", (1000, 1000), 8_000_000),
}


def make_workload(name: str, device="cuda", t: float = 21600.0, seed: int = 0, order: str = "node"):
    kind, args, agents = WORKLOADS[name]
    frm, to, n_nodes = (grid_links if kind == "grid" else ring_radial_links)(*args, device=device)
    frm, to = reorder_links(frm, to, order)
    g, Nmax = build_graph(frm, to, n_nodes)
    placed = warm_state(g, Nmax, agents, t, seed)
    return g, Nmax, placed


def population(g: Data, n_agents: int, t0: float, spread: int, seed: int = 0, device=None):
    """agent_features [A+1, 9] for a synthetic network: origin / destination intersections i.i.d.
    uniform with o != d (as SRC / DEST node ids), integer departure times uniform in [t0, t0+spread), row 0 = the
    dummy agent that never departs."""
    N, n_int = int(g.num_roads), int(g.num_intersections)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    af = torch.zeros(n_agents + 1, 9, dtype=torch.float32)
    af[0, 2] = 48 * 3600.0
    o = torch.randint(0, n_int, (n_agents,), generator=gen)
    d = (o + torch.randint(1, max(n_int, 2), (n_agents,), generator=gen)) % n_int
    af[1:, 0] = (N + 2 * o).to(torch.float32)
    af[1:, 1] = (N + 2 * d + 1).to(torch.float32)
    af[1:, 2] = t0 + torch.randint(0, max(spread, 1), (n_agents,), generator=gen).to(torch.float32)
    af[1:, 4] = torch.randint(18, 80, (n_agents,), generator=gen).to(torch.float32)
    return af.to(device if device is not None else g.x.device)


def write_scenario(root: str, name: str, kind: str = "grid", args=(5,), n_agents: int = 200, t0: int = 21540,
                   spread: int = 120, seed: int = 0):
    """data/<name>/network.xml + population.xml under `root`, in the MATSim dialect the readers accept
    (src/transportation_simulator.py:61-228, src/agents/base.py:38-242). Intersection ids are zero-padded so that
    their string order equals their numeric order."""
    import os
    frm, to, n_nodes = (grid_links if kind == "grid" else ring_radial_links)(*args)
    frm, to = reorder_links(frm, to, "node")
    width = len(str(n_nodes))
    nid = lambda v: f"{int(v):0{width}d}"
    side = int(round(n_nodes ** 0.5))
    d = os.path.join(root, "data", name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "network.xml"), "w") as f:
        f.write("<network>\n  <nodes>\n")
        for v in range(n_nodes):
            f.write(f'    <node id="{nid(v)}" x="{100.0 * (v % side)}" y="{100.0 * (v // side)}"/>\n')
        f.write(f'  </nodes>\n  <links effectivecellsize="{CELL}">\n')
        for i in range(frm.numel()):
            f.write(f'    <link id="{i}" from="{nid(frm[i])}" to="{nid(to[i])}" length="{LINK_LENGTH}" '
                    f'capacity="{LINK_CAPACITY}" freespeed="{LINK_FREESPEED}" permlanes="{LINK_LANES}"/>\n')
        f.write("  </links>\n</network>\n")
    gen = torch.Generator().manual_seed(seed)
    with open(os.path.join(d, "population.xml"), "w") as f:
        f.write("<population>\n")
        for p in range(n_agents):
            o = int(torch.randint(0, n_nodes, (1,), generator=gen))
            dd = (o + int(torch.randint(1, n_nodes, (1,), generator=gen))) % n_nodes
            dep = t0 + int(torch.randint(0, max(spread, 1), (1,), generator=gen))
            hh, mm, ss = dep // 3600, (dep % 3600) // 60, dep % 60
            f.write(f'  <person id="{p}" age="{20 + p % 50}" sex="{"f" if p % 2 else "m"}" car_avail="always">'
                    f'<plan><act type="h" link="{nid(o)}" end_time="{hh:02d}:{mm:02d}:{ss:02d}"/>'
                    f'<act type="w" link="{nid(dd)}"/></plan></person>\n')
        f.write("</population>\n")
    return d
