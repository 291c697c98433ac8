"""CPU oracle for the population operations either side of the core step and for the two step orders built on them.

TEST INFRASTRUCTURE ONLY. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this file.

Restates, in plain torch on the CPU, what the reference executes in
  Agents.insert_agent_into_network   /root/reference/src/agents/base.py:244-331
  Agents.withdraw_agent_from_network /root/reference/src/agents/base.py:334-403
  Agents.choice                      /root/reference/src/agents/base.py:446-494
  TransportationSimulator.run        /root/reference/src/transportation_simulator.py:294-351
  SimulatorEnv._step                 /root/reference/src/reinforcement_learning.py:222-309
with the declared divergences of SURVEY.md §8c: D3 (agents that enter the same road in the same step are ordered by
ascending agent id: the reference's default argsort is unstable), D4 (random draws injected: `choice` takes one
uniform per choosing node and picks neighbour number min(floor(u*deg), deg-1) in ascending road id, where the
reference calls torch.multinomial on the dense 0/1 row), D6 (`_step` always advances time).

Parity status: PINNED. tests/test_oracle_pinned.py checks it against the unmodified reference (run behind
oracle/shims with torch.argsort made stable and torch.multinomial replaced by the injected rule) whenever
/root/reference is present, and against tests/golden/sim_*.npz (oracle/gen_golden_sim.py) everywhere.
"""
from __future__ import annotations

import torch

import core_port
from core_port import Cols

ORIGIN, DESTINATION, DEPARTURE_TIME, ARRIVAL_TIME, AGE, SEX, EMPLOYED, ON_WAY, DONE = range(9)


def insert(x: torch.Tensor, af: torch.Tensor, t, c: Cols, cc: torch.Tensor | None):
    """base.py:244-331 on the full node table x [N_tot, F] and agent_features af, both mutated in place."""
    ready = (af[:, DEPARTURE_TIME] <= t) & (af[:, ON_WAY] == 0) & (af[:, DONE] == 0)                    # :247-251
    ids = torch.nonzero(ready).flatten()
    if ids.numel() == 0:
        return
    road = x[af[ids, ORIGIN].long(), c.SEL].long()                                                    # :258-259
    room = (x[road, c.MAXN] - c.JAM_BUFFER - x[road, c.NUM]).long()                                   # :262-266
    fits = room > 0
    ids, road, room = ids[fits], road[fits], room[fits]
    if ids.numel() == 0:
        return
    order = torch.argsort(road, stable=True)                                                          # :275, D3
    ids, road, room = ids[order], road[order], room[order]
    new_group = torch.ones_like(road, dtype=torch.bool)
    new_group[1:] = road[1:] != road[:-1]
    first = torch.nonzero(new_group).flatten()
    group = torch.cumsum(new_group.long(), 0) - 1
    rank = torch.arange(road.numel()) - first[group]
    admitted = rank < room                                                                            # :284-291
    ids, road, rank = ids[admitted], road[admitted], rank[admitted]
    if ids.numel() == 0:
        return
    n0 = x[road, c.NUM].long()                                                                        # :281
    pos = n0 + rank                                                                                   # :308
    x[road, c.ID0 + pos] = ids.float()                                                                # :310
    x[road, c.ARR0 + pos] = float(t)                                                                  # :311
    if cc is not None:
        t_cong = cc[road].to(x.dtype) / (x[road, c.MAXN] + 10 - n0.to(x.dtype))                       # :315-317
    else:
        t_cong = torch.zeros_like(n0, dtype=x.dtype)
    tt = torch.max(torch.stack((x[road, c.FFTT], t_cong)), dim=0).values                              # :321-323
    x[road, c.DEP0 + pos] = float(t) + tt                                                             # :324-325
    uniq, cnt = torch.unique_consecutive(road, return_counts=True)
    x[uniq, c.NUM] += cnt.to(x.dtype)                                                                 # :327
    af[ids, ON_WAY] = 1.0                                                                             # :328


def withdraw(x: torch.Tensor, af: torch.Tensor, t, c: Cols, adj: torch.Tensor, num_roads: int):
    """base.py:334-403. adj: dense bool [N_tot, N_tot]. Returns the withdrawn mask bool[num_roads]."""
    Nmax = c.Nmax
    rows = x[:, c.RIDX].long()
    ids = x[:, c.ID0:c.ID0 + Nmax].long()
    dest = af[ids, DESTINATION].long()
    ok = (adj[rows.unsqueeze(1), dest] > 0) & (x[:, c.DEP0:c.DEP0 + Nmax] <= t) \
        & (torch.arange(Nmax) < x[:, c.NUM].unsqueeze(1))                                             # :361-367
    prefix = torch.cumprod(ok.long(), dim=1).bool()                                                   # :370
    count = prefix.sum(dim=1)
    mask = count[:num_roads] > 0
    if bool(count.any()):
        gone = ids[prefix]
        shift = torch.arange(Nmax).unsqueeze(0) + count.unsqueeze(1)                                  # :377-379
        inside = shift < Nmax
        src = shift.clamp(max=Nmax - 1)
        for lo in (c.ID0, c.ARR0, c.DEP0):
            seg = x[:, lo:lo + Nmax].gather(1, src)
            seg[~inside] = 0
            x[:, lo:lo + Nmax] = seg
        x[:, c.NUM] -= count                                                                          # :396
        af[gone, DONE] = 1
        af[gone, ON_WAY] = 0
        af[gone, ARRIVAL_TIME] = t                                                                    # :398-400
    return mask.clone()


def choosers_and_neighbours(edge_index: torch.Tensor, num_roads: int, n_nodes: int):
    """The rows `choice` samples for and what it samples from, sparse: roads with a downstream road (rows of
    adj[:N,:N]) then SRC nodes with an outgoing road (rows of src_adj), each with its neighbours in ascending id."""
    src, dst = edge_index[0], edge_index[1]
    keep = dst < num_roads
    pairs = torch.unique(src[keep] * num_roads + dst[keep])
    s, d = pairs // num_roads, pairs % num_roads
    is_src_row = (s >= num_roads) & ((s - num_roads) % 2 == 0)
    keep = (s < num_roads) | is_src_row
    s, d = s[keep], d[keep]
    deg = torch.bincount(s, minlength=n_nodes)
    ptr = torch.zeros(n_nodes + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(deg, 0)
    return torch.nonzero(deg > 0).flatten(), ptr, d


def choice(x: torch.Tensor, c: Cols, edge_index: torch.Tensor, num_roads: int, u: torch.Tensor):
    """base.py:446-494 with the draw injected (D4): u holds one uniform per choosing node, in ascending node id."""
    nodes, ptr, nbr = choosers_and_neighbours(edge_index, num_roads, x.size(0))
    if nodes.numel() == 0:
        return
    deg = ptr[nodes + 1] - ptr[nodes]
    k = (u.to(torch.float32) * deg.to(torch.float32)).long().clamp(min=0)
    k = torch.minimum(k, deg - 1)
    x[nodes, c.SEL] = nbr[ptr[nodes] + k].to(x.dtype)                                                 # :486,:491


def multinomial_rule(u: torch.Tensor):
    """The stand-in for torch.multinomial that the pin tests / golden generator install into the reference run:
    row i picks its positive entry number min(floor(u_i*deg_i), deg_i-1), positives counted in ascending column."""
    def fake(probs, num_samples=1, **kw):
        assert num_samples == 1 and probs.size(0) == u.numel(), (probs.shape, u.shape)
        pos = probs > 0
        deg = pos.sum(dim=1)
        k = (u.to(torch.float32) * deg.to(torch.float32)).long().clamp(min=0)
        k = torch.minimum(k, deg - 1)
        nth = torch.cumsum(pos.long(), dim=1) - 1
        hit = pos & (nth == k.unsqueeze(1))
        return hit.float().argmax(dim=1, keepdim=True)
    return fake


def run_step(x, af, t, c: Cols, graph: dict, u_choice, u_core, timestep=1):
    """TransportationSimulator.run (transportation_simulator.py:294-342): insert, withdraw, choice, core. `graph`:
    dict(edge_index, edge_index_routes, edge_attr_routes, num_roads, adj_matrix, congestion_constant | None)."""
    N = graph["num_roads"]
    cc = graph.get("congestion_constant")
    insert(x, af, t, c, cc)
    wmask = withdraw(x, af, t, c, graph["adj_matrix"], N)
    choice(x, c, graph["edge_index"], N, u_choice)
    out = core_port.core_step(x[:N], graph["edge_index_routes"], graph["edge_attr_routes"], t, c.Nmax, u_core,
                              cc[:N] if cc is not None else None)
    out["withdrawn"] = wmask
    return out


def env_step(x, af, t, c: Cols, graph: dict, action, u_core, timestep=1):
    """SimulatorEnv._step (reinforcement_learning.py:222-276): action -> SELECTED_ROAD, core, withdraw, insert,
    reward = -sum NUM. Time advances every step (D6). Returns dict(reward, withdrawn, delta_tt, pop, done, t_next)."""
    N = graph["num_roads"]
    cc = graph.get("congestion_constant")
    ei = graph["edge_index"]
    on = action.to(torch.bool)
    x[ei[0][on], c.SEL] = ei[1][on].to(torch.float)                                                   # :223-231
    out = core_port.core_step(x[:N], graph["edge_index_routes"], graph["edge_attr_routes"], t, c.Nmax, u_core,
                              cc[:N] if cc is not None else None)
    out["withdrawn"] = withdraw(x, af, t, c, graph["adj_matrix"], N)
    insert(x, af, t, c, cc)
    out["reward"] = -torch.sum(x[:, c.NUM])                                                           # :266
    out["t_next"] = t + timestep
    out["done"] = out["t_next"] > 7 * 3600                                                            # :273
    return out
