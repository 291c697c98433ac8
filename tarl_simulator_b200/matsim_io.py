"""MATSim XML readers (host side, stdlib ElementTree + gzip): the network → graph tensors and the population →
agent_features rows. Restates what the reference parses in src/transportation_simulator.py:61-228 (config_network)
and src/agents/base.py:38-242 (config_agents_from_xml); the device never sees XML, so this is plain Python.

Unlike the reference, no dense [N_tot, N_tot] adjacency is required downstream (the kernels use CSR forms), which is
what makes the 40k-link and 1M-link networks constructible at all; `adj_matrix` / `src_adj` are still produced for
networks small enough to hold them (<= DENSE_LIMIT nodes) because user code of the reference may read them.
"""
from __future__ import annotations

import gzip
import os
import xml.etree.ElementTree as ET
from collections import defaultdict
from datetime import datetime

import numpy as np
import torch

from .data import Data
from .feature_helpers import FeatureHelpers

DENSE_LIMIT = 4096


def actual_path(prefix: str) -> str:
    """`prefix`.xml.gz if present, else `prefix`.xml (src/transportation_simulator.py:75-83)."""
    gz_path, xml_path = prefix + ".xml.gz", prefix + ".xml"
    if os.path.exists(gz_path):
        return gz_path
    if os.path.exists(xml_path):
        return xml_path
    raise FileNotFoundError(f"Neither {gz_path} nor {xml_path} exists.")


def parse_xml(path: str):
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return ET.parse(f).getroot()
    return ET.parse(path).getroot()


def graph_from_links(frm, to, length, capacity, freespeed, permlanes, effective_cell_size=7.5, dense=None):
    """The graph config_network builds, from per-link arrays (file order). `frm` / `to`: intersection ids (any
    sortable hashables; SRC/DEST node numbering follows their sorted order, :142). Returns (Data on cpu, Nmax)."""
    n = len(frm)
    length = np.asarray(length, dtype=np.float32)
    cap32 = np.asarray(capacity, dtype=np.float32)
    fftt = length / np.asarray(freespeed, dtype=np.float32)                       # :121, fp32 division
    maxn = (length * np.asarray(permlanes, dtype=np.float32) / np.float32(effective_cell_size)).astype(np.int64) + 1   # :122-124
    Nmax = int(maxn.max() + 1) if n else 1                                         # :128
    h = FeatureHelpers(Nmax=Nmax)
    inters = sorted(set(frm) | set(to))
    rank = {v: k for k, v in enumerate(inters)}
    n_nodes = n + 2 * len(inters)
    x = torch.zeros((n_nodes, h.num_features), dtype=torch.float32)
    x[:n, h.ROAD_INDEX] = torch.arange(n, dtype=torch.float32)
    x[:n, h.LENGHT_OF_ROAD] = torch.from_numpy(length)
    x[:n, h.MAX_FLOW] = torch.from_numpy(cap32)
    x[:n, h.FREE_FLOW_TIME_TRAVEL] = torch.from_numpy(fftt)
    x[:n, h.MAX_NUMBER_OF_AGENT] = torch.from_numpy(maxn.astype(np.float32))
    x[n:, h.ROAD_INDEX] = -1.0                                                    # :139-146

    outgoing, incoming = defaultdict(list), defaultdict(list)
    for i in range(n):
        outgoing[frm[i]].append(i)
        incoming[to[i]].append(i)
    r_from, r_to, r_attr = [], [], []
    cap64 = [float(c) for c in capacity]
    for j in range(n):                                                            # :153-168
        outs = outgoing.get(to[j], [])
        total = 0.0
        for _ in outs:
            total += cap64[j]
        for d in outs:
            r_from.append(j)
            r_to.append(d)
            r_attr.append(cap64[j] / (total if total > 0 else 1.0))
    f_from, f_to = list(r_from), list(r_to)
    for v in inters:                                                              # SRC(v) -> outgoing roads, :179-183
        for road in outgoing.get(v, []):
            f_from.append(n + 2 * rank[v])
            f_to.append(road)
    for v in inters:                                                              # incoming roads -> DEST(v), :186-190
        for road in incoming.get(v, []):
            f_from.append(road)
            f_to.append(n + 2 * rank[v] + 1)
    edge_index_routes = torch.tensor([r_from, r_to], dtype=torch.long).reshape(2, -1)
    edge_attr_routes = torch.tensor(r_attr, dtype=torch.float32).view(-1, 1)
    edge_index = torch.tensor([f_from, f_to], dtype=torch.long).reshape(2, -1)
    edge_attr = torch.tensor(r_attr + [0.0] * (len(f_from) - len(r_from)), dtype=torch.float32).view(-1, 1)

    critical_number = x[:, h.MAX_FLOW] * x[:, h.FREE_FLOW_TIME_TRAVEL] / 3600      # :207-210
    congestion_constant = x[:, h.FREE_FLOW_TIME_TRAVEL] * (x[:, h.MAX_NUMBER_OF_AGENT] + 10 - critical_number)
    g = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, edge_index_routes=edge_index_routes,
             edge_attr_routes=edge_attr_routes, num_roads=n, critical_number=critical_number,
             congestion_constant=congestion_constant)
    if dense is None:
        dense = n_nodes <= DENSE_LIMIT
    if dense:                                                                     # :196-204
        adj = torch.zeros((n_nodes, n_nodes), dtype=torch.bool)
        adj[edge_index[0], edge_index[1]] = True
        src_rows = torch.arange(n, n_nodes, 2, dtype=torch.long)
        src_adj = adj[src_rows, :n].to(torch.float32)
        deg = src_adj.sum(dim=1, keepdim=True)
        g.adj_matrix = adj
        g.src_adj = torch.where(deg > 0, src_adj / deg, torch.zeros_like(src_adj))
    g.intersections = inters
    return g, Nmax


def network_from_xml(prefix: str, dense=None):
    """config_network's parse (src/transportation_simulator.py:61-228). Returns (Data on cpu, Nmax)."""
    root = parse_xml(actual_path(prefix))
    links = root.find("links")
    try:
        cell = float(links.get("effectivecellsize"))
    except (TypeError, ValueError):
        cell = 7.5
    frm, to, length, cap, speed, lanes = [], [], [], [], [], []
    for link in links:
        a = link.attrib
        frm.append(a["from"]); to.append(a["to"])
        length.append(float(a["length"])); cap.append(float(a["capacity"]))
        speed.append(float(a["freespeed"])); lanes.append(float(a["permlanes"]))
    return graph_from_links(frm, to, length, cap, speed, lanes, cell, dense=dense)


def _departure_seconds(act) -> int:
    s = act.get("end_time")
    if not s:
        return 0
    for fmt in ("%H:%M:%S", "%H:%M"):
        try:
            t = datetime.strptime(s, fmt)
            return t.hour * 3600 + t.minute * 60 + t.second
        except ValueError:
            continue
    return 0


def _person_attributes(person) -> dict:
    attrs = dict(person.attrib)
    block = person.find("attributes")
    if block is not None:
        for a in block.findall("attribute"):
            if a.get("name") and a.text:
                attrs[a.get("name")] = a.text
    attrs.setdefault("car_avail", attrs.get("carAvail", "always"))
    attrs.setdefault("sex", "m")
    attrs.setdefault("employed", "no")
    attrs.setdefault("age", "20")
    return attrs


def population_from_xml(scenario_dir: str, verbose: bool = True):
    """Rows of agent_features (src/agents/base.py:38-242): one row per trip of every person with car_avail ==
    "always"; ORIGIN = SRC node of the activity's intersection, DESTINATION = DEST node of the next activity's;
    an activity whose `link` is not an intersection id falls back to the nearest intersection of its x/y."""
    population = parse_xml(actual_path(os.path.join(scenario_dir, "population")))
    network = parse_xml(actual_path(os.path.join(scenario_dir, "network")))
    nodes, links = network.find("nodes"), network.find("links")
    if nodes is None:
        raise ValueError("The XML file does not contain a 'nodes' element.")
    if links is None:
        raise ValueError("The XML file does not contain a 'links' element.")
    pos = {n.get("id"): (float(n.get("x")), float(n.get("y"))) for n in nodes}
    num_links = len(links)
    inters = sorted({l.get("from") for l in links} | {l.get("to") for l in links})
    index = {v: (num_links + 2 * k, num_links + 2 * k + 1) for k, v in enumerate(inters)}
    coords = np.array([pos[v] for v in inters], dtype=np.float64).reshape(-1, 2)

    def nearest(xs, ys):
        d = ((coords - np.array([float(xs), float(ys)])) ** 2).sum(axis=1)
        return inters[int(np.argmin(d))]

    rows = [[0.0, 0.0, 25 * 3600, 0.0, 20.0, 0.0, 0.0, 0.0, 0.0]]                 # dummy agent, :131-133
    persons = selected = 0
    for person in population:
        persons += 1
        attrs = _person_attributes(person)
        if attrs.get("car_avail", attrs.get("carAvail", "")).lower() != "always":
            continue
        plan = person.find("plan")
        if plan is None:
            continue
        acts = plan.findall("act") or plan.findall("activity")
        if len(acts) < 2:
            continue
        sex = 1 if attrs.get("sex", "m").lower() == "f" else 0
        employed = 1 if attrs.get("employed", "no").lower() == "yes" else 0
        age = float(attrs.get("age", 0))
        trips = 0
        for a, b in zip(acts[:-1], acts[1:]):
            o, d = a.get("link"), b.get("link")
            if o not in index and a.get("x") is not None and a.get("y") is not None:
                try:
                    o = nearest(a.get("x"), a.get("y"))
                except (TypeError, ValueError):
                    pass
            if d not in index and b.get("x") is not None and b.get("y") is not None:
                try:
                    d = nearest(b.get("x"), b.get("y"))
                except (TypeError, ValueError):
                    pass
            if o not in index or d not in index:
                continue
            rows.append([float(index[o][0]), float(index[d][1]), float(_departure_seconds(a)), 0.0, age, float(sex),
                         float(employed), 0.0, 0.0])
            trips += 1
        selected += trips > 0
    if verbose:
        print(f"population: {selected}/{persons} persons selected, {len(rows) - 1} trips")
    return rows
