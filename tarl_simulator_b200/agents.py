"""Agents: the population table and the per-step insert / withdraw / choice operations around the core step.

Drop-in for the reference's src/agents/base.py (`Agents`): same constructor, attributes (`agent_features`, `time`,
`withdraw_history`) and method signatures. The population table `agent_features` is fp32 [A+1, 9] with the columns
of AgentFeatureHelpers; row 0 is a dummy agent that never departs. The three per-step operations run as CUDA kernels
behind the C ABI (csrc/agents.cu), in place on `graph.x` and `agent_features`; there is no PyTorch or CPU
implementation of them in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch

from . import _cabi
from ._arena import StepArena
from .feature_helpers import AgentFeatureHelpers


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class PopulationIndex:
    """The population indexed by ORIGIN node (struct tarl_agent_index): agents of one origin in ascending agent id.
    ORIGIN is static, so this is built once per `agent_features` tensor (torch ops on its device: plumbing)."""

    def __init__(self, agent_features: torch.Tensor, n_nodes: int):
        origin = agent_features[..., AgentFeatureHelpers.ORIGIN].reshape(-1, agent_features.size(-2))[0].long()
        if origin.numel() and (int(origin.min()) < 0 or int(origin.max()) >= n_nodes):
            raise IndexError("agent ORIGIN outside the graph's nodes")
        dev = agent_features.device
        self.org_agent = torch.argsort(origin, stable=True).to(torch.int32)
        counts = torch.bincount(origin, minlength=n_nodes)
        self.org_ptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=dev)
        self.org_ptr[1:] = torch.cumsum(counts, 0)
        self.origins = torch.nonzero(counts > 0).flatten().to(torch.int32)
        self.n_origins = int(self.origins.numel())
        self.n_nodes = n_nodes
        # DEPARTURE_TIME of every origin's agents in ascending order (static: lets insertion count the departed
        # agents of an origin without reading their rows, see tarl_agents_insert's `inserted`)
        dep = agent_features[..., AgentFeatureHelpers.DEPARTURE_TIME].reshape(-1, agent_features.size(-2))[0].to(torch.float32)
        by_dep = torch.argsort(dep, stable=True)
        order = by_dep[torch.argsort(origin[by_dep], stable=True)]
        self.dep_sorted = dep[order].contiguous()
        self.struct = _cabi.AgentIndex(n_nodes, self.n_origins, self.org_ptr.data_ptr(), self.org_agent.data_ptr(),
                                       self.origins.data_ptr(), self.dep_sorted.data_ptr())

    def ref(self):
        return C.byref(self.struct)


class GraphSideTables:
    """Sparse forms of what the reference reads from dense [N_tot, N_tot] matrices, derived once per graph from the
    full edge_index (which is what config_network builds adj_matrix from, transportation_simulator.py:196-198):
    `adj` = CSR by source node (adj_matrix[ROAD_INDEX, DESTINATION] of base.py:361), and the neighbour lists `choice`
    draws from (adj_matrix[:N,:N] rows and src_adj rows, base.py:461-476)."""

    def __init__(self, graph):
        N = int(graph.num_roads)
        n_nodes = graph.x.size(-2)
        dev = graph.x.device
        ei = graph.edge_index if hasattr(graph, "edge_index") and graph.edge_index is not None else graph.edge_index_routes
        ei = ei.to(dev)
        src, dst = ei[0].long(), ei[1].long()
        order = torch.argsort(src * max(n_nodes, 1) + dst)
        self.adj_idx = dst[order].to(torch.int32)
        self.adj_ptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=dev)
        self.adj_ptr[1:] = torch.cumsum(torch.bincount(src, minlength=n_nodes), 0)
        self.adj = _cabi.CSR(n_nodes, int(src.numel()), self.adj_ptr.data_ptr(), self.adj_idx.data_ptr(), None)
        # neighbour CSR over all nodes: unique (node, road) pairs, ascending road id inside a node
        # (road rows of adj_matrix[:N,:N], and the SRC rows N, N+2, ... that src_adj covers — base.py:461,470)
        to_road = (dst < N) & ((src < N) | ((src - N) % 2 == 0))
        key = torch.unique(src[to_road] * N + dst[to_road])          # sorted: by node, then by road
        nsrc, ndst = key // max(N, 1), key % max(N, 1)
        self.nbr_idx = ndst.to(torch.int32)
        deg = torch.bincount(nsrc, minlength=n_nodes)
        self.nbr_ptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=dev)
        self.nbr_ptr[1:] = torch.cumsum(deg, 0)
        self.choosers = torch.nonzero(deg > 0).flatten().to(torch.int32)
        self.n_choosers = int(self.choosers.numel())
        self.nbr = _cabi.CSR(n_nodes, int(key.numel()), self.nbr_ptr.data_ptr(), self.nbr_idx.data_ptr(), None)
        self.src32, self.dst32 = src.to(torch.int32).contiguous(), dst.to(torch.int32).contiguous()
        self.N, self.n_nodes = N, n_nodes


def side_tables_for(graph) -> GraphSideTables:
    ei = graph.edge_index if hasattr(graph, "edge_index") and graph.edge_index is not None else graph.edge_index_routes
    key = (id(ei), ei._version, ei.data_ptr(), int(graph.num_roads), graph.x.size(-2), str(graph.x.device))
    hit = getattr(graph, "_tarl_side", None)
    if hit is None or hit[0] != key:
        hit = (key, GraphSideTables(graph))
        graph._tarl_side = hit
    return hit[1]


def rows_state(graph, Nmax: int, with_cc: bool = True) -> _cabi.AgentState:
    """struct tarl_agent_state for the reference row layout of graph.x ([N_tot, F] or [R, N_tot, F]). Every caller
    hands the raw pointer to a kernel that writes the rows: the graph is told (Data.rows_written) so that a resident
    link store shadowing them (core.py) knows its copy is stale."""
    x = graph.x
    if hasattr(graph, "rows_written"):
        graph.rows_written()
    if not x.is_cuda:
        raise RuntimeError("tarl_simulator_b200 computes on CUDA devices only (no CPU fallback): move graph.x to cuda")
    F = 3 * Nmax + 7
    if x.dtype != torch.float32 or x.size(-1) != F or (x.numel() and x.stride(-1) != 1):
        raise ValueError(f"graph.x must be fp32 [.., N_tot, {F}] with contiguous rows")
    N = int(graph.num_roads)
    n_nodes = x.size(-2)
    R = x.size(0) if x.dim() == 3 else 1
    st = _cabi.AgentState()
    st.x, st.x_row_stride = x.data_ptr(), (x.stride(-2) if n_nodes > 1 else F)
    st.x_replica_stride = x.stride(0) if x.dim() == 3 and R > 1 else 0
    st.n_links, st.nmax, st.n_replicas, st.n_nodes = N, Nmax, R, n_nodes
    cc = None
    if with_cc and hasattr(graph, "congestion_constant") and graph.congestion_constant is not None:   # base.py:314
        cc = graph.congestion_constant
        if cc.dtype != torch.float32 or not cc.is_contiguous() or cc.device != x.device:
            cc = cc.to(device=x.device, dtype=torch.float32).contiguous()
        graph._tarl_cc_keepalive = cc
    st.cc = cc.data_ptr() if cc is not None else None
    st.store, st.src_sel, st.t_garbage = None, None, 0.0
    return st


class Agents(AgentFeatureHelpers):
    def __init__(self, device):
        super().__init__()
        self.agent_features = None
        self.time = 0
        self.device = device
        self.withdraw_history: list = []
        self.keep_history = True        # False: the masks only feed on-device counters (metrics.LinkMetrics)
        self.last_withdrawn = None      # bool[N] of the latest withdraw call (device)
        self._index = None
        self._scratch = {}
        self._flags = None
        self._choice_calls = 0
        self.choice_seed = 0
        self._arena = StepArena()

    def set_time(self, time):
        self.time = time

    def reset(self):
        """src/agents/base.py:497-503"""
        self.agent_features[..., self.ON_WAY] = 0.0
        self.agent_features[..., self.DONE] = 0.0
        self.withdraw_history = []

    # --------------------------------------------------------------------------------------------------- plumbing
    def _table(self, R: int) -> _cabi.AgentTable:
        af = self.agent_features
        if af is None:
            raise RuntimeError("agent_features is not set (load a population first)")
        if not af.is_cuda:
            raise RuntimeError("tarl_simulator_b200 computes on CUDA devices only: move agent_features to cuda")
        if af.dtype != torch.float32 or af.size(-1) != 9 or not af.is_contiguous():
            raise ValueError("agent_features must be a contiguous fp32 [.., A+1, 9] tensor")
        if (af.size(0) if af.dim() == 3 else 1) != R:
            raise ValueError("agent_features and graph.x disagree on the number of replicas")
        return _cabi.AgentTable(af.data_ptr(), af.stride(0) if af.dim() == 3 else 0, af.size(-2), 0)

    def population_index(self, n_nodes: int) -> PopulationIndex:
        """Cached by tensor identity; in-place edits of ORIGIN need `invalidate_index()` (ON_WAY / DONE / times do
        not: the kernels read them live)."""
        af = self.agent_features
        key = (id(af), af.data_ptr(), tuple(af.shape), n_nodes)
        if self._index is None or self._index[0] != key or self._index[1]() is not af:
            self._index = (key, weakref.ref(af), PopulationIndex(af, n_nodes))
        return self._index[2]

    def invalidate_index(self):
        self._index = None

    def _buf(self, name, numel, dtype, device, fill=None):
        b = self._scratch.get(name)
        if b is None or b.numel() < numel or b.device != device:
            b = torch.empty(max(numel, 1), dtype=dtype, device=device)
            if fill is not None:
                b.fill_(fill)
            self._scratch[name] = b
        return b

    def _flag_words(self, device):
        if self._flags is None or self._flags.device != device:
            self._flags = torch.zeros(_cabi.FLAG_COUNT, dtype=torch.int32, device=device)
        return self._flags

    def check_errors(self):
        """Synchronises; raises if an insert / withdraw hit a data-dependent fault."""
        if self._flags is not None:
            bits = int(self._flags[_cabi.FLAG_ERROR])
            if bits:
                self._flags.zero_()
                raise IndexError(_cabi.decode_error_bits(bits))

    # ------------------------------------------------------------------------------------------------- operations
    def insert_agent_into_network(self, graph, h) -> torch.Tensor:
        """src/agents/base.py:244-331, in place on graph.x and agent_features."""
        st = rows_state(graph, h.Nmax)
        R, N, dev = st.n_replicas, st.n_links, graph.x.device
        tab = self._table(R)
        idx = self.population_index(st.n_nodes)
        head = self._buf("head", R * N, torch.int32, dev, fill=-1)
        nxt = self._buf("next", R * idx.n_origins, torch.int32, dev)
        cur = self._buf("cursor", R * idx.n_origins, torch.int32, dev)
        flags = self._flag_words(dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_agents_insert(C.byref(st), C.byref(tab), idx.ref(), float(self.time),
                                                head.data_ptr(), nxt.data_ptr(), cur.data_ptr(), None, None,
                                                flags.data_ptr(), None, None, None, None, None, _stream(dev))
        _cabi.check(rc, "tarl_agents_insert")
        return graph.x

    def withdraw_agent_from_network(self, graph, h) -> torch.Tensor:
        """src/agents/base.py:334-403, in place; appends (time, bool[N]) to withdraw_history every call."""
        st = rows_state(graph, h.Nmax, with_cc=False)
        R, N, dev = st.n_replicas, st.n_links, graph.x.device
        tab = self._table(R)
        side = side_tables_for(graph)
        mask, _ = self._arena.take(R * N, dev)
        flags = self._flag_words(dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_agents_withdraw(C.byref(st), C.byref(tab), C.byref(side.adj),
                                                  float(self.time), mask.data_ptr(), None, flags.data_ptr(), None, None,
                                                  _stream(dev))
        _cabi.check(rc, "tarl_agents_withdraw")
        self.last_withdrawn = mask if graph.x.dim() == 2 else mask.view(R, N)
        if self.keep_history:
            self.withdraw_history.append((self.time, self.last_withdrawn))
        return graph.x

    @torch.no_grad()
    def choice(self, graph, h, uniforms: torch.Tensor | None = None):
        """src/agents/base.py:446-494: a uniformly random downstream road for every road and SRC node that has one.
        `uniforms` (optional, [n_choosers] or [R, n_choosers], chooser order = ascending node id) injects the noise;
        otherwise an in-kernel Philox stream keyed by `self.choice_seed` and the call count is used."""
        st = rows_state(graph, h.Nmax, with_cc=False)
        dev = graph.x.device
        side = side_tables_for(graph)
        up = None
        if uniforms is not None:
            uniforms = uniforms.to(device=dev, dtype=torch.float32).contiguous()
            if uniforms.numel() != st.n_replicas * side.n_choosers:
                raise ValueError("uniforms must hold one value per replica and choosing node")
            up = uniforms.data_ptr()
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_agents_choice(C.byref(st), C.byref(side.nbr), side.choosers.data_ptr(),
                                                side.n_choosers, up, int(self.choice_seed), self._choice_calls,
                                                _stream(dev))
        _cabi.check(rc, "tarl_agents_choice")
        self._choice_calls += 1
        return graph

    # ------------------------------------------------------------------------------------------------ persistence
    def save(self, file_path: str) -> None:
        os.makedirs(os.path.dirname(file_path), exist_ok=True)
        torch.save(self.agent_features, file_path)

    def load(self, scenario: str) -> None:
        """save/<scenario>/population.pt (a bare tensor), else parse data/<scenario>/population.xml(.gz)
        (src/agents/base.py:420-444)."""
        file_path = os.path.join("save", scenario, "population.pt")
        try:
            obj = torch.load(file_path, weights_only=True, map_location=self.device)
            if not isinstance(obj, torch.Tensor):
                raise TypeError(f"Expected a Tensor for 'agent_features', got {type(obj)}.")
            self.agent_features = obj
        except FileNotFoundError:
            self.config_agents_from_xml(scenario)
            self.save(file_path)
        self.agent_features[0, self.DEPARTURE_TIME] = 48 * 3600     # agent 0 never joins the network
        self.invalidate_index()

    def config_agents_from_xml(self, scenario: str, *, verbose: bool = True) -> None:
        """src/agents/base.py:38-242 (host-side parsing, see matsim_io.py)."""
        from .matsim_io import population_from_xml
        rows = population_from_xml(os.path.join("data", scenario), verbose=verbose)
        self.agent_features = torch.tensor(rows, dtype=torch.float32, device=self.device)
        self.invalidate_index()
