"""RL environments over the simulator.

`SimulatorEnv` is the drop-in for the reference's src/reinforcement_learning.py:102-309 (one environment, state in
the reference's `graph.x` layout). torchrl / tensordict are not part of this stack: observations and step results
are plain dicts of tensors with the reference's keys ("node_features", "edge_features", "agent_index", "time",
"reward", "done", "terminated"), and `reset()` / `step()` / `rollout()` stand where EnvBase's would.

`BatchedSimulatorEnv` is the B200-first form of the same environment for PPO rollouts: R independent replicas of one
network advance together, state in the resident link store (engine.LinkStore) — per step one launch each for
action → SELECTED_ROAD, the two core-step kernels, withdraw, the two insert kernels and the observation/reward
kernel, for all replicas at once, with no host synchronisation.

One `_step` = apply action → core step → withdraw → insert → reward = −Σ NUM → time += timestep
(src/reinforcement_learning.py:222-276; time advances every step: declared divergence D6, the reference's
"only if nothing moved" test compares a view with itself).
"""
from __future__ import annotations

import ctypes as C
import time as _time
from types import SimpleNamespace

import torch

from . import _cabi
from .agents import Agents, PopulationIndex, rows_state, side_tables_for
from .engine import LinkStore
from .topology import group_csr_for
from .transportation_simulator import TransportationSimulator

EPISODE_START = 3600 * 6 - 60       # _reset, :203
EPISODE_END = 7 * 3600              # done test, :273


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _action_code(action: torch.Tensor):
    """2-D [R, E] view of the action (any strides) in a dtype the kernel reads directly."""
    if action.dtype == torch.bool:
        return action.view(torch.uint8), _cabi.ACTION_U8
    if action.dtype == torch.uint8:
        return action, _cabi.ACTION_U8
    if action.dtype == torch.int64:
        return action, _cabi.ACTION_I64
    return action.to(torch.float32), _cabi.ACTION_F32


class SimulatorEnv:
    def __init__(self, device: str = "cuda", timestep_size: int = 1, start_time: int = 0, scenario: str = "Easy",
                 torch_compile: bool = False, simulator: TransportationSimulator | None = None):
        self.device = torch.device(device)
        if simulator is None:
            simulator = TransportationSimulator(device=device, torch_compile=torch_compile)
            simulator.load_network(scenario=scenario)
        self.simulator = simulator
        self.simulator.config_parameters(timestep_size=timestep_size, start_time=start_time)
        g = self.simulator.graph
        self.num_edge = g.edge_index.size(1)
        self.num_node = g.x.size(0)
        self.num_obs = 7
        self.action_spec = SimpleNamespace(shape=torch.Size([self.num_edge]), dtype=torch.bool)
        self.reward_spec = SimpleNamespace(shape=torch.Size([1]), dtype=torch.float32, low=-1e6, high=1e6)
        self.observation_spec = {
            "node_features": SimpleNamespace(shape=(self.num_node, self.num_obs), dtype=torch.float32),
            "edge_features": SimpleNamespace(shape=(self.num_edge, 1), dtype=torch.float32),
            "agent_index": SimpleNamespace(shape=(self.num_node,), dtype=torch.int64),
            "time": SimpleNamespace(shape=(1,), dtype=torch.float32),
        }
        self.batch_size = torch.Size([])
        self.state = self.simulator.state()
        self.noise = None              # optional injected core-step uniforms for the next step ([E])

    def set_seed(self, seed):
        self.rng = torch.Generator(device=self.device)
        self.rng.manual_seed(seed)
        return seed

    _set_seed = set_seed

    def _observation(self):
        x, edge_attr, _, agent_index = self.simulator.state()
        return {"node_features": x, "edge_features": edge_attr, "agent_index": agent_index,
                "time": torch.tensor([self.simulator.time], dtype=torch.float32, device=self.device)}

    def _reset(self, tensordict=None):
        sim = self.simulator
        sim.reset()
        sim.inserting_time = sim.choice_time = sim.core_time = sim.withdraw_time = 0
        sim.leg_histogram_values = []
        sim.road_optimality_values = []
        sim.on_way_before = 0
        sim.done_before = 0
        sim.model_core.response_mpnn.update_history = []
        if sim.metrics is not None:
            sim.metrics.reset()
        sim.set_time(EPISODE_START)
        out = self._observation()
        sim.agent.reset()
        out["terminated"] = torch.tensor([False])
        out["done"] = torch.tensor([False])
        return out

    reset = _reset

    def _step(self, tensordict):
        sim, h = self.simulator, self.simulator.h
        g = sim.graph
        dev = g.x.device
        action, code = _action_code(tensordict["action"].to(dev).reshape(1, -1))
        side = side_tables_for(g)
        b = _time.time()
        st = rows_state(g, h.Nmax, with_cc=False)
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_agents_apply_action(C.byref(st), side.src32.data_ptr(), side.dst32.data_ptr(),
                                                      self.num_edge, _cabi.rows(action), code, _stream(dev))
        _cabi.check(rc, "tarl_agents_apply_action")
        e = _time.time(); sim.choice_time += e - b; b = e
        sim.graph = sim.model_core(g) if self.noise is None else sim.model_core(g, noise=self.noise)
        e = _time.time(); sim.core_time += e - b; b = e
        g.x = sim.agent.withdraw_agent_from_network(g, h)
        e = _time.time(); sim.withdraw_time += e - b; b = e
        g.x = sim.agent.insert_agent_into_network(g, h)
        e = _time.time(); sim.inserting_time += e - b
        reward = (-torch.sum(g.x[:, h.NUMBER_OF_AGENT])).flatten()                                   # :266-267
        if sim.node_metrics:
            sim.link_metrics().record(sim.time, pop=sim.model_core.last_pop, withdrawn=sim.agent.last_withdrawn,
                                      delta_tt=sim.model_core.direction_mpnn.road_optimality_data["delta_travel_time"])
        sim.set_time(sim.time + sim.timestep)                                                        # D6
        done = torch.tensor(sim.time > EPISODE_END)
        af = sim.agent.agent_features
        value_on_way, value_done = torch.sum(af[:, sim.agent.ON_WAY]), torch.sum(af[:, sim.agent.DONE])
        sim.leg_histogram_values.append([value_on_way - sim.on_way_before + value_done - sim.done_before,
                                         value_done - sim.done_before, value_on_way, sim.time])
        sim.on_way_before, sim.done_before = value_on_way, value_done
        if sim.record_road_optimality:
            sim.road_optimality_values.append(
                (sim.time, sim.model_core.direction_mpnn.road_optimality_data["delta_travel_time"].cpu()))
        out = self._observation()
        out.update(reward=reward, terminated=done, done=done)
        return out

    def step(self, tensordict):
        out = dict(tensordict)
        out["next"] = self._step(tensordict)
        return out

    @torch.no_grad()
    def rollout(self, max_steps, policy=None, break_when_any_done=True):
        """Reset, then up to `max_steps` steps with `policy(obs) -> dict containing "action"`. Returns the list of
        transitions (dicts with the observation keys, "action", and "next")."""
        obs = self._reset()
        traj = []
        for _ in range(max_steps):
            td = dict(obs)
            td.update(policy(obs) if policy is not None else {"action": self.random_action()})
            td = self.step(td)
            traj.append(td)
            obs = {k: v for k, v in td["next"].items() if k not in ("reward",)}
            if break_when_any_done and bool(td["next"]["done"]):
                break
        return traj

    def random_action(self):
        """A uniformly random valid action: one out-edge per source node."""
        from .distribution import GraphDistribution
        logits = torch.zeros(self.num_edge, device=self.simulator.graph.x.device)
        return GraphDistribution(logits, self.simulator.graph.edge_index).sample(dtype=torch.bool)


class BatchedSimulatorEnv:
    """R replicas of one network on one GPU, state in the resident link store. Static inputs (topology, link
    attributes, the population's origins / destinations / departure times) are shared; per replica: queues,
    SELECTED_ROAD, the ON_WAY / DONE / ARRIVAL_TIME columns of its own agent_features copy, its noise stream.

    step(action [R, E_full] one-hot) -> dict(reward [R], done [R], occupancy int32 [R]); `observe()` materialises
    node_features [R, N_tot, 7] / agent_index [R, N_tot] on demand (the policy's active path only needs ROAD_INDEX,
    the value net only NUM: `num_agents()` is a strided view)."""

    def __init__(self, graph, Nmax: int, agent_features: torch.Tensor, replicas: int, timestep: float = 1,
                 seed: int = 0, cluster: bool | None = None, first_replica: int = 0):
        """first_replica: global index of this environment's first replica when a job shards its replicas over several
        ranks (parallel.shard_replicas): it enters every in-kernel noise stream, so that no two ranks draw the same
        actions or the same hand-off noise from one job-wide seed."""
        dev = graph.x.device
        if dev.type != "cuda":
            raise RuntimeError("BatchedSimulatorEnv lives on a CUDA device (no CPU fallback)")
        self.graph, self.Nmax, self.R, self.device = graph, int(Nmax), int(replicas), dev
        self.N, self.n_nodes = int(graph.num_roads), graph.x.size(0)
        self.E_full = graph.edge_index.size(1)
        self.timestep = timestep
        self.first_replica = int(first_replica)
        self.seed = int(seed)
        salt = (self.first_replica * 0x9E3779B97F4A7C15) & ((1 << 62) - 1)
        self.store = LinkStore.from_graph(graph, Nmax, replicas=replicas, seed=(int(seed) + salt) & ((1 << 62) - 1),
                                          cluster=cluster)
        self._salt = salt
        self.seed_words = None          # int64 [2] on the device once begin_rollout() was called: {core noise, sampling}
        F = 3 * self.Nmax + 7
        self.src_sel = graph.x[self.N:, F - 2].to(torch.float32).repeat(self.R, 1).contiguous()
        self.agent_features = agent_features.to(dev, torch.float32).unsqueeze(0).repeat(self.R, 1, 1).contiguous()
        self.index = PopulationIndex(self.agent_features, self.n_nodes)
        self.side = side_tables_for(graph)
        i32 = dict(dtype=torch.int32, device=dev)
        self._head = torch.full((max(self.R * self.N, 1),), -1, **i32)
        self._next = torch.empty(max(self.R * self.index.n_origins, 1), **i32)
        self._cursor = torch.empty(max(self.R * self.index.n_origins, 1), **i32)
        self._inserted = torch.zeros(max(self.R * self.index.n_origins, 1), **i32)   # per (replica, origin), see insert()
        self._work = torch.empty(max(self.R * self.index.n_origins, 1), **i32)       # listed origins of a step, compacted
        self._work_count = torch.zeros(self.R + self.index.n_origins, **i32)         # + departed-per-origin scratch
        self.counters = torch.zeros(self.R, 2, **i32)          # running totals {inserted, withdrawn} per replica
        self.occupancy = torch.zeros(self.R, **i32)
        self.withdrawn = torch.zeros(self.R, self.N, dtype=torch.bool, device=dev)
        self.delta_tt = None                                   # optional [R, E] output of the core step
        self.metrics = None                                    # metrics.LinkMetrics once enable_metrics() was called
        self.time = float(EPISODE_START)
        self._table = _cabi.AgentTable(self.agent_features.data_ptr(), self.agent_features.stride(0),
                                       self.agent_features.size(1), 0)
        self._road_origin = self._exclusive_road_origins()

    def _exclusive_road_origins(self):
        """int32 [N]: for every road the index (into the population index's origins) of the ONLY origin node that has an
        edge into it in the full graph, -1 if none — or None when some road can be reached from two origins (or
        TARL_NO_DIRECT_INSERT is set). config_network's graphs qualify: a road leaves one intersection, whose SRC node
        alone points at it. tarl_agents_insert then runs offer and admit as one kernel (identical results)."""
        import os
        if os.environ.get("TARL_NO_DIRECT_INSERT") or self.index.n_origins == 0:
            return None
        # an origin that is itself a road selects through store.sel and could name any road: only the dummy agent of row
        # 0 (src/agents/base.py: departs after the episode) may sit on one — should it ever be ready, the kernel raises
        from .feature_helpers import AgentFeatureHelpers
        real = self.agent_features[0, 1:, AgentFeatureHelpers.ORIGIN]
        if real.numel() and int(real.min()) < self.N:
            return None
        ei = self.graph.edge_index.to(self.device)
        origin_rank = torch.full((self.n_nodes,), -1, dtype=torch.int64, device=self.device)
        origin_rank[self.index.origins.long()] = torch.arange(self.index.n_origins, device=self.device)
        into_road = (ei[1] < self.N) & (ei[0] >= self.N)        # edges from non-road nodes (SRC / DEST) into roads
        src, dst = ei[0][into_road], ei[1][into_road]
        # every non-road node with an edge into a road may come to select it (actions / choice draw among out-edges);
        # those that own no agent never insert, but they do not make a road ambiguous either
        if src.numel() == 0:
            return None
        counts = torch.bincount(dst, minlength=self.N)
        if int(counts.max()) > 1:
            return None
        out = torch.full((self.N,), -1, dtype=torch.int32, device=self.device)
        out[dst] = origin_rank[src].to(torch.int32)
        return out

    def _state(self) -> _cabi.AgentState:
        s = self.store
        s._fill_struct()
        st = _cabi.AgentState()
        st.x, st.x_row_stride, st.x_replica_stride = None, 0, 0
        st.n_links, st.nmax, st.n_replicas, st.n_nodes = self.N, self.Nmax, self.R, self.n_nodes
        st.cc = None
        st.store = C.pointer(s._struct)
        st.src_sel = self.src_sel.data_ptr()
        st.t_garbage = float(s.t_last)
        return st

    def begin_rollout(self, seed: int):
        """Puts the noise streams of the coming rollout under ONE host number: the keys of the core step's hand-off
        noise and of the action sampling move into two device words (so that the launches of a rollout can be captured
        in a CUDA graph once and still draw new noise on every replay), and the per-rollout counters that enter the
        Philox counters (step id, draw id) restart at 0. Call before reset()."""
        mask = (1 << 62) - 1
        core = (int(seed) + self._salt) & mask
        sample = ((int(seed) ^ 0x5DEECE66D) + self._salt) & mask
        if self.seed_words is None:
            self.seed_words = torch.zeros(2, dtype=torch.int64, device=self.device)
            self.store.seed_dev = self.seed_words[0:1]
        self.seed_words[0].fill_(core)
        self.seed_words[1].fill_(sample)
        self.store.step_id = 0
        self.store.cur = 0              # a captured rollout has the ping-pong roles of its first step baked in
        sink = self.action_sink()
        if sink is not None:
            sink.seed_dev = self.seed_words[1:2]
            sink.draw_id = 0

    def sel_pair(self, which: int):
        """Rollouts alternate two pairs of SELECTED_ROAD buffers (links [R*N], other nodes [R, N_tot - N]) so that the
        action of step t+1 can be drawn while step t still reads its own: makes pair `which` (0 / 1) the one the
        environment reads and returns (this pair, the other pair)."""
        if getattr(self, "_sel_pairs", None) is None:
            self._sel_pairs = [(self.store.sel, self.src_sel), (self.store.sel.clone(), self.src_sel.clone())]
            self._sel_cur = 0
        self._sel_cur = int(which)
        cur, other = self._sel_pairs[self._sel_cur], self._sel_pairs[self._sel_cur ^ 1]
        self.store.sel, self.src_sel = cur
        return cur, other

    def sync_sel_pairs(self):
        """The pair not in use becomes a copy of the current one (entries no draw ever writes must agree in both)."""
        cur, other = self.sel_pair(getattr(self, "_sel_cur", 0))
        other[0].copy_(cur[0]); other[1].copy_(cur[1])

    def reset(self):
        """_reset (:186-219): empty queues, ON_WAY = DONE = 0, t = 06:00 − 60 s."""
        self.store.clear_queues()
        assert Agents.DONE == Agents.ON_WAY + 1
        self.agent_features[..., Agents.ON_WAY:Agents.DONE + 1] = 0.0      # one strided pass over the table, not two
        self.counters.zero_()
        self._inserted.zero_()
        if self.metrics is not None:
            self.metrics.reset()
        self.time = float(EPISODE_START)

    def set_time(self, t):
        self.time = float(t)

    def enable_metrics(self, optimality: bool = False):
        """Hourly hand-off / withdrawal counters per (replica, link) accumulated on the device every step — the
        batched counterpart of compute_node_metrics' histories (src/transportation_simulator.py:584-613).
        optimality=True also keeps the per-link road-optimality aggregate (needs the [R, E] delta_tt output)."""
        from .metrics import LinkMetrics
        if self.store.slot_link is not None:
            raise NotImplementedError("metrics need the store in link-id order (cluster=False)")
        self.metrics = LinkMetrics(self.graph.edge_index_routes, self.N, replicas=self.R, optimality=optimality,
                                   device=self.device)
        if optimality and self.delta_tt is None:
            self.delta_tt = torch.empty(self.R, self.store.E, dtype=torch.float32, device=self.device)
        return self.metrics

    def apply_action(self, action: torch.Tensor):
        a, code = _action_code(action.reshape(self.R, self.E_full))
        st = self._state()
        grp = group_csr_for(self.graph.edge_index, "source_rank")
        if getattr(self, "_group_nodes", None) is None:
            self._group_nodes = grp.nodes.to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_agents_apply_action_groups(C.byref(st), grp.ref(), self._group_nodes.data_ptr(),
                                                             self.side.dst32.data_ptr(), _cabi.rows(a), code,
                                                             _stream(self.device))
        _cabi.check(rc, "tarl_agents_apply_action_groups")

    def action_sink(self):
        """ActionSink over this environment's SELECTED_ROAD arrays (None when the store keeps the links in its own
        locality order): lets GraphDistribution.sample write the routing decisions in the pass that draws them."""
        if self.store.slot_link is not None:
            return None
        from .distribution import ActionSink
        grp = group_csr_for(self.graph.edge_index, "source_rank")
        sel = self.store.sel
        key = (sel.data_ptr(), id(grp))
        if getattr(self, "_sink_key", None) != key:
            old = getattr(self, "_sink", None)
            self._sink = ActionSink(self.graph.edge_index, grp, sel[: self.R * self.N].view(self.R, self.N),
                                    self.src_sel if self.n_nodes > self.N else None, self.N, self.n_nodes,
                                    row_offset=(self.first_replica // 4) * 4)
            self._sink.salt = self._salt
            if old is not None:
                self._sink.seed_dev, self._sink.draw_id = old.seed_dev, old.draw_id
            self._sink_key = key
        return self._sink

    def choice(self, uniforms: torch.Tensor | None = None, seed: int = 0):
        """Random routing for every replica (Agents.choice on the store)."""
        st = self._state()
        up = None
        if uniforms is not None:
            uniforms = uniforms.to(self.device, torch.float32).contiguous()
            up = uniforms.data_ptr()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_agents_choice(C.byref(st), C.byref(self.side.nbr), self.side.choosers.data_ptr(),
                                                self.side.n_choosers, up, seed, self.store.step_id, _stream(self.device))
        _cabi.check(rc, "tarl_agents_choice")

    def withdraw(self, num_out: torch.Tensor | None = None):
        """num_out (optional, contiguous fp32 [R, N_tot]): also leave NUMBER_OF_AGENT of every node after the
        withdrawal there and their sum in self.occupancy — pass the same buffer to the insert() that follows."""
        st = self._state()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_agents_withdraw(C.byref(st), C.byref(self._table), C.byref(self.side.adj), self.time,
                                                  self.withdrawn.data_ptr(), self.counters.data_ptr(),
                                                  self.store.flags.data_ptr(),
                                                  num_out.data_ptr() if num_out is not None else None,
                                                  self.occupancy.data_ptr() if num_out is not None else None,
                                                  _stream(self.device))
        _cabi.check(rc, "tarl_agents_withdraw")

    def insert(self, num_out: torch.Tensor | None = None, direct: bool = False):
        """direct=True: offer and admit as ONE kernel (tarl_agents_insert's road_origin) — only where the graph allows it
        (_exclusive_road_origins) AND the caller knows that every origin's SELECTED_ROAD is one of its own out-roads,
        i.e. it was written by an action or by choice() since the last reset (the initial values of graph.x need not
        be); an origin naming somebody else's road while it has a ready agent raises. Pays at small replica counts (one
        launch and one grid-wide dependency less: 55 -> 43 us per step of 128 grid100 replicas); with many replicas the
        compacted worklist of the two-kernel form is faster."""
        st = self._state()
        road_origin = self._road_origin if direct else None
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_agents_insert(C.byref(st), C.byref(self._table), self.index.ref(), self.time,
                                                self._head.data_ptr(), self._next.data_ptr(), self._cursor.data_ptr(),
                                                self.counters.data_ptr(), self._inserted.data_ptr(),
                                                self.store.flags.data_ptr(), self._work.data_ptr(),
                                                self._work_count.data_ptr(),
                                                num_out.data_ptr() if num_out is not None else None,
                                                self.occupancy.data_ptr() if num_out is not None else None,
                                                road_origin.data_ptr() if road_origin is not None else None,
                                                _stream(self.device))
        _cabi.check(rc, "tarl_agents_insert")

    def observe(self, node_features: bool = True, agent_index: bool = True, compact_out=None):
        """One pass over the link store: occupancy (always), and on request node_features [R, N_tot, 7], agent_index
        [R, N_tot], and/or the compact observation written into compact_out = (NUM [R, N_tot] fp32, SELECTED_ROAD
        [R, N_tot] fp32, head agent id [R, N_tot] int64) — contiguous buffers, e.g. a frame of a trajectory; the second
        and third may be None (a rollout whose nets read the occupancy only keeps 4 of the 16 bytes per node)."""
        st = self._state()
        nf = torch.empty(self.R, self.n_nodes, 7, dtype=torch.float32, device=self.device) if node_features else None
        ai = torch.empty(self.R, self.n_nodes, dtype=torch.int64, device=self.device) if agent_index else None
        num = sel = head = None
        if compact_out is not None:
            num, sel, head = compact_out
            for t_, dt in ((num, torch.float32), (sel, torch.float32), (head, torch.int64)):
                if t_ is not None and (t_.dtype != dt or t_.shape != (self.R, self.n_nodes) or not t_.is_contiguous()):
                    raise ValueError("compact_out must be contiguous (fp32, fp32, int64) [R, N_tot] buffers")
            if ai is None:
                ai = head
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_store_observe(C.byref(st), nf.data_ptr() if nf is not None else None,
                                                ai.data_ptr() if ai is not None else None, self.occupancy.data_ptr(),
                                                num.data_ptr() if num is not None else None,
                                                sel.data_ptr() if sel is not None else None, _stream(self.device))
        _cabi.check(rc, "tarl_store_observe")
        if head is not None and ai is not head:
            head.copy_(ai)
        return nf, (ai if agent_index else None)

    def step(self, action: torch.Tensor | None, noise: torch.Tensor | None = None, observe: bool = False,
             compact_out=None, lean: bool = False, after_core=None, direct_insert: bool = False):
        """One _step for every replica. action None = keep the current SELECTED_ROAD values. compact_out: see
        observe() — the post-step compact observation comes out of the same pass that computes the reward.
        lean=True returns nothing: reward = -self.occupancy (which may be pointed at a row of a caller-owned [T, R] int32
        buffer beforehand), done = self.time > EPISODE_END. after_core: called once the core step (+ fused withdrawal)
        is enqueued, before the insertion (rollouts record an event there for work they overlap with the insertion)."""
        if action is not None:
            self.apply_action(action)
        # occupancy-only observation (compact_out = (NUM frame, None, None)): withdraw holds every record anyway and
        # leaves NUM + the reward behind, insert patches the roads it touches — no separate observe pass
        fused = (not observe and compact_out is not None and compact_out[1] is None and compact_out[2] is None
                 and compact_out[0].dtype == torch.float32 and compact_out[0].shape == (self.R, self.n_nodes)
                 and compact_out[0].is_contiguous())
        if fused and self.n_nodes - self.N <= self.N and self.store.can_fuse_withdraw():
            # ... and the withdrawal itself rides on the response phase of the core step: one pass over the records
            self.store.step(self.time, noise=noise, delta_tt=self.delta_tt,
                            withdraw=dict(table=self._table, adjacency=self.side.adj, n_nodes=self.n_nodes,
                                          mask=self.withdrawn, counters=self.counters, num_out=compact_out[0],
                                          occupancy=self.occupancy))
        else:
            self.store.step(self.time, noise=noise, delta_tt=self.delta_tt)
            self.withdraw(num_out=compact_out[0] if fused else None)
        if after_core is not None:
            after_core()
        self.insert(num_out=compact_out[0] if fused else None, direct=direct_insert)
        if self.metrics is not None:
            self.metrics.record(self.time, pop=self.store.pop[: self.R * self.N], withdrawn=self.withdrawn,
                                delta_tt=self.delta_tt if self.metrics.optimality_now is not None else None)
        nf = ai = None
        if not fused:
            nf, ai = self.observe(node_features=observe, agent_index=observe, compact_out=compact_out)
        self.time += self.timestep
        if lean:            # rollouts: the caller reads self.occupancy / self.time itself (no per-step tensors, no launches)
            return None
        out = {"reward": -self.occupancy.to(torch.float32), "occupancy": self.occupancy,
               "done": torch.full((self.R,), self.time > EPISODE_END, dtype=torch.bool, device=self.device),
               "time": self.time}
        if observe:
            out["node_features"], out["agent_index"] = nf, ai
        return out

    def num_agents(self) -> torch.Tensor:
        return self.store.num_agents()

    def compact_state(self, out=None):
        """The dynamic observation columns without materialising [R, N_tot, 7]: (NUMBER_OF_AGENT [R, N_tot],
        SELECTED_ROAD [R, N_tot], head agent id int64 [R, N_tot]); plain copies out of the store's arrays, optionally
        into three given buffers."""
        R, N, M = self.R, self.N, self.n_nodes
        if out is None:
            out = (torch.empty(R, M, dtype=torch.float32, device=self.device),
                   torch.empty(R, M, dtype=torch.float32, device=self.device),
                   torch.empty(R, M, dtype=torch.int64, device=self.device))
        if all(t_ is None or t_.is_contiguous() for t_ in out) and out[0] is not None:
            self.observe(node_features=False, agent_index=False, compact_out=out)
            return out
        num, sel, head = out
        num[:, :N] = self.store.num_agents()
        num[:, N:] = 0.0
        head[:, :N] = self.store.head_agents()
        head[:, N:] = 0
        sel[:, :N] = self.store.selected_road()
        sel[:, N:] = self.src_sel
        return num, sel, head

    def export_x(self) -> torch.Tensor:
        """The full node table [R, N_tot, F] exactly as the reference would hold it."""
        F = 3 * self.Nmax + 7
        x = self.graph.x.unsqueeze(0).repeat(self.R, 1, 1)
        self.store.export_x(out=x[:, :self.N])
        x[:, self.N:, F - 2] = self.src_sel
        return x

    def check_errors(self):
        self.store.check_errors()
