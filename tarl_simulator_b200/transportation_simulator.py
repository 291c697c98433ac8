"""TransportationSimulator: the classical simulation loop around the core step.

Drop-in for the reference's src/transportation_simulator.py:17-366 plus the data side of its metrics
(`compute_node_metrics`, the CSV of `plot_daily_counts`, the per-link series behind `plot_road_optimality`; the
matplotlib figures themselves are out of scope): same constructor, attributes (`graph`, `agent`,
`model_core`, `time`, `timestep`, `Nmax`, `h`, the four phase timers, `leg_histogram_values`,
`road_optimality_values`) and methods (`config_network`,
`save_network`, `load_network`, `configure_core`, `config_parameters`, `set_time`, `run`, `reset`, `state`).
One `run()` = insert → withdraw → choice → core step → time += timestep (:294-351), each phase one or a few CUDA
kernel launches in place on `graph.x` / `agent_features`.
"""
from __future__ import annotations

import os
import time

import torch

from .agents import Agents
from .core import SimulationCoreModel
from .feature_helpers import FeatureHelpers
from .matsim_io import network_from_xml
from .metrics import LinkMetrics, counts_from_histories, node_metrics_from_counts


class TransportationSimulator:
    def __init__(self, device: str, torch_compile: bool = False):
        self.model_core = None
        self.agent = Agents(device)
        self.device = device
        self.torch_compile = torch_compile
        self.graph = None
        self.time = 0
        self.inserting_time = 0
        self.core_time = 0
        self.withdraw_time = 0
        self.choice_time = 0
        self.timestep = 1
        self.node_metrics = False
        self.leg_histogram_values = []
        self.road_optimality_values = []
        self.on_way_before = 0
        self.done_before = 0
        self.record_road_optimality = True     # False: skip the per-step [E] device→host copy of :351
        # node_metrics = True (the reference declares the switch at :52 and never reads it): hourly hand-off /
        # withdrawal counters and the road-optimality aggregate are accumulated on the device every step
        # (metrics.LinkMetrics) and compute_node_metrics reads them instead of the per-step histories.
        self.metrics = None

    # ------------------------------------------------------------------------------------------------ network
    def config_network(self, file_path: str) -> None:
        """MATSim network.xml(.gz) → graph (src/transportation_simulator.py:61-228)."""
        t0 = time.time()
        graph, self.Nmax = network_from_xml(file_path)
        self.h = FeatureHelpers(Nmax=self.Nmax)
        self.graph = graph.to(self.device)
        print(f"Network configured in {time.time() - t0:.2f} seconds "
              f"({graph.num_roads} links, {graph.edge_index_routes.size(1)} dual edges, Nmax={self.Nmax})")

    def save_network(self, file_path: str) -> None:
        os.makedirs(os.path.dirname(file_path), exist_ok=True)
        torch.save({"graph": self.graph, "Nmax": self.Nmax}, file_path)

    def load_network(self, scenario: str) -> None:
        """save/<scenario>/network.pt, else data/<scenario>/network.xml(.gz) (:246-267)."""
        file_path = os.path.join("save", scenario, "network.pt")
        try:
            d = torch.load(file_path, weights_only=False)
            self.graph = d["graph"].to(self.device)
            self.Nmax = d["Nmax"]
        except FileNotFoundError:
            self.config_network(os.path.join("data", scenario, "network"))
            self.save_network(file_path)
        self.h = FeatureHelpers(Nmax=self.Nmax)

    def configure_core(self):
        self.model_core = SimulationCoreModel(self.Nmax, self.device, self.time, torch_compile=self.torch_compile)

    def config_parameters(self, timestep_size: float = 1, start_time: int = 0):
        self.timestep = timestep_size
        self.time = start_time
        self.configure_core()

    def set_time(self, time):
        self.time = time
        self.agent.set_time(time)
        self.model_core.set_time(time)

    # --------------------------------------------------------------------------------------------------- step
    def run(self, noise=None, choice_uniforms=None):
        """One classical timestep (:294-351). `noise` ([E]) and `choice_uniforms` ([n_choosers]) inject the random
        draws of the core step and of `choice`; both default to on-device generators. The phase timers measure host
        enqueue time only (the kernels run asynchronously) unless `self.sync_timers` is set."""
        h = self.h
        sync = getattr(self, "sync_timers", False)
        t_step = self.time                      # the time update_history / withdraw_history stamp this step with

        def lap(b):
            if sync:
                torch.cuda.synchronize()
            return time.time() - b

        b = time.time()
        self.graph.x = self.agent.insert_agent_into_network(self.graph, h)
        self.inserting_time += lap(b)
        b = time.time()
        self.graph.x = self.agent.withdraw_agent_from_network(self.graph, h)
        self.withdraw_time += lap(b)
        b = time.time()
        self.graph = self.agent.choice(self.graph, h) if choice_uniforms is None else \
            self.agent.choice(self.graph, h, uniforms=choice_uniforms)
        self.choice_time += lap(b)
        b = time.time()
        self.graph = self.model_core(self.graph) if noise is None else self.model_core(self.graph, noise=noise)
        self.core_time += lap(b)
        self.set_time(self.time + self.timestep)

        af = self.agent.agent_features
        value_on_way = torch.sum(af[:, self.agent.ON_WAY])
        value_done = torch.sum(af[:, self.agent.DONE])
        self.leg_histogram_values.append([value_on_way - self.on_way_before + value_done - self.done_before,
                                          value_done - self.done_before, value_on_way, self.time])
        self.on_way_before = value_on_way
        self.done_before = value_done
        dtt = self.model_core.direction_mpnn.road_optimality_data["delta_travel_time"]
        if self.node_metrics:
            self.link_metrics().record(t_step, pop=self.model_core.last_pop, withdrawn=self.agent.last_withdrawn,
                                       delta_tt=dtt)
        if self.record_road_optimality:
            self.road_optimality_values.append((self.time, dtt.cpu()))

    def reset(self):
        """:353-358"""
        h = self.h
        torch.zero_(self.graph.x[:, h.AGENT_POSITION])
        torch.zero_(self.graph.x[:, h.AGENT_TIME_DEPARTURE])
        torch.zero_(self.graph.x[:, h.AGENT_TIME_ARRIVAL])
        torch.zero_(self.graph.x[:, h.NUMBER_OF_AGENT])

    def state(self):
        """:360-366 — views into graph.x, no copies."""
        h = self.h
        x = self.graph.x[:, h.MAX_NUMBER_OF_AGENT:]
        agent_index = (self.graph.x[:, h.HEAD_FIFO]).to(torch.int64)
        return x, self.graph.edge_attr, self.graph.edge_index, agent_index

    # ------------------------------------------------------------------------------------------------ metrics
    def link_metrics(self) -> LinkMetrics:
        if self.metrics is None:
            self.metrics = LinkMetrics(self.graph.edge_index_routes, int(self.graph.num_roads), replicas=1,
                                       device=self.graph.x.device)
        return self.metrics

    def hourly_counts(self):
        """int64 [N, num_hours]: steps of each hour on which a link handed off its head or had agents withdrawn —
        from the on-device counters when `node_metrics` is on, else reduced from the two histories exactly as
        compute_node_metrics does (:584-613). None when nothing was recorded."""
        if self.metrics is not None and self.metrics.steps > 0:
            return self.metrics.counts_per_node()
        update_history = getattr(self.model_core.response_mpnn, "update_history", []) if self.model_core else []
        withdraw_history = getattr(self.agent, "withdraw_history", [])
        return counts_from_histories(update_history, withdraw_history, device=self.graph.x.device)

    def compute_node_metrics(self, output_dir: str | None = "data/outputs"):
        """:563-669 — per link: hourly counts, mean and population std over the hours of count / MAX_FLOW; writes
        node_metrics.csv when output_dir is given; returns {link: {'avg_vc', 'std_vc', 'hourly_counts'}}."""
        counts = self.hourly_counts()
        if counts is None:
            print("No update history available for computing node metrics.")
            return {}
        return node_metrics_from_counts(counts, self.graph.x[:, self.h.MAX_FLOW], output_dir)

    def plot_daily_counts(self, expected_counts=None, output_dir: str | None = "data/outputs"):
        """The data side of :672-746 (no figure): simulated daily total per link beside the expected one; writes
        daily_counts.csv (link_id, simulated, expected, difference) when output_dir is given; returns the rows."""
        counts = self.hourly_counts()
        if counts is None or not expected_counts:
            if counts is None:
                print("No update history available for computing node metrics.")
            return {}
        sim_totals = counts.sum(dim=1).cpu()
        n = sim_totals.size(0)
        road_ids = sorted(expected_counts.keys())
        rows = {"link_id": road_ids, "simulated": [int(sim_totals[i]) for i in road_ids],
                "expected": [float(expected_counts[i]) if 0 <= i < n else 0.0 for i in road_ids]}
        rows["difference"] = [s_ - e_ for s_, e_ in zip(rows["simulated"], rows["expected"])]
        if output_dir is not None:
            import pandas as pd
            os.makedirs(output_dir, exist_ok=True)
            pd.DataFrame(rows).to_csv(os.path.join(output_dir, "daily_counts.csv"), index=False)
            print(f"Daily counts CSV saved as {os.path.join(output_dir, 'daily_counts.csv')}")
        return rows

    def road_optimality_series(self):
        """(times_h [T], agg [T, N]) of plot_road_optimality (:482-488): delta_travel_time summed over each link's
        outgoing turns, per recorded step — reduced on the device from the retained [E] vectors."""
        if not self.road_optimality_values:
            return None
        dev = self.graph.x.device
        origin = self.graph.edge_index_routes[0]
        times = torch.tensor([t for t, _ in self.road_optimality_values], dtype=torch.float32) / 3600.0
        v = torch.stack([v for _, v in self.road_optimality_values], dim=0).to(dev)
        agg = torch.zeros(v.size(0), int(self.graph.num_roads), device=dev, dtype=v.dtype)
        agg.scatter_add_(1, origin.unsqueeze(0).expand(v.size(0), -1), v)
        return times, agg

    # The reference's matplotlib figures (:387-561) are out of scope; these keep Runner.eval's and
    # ppo_train's call sequences working.
    def plot_leg_histogram(self, output_dir=None):
        return None

    def plot_road_optimality(self, output_dir=None, road_ids=()):
        return None

    def plot_computation_time(self, output_dir=None):
        return None
