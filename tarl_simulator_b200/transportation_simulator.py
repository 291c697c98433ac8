"""TransportationSimulator: the classical simulation loop around the core step.

Drop-in for the reference's src/transportation_simulator.py:17-366 (the plotting / CSV side of that file is out of
scope): same constructor, attributes (`graph`, `agent`, `model_core`, `time`, `timestep`, `Nmax`, `h`,
the four phase timers, `leg_histogram_values`, `road_optimality_values`) and methods (`config_network`,
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
        if self.record_road_optimality:
            self.road_optimality_values.append(
                (self.time, self.model_core.direction_mpnn.road_optimality_data["delta_travel_time"].cpu()))

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

    # The reference's plotting / CSV / node-metric methods (:387-763) are out of scope; these keep
    # Runner.eval's call sequence working.
    def plot_leg_histogram(self, output_dir=None):
        return None

    def plot_road_optimality(self, output_dir=None):
        return None

    def plot_computation_time(self, output_dir=None):
        return None

    def plot_daily_counts(self, expected_demand=None, output_dir=None):
        return None

    def compute_node_metrics(self, output_dir=None):
        return {}
