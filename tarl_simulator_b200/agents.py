"""Agents: the population table and the per-step insert / withdraw / choice operations around the core step.

Mirrors the reference's src/agents/base.py (`Agents`). The population table `agent_features` is fp32 [A+1, 9] with
the columns of AgentFeatureHelpers; row 0 is a dummy agent that never departs.
"""
from __future__ import annotations

import os

import torch

from .feature_helpers import AgentFeatureHelpers


class Agents(AgentFeatureHelpers):
    def __init__(self, device):
        super().__init__()
        self.agent_features = None
        self.time = 0
        self.device = device
        self.withdraw_history: list = []

    def set_time(self, time):
        self.time = time

    def reset(self):
        """src/agents/base.py:497-503"""
        self.agent_features[:, self.ON_WAY] = 0.0
        self.agent_features[:, self.DONE] = 0.0
        self.withdraw_history = []

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

    def config_agents_from_xml(self, scenario: str, *, verbose: bool = True) -> None:
        from .matsim_io import population_from_xml
        rows = population_from_xml(os.path.join("data", scenario), verbose=verbose)
        self.agent_features = torch.tensor(rows, dtype=torch.float32, device=self.device)
