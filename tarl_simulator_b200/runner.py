"""Runner: unified entry point for classical and RL experiments.

Drop-in for the reference's src/runner.py:1-226 (`RunnerArgs`, `Runner.setup/train/eval`) for the algorithms on the
hot path: "random" (classical loop with random routing), "mpnn" (rollout of the learned policy) and "mpnn+ppo" (PPO
training of it). "dijkstra" and the MSA user-equilibrium post-processing are outside the scope contract (SURVEY.md
§2) and raise. `steps` (declared divergence D5: the reference's README documents `--steps` but main.py lacks it)
bounds the number of simulated timesteps; `replicas` > 1 trains on a BatchedSimulatorEnv (R environments per GPU).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import torch

from .agents import Agents
from .reinforcement_learning import BatchedSimulatorEnv, SimulatorEnv
from .transportation_simulator import TransportationSimulator


@dataclass
class RunnerArgs:
    algo: str
    scenario: str
    mode: str
    timestep_size: int = 1
    start_end_time: list = (0, 86400)
    epochs: int = 1
    rollout_steps: int = 32
    seed: int = 0
    device: str = "cuda"
    output_dir: str = "runs"
    profile: bool = False
    torch_compile: bool = False
    steps: int | None = None
    replicas: int = 1


class Runner:
    def __init__(self, args: RunnerArgs):
        self.args = args
        if not torch.cuda.is_available():
            raise RuntimeError("tarl_simulator_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(args.device if args.device != "cpu" else "cuda")
        torch.manual_seed(args.seed)
        self.summary = {}

    def setup(self):
        a = self.args
        if a.algo == "random":
            self.simulator = TransportationSimulator(str(self.device), torch_compile=a.torch_compile)
            self.agent = Agents(str(self.device))
            self.simulator.load_network(scenario=a.scenario)
            self.agent.load(scenario=a.scenario)
            self.simulator.config_parameters(timestep_size=a.timestep_size, start_time=a.start_end_time[0])
            self.agent.set_time(a.start_end_time[0])
        elif a.algo in {"mpnn", "mpnn+ppo"}:
            from .mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
            self.env = SimulatorEnv(device=str(self.device), timestep_size=a.timestep_size,
                                    start_time=a.start_end_time[0], scenario=a.scenario, torch_compile=a.torch_compile)
            g = self.env.simulator.graph
            edge_index = g.edge_index
            num_nodes = g.x.size(0)
            free_flow = g.x[:, self.env.simulator.h.FREE_FLOW_TIME_TRAVEL][edge_index[1]]
            self.policy_net = MPNNPolicyNet(edge_index, num_nodes, free_flow, device=str(self.device))
            self.policy_net.load(a.scenario)
            self.value_net = MPNNValueNetSimple(edge_index, num_nodes, device=str(self.device))
            self.value_net.load(a.scenario)
            self.env.simulator.agent = self.policy_net
        elif a.algo == "dijkstra":
            raise NotImplementedError("DijkstraAgents is outside the scope of this implementation (SURVEY.md §2)")
        else:
            raise ValueError(f"Unknown algorithm {a.algo}")

    def _modules(self, return_log_prob=True):
        from .rl.ppo_trainer import PolicyModule, ValueModule
        ei = self.env.simulator.graph.edge_index
        return PolicyModule(self.policy_net, ei, return_log_prob=return_log_prob), ValueModule(self.value_net)

    def train(self):
        a = self.args
        if not (a.algo == "mpnn+ppo" and a.mode == "train"):
            raise RuntimeError("Training is only supported for algo 'mpnn+ppo'")
        from .rl.ppo_trainer import ppo_train
        policy_module, value_module = self._modules()
        out = Path(a.output_dir)
        out.mkdir(parents=True, exist_ok=True)
        sim = self.env.simulator
        if a.replicas > 1:
            train_env = BatchedSimulatorEnv(sim.graph, sim.Nmax, self.policy_net.agent_features, a.replicas,
                                            timestep=a.timestep_size, seed=a.seed)
        else:
            train_env = self.env
        eval_env = SimulatorEnv(device=str(self.device), timestep_size=a.timestep_size, start_time=a.start_end_time[0],
                                scenario=a.scenario, torch_compile=a.torch_compile)
        eval_env.simulator.agent = self.policy_net
        self.history = ppo_train(train_env, policy_module, value_module, total_frames=a.rollout_steps,
                                 frames_per_batch=a.rollout_steps, num_epochs=a.epochs, device=self.device,
                                 checkpoint_path=out / "policy.pt", log_dir=str(out), eval_env=eval_env,
                                 eval_interval=1, seed=a.seed, history=[])

    def _report(self, sim, agent):
        mask = agent.agent_features[:, agent.DONE] == 1
        avg = torch.mean(agent.agent_features[mask, agent.ARRIVAL_TIME] - agent.agent_features[mask, agent.DEPARTURE_TIME])
        total = sim.inserting_time + sim.choice_time + sim.core_time + sim.withdraw_time
        self.summary = {"average_travel_time": float(avg), "arrived": int(mask.sum()), "insert_s": sim.inserting_time,
                        "choice_s": sim.choice_time, "core_s": sim.core_time, "withdraw_s": sim.withdraw_time,
                        "total_s": total}
        print("\n=== Simulation Summary ===")
        print(f"{'Average travel time:':25} {float(avg):10.2f} s")
        print(f"{'Agent Insertion time:':25} {sim.inserting_time:10.2f} s")
        print(f"{'Route Choice time:':25} {sim.choice_time:10.2f} s")
        print(f"{'Core Model time:':25} {sim.core_time:10.2f} s")
        print(f"{'Agent Withdrawal time:':25} {sim.withdraw_time:10.2f} s")
        print("-" * 42)
        print(f"{'Total simulation time:':25} {total:10.2f} s")

    def eval(self):
        a = self.args
        n = (a.start_end_time[1] - a.start_end_time[0]) // a.timestep_size
        if a.steps is not None:
            n = min(n, a.steps)
        if a.algo == "random":
            self.simulator.agent = self.agent                       # run_episode, src/algorithms/base_runner.py:37
            self.simulator.sync_timers = True
            for _ in range(n):
                self.simulator.run()
            self.agent.check_errors()
            self._report(self.simulator, self.agent)
        else:
            policy_module, _ = self._modules(return_log_prob=False)
            self.env.simulator.sync_timers = True
            self.env.rollout(n, policy_module, break_when_any_done=False)
            self.env.simulator.agent.check_errors()
            self._report(self.env.simulator, self.env.simulator.agent)
