"""Scratch: per-kernel average durations (torch.profiler, CUDA activities) of an eager 32-step PPO rollout of R grid100
replicas. Usage: python profiles/rollout_kernels.py R [tag]"""
import os, sys, torch
os.environ["TARL_NO_ROLLOUT_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, _EnvAdapter, collect
dev = torch.device("cuda")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tag = sys.argv[2] if len(sys.argv) > 2 else ""
frm, to, n_nodes = synthetic.grid_links(100, device=dev)
frm, to = synthetic.reorder_links(frm, to, "node")
g, Nmax = synthetic.build_graph(frm, to, n_nodes)
af = synthetic.population(g, 100_000, 21540, 600, seed=7)
env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100)
policy = MPNNPolicyNet(g.edge_index, g.x.size(0), None, "cuda")
pm = PolicyModule(policy, g.edge_index)
ad = _EnvAdapter.of(env)
collect(ad, pm, 32, occupancy_only=True); collect(ad, pm, 32, occupancy_only=True)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    collect(ad, pm, 32, occupancy_only=True)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:9]
print(tag, "R", R, " | ".join(f"{e.key.split('(')[0].split('::')[-1][:28]} {e.self_device_time_total / max(e.count, 1):.1f}us x{e.count}" for e in rows))
