"""Scratch: two PPO rollouts of R grid100 replicas without any profiler of its own, for ncu captures of the rollout
kernels (k_gd_sample_bcast, k_ell_*, k_withdraw, k_insert_*, k_observe). Usage: python profiles/rollout_ncu.py [R]"""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, _EnvAdapter, collect
dev = torch.device("cuda")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frm, to, n_nodes = synthetic.grid_links(100, device=dev)
frm, to = synthetic.reorder_links(frm, to, "node")
g, Nmax = synthetic.build_graph(frm, to, n_nodes)
af = synthetic.population(g, 100_000, 21540, 600, seed=7)
env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100)
pm = PolicyModule(MPNNPolicyNet(g.edge_index, g.x.size(0), None, "cuda"), g.edge_index)
ad = _EnvAdapter(env)
for _ in range(2):
    collect(ad, pm, 8, occupancy_only=True)
torch.cuda.synchronize()
