"""Scratch: CUDA kernel breakdown of a PPO rollout (collect) of R grid100 replicas."""
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
policy = MPNNPolicyNet(g.edge_index, g.x.size(0), None, "cuda")
pm = PolicyModule(policy, g.edge_index)
ad = _EnvAdapter(env)
collect(ad, pm, 8, occupancy_only=True); collect(ad, pm, 8, occupancy_only=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); collect(ad, pm, 8, occupancy_only=True); e1.record(); torch.cuda.synchronize()
print("collect(8): %.2f ms" % e0.elapsed_time(e1))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    collect(ad, pm, 8, occupancy_only=True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
