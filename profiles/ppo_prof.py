import sys, time, torch, cProfile, pstats
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, _EnvAdapter, collect, ppo_train
dev = torch.device("cuda")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frm, to, n_nodes = synthetic.grid_links(100, device=dev)
frm, to = synthetic.reorder_links(frm, to, "node")
g, Nmax = synthetic.build_graph(frm, to, n_nodes)
af = synthetic.population(g, 100_000, 21540, 600, seed=7)
env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100)
policy = MPNNPolicyNet(g.edge_index, g.x.size(0), None, "cuda"); value = MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda")
pm, vm = PolicyModule(policy, g.edge_index), ValueModule(value)
ppo_train(env, pm, vm, total_frames=4, frames_per_batch=4, num_epochs=1, sub_batch_size=32, history=[])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter()
ppo_train(env, pm, vm, total_frames=8, frames_per_batch=8, num_epochs=1, sub_batch_size=32, history=[])
torch.cuda.synchronize()
print("iteration s", time.perf_counter() - t)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
