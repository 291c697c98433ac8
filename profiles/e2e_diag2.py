import sys, time, torch, cProfile, pstats
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.core import SimulationCoreModel
dev = torch.device("cuda")
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
N, E = int(g.num_roads), g.edge_index_routes.size(1)
model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=21600)
bank = [synthetic.random_out_neighbour(g, 1000 + i) for i in range(8)]
t = [21600.0]
def compute_only():
    model.set_time(t[0]); model(g, selected_road=bank[int(t[0]) % 8]); t[0] += 1
for _ in range(3): compute_only()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): compute_only()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(18)
