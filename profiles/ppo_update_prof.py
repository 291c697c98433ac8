"""Scratch: where one PPO iteration spends its time at R replicas per GPU (one GPU): rollout / update by CUDA events,
host enqueue time beside it, and a torch.profiler table of the update. Usage: python profiles/ppo_update_prof.py [R]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, _EnvAdapter, collect, occupancy_only, ppo_train

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = 32
dev = torch.device("cuda")
frm, to, n_nodes = synthetic.grid_links(100, device=dev)
frm, to = synthetic.reorder_links(frm, to, "node")
g, Nmax = synthetic.build_graph(frm, to, n_nodes)
af = synthetic.population(g, 100_000, 21540, 600, seed=7)
env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100)
torch.manual_seed(0)
policy = MPNNPolicyNet(g.edge_index, g.x.size(0), None, "cuda")
value = MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda")
pm, vm = PolicyModule(policy, g.edge_index), ValueModule(value)
ad = _EnvAdapter.of(env)
ppo_train(env, pm, vm, total_frames=3 * T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)
torch.cuda.synchronize()

def timed(fn, n=5):
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter(); e0.record(); fn(); e1.record(); h1 = time.perf_counter()
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1), (h1 - h0) * 1e3))
    return min(o[0] for o in out), min(o[1] for o in out)

slim = occupancy_only(pm, vm)
print("R", R, "rollout  gpu ms %.2f  host enqueue ms %.2f" % timed(lambda: collect(ad, pm, T, occupancy_only=slim)))
print("R", R, "iteration gpu ms %.2f  host ms %.2f" % timed(lambda: ppo_train(env, pm, vm, total_frames=T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)))
if len(sys.argv) > 2 and sys.argv[2] == "syncdebug":
    torch.cuda.set_sync_debug_mode("warn")
    import warnings
    warnings.simplefilter("always")
    ppo_train(env, pm, vm, total_frames=T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)
    torch.cuda.set_sync_debug_mode("default")
    sys.exit(0)
os.environ["TARL_NO_ROLLOUT_GRAPH"] = "1"
print("R", R, "rollout (eager) gpu ms %.2f  host enqueue ms %.2f" % timed(lambda: collect(ad, pm, T, occupancy_only=slim)))
del os.environ["TARL_NO_ROLLOUT_GRAPH"]
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    ppo_train(env, pm, vm, total_frames=T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=60))
