"""Scratch: where one PPO update goes at R replicas on one GPU (host time vs device time). usage: ppo_update_prof.py [R]"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
from tarl_simulator_b200.rl import ppo_trainer as pt
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
frm, to, n_nodes = synthetic.grid_links(100, device=dev)
frm, to = synthetic.reorder_links(frm, to, "node")
g, Nmax = synthetic.build_graph(frm, to, n_nodes)
af = synthetic.population(g, 100_000, 21540, 600, seed=7)
env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100)
torch.manual_seed(0)
policy = MPNNPolicyNet(g.edge_index, g.x.size(0), g.edge_attr, "cuda")
value = MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda")
pm, vm = pt.PolicyModule(policy, g.edge_index), pt.ValueModule(value)
T = 32
pt.ppo_train(env, pm, vm, total_frames=3 * T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)
torch.cuda.synchronize()
def once():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    pt.ppo_train(env, pm, vm, total_frames=T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) * 1e3, e0.elapsed_time(e1)
for _ in range(3): h, d = once(); print("iteration: host %.2f ms, device %.2f ms" % (h, d))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    once(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# the update = everything after the last rollout kernel (k_insert*/k_ell*)
last_roll = max(i for i, e in enumerate(ev) if "k_insert" in e.name or "k_ell_" in e.name)
upd = ev[last_roll + 1:]
t_begin, t_end = upd[0].time_range.start, upd[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in upd)
print("update: %d device ops, span %.1f us, busy %.1f us" % (len(upd), t_end - t_begin, busy))
agg = {}
for e in upd:
    k = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0].split("<")[0][-48:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("  %-50s x%3d %8.1f us" % (k, n, t))
