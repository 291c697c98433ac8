"""Scratch: where a graph-replayed PPO rollout of R grid100 replicas spends its time — reset vs replay (CUDA events) and
the kernels inside the replay (torch.profiler). Usage: python profiles/rollout_timeline.py R"""
import os, re, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, _EnvAdapter, collect
dev = torch.device("cuda")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
frm, to, n_nodes = synthetic.grid_links(100, device=dev)
frm, to = synthetic.reorder_links(frm, to, "node")
g, Nmax = synthetic.build_graph(frm, to, n_nodes)
af = synthetic.population(g, 100_000, 21540, 600, seed=7)
env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100)
torch.manual_seed(0)         # same initial policy and (below) the same noise keys in every run
policy = MPNNPolicyNet(g.edge_index, g.x.size(0), None, "cuda")
pm = PolicyModule(policy, g.edge_index)
ad = _EnvAdapter.of(env)
for _ in range(4):
    collect(ad, pm, 32, occupancy_only=True)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tot, rst = [], []
for _ in range(5):
    ev[0].record(); ad.reset(); ad.dynamic(out=(ad.trajectory_buffers(32, True)["num"][0], None, None)); ev[1].record()
    torch.cuda.synchronize()
    ev[1].record(); collect(ad, pm, 32, occupancy_only=True); ev[2].record(); torch.cuda.synchronize()
    rst.append(ev[0].elapsed_time(ev[1])); tot.append(ev[1].elapsed_time(ev[2]))
print(f"R {R}: reset+observe alone {min(rst):.3f} ms; collect (reset + replay) {min(tot):.3f} ms")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    collect(ad, pm, 32, occupancy_only=True)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
total = sum(e.self_device_time_total for e in rows)
print(f"kernel time in one collect: {total / 1e3:.3f} ms")
for e in rows[:16]:
    m = re.search(r"(k_\w+|Memcpy\w*|Memset\w*|\w+Functor\w*|\w+_kernel\w*)", e.key)
    print(f"  {(m.group(1) if m else e.key[:40]):36s} {e.self_device_time_total / max(e.count, 1):8.1f} us x{e.count:3d} = {e.self_device_time_total / 1e3:7.3f} ms")
b = collect(ad, pm, 32, occupancy_only=True)
torch.cuda.synchronize()
env.check_errors()
# same numbers whatever the schedule (TARL_ROLLOUT_DRAW_AT, TARL_ROLLOUT_SIDE_PRIORITY, TARL_NO_ROLLOUT_OVERLAP)
lp = b["sample_log_prob"]
print("checksum: num %.1f action %d log_prob %.6f (+ %d impossible frames of %d) reward %.1f" % (
      float(b["num"].double().sum()), int(b["action"].sum()), float(lp[torch.isfinite(lp)].double().sum()),
      int((~torch.isfinite(lp)).sum()), lp.numel(), float(b["reward"].double().sum())))
