"""Scratch timing harness for kernel tuning (not part of the product or the bench contract): builds the ring_radial_1m
link store once and times store.run(n) repeatedly. Usage: python profiles/tune_step.py [reps] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tarl_simulator_b200 import synthetic  # noqa: E402
from tarl_simulator_b200.engine import LinkStore  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
workload = sys.argv[3] if len(sys.argv) > 3 else "ring_radial_1m"
R = int(sys.argv[4]) if len(sys.argv) > 4 else 1
g, Nmax, placed = synthetic.make_workload(workload, device="cuda", t=21600.0, seed=0)
store = LinkStore.from_graph(g, Nmax, replicas=R, seed=1234)
E = g.edge_index_routes.size(1)
dtt = torch.empty(R, E, device="cuda")
bank = [synthetic.random_out_neighbour(g, 1000 + i).repeat(R) for i in range(8)]
t = 21600.0
store.run(t, 10, sel_bank=bank, delta_tt=dtt); t += 10
torch.cuda.synchronize()
out = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    store.run(t, steps, sel_bank=bank, delta_tt=dtt)
    e1.record()
    torch.cuda.synchronize()
    t += steps
    out.append(e0.elapsed_time(e1) / steps * 1e3)
store.check_errors()
print(os.environ.get("TARL_TUNE", ""), "PDL off" if os.environ.get("TARL_NO_PDL") else "PDL on", workload, "R", R,
      "us/step:", " ".join(f"{v:.2f}" for v in out), "| min", f"{min(out):.2f}")
