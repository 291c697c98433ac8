"""Scratch timing harness for kernel tuning (not part of the product or the bench contract): builds the link store of
a workload once, and per repetition restores the SAME warm state and times store.run(steps) — the regime bench.py
measures (BENCH_r01: --steps 20 --warmup 5; ~25 % of the links pop per step, ~3 % are contested).
Usage: python profiles/tune_step.py [reps] [steps] [workload] [replicas] [warmup]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tarl_simulator_b200 import synthetic  # noqa: E402
from tarl_simulator_b200.engine import LinkStore  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
workload = sys.argv[3] if len(sys.argv) > 3 else "ring_radial_1m"
R = int(sys.argv[4]) if len(sys.argv) > 4 else 1
warm = int(sys.argv[5]) if len(sys.argv) > 5 else 5
g, Nmax, placed = synthetic.make_workload(workload, device="cuda", t=21600.0, seed=0)
N = int(g.num_roads)
store = LinkStore.from_graph(g, Nmax, replicas=R, seed=1234)
bank = [synthetic.random_out_neighbour(g, 1000 + i).repeat(R) for i in range(8)]
out = []
for _ in range(reps):
    store.import_x(g.x[:N], g.congestion_constant[:N], broadcast=True)
    store.step_id = 0
    t = 21600.0
    store.run(t, warm, sel_bank=bank); t += warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    store.run(t, steps, sel_bank=bank[warm % 8:] + bank[:warm % 8], pop_bits=True, delta_tt_link=True)
    e1.record()
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / steps * 1e3)
try:
    store.check_errors()
except RuntimeError as exc:
    print("faults:", str(exc)[:80])
print(os.environ.get("TARL_TUNE", ""), "ahead", os.environ.get("TARL_AHEAD_SELECT", "-"), os.environ.get("TARL_AHEAD_RESPOND", "-"), "PDL off" if os.environ.get("TARL_NO_PDL") else "PDL on", workload, "R", R,
      f"{steps} steps after {warm}: us/step:", " ".join(f"{v:.2f}" for v in out), "| min", f"{min(out):.2f}",
      "| pops last step", int(store.pop[: N * R].sum()) / (N * R))
