"""Scratch: time of the three value-MLP kernels (torch.profiler) at M rows."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200.mpnn_agent import MPNNValueNetSimple
M, N = int(sys.argv[1]), int(sys.argv[2])
net = MPNNValueNetSimple(torch.zeros(2, 1, dtype=torch.long, device="cuda"), N, "cuda")
num = torch.rand(M, N, device="cuda")
if len(sys.argv) > 3 and sys.argv[3] == "int": num = torch.randint(0, 16, (M, N), device="cuda").float()   # occupancies
time = torch.rand(M, 1, device="cuda")
with torch.no_grad():
    for _ in range(3): net.forward_occupancy(num, time)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(10): net.forward_occupancy(num, time)
        torch.cuda.synchronize()
for e in prof.key_averages():
    if "k_value" in e.key: print("%-24s %8.1f us" % (e.key.split("::")[-1][:22], e.device_time_total / e.count))
with torch.no_grad():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): net.forward_occupancy(num, time)
    e1.record(); torch.cuda.synchronize()
    print("forward, 50 calls back to back: %.1f us per call" % (e0.elapsed_time(e1) * 1000 / 50))
import time as _t
with torch.no_grad():
    torch.cuda.synchronize(); t0 = _t.perf_counter()
    for _ in range(50): net.forward_occupancy(num, time)
    t1 = _t.perf_counter(); torch.cuda.synchronize()
    print("host time to enqueue one forward: %.1f us" % ((t1 - t0) * 1e6 / 50))
