import sys, time, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.core import SimulationCoreModel
dev = torch.device("cuda")
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
N, E = int(g.num_roads), g.edge_index_routes.size(1)
model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=21600)
bank = [synthetic.random_out_neighbour(g, 1000 + i).cpu().pin_memory() for i in range(8)]
sel_h = bank[0]
sel_d = [bank[0].to(dev) for _ in range(2)]
dtt_h = [torch.empty(E).pin_memory() for _ in range(2)]
pop_h = [torch.empty(N, dtype=torch.bool).pin_memory() for _ in range(2)]
main = torch.cuda.current_stream(dev)
cs = torch.cuda.Stream(dev)
def t_ms(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - a) / n * 1e3
t = [21600.0]
def compute_only():
    model.set_time(t[0]); model(g, selected_road=sel_d[0]); t[0] += 1
def d2h_only():
    dtt_h[0].copy_(model.direction_mpnn.road_optimality_data["delta_travel_time"], non_blocking=True)
    pop_h[0].copy_(model.last_pop, non_blocking=True)
def h2d_only():
    sel_d[0].copy_(bank[int(t[0]) % 8], non_blocking=True)
def serial():
    h2d_only(); compute_only(); d2h_only()
k = [0]
ev_in = [torch.cuda.Event() for _ in range(2)]
def overlapped(record=True):
    i = k[0]
    main.wait_event(ev_in[i % 2])
    model.set_time(t[0]); model(g, selected_road=sel_d[i % 2]); t[0] += 1
    dtt = model.direction_mpnn.road_optimality_data["delta_travel_time"]; pop = model.last_pop
    done = torch.cuda.Event(); done.record(main)
    with torch.cuda.stream(cs):
        sel_d[(i + 1) % 2].copy_(bank[(i + 1) % 8], non_blocking=True); ev_in[(i + 1) % 2].record(cs)
        cs.wait_event(done)
        dtt_h[i % 2].copy_(dtt, non_blocking=True); pop_h[i % 2].copy_(pop, non_blocking=True)
        if record:
            dtt.record_stream(cs); pop.record_stream(cs)
    k[0] += 1
print("serial", t_ms(serial)); print("d2h", t_ms(d2h_only)); print("h2d", t_ms(h2d_only))
with torch.cuda.stream(cs):
    sel_d[0].copy_(sel_h, non_blocking=True); ev_in[0].record(cs)
print("overlapped", t_ms(overlapped))
model.response_mpnn.update_history.resolve()
import ctypes as C
from tarl_simulator_b200 import _cabi
def ev_ms(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); b = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (b - a) / n * 1e3
print("compute_only (gpu ms, host enqueue ms)", ev_ms(compute_only))
print("rand only", ev_ms(lambda: torch.rand(E, device=dev)))
from tarl_simulator_b200.topology import topology_for
print("topology_for", ev_ms(lambda: topology_for(g.edge_index_routes, N)))
