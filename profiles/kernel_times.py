"""Scratch: per-kernel durations without Python launch overhead.
(1) k_ell_select_append alone, 20 launches captured in one CUDA graph and replayed (vs. launched from Python).
(2) torch.profiler table of one MPNN fwd+bwd iteration (policy+distribution, value net) at B rows."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.engine import PHASE_RESPOND_POP, PHASE_SELECT_APPEND, LinkStore
from tarl_simulator_b200.distribution import GraphDistribution
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNet
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
N, E = int(g.num_roads), g.edge_index_routes.size(1)
store = LinkStore.from_graph(g, Nmax, replicas=1, seed=1)
store.sel = synthetic.random_out_neighbour(g, 1000)
dtt = torch.empty(1, E, device=dev)
store.run(21600.0, 5, delta_tt=dtt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def py_loop(mask, n=20):
    e0.record()
    for _ in range(n): store.step(21605.0, delta_tt=dtt, phase_mask=mask)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("select_append from python loop: %.1f us" % py_loop(PHASE_SELECT_APPEND))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for _ in range(20): store.step(21605.0, delta_tt=dtt, phase_mask=PHASE_SELECT_APPEND)
for _ in range(3): gr.replay()
torch.cuda.synchronize()
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print("select_append x20 in a CUDA graph: %.1f us per launch" % (e0.elapsed_time(e1) / 20 * 1e3))
store.step(21605.0, delta_tt=dtt, phase_mask=PHASE_RESPOND_POP)

ei = g.edge_index; Ef, Nt = ei.size(1), g.x.size(0)
nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
policy = MPNNPolicyNet(ei, Nt, None, "cuda")
with torch.no_grad():
    d0 = GraphDistribution(policy(nf, None, None), ei)
    action = d0.sample(dtype=torch.bool)
adv = torch.randn(B, device=dev)
def policy_iter():
    policy.nodes_embedding.weight.grad = None
    dd = GraphDistribution(policy(nf, None, None), ei)
    lp = dd.log_prob(action); ent = dd.entropy()
    (-(lp * adv).mean() - 0.01 * ent.mean()).backward()
value = MPNNValueNet(ei, Nt, "cuda"); value.agent_features = torch.rand(1024, 9, device=dev); value.eval()
ef = g.edge_attr.reshape(1, Ef, 1).expand(B, -1, -1)
ai = torch.randint(0, 1024, (B, Nt), device=dev); tm = torch.full((B, 1), 21600.0, device=dev); wv = torch.randn(B, 1, device=dev)
def value_iter():
    for p_ in value.parameters(): p_.grad = None
    (value(nf, ef, ai, tm) * wv).sum().backward()
for name, fn in (("policy", policy_iter), ("value", value_iter)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print("=====", name, "B=%d  %.3f ms/iter" % (B, e0.elapsed_time(e1) / 10))
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5): fn()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
