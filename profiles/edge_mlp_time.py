"""Scratch: where the time of MPNNPolicyNet.edge_logits goes at 8 rows on ring_radial_1m (torch.profiler)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
dev = torch.device("cuda")
B = 8
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
ei = g.edge_index; E, N = ei.size(1), g.x.size(0)
nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
net = MPNNPolicyNet(ei, N, None, "cuda"); net.agent_features = torch.rand(1024, 9, device=dev)
ai = torch.randint(0, 1024, (B, N), device=dev)
ef = g.edge_attr.reshape(1, E, 1).expand(B, -1, -1)
with torch.no_grad():
    for _ in range(3): net.edge_logits(nf, ef, ai, tensor_cores=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): net.edge_logits(nf, ef, ai, tensor_cores=True)
    e1.record(); torch.cuda.synchronize()
    print("tc per call ms", e0.elapsed_time(e1) / 5)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3): net.edge_logits(nf, ef, ai, tensor_cores=True)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=50))
w = torch.randn(B, E, device=dev)
def train():
    for p_ in net.edge_mlp.parameters(): p_.grad = None
    (net.edge_logits(nf, ef, ai, tensor_cores=True) * w).sum().backward()
for _ in range(2): train()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof2:
    for _ in range(2): train()
    torch.cuda.synchronize()
print(prof2.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=50))
