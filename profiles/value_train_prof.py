"""Scratch: per-kernel times of MPNNValueNet fwd+bwd in TRAIN mode (message dropout) at 32 rows on ring_radial_1m."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.mpnn_agent import MPNNValueNet
dev = torch.device("cuda")
B = 32
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
ei = g.edge_index; Ef, Nt = ei.size(1), g.x.size(0)
nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
value = MPNNValueNet(ei, Nt, "cuda"); value.agent_features = torch.rand(1024, 9, device=dev); value.train()
ef = g.edge_attr.reshape(1, Ef, 1).expand(B, -1, -1)
ai = torch.randint(0, 1024, (B, Nt), device=dev); tm = torch.full((B, 1), 21600.0, device=dev); wv = torch.randn(B, 1, device=dev)
def it():
    for p_ in value.parameters(): p_.grad = None
    (value(nf, ef, ai, tm) * wv).sum().backward()
for _ in range(3): it()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3): it()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
