"""Scratch: two fwd+bwd iterations of MPNNValueNet in TRAIN mode at 32 rows on ring_radial_1m (for ncu)."""
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
for _ in range(2):
    for p_ in value.parameters(): p_.grad = None
    (value(nf, ef, ai, tm) * wv).sum().backward()
torch.cuda.synchronize()
