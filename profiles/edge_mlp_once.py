"""Scratch: MPNNPolicyNet.edge_mlp forward (tcgen05 and fp32 pipe) + backward once at 8 rows on ring_radial_1m (for ncu)."""
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
w = torch.randn(B, E, device=dev)
with torch.no_grad():
    net.edge_logits(nf, ef, ai, tensor_cores=True)
    net.edge_logits(nf, ef, ai, tensor_cores=False)
(net.edge_logits(nf, ef, ai, tensor_cores=True) * w).sum().backward()
torch.cuda.synchronize()
