"""Scratch: a few MPNN fwd+bwd iterations on the full ring_radial_1m graph, for ncu captures of the k_gd_* / k_policy_* /
k_value_* kernels (ncu -k regex:... --launch-skip N)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.distribution import GraphDistribution
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNet, MPNNValueNetSimple
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
ei = g.edge_index; Ef, Nt = ei.size(1), g.x.size(0)
nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
policy = MPNNPolicyNet(ei, Nt, None, "cuda")
with torch.no_grad():
    d0 = GraphDistribution(policy(nf, None, None), ei)
    action = d0.sample(dtype=torch.bool)
adv = torch.randn(B, device=dev)
value = MPNNValueNet(ei, Nt, "cuda"); value.agent_features = torch.rand(1024, 9, device=dev); value.eval()
ef = g.edge_attr.reshape(1, Ef, 1).expand(B, -1, -1)
ai = torch.randint(0, 1024, (B, Nt), device=dev); tm = torch.full((B, 1), 21600.0, device=dev); wv = torch.randn(B, 1, device=dev)
vs = MPNNValueNetSimple(torch.zeros(2, 1, dtype=torch.long, device=dev), 59600, "cuda")
occ = torch.randint(0, 12, (1024, 59600), device=dev).float(); tv = torch.full((1024, 1), 21600.0, device=dev)
torch.cuda.synchronize()
print("MARK setup done")
for _ in range(iters):
    policy.nodes_embedding.weight.grad = None
    dd = GraphDistribution(policy(nf, None, None), ei)
    lp = dd.log_prob(action); ent = dd.entropy()
    (-(lp * adv).mean() - 0.01 * ent.mean()).backward()
    for p_ in value.parameters(): p_.grad = None
    (value(nf, ef, ai, tm) * wv).sum().backward()
    value.train(); value.time_net[1].p = 0.0; value.time_net[4].p = 0.0          # message dropout kernels
    for p_ in value.parameters(): p_.grad = None
    (value(nf, ef, ai, tm) * wv).sum().backward()
    value.eval()
    with torch.no_grad():
        vs.forward_occupancy(occ, tv)
        d1 = GraphDistribution(policy(nf, None, None), ei)
        d1.sample(dtype=torch.bool, return_log_prob=True)
torch.cuda.synchronize()
