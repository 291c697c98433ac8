"""Scratch: torch.profiler table of one MPNN fwd+bwd iteration on the full ring_radial_1m graph."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200 import synthetic
from tarl_simulator_b200.distribution import GraphDistribution
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNet
dev = torch.device("cuda")
g, Nmax, _ = synthetic.make_workload("ring_radial_1m", device=dev, t=21600.0)
B = 4
ei = g.edge_index; E, N = ei.size(1), g.x.size(0)
nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
policy = MPNNPolicyNet(ei, N, None, "cuda")
with torch.no_grad():
    d0 = GraphDistribution(policy(nf, None, None), ei)
    action = d0.sample(dtype=torch.bool)
adv = torch.randn(B, device=dev)
def policy_iter():
    policy.nodes_embedding.weight.grad = None
    dd = GraphDistribution(policy(nf, None, None), ei)
    lp = dd.log_prob(action); ent = dd.entropy()
    (-(lp * adv).mean() - 0.01 * ent.mean()).backward()
value = MPNNValueNet(ei, N, "cuda"); value.agent_features = torch.rand(1024, 9, device=dev); value.eval()
ef = g.edge_attr.reshape(1, E, 1).expand(B, -1, -1)
ai = torch.randint(0, 1024, (B, N), device=dev); tm = torch.full((B, 1), 21600.0, device=dev); wv = torch.randn(B, 1, device=dev)
def value_iter():
    for p_ in value.parameters(): p_.grad = None
    (value(nf, ef, ai, tm) * wv).sum().backward()
for name, fn in (("policy", policy_iter), ("value", value_iter)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(5): fn()
        torch.cuda.synchronize()
    print("=====", name)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
