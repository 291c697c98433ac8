import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200.mpnn_agent import MPNNValueNetSimple
M, N = 1024, 59600
net = MPNNValueNetSimple(torch.zeros(2, 1, dtype=torch.long, device="cuda"), N, "cuda")
num = torch.rand(M, N, device="cuda"); time = torch.rand(M, 1, device="cuda")
with torch.no_grad():
    for _ in range(3): net.forward_occupancy(num, time)
    torch.cuda.synchronize()
