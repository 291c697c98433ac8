"""Scratch: error anatomy of the tcgen05 value MLP's first layer at K = 59600 (vs fp64), and its timing vs cuBLAS."""
import sys, torch
sys.path.insert(0, "/root/repo")
from tarl_simulator_b200.mpnn_agent import MPNNValueNetSimple
M, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 59600
g = torch.Generator(device="cuda").manual_seed(3)
net = MPNNValueNetSimple(torch.zeros(2, 1, dtype=torch.long, device="cuda"), N, "cuda")
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.randn(p.shape, device="cuda", generator=g) * (0.05 if p.dim() == 2 and p.size(1) > 64 else 0.3))
num = torch.randint(0, 15, (M, N), device="cuda", generator=g).float() * (torch.rand(M, N, device="cuda", generator=g) < 0.7)
num[:, ::7] += torch.rand(M, num[:, ::7].size(1), device="cuda", generator=g)
time = (torch.arange(M, device="cuda", dtype=torch.float32).reshape(M, 1) % 7.0)
with torch.no_grad():
    # make the tail the identity on h1[0] so that the first layer is visible: w2 = I, b2 = 0, w3 = e_j
    l1, l2, l3 = net.final_mlp[0], net.final_mlp[2], net.final_mlp[4]
    x = torch.cat((num, time), -1)
    pre = x.double() @ l1.weight.double().t() + l1.bias.double()            # [M, 64] exact
    pre32 = x @ l1.weight.t() + l1.bias
    absum = (x.double().abs() @ l1.weight.double().abs().t())
    l2.weight.copy_(torch.eye(64, device="cuda")); l2.bias.zero_(); l3.bias.zero_()
    errs = []
    for j in range(0, 64, 9):
        l3.weight.zero_(); l3.weight[0, j] = 1.0
        got = net.forward_occupancy(num, time).double().reshape(-1)
        ref = torch.relu(pre[:, j])
        errs.append(float(((got - ref).abs() / absum[:, j]).max()))
    print("first layer: max |err| / sum|a w| per probed column:", ["%.2e" % e for e in errs])
    print("cuBLAS fp32:  max |err| / sum|a w|:", "%.2e" % float(((pre32.double() - pre).abs() / absum).max()))
    print("typical |pre| / sum|a w|:", "%.2e" % float((pre.abs() / absum).median()))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("tcgen05", lambda: net.forward_occupancy(num, time)), ("library", lambda: net.final_mlp(torch.cat((num, time), -1))),
                     ("library, no cat", lambda: net.final_mlp[0](x))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ev0.record()
        for _ in range(20): fn()
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 20
        print(f"{name}: {ms*1e3:.1f} us for M={M}, K={N+1}: A bytes {M*N*4/1e6:.0f} MB -> {M*N*4/ms/1e6:.0f} GB/s")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5): net.forward_occupancy(num, time)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=60))
