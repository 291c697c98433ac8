"""Raw pinned-memory PCIe bandwidth of the box next to bench.py's e2e figure (which is PCIe-bound): tells a slow box
from a slow pipeline. Usage: python profiles/pcie_diag.py"""
import time, torch
dev = torch.device("cuda")
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {20 * n / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s pinned, 64 MiB copies")
# both directions at once (two copy engines)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print(f"bidirectional: {2 * 20 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s total")
