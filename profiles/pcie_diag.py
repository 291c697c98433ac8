"""Raw pinned-memory PCIe bandwidth of the box next to bench.py's e2e figure (which is PCIe-bound): tells a slow box
from a slow pipeline. Under torchrun every rank copies to / from its own GPU AT THE SAME TIME (barrier first) and rank 0
prints the per-rank and the aggregate figures: what the e2e arm of bench.py can reach at N GPUs at the very most.
Usage: python profiles/pcie_diag.py   |   python -m torch.distributed.run --nproc-per-node 8 profiles/pcie_diag.py"""
import os, time, torch
import torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 4 << 20                                   # the e2e step's copy size: 4 MB each way
reps = 200
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
out = {}
for name in ("h2d", "d2h", "both"):
    for timed in (False, True):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps if timed else 10):
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    out[name] = (2 if name == "both" else 1) * reps * n / dt / 1e9
vals = torch.tensor([out["h2d"], out["d2h"], out["both"]], device=dev)
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    for i, nm in enumerate(("h2d", "d2h", "both directions at once")):
        per = [float(v[i]) for v in allv]
        print(f"{nm}: 4 MiB pinned copies, {world} GPU(s) at once: per GPU {min(per):.1f} - {max(per):.1f} GB/s, aggregate {sum(per):.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
