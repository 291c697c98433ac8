import os, time, torch
print("env", {k: v for k, v in os.environ.items() if "PYTORCH" in k or "CUDA" in k})
dev = torch.device("cuda")
x = torch.zeros(1500000, 52, device=dev)
def f():
    a = torch.empty(3993000, device=dev); b = torch.empty(999000, dtype=torch.bool, device=dev); return a, b
for _ in range(3): f()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(100): f()
print("empty pair us:", (time.perf_counter() - t) / 100 * 1e6)
keep = []
t = time.perf_counter()
for _ in range(50):
    keep.append(f()[1])
print("empty pair keeping masks us:", (time.perf_counter() - t) / 50 * 1e6)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_reserved() / 1e6)
