"""Developer tool: plain pinned H2D copy bandwidth (the ceiling for the e2e figure)."""
import torch, time
n = 8 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(4): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print(f"pinned H2D {4*n/(time.perf_counter()-t)/1e9:.2f} GB/s (8 GiB copies)")
