import torch, time
n = 214_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s", 10 * n / e0.elapsed_time(e1) / 1e6)
# chunked 8192 windows * 3264 B
c = 8192 * 3264
e0.record()
for _ in range(10):
    for o in range(0, n - c, c): d[o:o + c].copy_(h[o:o + c], non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D chunked GB/s", 10 * (n // c) * c / e0.elapsed_time(e1) / 1e6)
