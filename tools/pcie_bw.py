import torch, time
n = 1<<30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print("H2D GB/s", 5*n/(time.perf_counter()-t)/1e9)
t=time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print("D2H GB/s", 5*n/(time.perf_counter()-t)/1e9)
