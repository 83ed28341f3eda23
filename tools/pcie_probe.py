import torch, time
x = torch.empty(324_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
y = torch.empty(120_000_000, dtype=torch.uint8).pin_memory()
dy = torch.empty_like(y, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print(f"H2D 324MB: {dt*1e3:.2f} ms = {0.324/dt:.1f} GB/s")
t=time.perf_counter(); y.copy_(dy, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print(f"D2H 120MB: {dt*1e3:.2f} ms = {0.120/dt:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
t=time.perf_counter()
with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
with torch.cuda.stream(s2): y.copy_(dy, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print(f"both directions: {dt*1e3:.2f} ms")
