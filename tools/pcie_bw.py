import torch, time
x=torch.empty(256*1024*1024, dtype=torch.uint8).pin_memory(); y=torch.empty_like(x, device='cuda')
for _ in range(3): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("H2D pinned GB/s", 10*x.numel()/dt/1e9)
z=torch.empty_like(x).pin_memory()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): z.copy_(y, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("D2H pinned GB/s", 10*x.numel()/dt/1e9)
