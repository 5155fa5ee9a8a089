import torch, time
n=150_000_000
h=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device='cuda')
for _ in range(3): d.copy_(h,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("pinned H2D GB/s", n/dt/1e9, "ms", dt*1e3)
