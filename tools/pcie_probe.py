import torch, time
d=torch.empty(110<<20,dtype=torch.uint8,device='cuda'); h=torch.empty(110<<20,dtype=torch.uint8).pin_memory()
for _ in range(3): h.copy_(d,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): h.copy_(d,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("D2H GB/s", (110<<20)/dt/1e9)
t=time.perf_counter()
for _ in range(10): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("H2D GB/s", (110<<20)/dt/1e9)
