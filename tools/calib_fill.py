"""How fast does this B200 write: a 3.2 GB fill (pure write) and an in-place add (1 : 1 read : write), next to the
driver-written copy bandwidth of MEASURED_PEAKS.json that the rooflines use (profiles/r1_calib_fill.log)."""
import torch
dev=torch.device("cuda:0")
n=16384*16384*3
a=torch.empty(n,dtype=torch.float32,device=dev)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for name,fn,byt in (("fill_ (pure write, 3.2 GB)", lambda: a.fill_(1.0), 4*n),("cudaMemsetAsync (zero_)", lambda: a.zero_(), 4*n)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f"{name}: {ms:.4f} ms, {byt/ms/1e6:.0f} GB/s")
def mix():
    torch.add(a, 0.0, out=a)
for _ in range(3): mix()
torch.cuda.synchronize(); e0.record()
for _ in range(10): mix()
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
print(f"in-place add (3.2 GB read + 3.2 GB write): {ms:.4f} ms, {8*n/ms/1e6:.0f} GB/s")
