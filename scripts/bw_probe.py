import sys; sys.path.insert(0, '/root/repo')
import torch
dev = torch.device('cuda')
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
a = torch.empty(1 << 30, dtype=torch.float32, device=dev)   # 4 GiB
b = torch.empty(1 << 30, dtype=torch.float32, device=dev)
ms = timeit(lambda: a.zero_()); print(f"zero_ 4GiB: {ms:.3f} ms {4.295/ms*1e3:.0f} GB/s")
ms = timeit(lambda: a.fill_(1.5)); print(f"fill_ 4GiB: {ms:.3f} ms {4.295/ms*1e3:.0f} GB/s")
ms = timeit(lambda: b.copy_(a)); print(f"copy 4GiB->4GiB: {ms:.3f} ms {8.59/ms*1e3:.0f} GB/s")
ms = timeit(lambda: torch.sum(a)); print(f"sum (read) 4GiB: {ms:.3f} ms {4.295/ms*1e3:.0f} GB/s")
h = a.view(torch.bfloat16); hb = b.view(torch.bfloat16)
ms = timeit(lambda: hb.copy_(h)); print(f"copy bf16 view: {ms:.3f} ms {8.59/ms*1e3:.0f} GB/s")
c = torch.empty((100, 2048, 1813), dtype=torch.float32, device=dev)
ms = timeit(lambda: c.fill_(2.0)); print(f"fill (100,2048,1813): {ms:.3f} ms {c.numel()*4/ms/1e6:.0f} GB/s")
