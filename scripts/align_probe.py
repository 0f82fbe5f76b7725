import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import biahub_b200 as b2
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(0)
srcs = [torch.randint(0, 65536, (800, 300, 2048), generator=g, device=dev, dtype=torch.int32).to(torch.uint16) for _ in range(8)]
for align in (1, 4, 32, 1, 32):
    for _ in range(3):
        for s in srcs: o = b2.fast_deskew_zyx(s, 30.0, 0.386, False, 3, row_align=align)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for s in srcs: o = b2.fast_deskew_zyx(s, 30.0, 0.386, False, 3, row_align=align)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 80
    print(f"row_align={align:3d} pitch={o.stride(1)} {ms:.4f} ms/launch  {2468.2496/ms:.1f} GB/s  frac {2468.2496/ms/6534.1:.3f}")
