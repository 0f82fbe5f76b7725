import sys; sys.path.insert(0, '/root/repo')
import torch, biahub_b200 as b2
dev = torch.device('cuda'); g = torch.Generator(device=dev); g.manual_seed(0)
srcs = [torch.randint(0, 65536, (800, 300, 2048), generator=g, device=dev, dtype=torch.int32).to(torch.uint16) for _ in range(6)]
for N in (1, 2, 3):
    for _ in range(3):
        for s in srcs: o = b2.fast_deskew_zyx(s, 30.0, 0.386, False, N)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for s in srcs: o = b2.fast_deskew_zyx(s, 30.0, 0.386, False, N)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 60
    byt = 800*300*2048*2 + o.numel()*4
    print(f"N={N} out={tuple(o.shape)} {ms:.4f} ms  {byt/ms/1e6:.1f} GB/s frac {byt/ms/1e6/6534.1:.3f}")
