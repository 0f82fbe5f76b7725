import sys; sys.path.insert(0, '/root/repo')
import torch, biahub_b200 as b2
dev = torch.device('cuda'); g = torch.Generator(device=dev); g.manual_seed(0)
srcs = [torch.randint(0, 65536, (800, 300, 2048), generator=g, device=dev, dtype=torch.int32).to(torch.uint16) for _ in range(2)]
for _ in range(3):
    for s in srcs: o = b2.fast_deskew_zyx(s, 30.0, 0.386, False, 1)
torch.cuda.synchronize()
