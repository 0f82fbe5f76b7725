import sys; sys.path.insert(0, "/root/repo")
import torch, biahub_b200 as b2
g = torch.Generator(device="cuda").manual_seed(0)
v = torch.randint(1, 65536, (800, 300, 2048), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
for _ in range(3): o = b2.fast_deskew_zyx(v, 30.0, 0.386, average_n_slices=3, keep_overhang=True, overhang_fill="mean")
torch.cuda.synchronize()
