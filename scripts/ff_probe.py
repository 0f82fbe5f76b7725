import sys; sys.path.insert(0, "/root/repo")
import torch, biahub_b200 as b2
g = torch.Generator(device="cuda").manual_seed(0)
v = torch.randint(90, 1200, (800, 300, 2048), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
for _ in range(3): o = b2.flat_field._flatfield_tensor(v, torch.float32)
torch.cuda.synchronize()
