"""Time the flat-field median + apply kernels on the mantis-sized volume (device-resident)."""
import sys
sys.path.insert(0, "/root/repo")
import torch
import biahub_b200 as b2
Z, Y, X = 800, 300, 2048
g = torch.Generator(device="cuda").manual_seed(0)
vols = [torch.randint(90, 1200, (Z, Y, X), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16) for _ in range(4)]
for _ in range(2):
    for v in vols: o = b2.flat_field._flatfield_tensor(v, torch.float32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    for v in vols: o = b2.flat_field._flatfield_tensor(v, torch.float32)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
alg = Z * Y * X * (2 + 4)
print(f"{ms:.3f} ms/volume  {Z*Y*X/ms/1e6:.1f} Gvox/s  algorithmic {alg/ms/1e6:.0f} GB/s ({alg/ms/1e6/6534.1:.2f} of HBM peak)")
