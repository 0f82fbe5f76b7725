"""Time the flat-field kernels on a mantis-sized volume (uint16 (800,300,2048)); prints GB/s."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
from biahub_b200._device import pinned_empty
Z, Y, X = 800, 300, 2048
g = torch.Generator(device="cuda").manual_seed(0)
vols = [torch.randint(90, 1200, (Z, Y, X), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16) for _ in range(4)]
for _ in range(2):
    for v in vols: o = b2.flat_field._flatfield_tensor(v, torch.float32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    for v in vols: o = b2.flat_field._flatfield_tensor(v, torch.float32)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 12
alg = Z * Y * X * (2 + 4)
print(f"device: {ms:.3f} ms/volume  {Z*Y*X/ms/1e6:.1f} Gvox/s  algorithmic {alg/ms/1e6:.0f} GB/s ({alg/ms/1e6/6534.1:.2f} of HBM peak)")
h = pinned_empty((1, Z, Y, X), np.uint16)
torch.from_numpy(h.view(np.int16)).copy_(vols[0].view(torch.int16)[None])
for rep in range(4):
    res = None
    t0 = time.perf_counter()
    res = b2._flat_field_czyx(h, [0])
    dt = time.perf_counter() - t0
print(f"host API (pinned in, pooled out): {dt*1e3:.1f} ms/volume  {Z*Y*X/dt/1e9:.2f} Gvox/s")
hp = np.array(h, copy=True)
for rep in range(3):
    res = None
    t0 = time.perf_counter()
    res = b2._flat_field_czyx(hp, [0])
    dt = time.perf_counter() - t0
print(f"host API (pageable in): {dt*1e3:.1f} ms/volume  {Z*Y*X/dt/1e9:.2f} Gvox/s")
