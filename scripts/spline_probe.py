"""method="scipy" (cubic spline) on the C3 shape: kernel time of prefilter + evaluation and the
end-to-end call, next to scipy itself on a slab (one thread)."""
import sys, time; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
from biahub_b200 import _cabi
shape = (120, 2048, 2048)
g = torch.Generator(device="cuda").manual_seed(0)
vol = torch.rand(shape, generator=g, device="cuda") * 4095
T = np.eye(4); T[:3, 3] = (0.4, 3.25, -11.5)
M = T @ b2.get_3D_rotation_matrix(shape, 7.3) @ b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07))
for _ in range(2): o = b2.spline_warp(vol, M)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): o = b2.spline_warp(vol, M)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"spline_warp device f32 {shape}: {ms:.2f} ms per volume = {np.prod(shape)/ms/1e6:.1f} Gvox/s (prefilter x3 + evaluation)")
u = (vol[:60]).to(torch.int32).to(torch.uint16)
for _ in range(2): o = b2.spline_warp(u, M)
torch.cuda.synchronize(); e0.record()
for _ in range(3): o = b2.spline_warp(u, M)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"spline_warp device u16 {tuple(u.shape)}: {ms:.2f} ms = {u.numel()/ms/1e6:.1f} Gvox/s")
h = vol.cpu().numpy()
b2.apply_affine_transform(h, M, shape, method="scipy")
t0 = time.perf_counter(); r = b2.apply_affine_transform(h, M, shape, method="scipy"); dt = time.perf_counter() - t0
print(f"apply_affine_transform(method='scipy') host f32 {shape}: {dt*1e3:.1f} ms = {np.prod(shape)/dt/1e9:.2f} Gvox/s end to end (pageable in, pooled pinned out)")
import scipy.ndimage
slab = h[:8]
t0 = time.perf_counter(); w = scipy.ndimage.affine_transform(slab, M, slab.shape); dt = time.perf_counter() - t0
print(f"scipy.ndimage.affine_transform order 3 on {slab.shape}, 1 thread: {dt:.2f} s = {slab.size/dt/1e9:.4f} Gvox/s")
g2 = b2.apply_affine_transform(slab, M, slab.shape, method="scipy")
print("max |gpu - scipy| / range on the slab:", float(np.abs(g2 - w).max() / 4095))
