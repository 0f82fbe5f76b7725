"""One cubic-spline warp of a (60, 1024, 1024) float32 volume (for an ncu launch list)."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
shape = (60, 1024, 1024)
vol = torch.rand(shape, device="cuda") * 4095
T = np.eye(4); T[:3, 3] = (0.4, 3.25, -11.5)
M = T @ b2.get_3D_rotation_matrix(shape, 7.3) @ b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07))
for _ in range(2):
    o = b2.spline_warp(vol, M)
torch.cuda.synchronize()
print("ok", float(o.sum()))
