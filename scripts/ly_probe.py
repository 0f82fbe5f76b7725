"""One matrix of the manual-registration family through the lanes-along-y zsep kernel (for ncu).
argv[1]: rot90 | scaled (default: scaling @ rotate90 @ fliplr)"""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
shape = (120, 2048, 2048)
g = torch.Generator(device="cuda").manual_seed(0)
v = torch.rand(shape, generator=g, device="cuda") * 4095
if len(sys.argv) > 1 and sys.argv[1] == "rot90":
    M = b2.get_3D_rotation_matrix(shape, 90)
else:
    M = b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)) @ b2.get_3D_rotation_matrix(shape, 90) @ b2.get_3D_fliplr_matrix(shape)
for _ in range(3): o = b2.affine_warp(v, M, shape, order=1, boundary="itk")
torch.cuda.synchronize()
