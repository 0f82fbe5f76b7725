"""One matrix of the manual-registration family through the lanes-along-y zsep kernel (for ncu).
argv[1]: rot90 | generic (the family with 0.5/0.3 degree out-of-plane tilts: brick kernel) | scaled (default: scaling @ rotate90 @ fliplr)"""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
shape = (120, 2048, 2048)
g = torch.Generator(device="cuda").manual_seed(0)
v = torch.rand(shape, generator=g, device="cuda") * 4095
if len(sys.argv) > 1 and sys.argv[1] == "rot90":
    M = b2.get_3D_rotation_matrix(shape, 90)
elif len(sys.argv) > 1 and sys.argv[1] == "generic":
    c = (np.array(shape) - 1) / 2.0
    a, b = np.radians(0.5), np.radians(0.3)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1]])
    R = Ry @ Rx; T = np.eye(4); T[:3, :3] = R; T[:3, 3] = c - R @ c
    M = (b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)) @ b2.get_3D_rotation_matrix(shape, 90) @ b2.get_3D_fliplr_matrix(shape)) @ T
else:
    M = b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)) @ b2.get_3D_rotation_matrix(shape, 90) @ b2.get_3D_fliplr_matrix(shape)
for _ in range(3): o = b2.affine_warp(v, M, shape, order=1, boundary="itk")
torch.cuda.synchronize()
