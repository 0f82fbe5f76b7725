"""Affine kernels on uint16 sources (shipped as uint16, converted on the device)."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
shape = (120, 2048, 2048)
g = torch.Generator(device="cuda").manual_seed(0)
vu = [torch.randint(0, 65536, shape, generator=g, device="cuda", dtype=torch.int32).to(torch.uint16) for _ in range(3)]
vf = [v.view(torch.int16).to(torch.int32).bitwise_and(0xFFFF).to(torch.float32) for v in vu]
def run(name, vols, M, esz):
    for _ in range(2):
        for v in vols: o = b2.affine_warp(v, M, shape, order=1, boundary="itk")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for v in vols: o = b2.affine_warp(v, M, shape, order=1, boundary="itk")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 9
    print(f"{name:34s} {ms:7.3f} ms  {np.prod(shape)/ms/1e6:7.1f} Gvox/s  ({(4+esz)*np.prod(shape)/ms/1e6/6534.1:.2f} of its roofline)")
    return o
T = np.eye(4); T[:3, 3] = (0.4, 3.25, -11.5)
C3 = T @ b2.get_3D_rotation_matrix(shape, 7.3) @ b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07))
a = run("C3 float32", vf, C3, 4)
b = run("C3 uint16", vu, C3, 2)
print("max |u16 - f32| :", float((a - b).abs().max()))
c = (np.array(shape) - 1) / 2.0
al, be = np.radians(0.5), np.radians(0.3)
Ry = np.array([[np.cos(al), 0, np.sin(al)], [0, 1, 0], [-np.sin(al), 0, np.cos(al)]])
Rx = np.array([[np.cos(be), -np.sin(be), 0], [np.sin(be), np.cos(be), 0], [0, 0, 1]])
R = Ry @ Rx; Mo = np.eye(4); Mo[:3, :3] = R; Mo[:3, 3] = c - R @ c
run("generic float32", vf, C3 @ Mo, 4)
run("generic uint16", vu, C3 @ Mo, 2)
