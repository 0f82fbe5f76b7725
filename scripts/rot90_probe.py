"""zsep kernel on the manual-registration matrix family (reference estimate_registration.py:174-189:
scaling @ rotate90 @ fliplr): throughput vs the near-identity C3 matrix."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
shape = (120, 2048, 2048)
g = torch.Generator(device="cuda").manual_seed(0)
vols = [torch.rand(shape, generator=g, device="cuda") * 4095 for _ in range(4)]
def run(name, M):
    for _ in range(2):
        for v in vols: o = b2.affine_warp(v, M, shape, order=1, boundary="itk")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for v in vols: o = b2.affine_warp(v, M, shape, order=1, boundary="itk")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    nz = float((o != 0).float().mean())
    print(f"{name:34s} {ms:7.3f} ms  {np.prod(shape)/ms/1e6:7.1f} Gvox/s  ({8*np.prod(shape)/ms/1e6/6534.1:.2f} of roofline)  nonzero {nz:.2f}")
T = np.eye(4); T[:3, 3] = (0.4, 3.25, -11.5)
run("C3: rot 7.3 scale 1.07", T @ b2.get_3D_rotation_matrix(shape, 7.3) @ b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)))
run("rot 90", b2.get_3D_rotation_matrix(shape, 90))
run("scale 1.07 @ rot 90", b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)) @ b2.get_3D_rotation_matrix(shape, 90))
run("scale 1.07 @ rot 90 @ fliplr", b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)) @ b2.get_3D_rotation_matrix(shape, 90) @ b2.get_3D_fliplr_matrix(shape))
run("rot 45", b2.get_3D_rotation_matrix(shape, 45))
run("rot 180", b2.get_3D_rotation_matrix(shape, 180))
run("identity + frac shift", T)
# generic (non z-separable) matrices: brick kernel
def oop(shape, a_deg, b_deg):
    c = (np.array(shape) - 1) / 2.0
    a, b = np.radians(a_deg), np.radians(b_deg)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1]])
    R = Ry @ Rx; M = np.eye(4); M[:3, :3] = R; M[:3, 3] = c - R @ c
    return M
C3 = T @ b2.get_3D_rotation_matrix(shape, 7.3) @ b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07))
R90 = b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)) @ b2.get_3D_rotation_matrix(shape, 90) @ b2.get_3D_fliplr_matrix(shape)
run("generic: C3 @ tilt(0.5,0.3)", C3 @ oop(shape, 0.5, 0.3))
run("generic: rot90 family @ tilt", R90 @ oop(shape, 0.5, 0.3))
