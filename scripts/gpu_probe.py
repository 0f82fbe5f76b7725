"""Run each kernel family in its own process (a CUDA fault is sticky per process)."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
"deskew_gather": "t=cu(rnd((64,30,128)));o=b2.fast_deskew_zyx(t,30.0,0.386,False,3,_path=1);torch.cuda.synchronize();print(o.shape,float(o.sum()))",
"deskew_tma": "t=cu(rnd((64,30,128)));o=b2.fast_deskew_zyx(t,30.0,0.386,False,3,_path=2);torch.cuda.synchronize();print(o.shape,float(o.sum()))",
"affine_gather": "v=torch.rand((12,40,48),device='cuda');o=b2.affine_warp(v,M((12,40,48)),(12,40,48),_path=1);torch.cuda.synchronize();print(float(o.sum()))",
"affine_zsep": "v=torch.rand((12,40,48),device='cuda');o=b2.affine_warp(v,M((12,40,48)),(12,40,48),_path=2);torch.cuda.synchronize();print(float(o.sum()))",
"affine_zsep_big": "v=torch.rand((12,200,264),device='cuda');o=b2.affine_warp(v,M((12,200,264)),(12,200,264),_path=2);torch.cuda.synchronize();print(float(o.sum()))",
"deskew_tma_smallz": "t=cu(rnd((16,4,64)));o=b2.fast_deskew_zyx(t,30.0,0.386,True,2,_path=2);torch.cuda.synchronize();print(o.shape,float(o.sum()))",
"deskew_tma_f32": "t=torch.rand((64,30,128),device='cuda');o=b2.fast_deskew_zyx(t,30.0,0.386,False,3,_path=2);torch.cuda.synchronize();print(o.shape,float(o.sum()))",
"fill": "t=cu(rnd((64,30,128)));o=b2.fast_deskew_zyx(t,30.0,0.386,True,3,overhang_fill='mean',_path=1);torch.cuda.synchronize();print(float(o.sum()))",
"host_deskew": "o=b2._fast_deskew_czyx(rnd((64,30,128))[None],ls_angle_deg=30.0,px_to_scan_ratio=0.386,keep_overhang=False,average_n_slices=3);print(o.shape,float(o.sum()))",
}
PRE = """
import sys; sys.path.insert(0, %r)
import numpy as np, torch
import biahub_b200 as b2
from oracle import affine_oracle as ao
def rnd(s): return np.random.default_rng(0).integers(0,65536,size=s,dtype=np.uint16)
def cu(a): return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
def M(s): return ao.register_matrix_c3(s)
""" % ROOT
names = sys.argv[1:] or list(CASES)
for n in names:
    env = dict(os.environ)
    if "@" in n:
        n, dbg = n.split("@"); env["B2_ZSEP_DEBUG"] = dbg
    r = subprocess.run([sys.executable, "-c", PRE + CASES[n]], capture_output=True, text=True, timeout=300, env=env)
    tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or [""])[-1]
    print(f"{n:16s} rc={r.returncode} {tail}", flush=True)
