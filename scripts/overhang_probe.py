"""Deskew with keep_overhang=True and overhang_fill (the shipped example YAML): kernel + fill time."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import biahub_b200 as b2
g = torch.Generator(device="cuda").manual_seed(0)
vols = [torch.randint(1, 65536, (800, 300, 2048), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16) for _ in range(3)]
def run(name, **kw):
    for _ in range(2):
        for v in vols: o = b2.fast_deskew_zyx(v, 30.0, 0.386, average_n_slices=3, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for v in vols: o = b2.fast_deskew_zyx(v, 30.0, 0.386, average_n_slices=3, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 9
    byt = vols[0].numel() * 2 + o.numel() * 4
    print(f"{name:40s} out {tuple(o.shape)} {ms:7.3f} ms {o.numel()/ms/1e6:7.1f} Gvox/s  ({byt/ms/1e6/6534.1:.2f} of roofline, deskew bytes only)")
run("crop", keep_overhang=False)
run("keep_overhang, fill 0", keep_overhang=True)
run("keep_overhang, fill mean", keep_overhang=True, overhang_fill="mean")
run("keep_overhang, fill 100", keep_overhang=True, overhang_fill=100.0)
