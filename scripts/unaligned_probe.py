"""Deskew on source rows that are not 16-byte aligned (TMA-ineligible -> gather kernel)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, biahub_b200 as b2
g = torch.Generator(device="cuda").manual_seed(0)
for X in (2048, 2044, 2047):
    vols = [torch.randint(0, 65536, (800, 300, X), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16) for _ in range(3)]
    for _ in range(2):
        for v in vols: o = b2.fast_deskew_zyx(v, 30.0, 0.386, False, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for v in vols: o = b2.fast_deskew_zyx(v, 30.0, 0.386, False, 3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 9
    byt = vols[0].numel() * 2 + o.numel() * 4
    print(f"X={X}: {ms:.3f} ms  {o.numel()/ms/1e6:.0f} Gvox/s  ({byt/ms/1e6/6534.1:.2f} of roofline)")
