"""Write profiles/<round>_*.txt from the ncu captures in gpurun_out/ plus SASS evidence.

    python scripts/make_profile_summary.py r1b r1
"""
import csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

def launches(wl):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{wl}_{tag}.csv")
    if not os.path.exists(path):
        return None
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")[:70]
        dur = float(r[-1]) / 1e3  # ns -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += dur
    total = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list ({wl}; bench.py --steps 2 --warmup 3 --no-e2e --no-extra --units 2; cold-cache serialised times: compare SHARES)",
           f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'share':>7s}"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{name:70s} {n:8d} {t:10.1f} {100*t/total:6.1f}%")
    return "\n".join(out)

def full(wl):
    path = os.path.join(ROOT, "gpurun_out", f"prof_{wl}_{tag}.ncu-rep")
    if not os.path.exists(path):
        return None
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), path], capture_output=True, text=True)
    return f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(path)})\n" + r.stdout

def sass():
    out = ["# SASS evidence (cuobjdump -sass of the built objects): mnemonic counts per kernel family"]
    for obj, pat in (("b2_deskew.o", "deskew_tma_kernelItLi3ELi256"), ("b2_deskew.o", "deskew_stage_kernelILi1ELi256"),
                     ("b2_affine_zsep.o", "affine_zsep_kernelIfLi1ELi1ELb1ELb0"), ("b2_affine_zsep.o", "affine_zsep_kernelIfLi1ELi1ELb1ELb1"),
                     ("b2_affine_brick.o", "affine_brick_kernelIfLi1ELi1ELb1ELb0"),
                     ("b2_flatfield.o", "flatfield_median_kernel"), ("b2_flatfield.o", "flatfield_apply_kernelIf"),
                     ("b2_fill.o", "fill_bits_kernel"), ("b2_fill.o", "fill_dilate_z_kernelILb1")):
        p = os.path.join(ROOT, "biahub_b200", "_lib", "obj", obj)
        txt = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True).stdout
        m = re.search(r"Function : (\S*%s\S*)(.*?)(?=Function :|\Z)" % pat, txt, re.S)
        if not m: continue
        body = m.group(2)
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body, re.M)
        cnt = {}
        for o in ops: cnt[o] = cnt.get(o, 0) + 1
        keys = [k for k in cnt if re.match(r"UTMALDG|UBLKCP|SYNCS|LDS|STS|STG|LDG|TEX|TLD|PRMT|FFMA|FMUL|FADD|I2F|DFMA|DMUL|DADD|VOTE|REDUX|SHF|BAR", k)]
        out.append(f"\n{m.group(1)}\n  total SASS instructions: {len(ops)}")
        for k in sorted(keys, key=lambda k: -cnt[k]): out.append(f"  {k:34s} {cnt[k]}")
        out.append("  texture instructions (TEX/TLD): %d" % sum(v for k, v in cnt.items() if k.startswith(("TEX", "TLD"))))
    return "\n".join(out)

for wl in ("deskew_c2", "deskew_c1", "register_c3", "stabilize_c4", "register_generic"):
    parts = [p for p in (launches(wl), full(wl)) if p]
    if parts:
        with open(os.path.join(ROOT, "profiles", f"{rnd}_{wl}.txt"), "w") as fh:
            fh.write("\n\n".join(parts) + "\n")
        src = os.path.join(ROOT, "gpurun_out", f"launches_{wl}_{tag}.csv")
        if os.path.exists(src):
            import shutil; shutil.copy(src, os.path.join(ROOT, "profiles", f"{rnd}_launches_{wl}.csv"))
with open(os.path.join(ROOT, "profiles", f"{rnd}_sass_evidence.txt"), "w") as fh:
    fh.write(sass() + "\n")
print(os.listdir(os.path.join(ROOT, "profiles")))
