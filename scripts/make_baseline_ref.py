#!/usr/bin/env python
"""Stage the UNMODIFIED reference package for the GPU box: copies ``/root/reference/biahub``
(Python sources only) to the git-ignored ``baseline/_ref/biahub`` so that ``bench.py --impl
reference`` and the ``cpu_baseline`` leg can run the reference's own
``biahub.deskew._fast_deskew_czyx(device="cpu")`` (reference biahub/deskew.py:551-579) there —
``/root/reference`` does not exist on the GPU box, ``baseline/_ref`` travels with the snapshot.

    python scripts/make_baseline_ref.py            # no-op when /root/reference is absent

Nothing is modified and nothing is committed (``.gitignore`` lists ``baseline/_ref/``); the files
are loaded by ``oracle/ref_loader.py`` with ``BIAHUB_REFERENCE_ROOT=baseline/_ref`` and inert
stand-ins for the third-party packages that are not installed.  A
``pip install`` of the reference is not possible offline: iohub / monai / antspyx / submitit
are absent from /opt/wheelhouse.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.environ.get("BIAHUB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose=True) -> bool:
    src_pkg = os.path.join(SRC, "biahub")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"{src_pkg} not present: nothing staged")
        return False
    dst_pkg = os.path.join(DST, "biahub")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    manifest = {}
    for dirpath, dirnames, filenames in os.walk(src_pkg):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        for name in filenames:
            if not name.endswith(".py"):
                continue
            s = os.path.join(dirpath, name)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            with open(s, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()[:16]
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() or True else 1)
