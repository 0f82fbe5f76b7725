"""Host<->device copy ceiling of a multi-GPU box, lever by lever (VERDICT r1 "Next" #3):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe2.py

For every combination of
  * pinned allocation: cudaHostAlloc (torch pin_memory) | anonymous mmap + MADV_HUGEPAGE +
    cudaHostRegister (transparent huge pages) | cudaHostAlloc write-combined (H2D source only)
  * rank -> device mapping: identity | spread (rank r -> device r * (visible // N))
  * direction schedule: h2d alone | d2h alone | both at once (duplex) | phased (all ranks copy
    H2D, barrier, all ranks copy D2H: the two directions never overlap)
it prints the aggregate GB/s over all ranks with the C2 unit's byte mix (0.98 GB up, 1.49 GB
down per unit) and the Gvoxels/s that mix would allow.  One cudaMemcpyAsync per copy.
"""
import ctypes
import mmap
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
visible = torch.cuda.device_count()
UP, DOWN = 983_040_000, 1_485_209_600        # bytes of one C2 unit: uint16 in, float32 out
OUT_VOX = 100 * 2048 * 1813

if world > 1:
    dist.init_process_group("gloo")


def barrier():
    if world > 1:
        dist.barrier()


def gather_max(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class Pinned:
    """nbytes of page-locked host memory, allocated one of three ways."""

    def __init__(self, nbytes, how):
        self.how = how
        if how == "hostalloc":
            self.t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self.t.zero_()
        elif how == "thp":
            self.m = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            self.m.madvise(mmap.MADV_HUGEPAGE)
            arr = np.frombuffer(self.m, dtype=np.uint8)
            arr[:: 4096] = 0                      # first touch: fault the (huge) pages in
            self.t = torch.from_numpy(arr)
            rc = torch.cuda.cudart().cudaHostRegister(self.t.data_ptr(), nbytes, 0)
            assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
        else:
            raise ValueError(how)

    def close(self):
        if self.how == "thp":
            torch.cuda.cudart().cudaHostUnregister(self.t.data_ptr())
            del self.t
            self.m.close()


def thp_kb():
    try:
        for line in open("/proc/self/smaps_rollup"):
            if line.startswith("AnonHugePages"):
                return int(line.split()[1])
    except Exception:
        pass
    return -1


def run_case(how, mapping):
    dev = (lr * max(1, visible // world)) % visible if mapping == "spread" else lr % visible
    torch.cuda.set_device(dev)
    h_in, h_out = Pinned(UP, how), Pinned(DOWN, how)
    d_in = torch.empty(UP, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def up():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in.t, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_out.t.copy_(d_out, non_blocking=True)

    def timed(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return gather_max(time.perf_counter() - t0) / reps

    def phased():
        up()
        torch.cuda.synchronize()
        barrier()
        down()
        torch.cuda.synchronize()
        barrier()

    res = {}
    res["h2d"] = world * UP / timed(up) / 1e9
    res["d2h"] = world * DOWN / timed(down) / 1e9
    t_both = timed(lambda: (up(), down()))
    t_ph = timed(phased)
    res["duplex_unit_s"], res["phased_unit_s"] = t_both, t_ph
    if rank == 0:
        print(f"alloc={how:9s} map={mapping:8s} N={world}: h2d {res['h2d']:6.1f} GB/s  d2h {res['d2h']:6.1f} GB/s | "
              f"C2 unit per rank: duplex {t_both * 1e3:6.1f} ms = {world * OUT_VOX / t_both / 1e9:5.1f} Gvox/s "
              f"({world * UP / t_both / 1e9:5.1f} up + {world * DOWN / t_both / 1e9:5.1f} down GB/s), "
              f"phased {t_ph * 1e3:6.1f} ms = {world * OUT_VOX / t_ph / 1e9:5.1f} Gvox/s"
              f"  [AnonHugePages {thp_kb()} kB, device {dev}]", flush=True)
    del d_in, d_out
    h_in.close()
    h_out.close()
    torch.cuda.empty_cache()


if rank == 0:
    print(f"# pcie_probe2: {world} ranks, {visible} visible GPUs, host cpus {len(os.sched_getaffinity(0))}", flush=True)
mappings = ["identity", "spread"] if visible > world else ["identity"]
for how in ("hostalloc", "thp"):
    for mapping in mappings:
        try:
            run_case(how, mapping)
        except Exception as exc:  # noqa: BLE001 - report and go on with the next lever
            if rank == 0:
                print(f"alloc={how} map={mapping}: FAILED {exc!r}", flush=True)
            barrier()
if world > 1:
    dist.destroy_process_group()
