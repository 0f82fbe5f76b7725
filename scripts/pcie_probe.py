"""Host<->device copy bandwidth with pinned memory: one rank alone, then all ranks at once.
    python -m torch.distributed.run --nproc-per-node N scripts/pcie_probe.py
Gives the PCIe/host-memory ceiling the end-to-end numbers of bench.py run against."""
import os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
GB = 1 << 30
h_in = torch.empty(GB, dtype=torch.uint8, pin_memory=True); h_in.zero_()
h_out = torch.empty(GB, dtype=torch.uint8, pin_memory=True); h_out.zero_()
d_a = torch.empty(GB, dtype=torch.uint8, device="cuda"); d_b = torch.empty(GB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(mode, reps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * GB / dt / 1e9

def barrier():
    if world > 1: dist.barrier()

res = {}
for mode in ("h2d", "d2h", "both"):
    run(mode, 2)
    # alone: rank 0 only
    barrier()
    solo = run(mode) if rank == 0 else 0.0
    barrier()
    allr = run(mode)
    t = torch.tensor([solo, allr], device="cuda", dtype=torch.float64)
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
    else:
        lst = [t]
    if rank == 0:
        per = [float(x[1]) for x in lst]
        print(f"{mode:5s}: rank0 alone {float(lst[0][0]):6.1f} GB/s per direction | all {world} ranks: "
              f"sum {sum(per):7.1f} GB/s per direction, per rank min {min(per):.1f} max {max(per):.1f}", flush=True)
if world > 1: dist.destroy_process_group()
