#!/usr/bin/env python
"""Benchmark of the B200-native 3-D affine resampling path (BASELINE.json metric:
output Gvoxels/s; % of HBM bandwidth).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic (position, t, c) units that is
resident in HBM when the timed region starts; ``value`` is whole-job output Gvoxels/s (all
ranks' voxels / max-over-ranks device time).  ``e2e`` is the same metric through the
reference-facing Python call (``_fast_deskew_czyx`` / ``apply_affine_transform`` /
``apply_stabilization_transform`` bodies) with HOST buffers: pinned H2D of every input and D2H
of every output inside the timed region; ``e2e.ceiling`` is what plain pinned cudaMemcpyAsync
of the same bytes (both directions at once, all ranks at once) reaches on this host.
``roofline`` compares the dominant kernel's algorithmic bytes / launch time with the measured
HBM peak (MEASURED_PEAKS.json).  ``cpu_baseline`` times the reference's CPU path on rank 0.

The default run reports BASELINE.json configs[1] (C2 mantis deskew) as the headline and adds a
``workloads`` array with a short run of every other config (C1 deskew, C3 register, C4 stabilize,
C5 chained unit, generic 3-D affine): kernel-only value, launch time, roofline fraction, e2e,
clock record each.  ``--workload plate_c5`` runs configs[4] at plate level: 8 positions x T=32
units from host arrays through ``units_for_rank`` + ``run_units_overlapped`` +
``deskew_then_register`` (strong scaling: the plate is fixed, ranks share it).

Multi-GPU: one process per GPU (torchrun); units are sharded over ranks with NO data-path
collective (the reference's parallelism is independent (position, t, c) units).
torch.distributed is used only for the barrier and the max-over-ranks of the times.  When more
GPUs are visible than ranks, ranks are spread over the device list (rank r -> device
r * (visible // N)) so that they sit on different PCIe switches of an HGX board.

``--impl reference`` times the reference's own CPU implementation on the same workload/metric:
for the deskew workloads that is the UNMODIFIED ``biahub.deskew._fast_deskew_czyx(device="cpu")``
loaded from ``baseline/_ref`` (scripts/make_baseline_ref.py; kind "reference"), one WHOLE unit
per step with all host threads; the warps time ``scipy.ndimage.affine_transform(order=1)`` — the
library behind the reference's ``method="scipy"`` branch — one process per core (kind "port":
the default ANTs branch is not installable).  Rank 0 only.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "output Gvoxels/s (deskew/register/stabilize resampling); % of HBM BW"

WORKLOADS = {
    # BASELINE.json configs[1]: mantis-sized deskew, average_n_slices=3, uint16 (T=8,C=2,800,300,2048)
    "deskew_c2": dict(kind="deskew", shape=(800, 300, 2048), dtype="uint16", units=16,
                      ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                      average_n_slices=3, e2e_units=8, ncu_traffic=2.3533e9,
                      desc="C2 mantis deskew uint16 (T=8,C=2,Z=800,Y=300,X=2048) theta=30 px=0.386 N=3 crop"),
    # configs[0]
    "deskew_c1": dict(kind="deskew", shape=(256, 256, 512), dtype="uint16", units=16,
                      ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                      average_n_slices=1, e2e_units=8, ncu_traffic=2.273e8,
                      desc="C1 deskew uint16 (Z=256,Y=256,X=512) theta=30 px=0.386 N=1 crop; 16 distinct volumes"),
    # configs[2]
    "register_c3": dict(kind="register", shape=(120, 2048, 2048), dtype="float32", units=8,
                        e2e_units=4, ncu_traffic=4.3879e9,
                        desc="C3 register float32 (Z=120,Y=2048,X=2048) rot 7.3deg scale 1.07 shift (0.4,3.25,-11.5) order 1"),
    # C3 geometry with a NON z-separable matrix (small 3-D rotation about Y and X on top of C3):
    # what an ESTIMATED registration matrix looks like; exercises the generic kernel
    "register_generic": dict(kind="register", shape=(120, 2048, 2048), dtype="float32", units=8,
                             e2e_units=4, generic=True, ncu_traffic=3.9289e9,
                             desc="C3 shape float32 (120,2048,2048), generic 3-D affine (C3 matrix + 0.5/0.3 deg out-of-plane rotations), order 1"),
    # configs[3]
    "stabilize_c4": dict(kind="stabilize", shape=(64, 2048, 2048), dtype="float32", units=16,
                         e2e_units=8, ncu_traffic=2.1203e9,
                         desc="C4 stabilize float32 (Z=64,Y=2048,X=2048) fractional XYZ translations"),
    # the pipeline stage before deskew (not part of BASELINE's metric; here for
    # its roofline line): median over Z + exact float64 divide, uint16 -> float32
    "flatfield": dict(kind="flatfield", shape=(800, 300, 2048), dtype="uint16", units=8, e2e_units=2,
                      desc="flat-field (median over Z, divide) uint16 (Z=800,Y=300,X=2048) -> float32"),
    # configs[4] (per position/timepoint unit): deskew (C2 parameters) then register the deskewed
    # float32 (100,2048,1813) volume onto the same shape, intermediate resident in HBM
    "chain_c5": dict(kind="chain", shape=(800, 300, 2048), dtype="uint16", units=8,
                     ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                     average_n_slices=3, e2e_units=4,
                     desc="C5 unit: deskew uint16 (800,300,2048) N=3 then register f32 (100,2048,1813) rot 7.3deg scale 1.07, intermediate on device"),
    # configs[4] at plate level
    "plate_c5": dict(kind="chain", plate=True, shape=(800, 300, 2048), dtype="uint16", units=8,
                     positions=8, timepoints=32,
                     ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                     average_n_slices=3, e2e_units=4,
                     desc="C5 plate: 8 positions x T=32 units, deskew uint16 (800,300,2048) N=3 then register f32 (100,2048,1813), sharded over (position,t) by units_for_rank"),
}
EXTRA_WORKLOADS = ("deskew_c1", "register_c3", "stabilize_c4", "chain_c5", "register_generic")


# --------------------------------------------------------------------------------------------
def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi takes ~1 s to start on an 8-GPU box: wait until it reports."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)
        return time.perf_counter()

    def summary(self, since=0.0, until=None):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, row in list(self.rows):
            if stamp < since or (until is not None and stamp > until):
                continue
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def stabilize_matrices(n, seed=3000):
    rng = np.random.default_rng(seed)
    walk = np.cumsum(rng.normal(0.0, 1.5, size=(n, 3)), axis=0)
    mats = []
    for s in walk:
        m = np.eye(4)
        m[:3, 3] = s
        mats.append(m)
    return mats


def register_matrix_c3(shape, angle_deg=7.3, scale_yx=1.07, shift_zyx=(0.4, 3.25, -11.5)):
    """C3 matrix from the package's own builders (reference register.py:32-111)."""
    import biahub_b200 as b2

    T = np.eye(4)
    T[:3, 3] = shift_zyx
    return (T @ b2.get_3D_rotation_matrix(shape, angle_deg)
            @ b2.get_3D_rescaling_matrix(shape, (1.0, scale_yx, scale_yx)))


def out_of_plane_rotation(shape, deg_about_y, deg_about_x):
    """Small rotations that mix Z with X and Z with Y about the volume centre."""
    c = (np.array(shape) - 1) / 2.0
    a, b = np.radians(deg_about_y), np.radians(deg_about_x)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1]])
    R = Ry @ Rx
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = c - R @ c
    return M


def deskew_out_shape(w):
    """Deskewed (Z', Y', X') from the float64 shape logic (reference deskew.py:213-274; the same
    expressions as biahub_b200.get_deskewed_data_shape, restated so the CPU arm needs no GPU lib)."""
    Z, Y, X = w["shape"]
    ct = np.cos(w["ls_angle_deg"] * np.pi / 180)
    px = w["px_to_scan_ratio"]
    Xp = int(np.ceil(Z / px + Y * ct)) if w["keep_overhang"] else int(np.ceil(Z / px - Y * ct))
    return (int(np.ceil(Y / w["average_n_slices"])), X, Xp)


def unit_geometry(w):
    """(out_shape, algorithmic bytes per unit, output voxels per unit)."""
    Z, Y, X = w["shape"]
    esz = 2 if w["dtype"] == "uint16" else 4
    out_shape = deskew_out_shape(w) if w["kind"] in ("deskew", "chain") else tuple(w["shape"])
    out_vox = int(np.prod(out_shape))
    # every source voxel once + every output voxel once
    bytes_unit = Z * Y * X * esz + out_vox * 4
    if w["kind"] == "chain":  # + the register pass over the deskewed volume (8 B per voxel)
        bytes_unit += out_vox * 8
    return tuple(int(v) for v in out_shape), bytes_unit, out_vox


def workload_matrices(w, out_shape, units):
    if w["kind"] == "chain":
        return [register_matrix_c3(out_shape)] * units
    if w["kind"] == "register":
        M = register_matrix_c3(w["shape"])
        if w.get("generic"):
            M = M @ out_of_plane_rotation(w["shape"], 0.5, 0.3)
        return [M] * units
    if w["kind"] == "stabilize":
        return stabilize_matrices(units)
    return [None] * units


def kernel_name(w):
    return {"deskew": "deskew_tma_kernel",
            "chain": "deskew_tma_kernel + affine_zsep_kernel (bytes and time of both)",
            "flatfield": "flatfield_median_kernel + flatfield_apply_kernel (time of both; algorithmic bytes "
                         "= one read + one write; the radix select makes two sweeps over the source)"}.get(
        w["kind"], "affine_brick_kernel" if w.get("generic") else "affine_zsep_kernel")


def e2e_api(w):
    return ({"flatfield": "biahub_b200._flat_field_czyx",
             "deskew": "biahub_b200._fast_deskew_czyx",
             "chain": "biahub_b200.deskew_then_register (b2h_deskew_affine3d)"}.get(
        w["kind"], "biahub_b200.affine_warp (apply_affine_transform/apply_stabilization_transform body)")
        + " with pinned host in/out -> b2h_* C-ABI")


# --------------------------------------------------------------------------------------------
class Ctx:
    """Per-process state of the GPU arm: device, process group, clock sampler, pinned arenas."""

    def __init__(self, rank, world, local_rank):
        import torch

        from biahub_b200 import _cabi
        from biahub_b200._device import bind_to_gpu_numa

        self.rank, self.world = rank, world
        visible = torch.cuda.device_count()
        stride = max(1, visible // world) if visible > world else 1
        self.dev_index = (local_rank * stride) % max(visible, 1)
        torch.cuda.set_device(self.dev_index)
        self.dev = torch.device("cuda", self.dev_index)
        self.numa = bind_to_gpu_numa(self.dev_index) if world > 1 else {"numa_node": None}
        _cabi.require_device(self.dev_index)
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            self.dist = dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.dev_index)
        self.sampler.start()
        self.sampler.wait_first_sample()
        self._arena = {}

    def barrier(self):
        import torch

        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, value):
        import torch

        t = torch.tensor([float(value)], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def arena(self, name, nbytes):
        """Pinned host bytes reused by every workload of this run (cudaHostAlloc is slow)."""
        import torch

        have = self._arena.get(name)
        if have is None or have.numel() < nbytes:
            self._arena[name] = None
            have = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
            self._arena[name] = have
        return have

    def close(self):
        self.sampler.stop()
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def make_sources(w, ctx, units):
    """Distinct synthetic volumes resident in HBM (total footprint >> 126 MB L2)."""
    import torch

    Z, Y, X = w["shape"]
    gen = torch.Generator(device=ctx.dev)
    srcs = []
    for u in range(units):
        gen.manual_seed(1000 + ctx.rank * units + u)
        if w["dtype"] == "uint16":
            t = torch.randint(0, 65536, (Z, Y, X), generator=gen, device=ctx.dev, dtype=torch.int32)
            srcs.append(t.to(torch.uint16))
            del t
        else:
            srcs.append(torch.rand((Z, Y, X), generator=gen, device=ctx.dev, dtype=torch.float32) * 4095.0)
    return srcs


def pcie_ceiling(ctx, in_bytes, out_bytes, h_in, h_out, reps=3):
    """Plain pinned cudaMemcpyAsync of one unit's bytes, H2D and D2H at once, all ranks at once:
    the host-side ceiling of the end-to-end number on this box."""
    import torch

    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=ctx.dev)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in[:in_bytes], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[:out_bytes].copy_(d_out, non_blocking=True)

    once()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0) / reps
    del d_in, d_out
    return dt


def measure(args, w, ctx, steps, warmup, with_cpu=False, pageable=False, min_load_s=0.5):
    """Kernel-only step + end-to-end step of one workload on this rank; returns the result dict
    (rank-0 values are the max over ranks)."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    out_shape, bytes_unit, out_vox = unit_geometry(w)
    units = w["units"]
    Z, Y, X = w["shape"]
    esz = 2 if w["dtype"] == "uint16" else 4
    srcs = make_sources(w, ctx, units)
    mats = workload_matrices(w, out_shape, units)
    outs = [None] * units

    def device_step():
        for u in range(units):
            if w["kind"] == "deskew":
                outs[u] = b2.fast_deskew_zyx(srcs[u], w["ls_angle_deg"], w["px_to_scan_ratio"],
                                             w["keep_overhang"], w["average_n_slices"])
            elif w["kind"] == "chain":
                outs[u] = b2.deskew_then_register(
                    srcs[u], mats[u], out_shape, ls_angle_deg=w["ls_angle_deg"],
                    px_to_scan_ratio=w["px_to_scan_ratio"], keep_overhang=w["keep_overhang"],
                    average_n_slices=w["average_n_slices"])
            elif w["kind"] == "flatfield":
                outs[u] = b2.flat_field._flatfield_tensor(srcs[u], torch.float32)
            else:
                outs[u] = b2.affine_warp(srcs[u], mats[u], out_shape, order=1, boundary="itk")

    load_start = time.perf_counter()
    warm_done = 0
    while warm_done < warmup or time.perf_counter() - load_start < min_load_s:
        device_step()          # untimed warm-up: at least W steps and some load for the sampler
        warm_done += 1
        if warm_done % 4 == 0:
            torch.cuda.synchronize()
    ctx.barrier()
    launches0 = _cabi.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(steps):
        device_step()
    e1.record()
    ctx.barrier()
    t_end = time.perf_counter()
    kernel_launches = _cabi.launch_count() - launches0
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / steps
    value = ctx.world * units * out_vox / (ms_per_step * 1e-3) / 1e9
    # one kernel launch per unit (two for the chained and flat-field workloads: bytes and time of both)
    launch_ms = ms_total / max(steps * units, 1)
    time.sleep(0.12)  # let the 100 ms sampler see the tail of the timed region
    clocks = ctx.sampler.summary(since=load_start + 0.2 * min_load_s, until=t_end + 0.1)

    # ---- end to end through the reference-facing call with host buffers (pinned in, pinned out)
    e2e_units = 0 if args.no_e2e else w["e2e_units"]
    n_buf = min(4, e2e_units, units)   # distinct host volumes, cycled over the e2e units
    in_bytes, out_bytes = Z * Y * X * esz, out_vox * 4
    e2e = None
    if e2e_units:
        a_in = ctx.arena("in", n_buf * in_bytes)
        a_out = ctx.arena("out", n_buf * out_bytes)
        np_dt = np.uint16 if w["dtype"] == "uint16" else np.float32
        h_in, h_out = [], []
        for u in range(n_buf):
            seg = a_in[u * in_bytes:(u + 1) * in_bytes]
            seg.copy_(srcs[u].view(torch.uint8).reshape(-1))
            h_in.append(seg.numpy().view(np_dt).reshape(Z, Y, X))
            h_out.append(a_out[u * out_bytes:(u + 1) * out_bytes].numpy().view(np.float32).reshape(out_shape))
        torch.cuda.synchronize()
    del srcs, outs
    torch.cuda.empty_cache()
    if e2e_units:
        res_ff = [None]  # flat-field returns its (pooled, pinned) result instead of filling `out`

        def e2e_step():
            for k in range(e2e_units):
                u = k % n_buf
                if w["kind"] == "chain":
                    b2.deskew_then_register(
                        h_in[u], mats[u], out_shape, ls_angle_deg=w["ls_angle_deg"],
                        px_to_scan_ratio=w["px_to_scan_ratio"], keep_overhang=w["keep_overhang"],
                        average_n_slices=w["average_n_slices"], device=ctx.dev_index, out=h_out[u])
                elif w["kind"] == "flatfield":
                    res_ff[0] = None
                    res_ff[0] = b2._flat_field_czyx(h_in[u][None], [0], device=ctx.dev_index)
                elif w["kind"] == "deskew":
                    b2._fast_deskew_czyx(h_in[u][None], device=f"cuda:{ctx.dev_index}", out=h_out[u],
                                         ls_angle_deg=w["ls_angle_deg"],
                                         px_to_scan_ratio=w["px_to_scan_ratio"],
                                         keep_overhang=w["keep_overhang"],
                                         average_n_slices=w["average_n_slices"])
                else:
                    b2.affine_warp(h_in[u], mats[u], out_shape, order=1, boundary="itk",
                                   device=ctx.dev_index, out=h_out[u])

        e2e_steps = max(1, min(steps, 5))
        e2e_step()   # warm-up: grows the library's device buffers and rings
        ctx.barrier()
        launches_e2e0 = _cabi.launch_count()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()          # each call returns only when its output is complete on the host
        torch.cuda.synchronize()
        e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
        launches_e2e = _cabi.launch_count() - launches_e2e0
        e2e_value = ctx.world * e2e_units * e2e_steps * out_vox / e2e_s / 1e9
        first_out = (res_ff[0][0] if w["kind"] == "flatfield" and res_ff[0] is not None else h_out[0])
        check = float(first_out.ravel()[:: max(1, first_out.size // 1000)].astype(np.float64).sum())
        ceil_s = pcie_ceiling(ctx, in_bytes, out_bytes, a_in, a_out)
        # same call with ordinary (pageable) numpy arrays in and a fresh array out: what the
        # reference's process_single_position hands over; staged through the library's pinned rings
        pageable_value = None
        if pageable and ctx.rank == 0 and ctx.world == 1 and w["kind"] in ("deskew", "register", "stabilize"):
            src_pg = np.array(h_in[0], copy=True)
            res = None
            best_pg = None
            for rep in range(5):  # the first two calls grow the pinned result pool; steady state after
                res = None        # drop the previous result, as process_single_position does
                t0 = time.perf_counter()
                if w["kind"] == "deskew":
                    res = b2._fast_deskew_czyx(src_pg[None], device=f"cuda:{ctx.dev_index}",
                                               ls_angle_deg=w["ls_angle_deg"],
                                               px_to_scan_ratio=w["px_to_scan_ratio"],
                                               keep_overhang=w["keep_overhang"],
                                               average_n_slices=w["average_n_slices"])
                else:
                    res = b2.affine_warp(src_pg, mats[0], out_shape, order=1, boundary="itk",
                                         device=ctx.dev_index)
                dt_pg = time.perf_counter() - t0
                if rep >= 2:
                    best_pg = dt_pg if best_pg is None else min(best_pg, dt_pg)
            pageable_value = out_vox / best_pg / 1e9
            del res, src_pg
        e2e = {"value": round(e2e_value, 3), "unit": "Gvoxels/s",
               "h2d_bytes_per_step": int(e2e_units * in_bytes),
               "d2h_bytes_per_step": int(e2e_units * out_bytes),
               "steps": e2e_steps, "units_per_step_per_gpu": e2e_units, "distinct_host_volumes": n_buf,
               "api": e2e_api(w), "gpu_launches": int(launches_e2e), "checksum": check,
               "numa_node": ctx.numa.get("numa_node"),
               # plain pinned cudaMemcpyAsync of one unit's bytes, both directions at once, all
               # ranks at once: the host-side ceiling of this number on this box
               "ceiling": {"value": round(ctx.world * out_vox / ceil_s / 1e9, 3), "unit": "Gvoxels/s",
                           "h2d_gbs": round(ctx.world * in_bytes / ceil_s / 1e9, 1),
                           "d2h_gbs": round(ctx.world * out_bytes / ceil_s / 1e9, 1)},
               "pageable_value": None if pageable_value is None else round(pageable_value, 3)}

    peak, peak_src = read_peaks()
    achieved = bytes_unit / (launch_ms * 1e-3) / 1e9
    res = {
        "value": round(value, 3), "unit": "Gvoxels/s", "ms_per_step": round(ms_per_step, 4),
        "steps": steps, "warmup": warmup,
        "config": {"workload": w["desc"], "source_dtype": w["dtype"], "units_per_step_per_gpu": units,
                   "out_shape": list(out_shape),
                   "sharding": "independent (position,t,c) units per rank, no collective",
                   "l2": f"inputs+outputs resident per step = {units * bytes_unit / 1e9:.1f} GB >> 126 MB L2 (no flush needed)"},
        "clocks": clocks, "gpu_launches": int(kernel_launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4),
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel in the
                     # committed ncu --set full capture (profiles/): a constant, NOT measured in this run
                     "traffic": w.get("ncu_traffic"), "traffic_source": "profiles/r2_*.txt (ncu --set full capture of this kernel; a constant, not measured in this run)",
                     "kernel": kernel_name(w), "algorithmic_bytes_per_launch": int(bytes_unit),
                     "launch_ms": round(launch_ms, 4), "peak_source": peak_src},
    }
    if e2e is not None:
        res["e2e"] = e2e
    return res


# --------------------------------------------------------------------------------------------
# plate-level C5 (BASELINE.json configs[4]): 8 positions x T=32, deskew -> register, whole-box
# sharding over (position, t)
# --------------------------------------------------------------------------------------------
def plate_unit(czyx, matrix=None, output_shape_zyx=None, device=None, **deskew_kwargs):
    """Module-level per-unit callable (what ``process_single_position`` would be handed)."""
    import biahub_b200 as b2

    return b2.deskew_then_register(czyx[0], matrix, output_shape_zyx, device=device, **deskew_kwargs)[None]


def run_plate(args, w, ctx):
    """One timed step = the whole plate once: every rank takes its share of the 256 units
    (``units_for_rank``) and streams them through ``run_units_overlapped`` (reader / GPU / writer
    threads) calling ``deskew_then_register`` with host arrays in and out."""
    import torch

    from biahub_b200 import _cabi
    from biahub_b200.sharding import enumerate_units, run_units_overlapped, units_for_rank

    out_shape, bytes_unit, out_vox = unit_geometry(w)
    Z, Y, X = w["shape"]
    units = enumerate_units(w["positions"], range(w["timepoints"]), [0])
    mine = units_for_rank(units, ctx.rank, ctx.world)
    in_bytes = Z * Y * X * 2
    n_src = 4
    a_in = ctx.arena("in", n_src * in_bytes)
    gen = torch.Generator(device=ctx.dev)
    pool = []
    for u in range(n_src):
        gen.manual_seed(4000 + ctx.rank * n_src + u)
        t = torch.randint(0, 65536, (Z, Y, X), generator=gen, device=ctx.dev, dtype=torch.int32).to(torch.uint16)
        seg = a_in[u * in_bytes:(u + 1) * in_bytes]
        seg.copy_(t.view(torch.uint8).reshape(-1))
        pool.append(seg.numpy().view(np.uint16).reshape(1, Z, Y, X))
        del t
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    M = register_matrix_c3(out_shape)
    kw = dict(matrix=M, output_shape_zyx=out_shape, device=ctx.dev_index,
              ls_angle_deg=w["ls_angle_deg"], px_to_scan_ratio=w["px_to_scan_ratio"],
              keep_overhang=w["keep_overhang"], average_n_slices=w["average_n_slices"])
    T = w["timepoints"]
    sums = [0.0]

    def read_unit(p, t, c):     # stands in for the zarr chunk decode: a (C, Z, Y, X) host block
        return pool[(p * T + t) % n_src]

    def write_unit(p, t, c, out):   # stands in for the zarr write: consume, then drop the array
        sums[0] += float(out[0, 0, 0, 0]) + float(out[0, -1, -1, -1])

    def plate_pass(sel):
        return run_units_overlapped(plate_unit, read_unit, write_unit, sel, prefetch=2, **kw)

    plate_pass(mine[: min(3, len(mine))])   # warm-up: device buffers, rings, result pool
    ctx.barrier()
    steps = max(1, args.plate_steps)
    l0 = _cabi.launch_count()
    t_begin = time.perf_counter()
    for _ in range(steps):
        done = plate_pass(mine)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t_begin) / steps
    t_end = time.perf_counter()
    launches = _cabi.launch_count() - l0
    time.sleep(0.12)
    clocks = ctx.sampler.summary(since=t_begin, until=t_end + 0.1)
    total_vox = len(units) * out_vox
    ceil_s = pcie_ceiling(ctx, in_bytes, out_vox * 4, a_in, ctx.arena("out", out_vox * 4))
    return {
        "value": round(total_vox / dt / 1e9, 3), "unit": "Gvoxels/s", "seconds_per_plate": round(dt, 3),
        "units_total": len(units), "units_this_rank": len(mine), "units_done_last_pass": int(done),
        "steps": steps, "scaling": "strong",
        "h2d_bytes_per_step": int(len(units) * in_bytes), "d2h_bytes_per_step": int(len(units) * out_vox * 4),
        "api": "sharding.units_for_rank + sharding.run_units_overlapped(plate_unit -> "
               "biahub_b200.deskew_then_register -> b2h_deskew_affine3d), pinned host volumes in, "
               "pooled pinned arrays out",
        "distinct_host_volumes": n_src, "gpu_launches": int(launches), "checksum": sums[0],
        "ceiling": {"value": round(ctx.world * out_vox / ceil_s / 1e9, 3), "unit": "Gvoxels/s"},
        "clocks": clocks,
    }


# --------------------------------------------------------------------------------------------
def run_b200(args, w, rank, world, local_rank):
    ctx = Ctx(rank, world, local_rank)
    warm = max(args.warmup, 3)
    head = measure(args, w, ctx, args.steps, warm, pageable=True)
    result = {
        "metric": METRIC, "value": head["value"], "unit": "Gvoxels/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        # the arithmetic type of the path (sources are stored as uint16 or float32, see config)
        "dtype": "f32" if w["kind"] != "flatfield" else "f64",
        "data": "synthetic (uniform noise, seeded, generated on device; e2e copies of the same volumes in pinned host memory)",
        "config": head["config"], "clocks": head["clocks"], "gpu_launches": head["gpu_launches"],
        "device": ctx.dev_index,
    }
    if "e2e" in head:
        result["e2e"] = head["e2e"]
    result["roofline"] = head["roofline"]
    if w.get("plate"):
        result["plate"] = run_plate(args, w, ctx)
        # the plate pass is the end-to-end number of this workload
        result["e2e"] = {"value": result["plate"]["value"], "unit": "Gvoxels/s",
                         "h2d_bytes_per_step": result["plate"]["h2d_bytes_per_step"],
                         "d2h_bytes_per_step": result["plate"]["d2h_bytes_per_step"],
                         "api": result["plate"]["api"], "scaling": "strong",
                         "ceiling": result["plate"]["ceiling"]}
    extras = []
    if args.extra:
        for name in EXTRA_WORKLOADS:
            if name == args.workload:
                continue
            ew = dict(WORKLOADS[name])
            ew["units"] = min(ew["units"], 8 if ew["shape"][1] < 1024 else 4)
            ew["e2e_units"] = min(ew["e2e_units"], 4)
            r = measure(args, ew, ctx, steps=max(3, min(args.steps, 5)), warmup=3, min_load_s=0.3)
            entry = {"name": name, "workload": ew["desc"], "value": r["value"], "unit": "Gvoxels/s",
                     "ms_per_step": r["ms_per_step"], "steps": r["steps"], "warmup": r["warmup"],
                     "units_per_step_per_gpu": ew["units"], "launch_ms": r["roofline"]["launch_ms"],
                     "roofline": {k: r["roofline"][k] for k in ("frac", "achieved", "peak", "kernel",
                                                                "algorithmic_bytes_per_launch")},
                     "clocks": r["clocks"], "gpu_launches": r["gpu_launches"]}
            if "e2e" in r:
                entry["e2e"] = {k: r["e2e"][k] for k in ("value", "unit", "h2d_bytes_per_step",
                                                         "d2h_bytes_per_step", "steps",
                                                         "units_per_step_per_gpu", "ceiling")}
            extras.append(entry)
        result["workloads"] = extras
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(w, budget_s=24.0)
    ctx.close()
    return result


# --------------------------------------------------------------------------------------------
# CPU arms: the reference's own implementation on the host cores
# --------------------------------------------------------------------------------------------
def load_staged_reference_deskew():
    """The UNMODIFIED reference ``biahub.deskew`` from baseline/_ref (scripts/make_baseline_ref.py),
    or None when nothing has been staged (then the oracle port is timed, kind "port")."""
    try:
        from oracle import ref_loader

        if not ref_loader.use_staged_reference():
            return None
        return ref_loader.load_reference_deskew()
    except Exception as exc:  # noqa: BLE001 - report and fall back to the port
        sys.stderr.write(f"bench: staged reference not usable ({exc!r}); timing the oracle port\n")
        return None


def cpu_deskew_unit(w):
    """(callable, output voxels, threads, kind, description): ONE WHOLE unit of the workload through
    the reference's ``_fast_deskew_czyx(device="cpu")`` (biahub/deskew.py:551-579)."""
    import torch

    Z, Y, X = w["shape"]
    rng = np.random.default_rng(1000)
    raw = rng.integers(0, 65536, size=(1, Z, Y, X), dtype=np.uint16)
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    out_vox = int(np.prod(deskew_out_shape(w)))
    kw = dict(ls_angle_deg=w["ls_angle_deg"], px_to_scan_ratio=w["px_to_scan_ratio"],
              keep_overhang=w["keep_overhang"], average_n_slices=w["average_n_slices"])
    ref = load_staged_reference_deskew()
    if ref is not None:
        def fn():
            return ref._fast_deskew_czyx(raw, device="cpu", **kw)

        return fn, out_vox, threads, "reference", (
            f"1 whole unit (uint16 {Z}x{Y}x{X}) through the UNMODIFIED reference "
            f"biahub.deskew._fast_deskew_czyx(device='cpu') (biahub/deskew.py:551-579) loaded from "
            f"baseline/_ref, torch CPU, {threads} threads")
    from oracle import deskew_oracle as do

    def fn():
        return do.deskew_oracle_torch(raw[0], **kw)

    return fn, out_vox, threads, "port", (
        f"1 whole unit (uint16 {Z}x{Y}x{X}); baseline/_ref not staged: CPU torch port of reference "
        f"fast_deskew_zyx stages (biahub/deskew.py:505-536), {threads} threads")


_CPU_AFFINE_JOB = None


def _cpu_affine_init(job):
    global _CPU_AFFINE_JOB
    _CPU_AFFINE_JOB = job


def _cpu_affine_slab(_i):
    from oracle import affine_oracle as ao

    vol, M, shape = _CPU_AFFINE_JOB
    return float(ao.affine_oracle_scipy(vol, M, shape, 1)[0, 0, 0])


def cpu_affine_sample(shape, M, zs=4):
    """scipy is single-threaded C; the reference fans (t, c) units out over a process pool (iohub
    process_single_position, num_workers): one slab of `zs` output planes per host thread."""
    import multiprocessing as mp

    Z, Y, X = shape
    zs = min(Z, zs)
    threads = max(1, len(os.sched_getaffinity(0)))
    rng = np.random.default_rng(1000)
    vol = (rng.random((zs + 2, Y, X), dtype=np.float32) * 4095).astype(np.float32)
    # spawn, not fork: the GPU arm calls this after CUDA has been initialised in this process
    _cpu_affine_init((vol, M, (zs, Y, X)))
    pool = (mp.get_context("spawn").Pool(threads, initializer=_cpu_affine_init,
                                         initargs=((vol, M, (zs, Y, X)),)) if threads > 1 else None)

    def fn():
        if pool is None:
            return _cpu_affine_slab(0)
        return pool.map(_cpu_affine_slab, range(threads), chunksize=1)

    fn.close = (lambda: pool.terminate()) if pool is not None else (lambda: None)
    desc = (f"{threads} slabs of {zs} output planes of one unit (float32 {zs}x{Y}x{X} each), one per "
            f"host thread in a process pool; scipy.ndimage.affine_transform order=1 (library of the "
            f"reference's method='scipy' branch, biahub/register.py:272; the default ANTs branch is "
            f"not installable — the reference calls it ~10x faster than scipy, register.py:256-257)")
    return fn, threads * zs * Y * X, threads, "port", desc


def cpu_sample(w):
    """A bounded sample of the workload for the CPU arm:
    (callable, output voxels it produces, threads, kind, description)."""
    Z, Y, X = w["shape"]
    if w["kind"] == "deskew":
        return cpu_deskew_unit(w)
    if w["kind"] == "flatfield":
        from oracle import flatfield_oracle as fo

        xs = min(X, 128)
        raw = np.random.default_rng(1000).integers(90, 1200, size=(Z, Y, xs), dtype=np.uint16)

        def fn():
            return fo.flat_field_czyx_oracle(raw[None], [0])

        return fn, Z * Y * xs, 1, "port", (
            f"1 unit restricted to X={xs} of {X} columns (uint16 {Z}x{Y}x{xs}); numpy port of "
            f"reference _flat_field_czyx (biahub/flat_field.py:105-166), 1 thread")
    if w["kind"] == "chain":
        # one whole unit through the reference deskew, then the register pass of the deskewed
        # volume extrapolated from a slab sample (same voxels out): time = t_deskew + vox / rate
        out_shape = deskew_out_shape(w)
        d_fn, out_vox, threads, kind, d_desc = cpu_deskew_unit(w)
        from oracle import affine_oracle as ao

        a_fn, a_vox, _, _, a_desc = cpu_affine_sample(out_shape, ao.register_matrix_c3(out_shape))
        state = {"rate": None}

        def fn():
            d_fn()
            t0 = time.perf_counter()
            a_fn()
            state["rate"] = a_vox / (time.perf_counter() - t0)
            # spend the time the rest of the register pass would take at the sampled rate
            time.sleep(max(0.0, (out_vox - a_vox) / state["rate"]))

        fn.close = a_fn.close
        return fn, out_vox, threads, kind, (
            d_desc + "; THEN the register pass of the deskewed volume: " + a_desc +
            ", the remaining planes charged at the sampled rate")
    from oracle import affine_oracle as ao

    M = ao.register_matrix_c3((Z, Y, X)) if w["kind"] == "register" else stabilize_matrices(2)[1]
    if w.get("generic"):
        M = M @ out_of_plane_rotation((Z, Y, X), 0.5, 0.3)
    return cpu_affine_sample((Z, Y, X), M)


def cpu_baseline(w, budget_s=24.0):
    fn, vox, threads, kind, desc = cpu_sample(w)
    fn()  # warm
    times = []
    t_start = time.perf_counter()
    while len(times) < 3:
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    best = min(times)
    getattr(fn, "close", lambda: None)()
    return {"value": round(vox / best / 1e9, 5), "unit": "Gvoxels/s", "cores": threads,
            "kind": kind, "sample": desc + f"; best of {len(times)} after 1 warm-up"}


def run_reference(args, w, rank, world):
    """Reference arm: the reference's CPU path, all host threads it can use; rank 0 only."""
    if rank != 0:
        return None
    fn, vox, threads, kind, desc = cpu_sample(w)
    warm = max(args.warmup, 3)
    steps = args.steps
    # keep the whole run within a few minutes whatever K and W the driver passes: one probe call
    # sizes the number of warm-up / timed calls actually made (reported in the line)
    t0 = time.perf_counter()
    fn()
    probe = time.perf_counter() - t0
    budget = 240.0
    if probe * (warm + steps) > budget:
        scale = budget / (probe * (warm + steps))
        warm = max(1, int(warm * scale))
        steps = max(1, int(steps * scale))
    for _ in range(max(0, warm - 1)):   # the probe call was the first warm-up
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    getattr(fn, "close", lambda: None)()
    value = steps * vox / dt / 1e9
    out_shape, bytes_unit, _ = unit_geometry(w)
    return {
        "impl": "reference", "metric": METRIC,
        "value": round(value, 5), "unit": "Gvoxels/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (uniform noise, seeded)",
        # the GPU arm's config (same workload, same keys); what one CPU step covers is in cpu_baseline.sample
        "config": {"workload": w["desc"], "source_dtype": w["dtype"], "units_per_step_per_gpu": w["units"],
                   "out_shape": list(out_shape),
                   "sharding": "independent (position,t,c) units per rank, no collective",
                   "l2": f"inputs+outputs resident per step = {w['units'] * bytes_unit / 1e9:.1f} GB >> 126 MB L2 (no flush needed)"},
        "requested": {"steps": args.steps, "warmup": args.warmup},
        "cpu_baseline": {"value": round(value, 5), "unit": "Gvoxels/s", "cores": threads,
                         "kind": kind, "sample": desc},
        "e2e": {"value": round(value, 5), "unit": "Gvoxels/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="deskew_c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel-only (for short ncu captures)")
    ap.add_argument("--no-extra", dest="extra", action="store_false",
                    help="skip the short runs of the other BASELINE configs (`workloads` array)")
    ap.add_argument("--units", type=int, default=0, help="override units per step per GPU")
    ap.add_argument("--plate-steps", type=int, default=1, help="plate passes timed by --workload plate_c5")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.units > 0:
        w["units"] = args.units
        w["e2e_units"] = min(w["e2e_units"], args.units)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if w.get("plate"):
        # reader / GPU / writer stages keep up to four 1.5 GB results alive per rank
        os.environ.setdefault("BIAHUB_B200_PINNED_POOL_MB", "8192")
    if args.impl == "reference":
        res = run_reference(args, w, rank, world)
    else:
        res = run_b200(args, w, rank, world, local_rank)
    if rank == 0 and res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
