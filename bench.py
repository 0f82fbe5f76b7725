#!/usr/bin/env python
"""Benchmark of the B200-native 3-D affine resampling path (BASELINE.json metric:
output Gvoxels/s; % of HBM bandwidth).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic (position, t, c) units that is
resident in HBM when the timed region starts; ``value`` is whole-job output Gvoxels/s (all
ranks' voxels / max-over-ranks device time).  ``e2e`` is the same metric through the
reference-facing Python call (``_fast_deskew_czyx`` / ``apply_affine_transform`` /
``apply_stabilization_transform``) with HOST buffers: pinned H2D of every input and D2H of every
output inside the timed region.  ``roofline`` compares the dominant kernel's algorithmic bytes /
launch time with the measured HBM peak (MEASURED_PEAKS.json).  ``cpu_baseline`` times the CPU
restatement of the reference's algorithm (oracle/, kind "port": the reference is pure Python
whose arithmetic lives in torch/scipy, nothing to compile) on a bounded sample on rank 0.

Multi-GPU: one process per GPU (torchrun); units are sharded over ranks with NO data-path
collective (the reference's parallelism is independent (position, t, c) units) — weak scaling:
every rank processes the same number of units per step.  torch.distributed is used only for the
barrier and the max-over-ranks of the device time.

``--impl reference`` times the CPU path (oracle port of the reference's algorithm, all host
threads) on the same workload/metric, on a bounded sample per step; rank 0 only.
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

WORKLOADS = {
    # BASELINE.json configs[1]: mantis-sized deskew, average_n_slices=3, uint16 (T=8,C=2,800,300,2048)
    "deskew_c2": dict(kind="deskew", shape=(800, 300, 2048), dtype="uint16", units=16,
                      ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                      average_n_slices=3, e2e_units=2, ncu_traffic=2.3523e9,
                      desc="C2 mantis deskew uint16 (T=8,C=2,Z=800,Y=300,X=2048) theta=30 px=0.386 N=3 crop"),
    # configs[0]
    "deskew_c1": dict(kind="deskew", shape=(256, 256, 512), dtype="uint16", units=16,
                      ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                      average_n_slices=1, e2e_units=8, ncu_traffic=2.262e8,
                      desc="C1 deskew uint16 (Z=256,Y=256,X=512) theta=30 px=0.386 N=1 crop; 16 distinct volumes"),
    # configs[2]
    "register_c3": dict(kind="register", shape=(120, 2048, 2048), dtype="float32", units=8,
                        e2e_units=2, ncu_traffic=4.5084e9,
                        desc="C3 register float32 (Z=120,Y=2048,X=2048) rot 7.3deg scale 1.07 shift (0.4,3.25,-11.5) order 1"),
    # C3 geometry with a NON z-separable matrix (small 3-D rotation about Y and X on top of C3):
    # exercises the generic path
    "register_generic": dict(kind="register", shape=(120, 2048, 2048), dtype="float32", units=8,
                             e2e_units=2, generic=True, ncu_traffic=3.9261e9,
                             desc="C3 shape float32 (120,2048,2048), generic 3-D affine (C3 matrix + 0.5/0.3 deg out-of-plane rotations), order 1"),
    # configs[3]
    "stabilize_c4": dict(kind="stabilize", shape=(64, 2048, 2048), dtype="float32", units=16,
                         e2e_units=4, ncu_traffic=2.1182e9,
                         desc="C4 stabilize float32 (Z=64,Y=2048,X=2048) fractional XYZ translations"),
    # the pipeline stage before deskew (not part of BASELINE's metric; here for
    # its roofline line): median over Z + exact float64 divide, uint16 -> float32
    "flatfield": dict(kind="flatfield", shape=(800, 300, 2048), dtype="uint16", units=8, e2e_units=2,
                      desc="flat-field (median over Z, divide) uint16 (Z=800,Y=300,X=2048) -> float32"),
    # configs[4] (per position/timepoint unit): deskew (C2 parameters) then register the deskewed
    # float32 (100,2048,1813) volume onto the same shape, intermediate resident in HBM
    "chain_c5": dict(kind="chain", shape=(800, 300, 2048), dtype="uint16", units=8,
                     ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False,
                     average_n_slices=3, e2e_units=2,
                     desc="C5 unit: deskew uint16 (800,300,2048) N=3 then register f32 (100,2048,1813) rot 7.3deg scale 1.07, intermediate on device"),
}


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

    def stop(self, since=0.0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, row in self.rows:
            if stamp < since:
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
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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


def unit_geometry(w):
    """(out_shape, algorithmic bytes per unit, output voxels per unit)."""
    import biahub_b200 as b2

    Z, Y, X = w["shape"]
    esz = 2 if w["dtype"] == "uint16" else 4
    if w["kind"] in ("deskew", "chain"):
        out_shape, _ = b2.get_deskewed_data_shape(w["shape"], w["ls_angle_deg"], w["px_to_scan_ratio"],
                                                  w["keep_overhang"], w["average_n_slices"])
    else:
        out_shape = tuple(w["shape"])
    out_vox = int(np.prod(out_shape))
    # every source voxel once + every output voxel once
    bytes_unit = Z * Y * X * esz + out_vox * 4
    if w["kind"] == "chain":  # + the register pass over the deskewed volume (8 B per voxel)
        bytes_unit += out_vox * 8
    return tuple(int(v) for v in out_shape), bytes_unit, out_vox


# --------------------------------------------------------------------------------------------
def run_b200(args, w, rank, world, local_rank):
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi
    from biahub_b200._device import pinned_empty

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from biahub_b200._device import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"numa_node": None}
    _cabi.require_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    out_shape, bytes_unit, out_vox = unit_geometry(w)
    units = w["units"]
    Z, Y, X = w["shape"]

    # ---- synthetic units resident in HBM (distinct volumes; total footprint >> 126 MB L2)
    gen = torch.Generator(device=dev)
    srcs = []
    for u in range(units):
        gen.manual_seed(1000 + rank * units + u)
        if w["dtype"] == "uint16":
            t = torch.randint(0, 65536, (Z, Y, X), generator=gen, device=dev, dtype=torch.int32)
            srcs.append(t.to(torch.uint16))
            del t
        else:
            srcs.append(torch.rand((Z, Y, X), generator=gen, device=dev, dtype=torch.float32) * 4095.0)
    if w["kind"] == "chain":
        mats = [register_matrix_c3(out_shape)] * units
    elif w["kind"] == "register":
        M = register_matrix_c3(w["shape"])
        if w.get("generic"):
            M = M @ out_of_plane_rotation(w["shape"], 0.5, 0.3)
        mats = [M] * units
    elif w["kind"] == "stabilize":
        mats = stabilize_matrices(units)
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

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    load_start = sampler.wait_first_sample()
    warm_done = 0
    while warm_done < max(args.warmup, 3) or time.perf_counter() - load_start < 0.5:
        device_step()          # untimed warm-up: at least W steps and 0.5 s of load for the sampler
        warm_done += 1
        if warm_done % 4 == 0:
            torch.cuda.synchronize()
    barrier()
    launches0 = _cabi.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    clocks = sampler.stop(since=load_start + 0.2)
    kernel_launches = _cabi.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * units * out_vox / (ms_per_step * 1e-3) / 1e9
    # one kernel launch per unit (two for the chained and flat-field workloads: bytes and time of both)
    launch_ms = ms_total / max(args.steps * units, 1)

    # ---- end to end through the reference-facing call with host buffers (pinned in, pinned out)
    e2e_units = min(w["e2e_units"], units)
    if args.no_e2e:
        e2e_units = 0
    h_in = []
    for u in range(e2e_units):
        buf = pinned_empty((Z, Y, X), np.uint16 if w["dtype"] == "uint16" else np.float32)
        if w["dtype"] == "uint16":
            torch.from_numpy(buf.view(np.int16)).copy_(srcs[u].view(torch.int16))
        else:
            torch.from_numpy(buf).copy_(srcs[u])
        h_in.append(buf)
    del srcs, outs
    torch.cuda.empty_cache()
    h_out = [pinned_empty(out_shape, np.float32) for _ in range(e2e_units if w["kind"] != "flatfield" else 0)]
    res_ff = [None]  # flat-field returns its (pooled, pinned) result instead of filling `out`

    def e2e_step():
        for u in range(e2e_units):
            if w["kind"] == "chain":
                b2.deskew_then_register(
                    h_in[u], mats[u], out_shape, ls_angle_deg=w["ls_angle_deg"],
                    px_to_scan_ratio=w["px_to_scan_ratio"], keep_overhang=w["keep_overhang"],
                    average_n_slices=w["average_n_slices"], device=local_rank, out=h_out[u])
            elif w["kind"] == "flatfield":
                res_ff[0] = None
                res_ff[0] = b2._flat_field_czyx(h_in[u][None], [0], device=local_rank)
            elif w["kind"] == "deskew":
                b2._fast_deskew_czyx(h_in[u][None], device=f"cuda:{local_rank}", out=h_out[u],
                                     ls_angle_deg=w["ls_angle_deg"],
                                     px_to_scan_ratio=w["px_to_scan_ratio"],
                                     keep_overhang=w["keep_overhang"],
                                     average_n_slices=w["average_n_slices"])
            else:
                b2.affine_warp(h_in[u], mats[u], out_shape, order=1, boundary="itk",
                               device=local_rank, out=h_out[u])

    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(2 if e2e_units else 0):
        e2e_step()
    barrier()
    launches_e2e0 = _cabi.launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()          # each call returns only when its output is complete on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches_e2e = _cabi.launch_count() - launches_e2e0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = world * e2e_units * e2e_steps * out_vox / e2e_s / 1e9
    esz = 2 if w["dtype"] == "uint16" else 4
    first_out = res_ff[0][0] if w["kind"] == "flatfield" and res_ff[0] is not None else (h_out[0] if h_out else None)
    check = (float(first_out.ravel()[:: max(1, first_out.size // 1000)].astype(np.float64).sum())
             if e2e_units and first_out is not None else None)
    # same call with ordinary (pageable) numpy arrays in and a fresh array out: what the
    # reference's process_single_position hands over; staged through the library's pinned rings
    pageable_value = None
    if e2e_units and rank == 0 and world == 1 and w["kind"] in ("deskew", "register", "stabilize"):
        src_pg = np.array(h_in[0], copy=True)
        res = None
        best_pg = None
        for rep in range(5):  # the first two calls grow the pinned result pool; steady state after
            res = None        # drop the previous result, as process_single_position does
            t0 = time.perf_counter()
            if w["kind"] == "deskew":
                res = b2._fast_deskew_czyx(src_pg[None], device=f"cuda:{local_rank}",
                                           ls_angle_deg=w["ls_angle_deg"],
                                           px_to_scan_ratio=w["px_to_scan_ratio"],
                                           keep_overhang=w["keep_overhang"],
                                           average_n_slices=w["average_n_slices"])
            else:
                res = b2.affine_warp(src_pg, mats[0], out_shape, order=1, boundary="itk",
                                     device=local_rank)
            dt_pg = time.perf_counter() - t0
            if rep >= 2:
                best_pg = dt_pg if best_pg is None else min(best_pg, dt_pg)
        pageable_value = out_vox / best_pg / 1e9
        del res, src_pg

    peak, peak_src = read_peaks()
    achieved = bytes_unit / (launch_ms * 1e-3) / 1e9
    result = {
        "metric": "output Gvoxels/s (deskew/register/stabilize resampling); % of HBM BW",
        "value": round(value, 3),
        "unit": "Gvoxels/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        # the arithmetic type of the path (sources are stored as uint16 or float32, see config)
        "dtype": "f32" if w["kind"] != "flatfield" else "f64",
        "data": "synthetic (uniform noise, seeded, generated on device; e2e copies of the same volumes in pinned host memory)",
        "config": {"workload": w["desc"], "source_dtype": w["dtype"], "units_per_step_per_gpu": units,
                   "out_shape": list(out_shape), "sharding": "independent (position,t,c) units per rank, no collective",
                   "l2": f"inputs+outputs resident per step = {units * bytes_unit / 1e9:.1f} GB >> 126 MB L2 (no flush needed)"},
        "clocks": clocks,
        "gpu_launches": int(kernel_launches),
        "e2e": {"value": round(e2e_value, 3), "unit": "Gvoxels/s",
                "h2d_bytes_per_step": int(e2e_units * Z * Y * X * esz),
                "d2h_bytes_per_step": int(e2e_units * out_vox * 4),
                "steps": e2e_steps, "units_per_step_per_gpu": e2e_units,
                "api": ("biahub_b200._flat_field_czyx" if w["kind"] == "flatfield" else "biahub_b200._fast_deskew_czyx" if w["kind"] == "deskew" else "biahub_b200.deskew_then_register (b2h_deskew_affine3d)" if w["kind"] == "chain" else "biahub_b200.affine_warp (apply_affine_transform/apply_stabilization_transform body)")
                       + " with pinned host in/out -> b2h_* C-ABI",
                "gpu_launches": int(launches_e2e), "checksum": check, "numa_node": numa.get("numa_node"),
                "pageable_value": None if pageable_value is None else round(pageable_value, 3)},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4),
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the
                     # committed ncu --set full capture of this workload (profiles/r1_*.txt)
                     "traffic": w.get("ncu_traffic"),
                     "kernel": {"deskew": "deskew_tma_kernel", "chain": "deskew_tma_kernel + affine_zsep_kernel (bytes and time of both)", "flatfield": "flatfield_median_kernel + flatfield_apply_kernel (time of both; algorithmic bytes = one read + one write, the 5-pass radix select re-reads the source)"}.get(w["kind"], "affine_brick_kernel" if w.get("generic") else "affine_zsep_kernel"),
                     "algorithmic_bytes_per_launch": int(bytes_unit),
                     "launch_ms": round(launch_ms, 4), "peak_source": peak_src},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(w, budget_s=20.0)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return result


# --------------------------------------------------------------------------------------------
def cpu_sample(w):
    """A bounded sample of the workload for the CPU arm: (callable, output voxels, description)."""
    from oracle import affine_oracle as ao
    from oracle import deskew_oracle as do

    Z, Y, X = w["shape"]
    rng = np.random.default_rng(1000)
    if w["kind"] == "deskew":
        xs = min(X, 256)  # output rows (input X columns) are independent: an X-slab is a fair sample
        raw = rng.integers(0, 65536, size=(Z, Y, xs), dtype=np.uint16)
        import torch

        threads = len(os.sched_getaffinity(0))
        torch.set_num_threads(threads)
        out_shape, _ = do.deskewed_shape_oracle(raw.shape, w["ls_angle_deg"], w["px_to_scan_ratio"],
                                                w["keep_overhang"], w["average_n_slices"])

        def fn():
            return do.deskew_oracle_torch(raw, w["ls_angle_deg"], w["px_to_scan_ratio"],
                                          w["keep_overhang"], w["average_n_slices"])

        desc = (f"1 unit restricted to X={xs} of {X} coverslip columns (uint16 {Z}x{Y}x{xs}); CPU torch "
                f"port of reference fast_deskew_zyx stages (biahub/deskew.py:505-536), {threads} threads")
        return fn, int(np.prod(out_shape)), threads, desc
    if w["kind"] == "flatfield":
        from oracle import flatfield_oracle as fo

        xs = min(X, 128)
        raw = rng.integers(90, 1200, size=(Z, Y, xs), dtype=np.uint16)

        def fn():
            return fo.flat_field_czyx_oracle(raw[None], [0])

        desc = (f"1 unit restricted to X={xs} of {X} columns (uint16 {Z}x{Y}x{xs}); numpy port of "
                f"reference _flat_field_czyx (biahub/flat_field.py:105-166), 1 thread")
        return fn, Z * Y * xs, 1, desc
    # affine: scipy is single-threaded C; the reference fans (t, c) units out over a process pool
    # (iohub process_single_position, num_workers), so the CPU arm runs one slab of output planes
    # per host thread in a fork-ed pool and counts all of them
    import multiprocessing as mp

    zs = min(Z, 4)
    threads = max(1, len(os.sched_getaffinity(0)))
    vol = (rng.random((zs + 2, Y, X), dtype=np.float32) * 4095).astype(np.float32)
    M = ao.register_matrix_c3((Z, Y, X)) if w["kind"] == "register" else stabilize_matrices(2)[1]
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
            f"host thread in a process pool; scipy.ndimage.affine_transform order=1 (library of "
            f"the reference's method='scipy' branch, biahub/register.py:272; the reference's default "
            f"ANTs branch is not installable)")
    return fn, threads * zs * Y * X, threads, desc


_CPU_AFFINE_JOB = None


def _cpu_affine_init(job):
    global _CPU_AFFINE_JOB
    _CPU_AFFINE_JOB = job


def _cpu_affine_slab(_i):
    from oracle import affine_oracle as ao

    vol, M, shape = _CPU_AFFINE_JOB
    return float(ao.affine_oracle_scipy(vol, M, shape, 1)[0, 0, 0])


def cpu_baseline(w, budget_s=20.0):
    fn, vox, threads, desc = cpu_sample(w)
    fn()  # warm
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s / 2 and len(times) < 10):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    best = min(times)
    getattr(fn, "close", lambda: None)()
    return {"value": round(vox / best / 1e9, 5), "unit": "Gvoxels/s", "cores": threads,
            "kind": "port", "sample": desc + f"; best of {len(times)}"}


def run_reference(args, w, rank, world):
    """Reference arm: the CPU path (oracle port), all host threads it can use, bounded sample/step."""
    if rank != 0:
        return None
    fn, vox, threads, desc = cpu_sample(w)
    for _ in range(max(1, min(args.warmup, 2))):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    getattr(fn, "close", lambda: None)()
    value = args.steps * vox / dt / 1e9
    return {
        "impl": "reference",
        "metric": "output Gvoxels/s (deskew/register/stabilize resampling); % of HBM BW",
        "value": round(value, 5), "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (uniform noise, seeded)",
        "config": {"workload": w["desc"]},
        "cpu_baseline": {"value": round(value, 5), "unit": "Gvoxels/s", "cores": threads,
                         "kind": "port", "sample": desc},
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
    ap.add_argument("--units", type=int, default=0, help="override units per step per GPU")
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

    if args.impl == "reference":
        res = run_reference(args, w, rank, world)
    else:
        res = run_b200(args, w, rank, world, local_rank)
    if rank == 0 and res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
