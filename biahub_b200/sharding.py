"""Whole-box sharding of (position, t, c) units + a local stand-in for the per-position runner.

The reference parallelises by independent units only: one Slurm/local job per position
(reference biahub/deskew.py:733-749) and, inside a job, ``iohub.ngff.utils.process_single_position``
loops a process pool over (t, c), calling ``func(czyx, **kwargs)`` per unit and writing the
result to the output store.  No unit ever needs another unit's data, so the multi-GPU form is:
one process per GPU, a static partition of the unit list, **no collective**.

``run_units`` mirrors the contract of ``process_single_position`` that the hot-path callables rely
on: the callable receives a (C, Z, Y, X) block and the kwargs; ``extra_metadata``
is popped; ``input_time_index`` is injected when the callee's signature names it
(reference biahub/stabilize.py:32-37 vs :288-300).
"""

from __future__ import annotations

import inspect
import os
import queue
import threading
from typing import Callable, Iterable, Sequence


def rank_and_world():
    """(rank, world_size) from the torchrun / Slurm environment (defaults 0, 1)."""
    for r, w in (("RANK", "WORLD_SIZE"), ("SLURM_PROCID", "SLURM_NTASKS")):
        if r in os.environ and w in os.environ:
            return int(os.environ[r]), int(os.environ[w])
    return 0, 1


def enumerate_units(n_positions: int, time_indices: Sequence[int], channel_indices: Sequence[int]):
    """Flat, deterministic unit list [(p, t, c), ...] — position-major like the reference's job
    fan-out, then time, then channel."""
    return [(p, int(t), int(c)) for p in range(int(n_positions)) for t in time_indices
            for c in channel_indices]


def units_for_rank(units: Sequence, rank: int, world_size: int):
    """Round-robin partition: rank r takes units r, r+W, r+2W, ....  The shards
    are disjoint, cover every unit exactly once and differ in size by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(units[rank::world_size])


def run_units(func: Callable, read_unit: Callable, write_unit: Callable, units: Iterable,
              **kwargs):
    """Per-unit loop of ``process_single_position``: ``out = func(read_unit(p,t,c), **kwargs)``
    then ``write_unit(p,t,c,out)``.  ``read_unit`` returns a (C, Z, Y, X) block."""
    kwargs = dict(kwargs)
    kwargs.pop("extra_metadata", None)
    wants_time = "input_time_index" in inspect.signature(func).parameters
    done = 0
    for (p, t, c) in units:
        block = read_unit(p, t, c)
        call_kwargs = dict(kwargs)
        if wants_time and "input_time_index" not in call_kwargs:
            call_kwargs["input_time_index"] = t
        write_unit(p, t, c, func(block, **call_kwargs))
        done += 1
    return done


def run_units_overlapped(func: Callable, read_unit: Callable, write_unit: Callable, units: Iterable,
                         prefetch: int = 2, **kwargs):
    """``run_units`` as a three-stage pipeline: a reader thread decodes the next ``prefetch``
    units (zarr chunk I/O) and a writer thread stores finished ones while the calling thread
    keeps the GPU busy.  The compute call spends its time inside libbiahub_b200 (ctypes releases
    the GIL; the library itself overlaps H2D / kernel / D2H on three CUDA streams), so host I/O,
    PCIe and kernels all run concurrently.  Units complete in order; the first exception of any
    stage is re-raised after the other stages have been stopped."""
    kwargs = dict(kwargs)
    kwargs.pop("extra_metadata", None)
    wants_time = "input_time_index" in inspect.signature(func).parameters
    units = list(units)
    q_in: "queue.Queue" = queue.Queue(maxsize=max(1, int(prefetch)))
    q_out: "queue.Queue" = queue.Queue(maxsize=max(1, int(prefetch)))
    errors = []
    stop = threading.Event()

    def reader():
        try:
            for u in units:
                if stop.is_set():
                    break
                q_in.put((u, read_unit(*u)))
        except BaseException as exc:  # noqa: BLE001 - re-raised on the caller's thread
            errors.append(exc)
        finally:
            q_in.put(None)

    def writer():
        try:
            while True:
                item = q_out.get()
                if item is None:
                    break
                (p, t, c), out = item
                write_unit(p, t, c, out)
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)
            stop.set()
            while q_out.get() is not None:  # drain so the producer never blocks
                pass

    tr = threading.Thread(target=reader, name="b2-read", daemon=True)
    tw = threading.Thread(target=writer, name="b2-write", daemon=True)
    tr.start()
    tw.start()
    done = 0
    try:
        while True:
            item = q_in.get()
            if item is None or stop.is_set():
                break
            (p, t, c), block = item
            call_kwargs = dict(kwargs)
            if wants_time and "input_time_index" not in call_kwargs:
                call_kwargs["input_time_index"] = t
            q_out.put(((p, t, c), func(block, **call_kwargs)))
            done += 1
    except BaseException as exc:  # noqa: BLE001
        errors.append(exc)
    finally:
        stop.set()
        q_out.put(None)
        while tr.is_alive():  # unblock a reader stuck on a full queue
            try:
                q_in.get_nowait()
            except queue.Empty:
                tr.join(timeout=0.05)
        tw.join()
    if errors:
        raise errors[0]
    return done
