"""CPU tests of the N>1 path: unit partitioning and the per-unit runner, incl. a world-size-2
gloo run (the data path has no collective; torch.distributed only gathers what each rank did)."""
import os
import socket

import numpy as np
import pytest

from biahub_b200 import sharding


def test_partition_is_disjoint_and_complete():
    units = sharding.enumerate_units(8, range(32), [0])
    assert len(units) == 256 and units[0] == (0, 0, 0) and units[-1] == (7, 31, 0)
    for world in (1, 2, 4, 8, 3):
        shards = [sharding.units_for_rank(units, r, world) for r in range(world)]
        flat = sorted(u for s in shards for u in s)
        assert flat == sorted(units)
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
    with pytest.raises(ValueError):
        sharding.units_for_rank(units, 2, 2)


def test_run_units_mirrors_process_single_position():
    seen = []

    def stabilize_like(czyx, list_of_shifts, input_time_index, output_shape=None):
        seen.append(input_time_index)
        return czyx + list_of_shifts[input_time_index]

    def deskew_like(czyx, **kw):
        assert "extra_metadata" not in kw and "input_time_index" not in kw
        return czyx * kw["scale"]

    store = {}
    units = sharding.enumerate_units(1, [0, 2], [1])
    n = sharding.run_units(stabilize_like, lambda p, t, c: np.full((1, 2, 2, 2), t, np.float32),
                           lambda p, t, c, out: store.__setitem__((p, t, c), out), units,
                           list_of_shifts=[10, 20, 30], extra_metadata={"x": 1})
    assert n == 2 and seen == [0, 2]
    assert store[(0, 2, 1)].max() == 32
    sharding.run_units(deskew_like, lambda p, t, c: np.ones((1, 1, 1, 1)),
                       lambda p, t, c, out: store.__setitem__(("d", t), out), units, scale=3.0,
                       extra_metadata={})
    assert store[("d", 0)].item() == 3.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, out_q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.rank_and_world() == (rank, world)
        units = sharding.enumerate_units(2, range(n_units), [0, 1])
        mine = sharding.units_for_rank(units, rank, world)
        results = {}
        sharding.run_units(lambda czyx: czyx.sum(axis=(1, 2, 3)),
                           lambda p, t, c: np.full((1, 2, 2, 2), p * 100 + t * 10 + c, np.float64),
                           lambda p, t, c, out: results.__setitem__((p, t, c), float(out[0])), mine)
        gathered = [None] * world
        dist.all_gather_object(gathered, results)  # bookkeeping only: not a data-path collective
        if rank == 0:
            out_q.put(gathered)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for part in gathered:
        assert not (set(part) & set(merged))      # disjoint shards
        merged.update(part)
    assert len(merged) == 2 * 3 * 2               # every unit exactly once
    assert len(gathered[0]) == len(gathered[1]) == 6
    for (p, t, c), v in merged.items():
        assert v == 8 * (p * 100 + t * 10 + c)


def test_run_units_overlapped_pipelines_and_orders():
    import threading
    import time

    units = sharding.enumerate_units(1, range(6), [0])
    log, written = [], []
    lock = threading.Lock()

    def read(p, t, c):
        time.sleep(0.05)
        with lock:
            log.append(("r", t, time.perf_counter()))
        return np.full((1, 1, 1, 2), t, np.float32)

    def compute(czyx, input_time_index, gain):
        time.sleep(0.05)
        assert czyx[0, 0, 0, 0] == input_time_index
        return czyx * gain

    def write(p, t, c, out):
        time.sleep(0.05)
        written.append((t, float(out[0, 0, 0, 0])))

    t0 = time.perf_counter()
    n = sharding.run_units_overlapped(compute, read, write, units, prefetch=2, gain=2.0,
                                      extra_metadata={"ignored": True})
    dt = time.perf_counter() - t0
    assert n == 6
    assert written == [(t, 2.0 * t) for t in range(6)]        # in order, every unit once
    assert dt < 0.75 * (6 * 0.15)                              # stages overlapped (serial = 0.9 s)


def test_run_units_overlapped_propagates_errors():
    units = sharding.enumerate_units(1, range(5), [0])

    def bad_compute(czyx):
        if czyx[0, 0, 0, 0] == 2:
            raise RuntimeError("kernel failed")
        return czyx

    with pytest.raises(RuntimeError, match="kernel failed"):
        sharding.run_units_overlapped(bad_compute, lambda p, t, c: np.full((1, 1, 1, 1), t),
                                      lambda p, t, c, o: None, units)

    def bad_read(p, t, c):
        if t == 3:
            raise OSError("chunk missing")
        return np.zeros((1, 1, 1, 1))

    with pytest.raises(OSError, match="chunk missing"):
        sharding.run_units_overlapped(lambda x: x, bad_read, lambda p, t, c, o: None, units)

    def bad_write(p, t, c, o):
        raise IOError("disk full")

    with pytest.raises(IOError, match="disk full"):
        sharding.run_units_overlapped(lambda x: x, lambda p, t, c: np.zeros((1, 1, 1, 1)),
                                      bad_write, units)
