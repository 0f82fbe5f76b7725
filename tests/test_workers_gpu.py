"""Several worker PROCESSES against one box: the reference fans (t, c) units
out over a spawn-ed pool of up to 16 workers per position (biahub/deskew.py:693-695, iohub
process_single_position).  Each worker owns a CUDA context, device buffers, pinned rings and a
share of the box-wide pinned result-pool budget."""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(args):
    seed, n_units = args
    import numpy as np

    import biahub_b200 as b2
    from biahub_b200 import _device
    from biahub_b200.sharding import run_units
    from oracle import deskew_oracle as do

    rng = np.random.default_rng(seed)
    raws = {t: rng.integers(0, 65536, size=(1, 128, 24, 256), dtype=np.uint16) for t in range(n_units)}
    outs = {}
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    run_units(b2._fast_deskew_czyx, lambda p, t, c: raws[t], lambda p, t, c, out: outs.__setitem__(t, out),
              [(0, t, 0) for t in range(n_units)], **kw)
    worst = 0.0
    for t in range(n_units):
        want = do.deskew_oracle_numpy(raws[t][0], 30.0, 0.386, False, 3)
        worst = max(worst, float(np.abs(outs[t][0] - want).max()))
    pool = _device.result_pool()
    return dict(worst=worst, device=_device.default_device(), index=_device.worker_index(),
                cap=pool.cap_bytes, pinned=pool.total_bytes, pid=os.getpid())


def test_four_worker_processes_share_one_box(monkeypatch):
    import torch

    from biahub_b200 import _device

    monkeypatch.setenv("BIAHUB_B200_WORKERS", "4")
    monkeypatch.setenv("BIAHUB_B200_PINNED_POOL_MB", "2048")
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert _device.pinned_pool_cap_bytes() == (2048 << 20) // 4
    ctx = mp.get_context("spawn")
    with ctx.Pool(4) as pool:
        res = pool.map(_worker, [(100 + i, 3) for i in range(4)], chunksize=1)
    ngpu = torch.cuda.device_count()
    assert len({r["pid"] for r in res}) >= 2               # really several processes
    for r in res:
        assert r["worst"] <= 2e-7 * 65535, r
        assert r["cap"] == (2048 << 20) // 4                # each worker took its share of the budget
        assert r["pinned"] <= r["cap"]
        assert r["device"] == r["index"] % ngpu             # GPU picked by worker index, not pid
    assert sum(r["pinned"] for r in res) <= 2048 << 20      # box-wide pinned result memory is bounded
