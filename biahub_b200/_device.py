"""GPU selection for worker processes + array/tensor classification helpers.

The reference pins nothing: every spawned (t, c) worker targets ``cuda:0``
(reference biahub/virtual_stain.py:323-335) and ``DeskewSettings.device`` defaults to ``"cpu"``
(reference biahub/settings.py:356).  This package always computes on a B200, so a ``device``
argument only selects WHICH GPU:

* ``"cuda:K"``  → GPU K;  ``"cuda"`` → the per-process default below
* ``"cpu"`` / ``None`` → the per-process default below (there is no CPU path to fall back to;
  if no GPU is present the call raises)
* env ``BIAHUB_B200_DEVICE`` (``"K"`` or ``"cuda:K"``) overrides everything.

Per-process default: ``LOCAL_RANK`` if set (torchrun / one process per GPU), else
``SLURM_LOCALID``, else ``BIAHUB_B200_WORKER_INDEX``, else the index of this multiprocessing pool
worker (``SpawnPoolWorker-k`` → k-1: the spawn-ed workers of ``process_single_position`` are
numbered consecutively, so they spread evenly over the GPUs of a node where ``pid %`` could pile
several onto one), else ``os.getpid()`` — always modulo the device count.

Pinned host memory is a per-BOX budget: ``BIAHUB_B200_PINNED_POOL_MB`` (default 6144) is divided
by ``BIAHUB_B200_WORKERS`` (the number of sibling worker processes on the box, default 1; set it
to the ``num_workers`` the CLI runs with — the reference defaults to 16 per position,
biahub/deskew.py:693-695) to give each process's result-pool cap.
"""

from __future__ import annotations

import os

import numpy as np

from . import _cabi

_default_device = None


def default_device() -> int:
    global _default_device
    if _default_device is None:
        n = _cabi.device_count()
        if n <= 0:
            _cabi.require_device(0)  # raises with the library's message
        for var in ("LOCAL_RANK", "SLURM_LOCALID", "BIAHUB_B200_WORKER_INDEX"):
            v = os.environ.get(var)
            if v is not None and v.isdigit():
                _default_device = int(v) % n
                break
        else:
            _default_device = worker_index() % n
    return _default_device


def worker_index() -> int:
    """Index of this process among its siblings: the multiprocessing pool-worker number when
    there is one (``_identity`` of ``SpawnPoolWorker-k`` is ``(k,)``), else the pid."""
    try:
        import multiprocessing

        ident = multiprocessing.current_process()._identity
        if ident:
            return int(ident[0]) - 1
    except Exception:  # noqa: BLE001 - private attribute: fall back to the pid
        pass
    return os.getpid()


def pinned_pool_cap_bytes() -> int:
    """This process's share of the box-wide pinned result-pool budget."""
    total_mb = int(os.environ.get("BIAHUB_B200_PINNED_POOL_MB", "6144"))
    workers = os.environ.get("BIAHUB_B200_WORKERS", "1")
    workers = int(workers) if workers.isdigit() and int(workers) > 0 else 1
    return (total_mb << 20) // workers


def resolve_device(device=None) -> int:
    env = os.environ.get("BIAHUB_B200_DEVICE")
    if env:
        device = env
    if device is None:
        return default_device()
    if isinstance(device, int):
        return device
    name = str(device)
    if name.isdigit():
        return int(name)
    if name.startswith("cuda"):
        if ":" in name:
            return int(name.split(":", 1)[1])
        return default_device()
    if name == "cpu":
        return default_device()
    raise ValueError(f"unknown device {device!r}")


def is_torch_tensor(obj) -> bool:
    mod = type(obj).__module__
    return mod == "torch" or mod.startswith("torch.")


def host_source(arr):
    """Return (C-contiguous numpy array, B2 dtype code) ready for the b2h_* calls.

    uint16 and float32 go through unchanged (uint16 is shipped over PCIe as uint16); every other
    dtype gets the reference's ``astype(float32)`` (biahub/deskew.py:578, biahub/register.py:266).
    """
    arr = np.asarray(arr)
    if arr.dtype == np.uint16:
        code = _cabi.DTYPE_U16
    else:
        if arr.dtype != np.float32:
            with np.errstate(over="ignore"):
                arr = arr.astype(np.float32)
        code = _cabi.DTYPE_F32
    if not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr)
    if not arr.dtype.isnative:
        arr = arr.astype(arr.dtype.newbyteorder("="))
    return arr, code


def row_pitch(tensor):
    """Row pitch (elements) of a (Z, Y, X) tensor whose rows are dense and whose planes are
    Y*pitch apart (what ``padded_empty`` produces), else None."""
    if tensor.ndim != 3:
        return None
    sz, sy, sx = tensor.stride()
    Z, Y, X = tensor.shape
    if sx == 1 and sy >= X and (sz == sy * Y or Z == 1):
        return int(sy)
    return None


def padded_empty(shape, device, row_align=4):
    """float32 CUDA tensor view of ``shape`` whose row pitch is rounded up to ``row_align``
    elements (16-byte aligned rows → TMA-eligible source for a chained affine warp)."""
    import torch

    Z, Y, X = (int(v) for v in shape)
    pitch = -(-X // row_align) * row_align
    buf = torch.empty((Z, Y, pitch), dtype=torch.float32, device=device)
    return buf[:, :, :X]


def device_source(tensor, allow_pitched=False):
    """Return (CUDA tensor, B2 dtype code) for the b2_* calls: contiguous, or — when
    ``allow_pitched`` — a dense-row view with a row pitch (see ``row_pitch``)."""
    import torch

    if not tensor.is_cuda:
        raise RuntimeError(
            "biahub_b200 computes on the GPU only: pass a CUDA tensor (or a numpy array, which is "
            "staged through the pinned host pipeline); there is no CPU fallback"
        )
    if tensor.dtype == torch.uint16:
        code = _cabi.DTYPE_U16
    else:
        if tensor.dtype != torch.float32:
            tensor = tensor.to(torch.float32)
        code = _cabi.DTYPE_F32
    if allow_pitched and row_pitch(tensor) is not None:
        return tensor, code
    return tensor.contiguous(), code


def pinned_empty(shape, dtype):
    """Page-locked host array (numpy view of a pinned torch byte tensor): the b2h_* pipeline
    copies straight from/to such buffers, skipping its own staging copy."""
    import torch

    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    buf = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True)
    return buf.numpy()[:n].view(dtype).reshape(shape)


class _PinnedHolder:
    """Owner object of one result array: numpy keeps it as ``.base`` of the array and of every
    view of it, so the block returns to the pool when the last of them is garbage-collected."""

    __slots__ = ("block", "__array_interface__", "__weakref__")

    def __init__(self, block, shape, dtype):
        self.block = block
        self.__array_interface__ = {
            "shape": tuple(int(v) for v in shape),
            "typestr": np.dtype(dtype).str,
            "data": (block["ptr"], False),
            "version": 3,
        }


class PinnedResultPool:
    """Recycling pool of page-locked blocks for RESULT arrays.

    The reference's callers (``process_single_position``) take a fresh array per (t, c) unit, write
    it to zarr and drop it.  A fresh pageable ``np.empty`` costs a staging copy out of the pinned
    ring plus the first-touch page faults of 1.5 GB per mantis volume — 2-3x the PCIe time.  Blocks
    from this pool are the DMA target themselves; a block is handed out again once the array (and
    all its views) is gone.  Bounded by this process's share of ``BIAHUB_B200_PINNED_POOL_MB``
    (box-wide, default 6144; 0 disables; divided by ``BIAHUB_B200_WORKERS``): beyond that, results
    fall back to ordinary pageable arrays.
    """

    def __init__(self, alloc=None, cap_bytes=None):
        self._alloc = alloc or self._alloc_pinned
        if cap_bytes is None:
            cap_bytes = pinned_pool_cap_bytes()
        self.cap_bytes = cap_bytes
        self.free = []          # blocks: {"ptr": int, "nbytes": int, "keep": object}
        self.total_bytes = 0
        self.handed_out = 0
        import threading

        self._lock = threading.Lock()

    @staticmethod
    def _alloc_pinned(nbytes):
        import torch

        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return {"ptr": t.data_ptr(), "nbytes": nbytes, "keep": t}

    def _release(self, block):
        with self._lock:
            self.handed_out -= 1
            self.free.append(block)

    def empty(self, shape, dtype=np.float32):
        """A writeable C-contiguous array of ``shape`` in pinned memory, or None when the pool is
        disabled / exhausted."""
        import weakref

        dtype = np.dtype(dtype)
        need = max(int(np.prod(shape)) * dtype.itemsize, 1)
        block = None
        with self._lock:
            fits = [b for b in self.free if b["nbytes"] >= need]
            if fits:
                block = min(fits, key=lambda b: b["nbytes"])
                self.free.remove(block)
            elif self.total_bytes + need > self.cap_bytes:
                # make room by dropping idle blocks that are too small before giving up
                while self.free and self.total_bytes + need > self.cap_bytes:
                    self.total_bytes -= self.free.pop(0)["nbytes"]
                if self.total_bytes + need > self.cap_bytes:
                    return None
        if block is None:
            try:
                block = self._alloc(need)
            except Exception:  # noqa: BLE001 - no driver / pinned memory exhausted: pageable result
                return None
            with self._lock:
                self.total_bytes += block["nbytes"]
        holder = _PinnedHolder(block, shape, dtype)
        arr = np.asarray(holder)
        with self._lock:
            self.handed_out += 1
        weakref.finalize(holder, self._release, block)
        return arr


_result_pool = None


def result_pool() -> PinnedResultPool:
    global _result_pool
    if _result_pool is None:
        _result_pool = PinnedResultPool()
    return _result_pool


def check_out(out, shape, dtype=np.float32):
    """The result buffer: the caller's ``out`` (validated) or a fresh array — from the pinned
    result pool when possible, else an ordinary ``np.empty``."""
    dtype = np.dtype(dtype)
    if out is None:
        arr = None
        if int(np.prod(shape)) >= (1 << 20):  # small results are not worth a pinned block
            arr = result_pool().empty(shape, dtype)
        return arr if arr is not None else np.empty(shape, dtype=dtype)
    if not (isinstance(out, np.ndarray) and out.dtype == dtype and out.shape == tuple(shape)
            and out.flags.c_contiguous):
        raise ValueError(f"out must be a C-contiguous {dtype} array of shape {tuple(shape)}")
    return out


def bind_to_gpu_numa(device: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned
    staging buffers allocated afterwards (first touch) and the host copy threads are local to the
    GPU's PCIe root.  On multi-socket hosts remote pinned memory is what limits 8-GPU end-to-end
    throughput.  Best effort: returns what was done, never raises."""
    info = {"device": int(device), "numa_node": None, "cpus": None}
    try:
        import torch

        bdf = torch.cuda.get_device_properties(int(device)).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(int(device)), "pci_bus_id") else None
        if bdf is None:
            import ctypes

            buf = ctypes.create_string_buffer(32)
            rt = ctypes.CDLL("libcudart.so.12")
            if rt.cudaDeviceGetPCIBusId(buf, 32, int(device)) != 0:
                return info
            bdf = buf.value.decode()
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:      # 00000000:17:00.0 -> 0000:17:00.0
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(numa_node=node, cpus=len(allowed))
    except Exception:  # noqa: BLE001 - best effort only
        pass
    return info
