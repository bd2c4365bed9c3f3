"""Data-parallel plumbing for the one place the path shards: the batch dimension.

Every image is independent in inference (BatchNorm uses running statistics; `axis_name="batch"` has no
numeric effect), so N GPUs = N replicas of the weights, a contiguous split of the batch, and - only
when the caller wants the gathered output - ONE all-gather of the logits.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the launcher and rendezvous; no collective sits
inside the forward pass. On GPUs the gather itself is this library's own kernel over NVLink peer memory
(`LogitsAllGather` -> `eqxv_allgather_push`, csrc/p2p.cu); `all_gather_rows` uses it whenever the local block is
a CUDA tensor and falls back to `torch.distributed.all_gather` only for CPU tensors (the gloo tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced split of range(n): the first n % world ranks get one extra item"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


def shard(images, rank: Optional[int] = None, world: Optional[int] = None):
    """this rank's slice of a [B, ...] batch"""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist else 0
    if world is None:
        world = dist.get_world_size() if dist else 1
    lo, hi = shard_bounds(images.shape[0], rank, world)
    return images[lo:hi]


class LogitsAllGather:
    """All-gather of [rows, cols] blocks (one per rank, same shape) into [world * rows, cols] on every rank, through
    peer stores over NVLink (include/eqxv_b200.h, C1). Setup (once): every rank allocates a window, the CUDA-IPC
    handles are exchanged through torch.distributed and mapped. A call enqueues ONE kernel plus a device-to-device
    copy of the gathered rows into a fresh tensor on torch's current stream; nothing synchronises the host."""

    def __init__(self, rows: int, cols: int, dtype: torch.dtype = torch.float32, group=None):
        import ctypes as C

        from . import _lib

        dist = _dist()
        if dist is None:
            raise RuntimeError("LogitsAllGather needs an initialised torch.distributed process group")
        self._lib, self._C = _lib, C
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rows, self.cols, self.dtype = rows, cols, dtype
        es = torch.empty(0, dtype=dtype).element_size()
        self.slice_bytes = rows * cols * es
        if self.slice_bytes % 16 != 0:
            raise ValueError("LogitsAllGather: rows * cols * itemsize must be a multiple of 16 bytes")
        self.slot_bytes = self.slice_bytes
        self.buf_bytes = self.slot_bytes * self.world
        total = C.c_int64()
        _lib.call("eqxv_p2p_window_bytes", self.buf_bytes, C.byref(total))
        self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.init(self.device.index)
        own = C.c_void_p()
        _lib.call("eqxv_p2p_alloc", C.byref(own), total.value)
        self._own = own.value
        handle = C.create_string_buffer(64)
        _lib.call("eqxv_ipc_get_handle", self._own, handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._opened = []
        self._windows = (C.c_void_p * self.world)()
        for r, h in enumerate(handles):
            if r == self.rank:
                self._windows[r] = self._own
            else:
                p = C.c_void_p()
                _lib.call("eqxv_ipc_open_handle", C.create_string_buffer(h, 64), C.byref(p))
                self._opened.append(p.value)
                self._windows[r] = p.value
        self.launches = 0
        dist.barrier(group=group)   # every window is mapped everywhere before the first push

    def __call__(self, local: torch.Tensor) -> torch.Tensor:
        C, _lib = self._C, self._lib
        if not local.is_cuda or tuple(local.shape) != (self.rows, self.cols) or local.dtype != self.dtype:
            raise ValueError(f"LogitsAllGather: expected a CUDA [{self.rows}, {self.cols}] {self.dtype} block")
        local = local.contiguous()
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("eqxv_allgather_push", local.data_ptr(), self.slice_bytes, self._windows, self.rank, self.world,
                  self.slot_bytes, self.buf_bytes, st)
        self.launches += 1
        off = C.c_int64()
        _lib.call("eqxv_p2p_buffer_offset", self.launches & 1, self.buf_bytes, C.byref(off))
        out = torch.empty((self.world * self.rows, self.cols), dtype=self.dtype, device=self.device)
        _lib.call("eqxv_memcpy_async", out.data_ptr(), self._own + off.value, self.buf_bytes, st)
        return out

    def close(self):
        if self._own is None:
            return
        torch.cuda.synchronize()
        dist = _dist()
        if dist is not None:
            dist.barrier()          # nobody unmaps a window a peer may still push into
        for p in self._opened:
            self._lib.call("eqxv_ipc_close_handle", p)
        self._lib.call("eqxv_p2p_free", self._own)
        self._own, self._opened = None, []


_gatherers: dict = {}


def all_gather_rows(local: torch.Tensor, total_rows: int) -> torch.Tensor:
    """concatenate every rank's [rows_r, ...] block in rank order (uneven blocks are padded for the
    collective and trimmed afterwards). Single process: returns `local`."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total_rows, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    if local.is_cuda and local.dim() == 2:
        # this library's own collective: peer stores over NVLink (csrc/p2p.cu), one cached communicator per shape
        key = (max_rows, local.shape[1], local.dtype, torch.cuda.current_device())
        if key not in _gatherers:
            cols = local.shape[1]
            while (max_rows * cols * local.element_size()) % 16:
                cols += 1
            _gatherers[key] = (LogitsAllGather(max_rows, cols, local.dtype), cols)
        ag, cols = _gatherers[key]
        pad = torch.zeros((max_rows, cols), dtype=local.dtype, device=local.device)
        pad[: local.shape[0], : local.shape[1]] = local
        full = ag(pad).reshape(world, max_rows, cols)
        return torch.cat([full[r, : hi - lo, : local.shape[1]] for r, (lo, hi) in enumerate(sizes)], 0)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def data_parallel_forward(forward: Callable[[torch.Tensor], torch.Tensor], images, gather: bool = True):
    """run `forward` (e.g. `vmap(net, axis_name="batch")` bound to its keys) on this rank's shard of
    `images`; with `gather` every rank returns the logits of the whole batch."""
    local = forward(shard(images))
    return all_gather_rows(local, images.shape[0]) if gather else local
