"""PRN-major sharding of the search grid over the GPUs of one box (one process per GPU).

The (PRN, Doppler) rows are independent; the only shared input is the IF block and the only
cross-row step -- the per-PRN maximum over bins -- stays inside a GPU when whole PRNs are assigned
to ranks.  So the exchange is: broadcast the IF block from rank 0 (NCCL over
NVLink), search the local shard, all-gather the fixed-size result rows.  `torch.distributed` is the
plumbing; the numeric work is `libgnssacq.so`.

`backend` is injectable so the host logic (sharding, gather, merge) is testable with gloo on CPU:
the product backend is :class:`CudaShard`, which has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import api

ROW_BYTES = C.sizeof(api.Result)


def prn_shard(prns: Sequence[int], rank: int, world: int) -> List[int]:
    """Contiguous PRN-major split; ranks differ by at most one PRN."""
    prns = list(prns)
    lo = rank * len(prns) // world
    hi = (rank + 1) * len(prns) // world
    return prns[lo:hi]


def shard_sizes(n_prn: int, world: int) -> List[int]:
    return [(r + 1) * n_prn // world - r * n_prn // world for r in range(world)]


def rows_from_bytes(buf: bytes) -> List[api.Result]:
    n = len(buf) // ROW_BYTES
    return list((api.Result * n).from_buffer_copy(buf[: n * ROW_BYTES]))


class CudaShard:
    """One rank's share of the grid on its own GPU (torch tensors for HBM, NCCL for the exchange)."""

    def __init__(self, cfg_factory, prns: Sequence[int], rank: int, world: int, device: int):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.all_prns = list(prns)
        self.sizes = shard_sizes(len(self.all_prns), world)
        self.max_rows = max(self.sizes)
        mine = prn_shard(self.all_prns, rank, world)
        self.n_local = len(mine)
        self.searcher = api.Searcher(cfg_factory(mine, device)) if mine else None
        cfg0 = cfg_factory(self.all_prns[:1], device)
        self.if_bytes = int(api.lib.gnssacq_if_bytes(C.byref(cfg0)))
        self.d_if = torch.empty(self.if_bytes, dtype=torch.uint8, device=self.device)
        # equal-size slots so one all_gather_into_tensor moves every shard (NCCL needs equal counts)
        self.d_rows = torch.zeros(self.max_rows * ROW_BYTES, dtype=torch.uint8, device=self.device)
        self.d_all = torch.zeros(world * self.max_rows * ROW_BYTES, dtype=torch.uint8, device=self.device)
        self.h_all = torch.empty(world * self.max_rows * ROW_BYTES, dtype=torch.uint8).pin_memory()

    def bind_stream(self) -> None:
        """Kept for callers of r01; enqueue() now binds the handle to torch's current stream on every call."""
        if self.searcher:
            self.searcher.set_stream(self.torch.cuda.current_stream().cuda_stream)

    def enqueue(self, dist, h_if=None) -> None:
        """One acquisition, stream-ordered ON TORCH'S CURRENT STREAM: [H2D on rank 0] -> broadcast -> search shard
        -> all-gather.  The handle is (re)bound to that stream here, every call, so the search can never run
        unordered against the collectives and copies around it (whatever stream the caller has switched to)."""
        self.bind_stream()
        if h_if is not None and self.rank == 0:
            self.d_if.copy_(h_if, non_blocking=True)
        if self.world > 1:
            dist.broadcast(self.d_if, src=0)
        if self.searcher:
            self.searcher.enqueue_device_out(self.d_if.data_ptr(), self.if_bytes, self.d_rows.data_ptr())
        if self.world > 1:
            dist.all_gather_into_tensor(self.d_all, self.d_rows)
        else:
            self.d_all.copy_(self.d_rows, non_blocking=True)

    def fetch(self) -> List[api.Result]:
        """D2H of the assembled table + host sync; rows in the original PRN order."""
        self.h_all.copy_(self.d_all, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        return merge_table(self.h_all.numpy().tobytes(), self.sizes, self.max_rows)

    def close(self) -> None:
        if self.searcher:
            self.searcher.close()


def merge_table(gathered: bytes, sizes: Sequence[int], slot_rows: int) -> List[api.Result]:
    """Drop the padding of the equal-size all-gather slots; shards are already in PRN order."""
    rows: List[api.Result] = []
    for r, n in enumerate(sizes):
        off = r * slot_rows * ROW_BYTES
        rows += rows_from_bytes(gathered[off: off + n * ROW_BYTES])
    return rows


def gather_rows_host(dist, local_rows: Sequence[api.Result], sizes: Sequence[int]) -> List[api.Result]:
    """Backend-agnostic gather (works on gloo): used by the CPU tests of the host logic."""
    import torch
    slot = max(sizes)
    buf = bytearray(slot * ROW_BYTES)
    raw = b"".join(bytes(r) for r in local_rows)
    buf[: len(raw)] = raw
    mine = torch.frombuffer(buf, dtype=torch.uint8).clone()
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return merge_table(b"".join(t.numpy().tobytes() for t in out), sizes, slot)
