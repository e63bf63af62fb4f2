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


def merge_candidates(cands, n_samples: int, w: int, fmin: float, fstep: float, thr: float = 12.0):
    """Host statement of what K4 (finalize_kernel) does with one PRN's row candidates, whichever GPU produced
    them (bin-split shards): ``cands[b] = (peak, lag, sum_all, sum_win)`` for every Doppler bin b of
    the FULL grid.  Winner = the maximum peak; among equal peaks the LOWEST bin (acquisition.m:62) and the lowest
    code phase (:63); the noise floor comes from the winning bin's own tuple with the reference's clipped index set
    (:66-68).  Returns (code_phase, doppler_bin, doppler_hz, peak, noise_meansq, snr_db, acquired).  Used by the
    CPU tests of the multi-GPU merge; the product runs the same rule on the root GPU."""
    import math
    g = max(c[0] for c in cands)
    tied = [b for b, c in enumerate(cands) if c[0] == g]
    fbin = tied[0]
    cp = min(cands[b][1] for b in tied)
    cp1 = cp + 1
    cnt = max(cp1 - w, 0) + max(n_samples - cp1 - w + 1, 0)
    noise = (cands[fbin][2] - cands[fbin][3]) / cnt if cnt else float("nan")
    snr = 10.0 * math.log10(g * g / noise) if noise > 0 else float("nan")
    return cp, fbin, fmin + fstep * fbin, g, noise, snr, bool(snr >= thr)


def root_extra_for_wait(extra_permille: int, wait_ms: float, search_ms: float, world: int) -> int:
    """The root's weight (gnssacq_shard_plan_rows' root_extra_permille) that levels the finish times, given that with
    the weight `extra_permille` the root waited `wait_ms` for the other shards after `search_ms` of its own search.
    Moving x rows to the root costs it x row times and saves every other shard x/(world-1): equal finish at
    x = wait/T_row * (world-1)/world, i.e. the root's share grows by the factor 1 + (wait/search)(world-1)/world.
    (The step gets shorter by wait/world, not by wait: only ONE GPU was idle.)"""
    if world <= 1 or search_ms <= 0:
        return int(extra_permille)
    e0 = extra_permille / 1000.0
    share = (1 + e0) / (world + e0) * (1 + max(wait_ms, 0.0) / search_ms * (world - 1) / world)
    share = min(share, 0.9)
    return int(round(1000 * (share * world - 1) / (1 - share)))


def rows_from_bytes(buf: bytes) -> List[api.Result]:
    n = len(buf) // ROW_BYTES
    return list((api.Result * n).from_buffer_copy(buf[: n * ROW_BYTES]))


class PeerShard:
    """One rank's share of ONE acquisition, exchanged through peer memory instead of NCCL (gnssacq_xchg_*):
    rank 0 holds the IF buffer and the candidate table of the full grid; the other ranks' K1a pulls the IF
    block out of rank 0's HBM over NVLink, their K2 stores its row candidates straight into rank 0's table,
    and rank 0 runs K4 over the full table -- so the rows are byte-identical for every world size, also when
    there are fewer PRNs than GPUs (then the Doppler bins are split).  `torch.distributed` is used once, to
    hand rank 0's CUDA IPC handle to the other processes.  No CPU fallback."""

    def __init__(self, cfg_full: api.Config, rank: int, world: int, device: int, dist=None,
                 plan: str = "prn", root_extra_permille: int = 0):
        """plan "prn": gnssacq_shard_plan (whole PRNs per shard, or bin ranges when there are fewer PRNs than shards);
        plan "rows": gnssacq_shard_plan_rows (contiguous row ranges, the root's share weighted by
        `root_extra_permille`; see rebalance())."""
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self._dist = dist
        self._device = device
        torch.cuda.set_device(device)
        self._full = api.Config.from_buffer_copy(bytes(cfg_full))
        self._full.device = device
        self.plan = plan
        self.root_extra_permille = int(root_extra_permille)
        self.searcher = None
        self._build()

    def _build(self) -> None:
        full, rank, world, dist = self._full, self.rank, self.world, self._dist
        if self.plan == "rows":
            mine, self.shard = api.shard_plan_rows(full, rank, world, self.root_extra_permille)
        elif self.plan == "prn":
            mine, self.shard = api.shard_plan(full, rank, world)
        else:
            raise ValueError(f"unknown shard plan {self.plan!r}")
        self.n_prn_total = self.shard.n_prn_total
        self.rows_local = self.shard.n_rows
        self.searcher = api.Searcher(mine) if self.rows_local > 0 else None
        self.if_bytes = int(api.lib.gnssacq_if_bytes(C.byref(full)))
        ipc = [None]
        if rank == 0:
            if self.searcher is None:
                raise ValueError("rank 0 must own rows")
            ipc[0] = self.searcher.xchg_root(self.shard)
        if world > 1:
            dist.broadcast_object_list(ipc, src=0)
        if rank != 0 and self.searcher:
            self.searcher.xchg_attach(self.shard, ipc[0])
        self.d_if_ptr = self.searcher.xchg_if_buffer() if rank == 0 else 0

    def rebalance(self, steps: int = 12, host_if=None) -> int:
        """Level the finish times of the shards (plan "rows" only).  The other shards start every step later than
        the root by the time the IF block takes to reach them, so with equal shares the root idles at the end of
        each step (gnssacq_stats.gather_wait_ms).  Runs `steps` steps, reads on the root how long it waited relative
        to its own search, moves that share of the rows from the others to the root, and rebuilds the handles of
        every rank with the new plan.  Set-up work (one more create per rank): call it once, outside any timed
        region, with the IF block already in the exchange buffer (upload()) or passed as `host_if`.  Collective:
        every rank must call it.  Returns the new root_extra_permille.  Results do not depend on the plan."""
        if self.plan != "rows" or self.world == 1:
            return self.root_extra_permille
        torch, dist = self.torch, self._dist
        waits, searches = [], []
        for _ in range(steps):
            torch.cuda.synchronize()
            dist.barrier()
            self.enqueue(host_if)
            self.fetch()
            if self.rank == 0:
                st = self.searcher.last_stats
                waits.append(st.gather_wait_ms)
                searches.append(st.search_ms)
        msg = [None]
        if self.rank == 0:
            # root idles `wait` per `search` of its own rows: give it that many more (its rows' time grows by
            # wait * (world-1)/world, everybody else's shrinks by wait/world -> equal finish).  Medians: one slow
            # step (a late launch on some rank) must not skew the plan.
            wait = sorted(waits)[len(waits) // 2]
            search = sorted(searches)[len(searches) // 2]
            wait = max(wait - 0.003, 0.0)                      # the wait kernel's own latency when nobody is late
            msg[0] = (root_extra_for_wait(self.root_extra_permille, wait, search, self.world),)
        dist.broadcast_object_list(msg, src=0)
        new_extra = msg[0][0]
        if new_extra != self.root_extra_permille:
            saved = self._saved_if()
            self.close(_final=False)
            self.root_extra_permille = new_extra
            self._build()
            if self.rank == 0 and saved is not None:
                self._copy_into_if(saved)
                torch.cuda.synchronize()
        return self.root_extra_permille

    def _saved_if(self):
        """rank 0: a device copy of the exchange buffer's IF block (it moves with the handle)."""
        if self.rank != 0 or not self.d_if_ptr:
            return None
        arr = _DevArray(self.d_if_ptr, self.if_bytes)
        return self.torch.as_tensor(arr, device="cuda").clone()

    def bind_stream(self) -> None:
        if self.searcher:
            self.searcher.set_stream(self.torch.cuda.current_stream().cuda_stream)

    def upload(self, h_if) -> None:
        """rank 0: put an IF block into the exchange buffer (for device-resident timing loops)."""
        if self.rank == 0:
            t = self.torch.frombuffer(bytearray(h_if), dtype=self.torch.uint8) if not hasattr(h_if, "data_ptr") else h_if
            self._copy_into_if(t)

    def _copy_into_if(self, t) -> None:
        torch = self.torch
        # wrap the exchange block's IF buffer as a tensor (no ownership) and copy with torch's own stream ordering
        arr = _DevArray(self.d_if_ptr, self.if_bytes)
        dst = torch.as_tensor(arr, device="cuda")
        dst.copy_(t[: self.if_bytes], non_blocking=True)

    def enqueue(self, host_if=None) -> None:
        """One step on torch's current stream.  rank 0: `host_if` (bytes-like, pageable is fine: the library
        stages it) or None when the block is already in the exchange buffer."""
        self.bind_stream()
        if self.searcher is None:
            return
        self.searcher.xchg_enqueue(host_if if self.rank == 0 else None)
        if self.rank == 0:
            self.searcher.xchg_finish()

    def fetch(self) -> List[api.Result]:
        """rank 0: host sync + all rows (original PRN order); other ranks: host sync, []."""
        if self.searcher is None:
            return []
        return self.searcher.xchg_fetch(rows=(self.rank == 0))

    def close(self, _final: bool = True) -> None:
        """The other ranks unmap the root's exchange block BEFORE the root frees it."""
        if self.rank != 0 and self.searcher:
            self.searcher.close()
            self.searcher = None
        if self.world > 1 and self._dist is not None:
            self.torch.cuda.synchronize()
            self._dist.barrier()
        if self.searcher:
            self.searcher.close()
            self.searcher = None


class _DevArray:
    """Minimal __cuda_array_interface__ view of raw device memory (uint8)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class LocalMultiGpu:
    """The same exchange inside ONE process: shard r on devices[r] (what a MATLAB host can drive without
    torchrun).  If several shards share a device (tests on a one-GPU box), the steps are serialised with host
    syncs, because kernels of different shards that wait for one another must not share a GPU."""

    def __init__(self, cfg_full: api.Config, devices: Sequence[int], plan: str = "prn", root_extra_permille: int = 0):
        self.world = len(devices)
        self.serial = len(set(devices)) < len(devices)
        self.searchers: List[api.Searcher] = []
        self.shards = []
        for r, dev in enumerate(devices):
            full = api.Config.from_buffer_copy(bytes(cfg_full))
            full.device = dev
            if plan == "rows":
                mine, sh = api.shard_plan_rows(full, r, self.world, root_extra_permille)
            else:
                mine, sh = api.shard_plan(full, r, self.world)
            self.shards.append(sh)
            self.searchers.append(api.Searcher(mine) if sh.n_rows > 0 else None)
        root = self.searchers[0]
        root.xchg_root(self.shards[0])
        for s, sh in zip(self.searchers[1:], self.shards[1:]):
            if s:
                s.xchg_attach_local(sh, root)

    def search(self, if_bytes) -> List[api.Result]:
        root = self.searchers[0]
        root.xchg_enqueue(if_bytes)
        if self.serial:
            root.xchg_fetch(rows=False)
        for s in self.searchers[1:]:
            if s:
                s.xchg_enqueue(None)
                if self.serial:
                    s.xchg_fetch(rows=False)
        root.xchg_finish()
        return root.xchg_fetch()

    def close(self) -> None:
        for s in self.searchers[1:] + self.searchers[:1]:
            if s:
                s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


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
