"""Host logic of the multi-GPU path on CPU: world_size-2 gloo.  PRN-major sharding, equal-slot
all-gather and table merge are exercised with the ORACLE standing in for a rank's GPU (test double,
injected -- the product backend CudaShard has no CPU path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gnssacq
from gnssacq import api
from gnssacq.dist import prn_shard, shard_sizes, merge_table, gather_rows_host, merge_candidates, ROW_BYTES
from helpers import structs, small_spec, oracle_rows
from oracle.synth import synth_if


def test_shards_partition_the_prn_list():
    prns = list(range(1, 33))
    for world in (1, 2, 3, 4, 5, 8, 32, 40):
        parts = [prn_shard(prns, r, world) for r in range(world)]
        assert sum(parts, []) == prns
        assert [len(p) for p in parts] == shard_sizes(32, world)
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_merge_drops_slot_padding():
    rows = []
    for i in range(5):
        r = api.Result()
        r.prn, r.code_phase, r.peak = i + 1, 100 + i, 1.5 * i
        rows.append(r)
    sizes, slot = [3, 2], 3
    blob = b"".join(bytes(r) for r in rows[:3]) + b"".join(bytes(r) for r in rows[3:]) + bytes(ROW_BYTES)
    out = merge_table(blob, sizes, slot)
    assert [(r.prn, r.code_phase, r.peak) for r in out] == [(r.prn, r.code_phase, r.peak) for r in rows]


def _to_result(r):
    o = api.Result()
    o.prn, o.acquired, o.code_phase, o.doppler_bin = r.prn, int(r.acquired), r.code_phase, r.doppler_bin
    o.doppler_hz, o.peak, o.noise_meansq, o.snr_db, o.fine_freq_hz = r.doppler_hz, r.peak, r.noise_meansq, r.snr_db, float("nan")
    return o


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs, if_hz = 6e6, 1.25e6
    file, signal, acq = structs(fs, if_hz, datalen=1, freq_min=-500.0, freq_step=500.0, freq_num=3)
    n = int(signal.Sample)
    prns = [1, 3, 7, 22, 30]
    # rank 0 owns the IF block; broadcast it (the NCCL broadcast of the product path)
    raw = torch.frombuffer(bytearray(synth_if(small_spec(fs, if_hz, n), 0, 1)), dtype=torch.uint8).clone() \
        if rank == 0 else torch.empty(n * 2, dtype=torch.uint8)
    dist.broadcast(raw, src=0)
    mine = prn_shard(prns, rank, world)
    local = [_to_result(r) for r in oracle_rows(raw.numpy().tobytes(), file, signal, acq, mine)]
    table = gather_rows_host(dist, local, shard_sizes(len(prns), world))
    if rank == 0:
        q.put([(r.prn, r.code_phase, r.doppler_bin, r.acquired, r.peak) for r in table])
    dist.destroy_process_group()


def test_world2_gloo_table_equals_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    fs, if_hz = 6e6, 1.25e6
    file, signal, acq = structs(fs, if_hz, datalen=1, freq_min=-500.0, freq_step=500.0, freq_num=3)
    raw = synth_if(small_spec(fs, if_hz, int(signal.Sample)), 0, 1)
    want = [(r.prn, r.code_phase, r.doppler_bin, int(r.acquired), r.peak)
            for r in oracle_rows(raw, file, signal, acq, [1, 3, 7, 22, 30])]
    assert got == want


# ---------------------------------------------------------------- bin-split shards (fewer PRNs than GPUs)
def test_shard_plan_tiles_the_grid():
    """gnssacq_shard_plan (host only): the shards' (PRN range x bin range) rectangles tile the (PRN, bin) grid
    exactly once -- whole PRNs when n_prn >= world, all PRNs x a bin range otherwise."""
    for prns, bins, world in [(list(range(1, 33)), 41, 8), (list(range(1, 33)), 41, 5), ([7], 41, 8), ([3, 9, 30], 2001, 8),
                              ([3, 9, 30], 41, 2), ([5], 3, 8), (list(range(1, 9)), 41, 8)]:
        cfg = gnssacq.make_config(prns=prns, freq_num=bins, freq_step_hz=500.0)
        seen = np.zeros((len(prns), bins), dtype=int)
        sizes = []
        for r in range(world):
            mine, sh = api.shard_plan(cfg, r, world)
            assert (sh.rank, sh.world, sh.n_prn_total, sh.freq_num_total) == (r, world, len(prns), bins)
            assert list(mine.prn[: mine.n_prn]) == prns[sh.prn_first: sh.prn_first + sh.prn_count]
            assert (mine.bin_first, mine.bin_count) == (sh.bin_first, sh.bin_count)
            seen[sh.prn_first: sh.prn_first + sh.prn_count, sh.bin_first: sh.bin_first + sh.bin_count] += 1
            sizes.append(sh.prn_count * sh.bin_count)
            if len(prns) >= world:
                assert sh.bin_count == bins and sh.prn_count >= 1
            else:
                assert sh.prn_count == len(prns)
        assert (seen == 1).all()
        if len(prns) >= world or bins >= world:
            assert max(sizes) - min(sizes) <= max(len(prns), bins)
    with pytest.raises(gnssacq.GnssAcqError):
        api.shard_plan(gnssacq.make_config(prns=[1]), 2, 2)


def test_shard_plan_rows_tiles_the_grid():
    """gnssacq_shard_plan_rows (host only): contiguous ranges of the bin-major row list, every row exactly once; the
    root's share follows its weight, the other shards differ by at most one row; a shard may end up empty."""
    for prns, bins, world, extra in [(list(range(1, 33)), 41, 8, 0), (list(range(1, 33)), 41, 8, 45), (list(range(1, 33)), 41, 5, -200),
                                     ([7], 41, 8, 100), ([3, 9, 30], 2001, 8, 10), ([5], 3, 8, 0), ([5, 6], 1, 4, 300),
                                     (list(range(1, 33)), 41, 1, 70)]:
        cfg = gnssacq.make_config(prns=prns, freq_num=bins, freq_step_hz=500.0)
        total = len(prns) * bins
        nxt, sizes = 0, []
        for r in range(world):
            mine, sh = api.shard_plan_rows(cfg, r, world, extra)
            assert (sh.rank, sh.world, sh.n_prn_total, sh.freq_num_total, sh.plan_rows, sh.root_extra_permille) == (r, world, len(prns), bins, 1, extra)
            assert list(mine.prn[: mine.n_prn]) == prns and (mine.bin_first, mine.bin_count) == (0, 0)
            assert sh.n_rows == sh.row_count == mine.row_count
            if sh.row_count:
                assert sh.row_first == nxt == mine.row_first
            nxt += sh.row_count
            sizes.append(sh.row_count)
        assert nxt == total
        assert sizes[0] >= 1
        if world > 1:
            others = sizes[1:]
            assert max(others) - min(others) <= 1
            want_root = total * (1000 + extra) / (1000 * world + extra)
            assert abs(sizes[0] - max(1.0, want_root)) <= 1.0
    with pytest.raises(gnssacq.GnssAcqError):
        api.shard_plan_rows(gnssacq.make_config(prns=[1]), 0, 2, -1000)
    bad = gnssacq.make_config(prns=[1, 2], row_first=80, row_count=10)      # 2 x 41 grid has 82 rows
    with pytest.raises(gnssacq.GnssAcqError):
        api.shard_plan(bad, 0, 1)


def test_root_weight_from_the_measured_wait():
    """dist.root_extra_for_wait (PeerShard.rebalance's arithmetic): no wait keeps the weight; a wait of w after a search
    of s moves w/s * (G-1)/G of the root's rows' worth to the root; planning with the new weight and a linear cost
    model (rows x row time, the others delayed by a fixed amount) levels the finish times to within one row."""
    from gnssacq.dist import root_extra_for_wait
    assert root_extra_for_wait(0, 0.0, 1.0, 8) == 0
    assert root_extra_for_wait(70, 0.0, 1.0, 8) == 70
    assert root_extra_for_wait(0, 0.04, 0.9, 1) == 0
    cfg = gnssacq.make_config(prns=list(range(1, 33)), freq_num=41, freq_step_hz=500.0)
    t_row, delay = 5.5e-3, 40e-3                                             # ms per row, ms the others start late
    for world in (2, 4, 8):
        rows0 = [api.shard_plan_rows(cfg, r, world, 0)[1].row_count for r in range(world)]
        wait = delay + (max(rows0[1:]) - rows0[0]) * t_row                   # what the root measures with equal shares
        extra = root_extra_for_wait(0, wait, rows0[0] * t_row, world)
        rows1 = [api.shard_plan_rows(cfg, r, world, extra)[1].row_count for r in range(world)]
        assert sum(rows1) == 32 * 41 and rows1[0] > rows0[0]
        finish_root, finish_others = rows1[0] * t_row, delay + max(rows1[1:]) * t_row
        assert abs(finish_root - finish_others) <= 1.5 * t_row
        assert max(finish_root, finish_others) <= delay + max(rows0[1:]) * t_row  # never longer (whole rows: not always shorter)


def _cand_of_rows(surface, w):
    """(peak, first lag, sum of squares, windowed sum of squares) of every bin row: what K2/K3 emit per row."""
    out = []
    n = surface.shape[1]
    for row in surface:
        lag = int(np.argmax(row))
        lo, hi = max(lag - (w - 1), 0), min(lag + (w - 1), n - 1)
        out.append((float(row[lag]), lag, float(np.sum(row ** 2)), float(np.sum(row[lo:hi + 1] ** 2))))
    return out


def _bin_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import math
    import oracle
    from oracle.acquisition_ref import correlation_surface, samples_from_bytes
    fs, if_hz = 6e6, 1.25e6
    file, signal, acq = structs(fs, if_hz, datalen=2, freq_min=-1500.0, freq_step=500.0, freq_num=7)
    n = int(signal.Sample)
    raw = torch.frombuffer(bytearray(synth_if(small_spec(fs, if_hz, n), 0, 2)), dtype=torch.uint8).clone() \
        if rank == 0 else torch.empty(n * 2 * 2, dtype=torch.uint8)
    dist.broadcast(raw, src=0)
    cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=[7], freq_min_hz=-1500.0, freq_step_hz=500.0, freq_num=7, noncoh_blocks=2)
    mine, sh = api.shard_plan(cfg, rank, world)                       # one PRN, two ranks: bins are split
    assert sh.prn_count == 1 and sh.bin_count in (3, 4)
    sub = oracle.AcqParams(freqStep=500.0, freqMin=-1500.0 + 500.0 * sh.bin_first, freqNum=sh.bin_count, datalen=2)
    surf = correlation_surface(samples_from_bytes(raw.numpy().tobytes(), 2, 1), signal, sub, 7)   # the oracle stands in for this rank's GPU
    w = int(math.ceil(fs / 1.023e6))
    gathered = [None] * world
    dist.all_gather_object(gathered, (sh.bin_first, _cand_of_rows(surf, w)))
    if rank == 0:
        table = [None] * 7
        for b0, cands in gathered:
            table[b0: b0 + len(cands)] = cands
        q.put(merge_candidates(table, n, w, -1500.0, 500.0))
    dist.destroy_process_group()


def test_world2_gloo_bin_split_merge_equals_single_search():
    """One PRN on two ranks (P < G): every rank searches a range of Doppler bins, the winner tuples are merged with
    K4's rule (max peak, lowest bin, lowest code phase, noise from the winner's own row) -- same row as the
    single-process oracle search of the whole grid (acquisition.m:62-68)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bin_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cp, fbin, dop, peak, noise, snr, acquired = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    fs, if_hz = 6e6, 1.25e6
    file, signal, acq = structs(fs, if_hz, datalen=2, freq_min=-1500.0, freq_step=500.0, freq_num=7)
    raw = synth_if(small_spec(fs, if_hz, int(signal.Sample)), 0, 2)
    ref = oracle_rows(raw, file, signal, acq, [7])[0]
    assert (cp, fbin, dop, acquired) == (ref.code_phase, ref.doppler_bin, ref.doppler_hz, ref.acquired)
    assert abs(peak - ref.peak) <= 1e-12 * ref.peak and abs(snr - ref.snr_db) <= 1e-9


def _row_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import math
    import oracle
    from oracle.acquisition_ref import correlation_surface, samples_from_bytes
    fs, if_hz, prns, bins = 6e6, 1.25e6, [3, 7], 7
    file, signal, acq = structs(fs, if_hz, datalen=2, freq_min=-1500.0, freq_step=500.0, freq_num=bins)
    n = int(signal.Sample)
    raw = torch.frombuffer(bytearray(synth_if(small_spec(fs, if_hz, n), 0, 2)), dtype=torch.uint8).clone() \
        if rank == 0 else torch.empty(n * 2 * 2, dtype=torch.uint8)
    dist.broadcast(raw, src=0)
    cfg = gnssacq.make_config(fs_hz=fs, if_hz=if_hz, prns=prns, freq_min_hz=-1500.0, freq_step_hz=500.0, freq_num=bins, noncoh_blocks=2)
    mine, sh = api.shard_plan_rows(cfg, rank, world, 300)            # 14 rows: a heavy root (8) and one more shard (6)
    assert sh.plan_rows == 1 and mine.n_prn == 2 and mine.row_count == sh.row_count
    w = int(math.ceil(fs / 1.023e6))
    x = samples_from_bytes(raw.numpy().tobytes(), 2, 1)
    surf = {p: None for p in range(len(prns))}
    mine_rows = []
    for row in range(sh.row_first, sh.row_first + sh.row_count):    # row = bin * n_prn + prn index (the kernel's order)
        b, p = divmod(row, len(prns))
        if surf[p] is None:                                          # the oracle stands in for this rank's GPU
            surf[p] = correlation_surface(x, signal, oracle.AcqParams(freqStep=500.0, freqMin=-1500.0, freqNum=bins, datalen=2), prns[p])
        mine_rows.append((p, b, _cand_of_rows(surf[p][b: b + 1], w)[0]))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine_rows)
    if rank == 0:
        table = [[None] * bins for _ in prns]
        for part in gathered:
            for p, b, cand in part:
                assert table[p][b] is None                           # every row exactly once
                table[p][b] = cand
        q.put([merge_candidates(table[p], n, w, -1500.0, 500.0) for p in range(len(prns))])
    dist.destroy_process_group()


def test_world2_gloo_row_ranges_merge_equals_single_search():
    """gnssacq_shard_plan_rows on two ranks with a weighted root: both PRNs have bins on both ranks; the per-row
    candidates, gathered into the full table and merged with K4's rule, give the single-process oracle rows."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_row_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    fs, if_hz = 6e6, 1.25e6
    file, signal, acq = structs(fs, if_hz, datalen=2, freq_min=-1500.0, freq_step=500.0, freq_num=7)
    raw = synth_if(small_spec(fs, if_hz, int(signal.Sample)), 0, 2)
    refs = oracle_rows(raw, file, signal, acq, [3, 7])
    for (cp, fbin, dop, peak, noise, snr, acquired), ref in zip(merged, refs):
        assert (cp, fbin, dop, acquired) == (ref.code_phase, ref.doppler_bin, ref.doppler_hz, ref.acquired)
        assert abs(peak - ref.peak) <= 1e-12 * ref.peak and abs(snr - ref.snr_db) <= 1e-9
