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
from gnssacq.dist import prn_shard, shard_sizes, merge_table, gather_rows_host, ROW_BYTES
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
