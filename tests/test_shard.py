"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): sharding + the size all-gather must
give every rank the single-process offsets table."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "huffman-codec_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import shard  # noqa: E402
import synth  # noqa: E402


def _files(n):
    return [synth.image(synth.CLASSES[i % 4], 32, 900 + i, 24).reshape(-1) for i in range(n)]


def _sizes(files):
    import pyoracle
    ora = pyoracle.Oracle()
    return [int(ora.compress(f, diff=True, adapt=True, width=32)[1].size) for f in files]


def _worker(rank, world, port, nfiles, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = _files(nfiles)
    lo, hi = shard.shard_range(nfiles, rank, world)
    local = torch.tensor(_sizes(files[lo:hi]), dtype=torch.int64)      # this rank compresses only its shard
    sizes = shard.gather_sizes(local, nfiles)
    off, total = shard.global_offsets(sizes)
    q.put((rank, sizes.tolist(), off.tolist(), total))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, nfiles):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nfiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return sorted(res)


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 8, 4096, 4099):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_size_allgather_matches_single_process():
    nfiles = 11                                      # not divisible by 2 or 3: ragged shards
    ref_sizes = _sizes(_files(nfiles))
    ref_off, ref_total = shard.global_offsets(torch.tensor(ref_sizes, dtype=torch.int64))
    for world in (2, 3):
        for rank, sizes, off, total in _run(world, nfiles):
            assert sizes == ref_sizes, (world, rank)
            assert off == ref_off.tolist() and total == ref_total


def test_lpt_assignment_balances_and_is_deterministic():
    rng = np.random.default_rng(3)
    costs = list(rng.integers(1, 1000, 57)) + [100000, 90000]
    for world in (1, 2, 3, 8):
        owner, load = shard.lpt_assign(costs, world)
        assert owner == shard.lpt_assign(list(costs), world)[0]
        assert sorted(set(owner)) == list(range(world)) and sum(load) == sum(costs)
        # LPT is within 4/3 of the optimum, which is at least max(largest job, mean load)
        assert max(load) <= 4 / 3 * max(max(costs), sum(costs) / world) + 1
    # the two giants never share a rank when there is a choice; a class-sorted batch is spread evenly
    owner, _ = shard.lpt_assign(costs, 2)
    assert owner[-1] != owner[-2]
    sorted_costs = [262144] * 16 + [1000] * 48
    _, load = shard.lpt_assign(sorted_costs, 4)
    assert max(load) - min(load) <= 1000
    cont = shard.contiguous_assign(64, 4)
    assert sum(c for c, r in zip(sorted_costs, cont) if r == 0) == 16 * 262144


def _xchg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = [torch.full((100 * rank + 10 * d + 1,), 16 * rank + d, dtype=torch.uint8) for d in range(world)]
    got = shard.exchange_streams(parts)
    q.put((rank, [(int(t.numel()), int(t[0]) if t.numel() else -1) for t in got]))
    dist.destroy_process_group()


def test_stream_exchange_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_xchg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank in range(world):
        assert res[rank] == [(100 * s + 10 * rank + 1, 16 * s + rank) for s in range(world)]
