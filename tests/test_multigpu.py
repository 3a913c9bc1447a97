"""Multi-GPU parity (`-m gpu`, needs >= 2 visible devices; run with `gpurun --gpus 2`): the same batch
compressed on 1 GPU and sharded over N GPUs must give identical bytes and an identical offsets table
(SURVEY.md 4(4), 8e).  The size exchange goes through the C ABI (hc_shard_sizes_allgather) on an NCCL
communicator created with the NCCL C API -- no torch.distributed NCCL backend on the data path."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "huffman-codec_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import shard  # noqa: E402
import synth  # noqa: E402

pytestmark = pytest.mark.gpu
NFILES = 45          # not divisible by 2 or 4: ragged shards


def _files():
    kinds = synth.CLASSES + ("fib", "longrun")
    return [synth.image(kinds[i % len(kinds)], 96, 7000 + i, 64 + 8 * (i % 4)).reshape(-1) for i in range(NFILES)]


def _worker(rank, world, port, q):
    import hc_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)           # control plane only
    torch.cuda.set_device(rank)
    comm = shard.NcclComm(rank, world)                                       # NCCL C API communicator
    files = _files()
    lo, hi = shard.shard_range(NFILES, rank, world)
    cd = hc_b200.Codec(rank)
    outs, st = cd.compress(files[lo:hi], diff=True, adapt=True, width=96)
    assert not st.any()
    local = torch.tensor([o.size for o in outs], dtype=torch.int64, device="cuda")
    sizes, offs, total = comm.gather_sizes(local, NFILES, 16)
    back, st = cd.decompress(outs)
    assert not st.any() and all(np.array_equal(a, b) for a, b in zip(back, files[lo:hi]))
    q.put((rank, sizes.cpu().tolist(), offs.cpu().tolist(), int(total), [o.tobytes() for o in outs]))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_batch_equals_single_gpu(world, oracle):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import hc_b200
    files = _files()
    one, st = hc_b200.Codec(0).compress(files, diff=True, adapt=True, width=96)
    assert not st.any()
    for f, o in zip(files[:8], one):
        rc, exp = oracle.compress(f, diff=True, adapt=True, width=96)
        assert rc == 0 and np.array_equal(o, exp)
    ref_sizes = [o.size for o in one]
    ref_off, ref_total = shard.global_offsets(torch.tensor(ref_sizes, dtype=torch.int64), 16)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    blob = []
    for rank, sizes, offs, total, outs in res:
        assert sizes == ref_sizes and offs == ref_off.tolist() and total == ref_total, rank
        blob += outs
    assert len(blob) == NFILES and all(a == o.tobytes() for a, o in zip(blob, one))


def _lpt_worker(rank, world, port, q):
    import hc_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    side, n = 64, 22
    # class-sorted batch: the contiguous split puts every high-entropy file on the first rank
    kinds = ["random"] * 6 + ["walk"] * 5 + ["smooth"] * 6 + ["const"] * 5
    files = [synth.image(kinds[i], side, 8000 + i).reshape(-1) for i in range(n)]
    lo, hi = shard.shard_range(n, rank, world)
    d_in = torch.from_numpy(np.stack(files[lo:hi])).cuda()
    sc = shard.ShardedCompressor(hc_b200.lib(), rank, world, torch.device("cuda", rank), side, use_adapt=True)
    res = {}
    for policy in ("contiguous", "lpt"):
        r = sc.compress(d_in, lo, n, policy)
        torch.cuda.synchronize()
        lens = r["out_len"].cpu().tolist()
        offs = r["out_off"].cpu().tolist()
        blob = r["out"].cpu().numpy()
        res[policy] = {"ids": r["ids"], "bytes": [blob[o:o + ln].tobytes() for o, ln in zip(offs, lens)], "sizes": r["sizes"].cpu().tolist(),
                       "offsets": r["offsets"].cpu().tolist(), "load": r["load"]}
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cost_aware_sharding_gives_identical_files(world, oracle):
    """LPT placement of the FGK stage + migration of the symbol streams over NVLink: same bytes, same offsets table,
    better balance than contiguous slices on a class-sorted batch."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    side, n = 64, 22
    kinds = ["random"] * 6 + ["walk"] * 5 + ["smooth"] * 6 + ["const"] * 5
    files = [synth.image(kinds[i], side, 8000 + i).reshape(-1) for i in range(n)]
    exp = [oracle.compress(f, diff=True, adapt=True, width=side)[1].tobytes() for f in files]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lpt_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for policy in ("contiguous", "lpt"):
        seen = {}
        for rank in range(world):
            r = res[rank][policy]
            assert r["sizes"] == [len(e) for e in exp], (policy, rank)
            for g, b in zip(r["ids"], r["bytes"]):
                assert g not in seen
                seen[g] = b
        assert sorted(seen) == list(range(n)) and all(seen[g] == exp[g] for g in range(n)), policy
    lc, ll = res[0]["contiguous"]["load"], res[0]["lpt"]["load"]
    assert max(ll) < max(lc) and max(ll) <= 4 / 3 * max(sum(ll) / world, max(len(e) for e in exp) * 8)
