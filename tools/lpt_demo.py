#!/usr/bin/env python3
"""Cost-aware sharding demo, run under torchrun on N GPUs of one box:
the C3 batch SORTED BY CLASS (all high-entropy files first: the worst case for contiguous shards) is compressed with
contiguous placement and with LPT placement of the FGK stage (symbol streams migrate over NVLink).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "huffman-codec_b200")):
    sys.path.insert(0, p)
import hc_b200  # noqa: E402
import shard  # noqa: E402
import synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sys.stdout.flush()
saved = os.dup(1)
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.barrier()
os.dup2(saved, 1)
lo, hi = shard.shard_range(total, rank, world)
per = total // 4
cls = lambda g: synth.CLASSES[(2, 0, 1, 3)[min(g // per, 3)]]           # random, walk, smooth, const blocks
files = np.stack([synth.image(cls(g), 512, 1234 + g).reshape(-1) for g in range(lo, hi)])
d_in = torch.from_numpy(files).to(dev)
sc = shard.ShardedCompressor(hc_b200.lib(), rank, world, dev, 512, use_adapt=True)
out = {}
for policy in ("contiguous", "lpt"):
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = sc.compress(d_in, lo, total, policy)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[policy] = {"seconds": float(t.item()), "GB_per_s": total * 512 * 512 / float(t.item()) / 1e9, "fgk_symbols_per_rank": r["load"],
                   "bytes_total": r["total"]}
assert out["contiguous"]["bytes_total"] == out["lpt"]["bytes_total"]
if rank == 0:
    out["workload"] = "%d x 512x512 C3 images sorted by class (random | walk | smooth | const), -m -a, compress only, %d GPUs" % (total, world)
    out["speedup"] = out["contiguous"]["seconds"] / out["lpt"]["seconds"]
    print(json.dumps(out))
dist.destroy_process_group()
