"""Multi-GPU plumbing of the batched codec.

Files are independent (the FGK tree is built per file, reference src/transform.cpp:366,388), so a
batch shards over ranks with NO data-path collective: every rank runs the full pipeline on its own
contiguous slice of the batch.  The only exchange is one all-gather of the per-file output sizes
(a few KB), from which every rank derives the same global offsets table -- the index of the
concatenated output container.  Works with any torch.distributed backend (NCCL on GPUs, gloo on
CPU for the tests)."""
import torch
import torch.distributed as dist


def shard_range(nfiles, rank, world):
    """Contiguous, balanced slice [lo, hi) of the batch owned by `rank`."""
    base, extra = divmod(nfiles, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_sizes(local_sizes, nfiles, group=None):
    """All-gather the per-file output sizes of every rank's shard.

    local_sizes: 1-D int64 tensor (this rank's shard, on the backend's device).
    Returns a 1-D int64 tensor of nfiles entries in global file order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_sizes.clone()
    per = -(-nfiles // world)                      # shards differ by at most one file: pad to equal
    padded = torch.zeros(per, dtype=torch.int64, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    out = torch.empty(per * world, dtype=torch.int64, device=local_sizes.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(nfiles, r, world)
        parts.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(parts)


def global_offsets(sizes, align=16):
    """Exclusive scan of the aligned sizes -> (offsets, total) of the concatenated container."""
    al = (sizes + (align - 1)) // align * align
    off = torch.cumsum(al, 0) - al
    return off, int(al.sum().item())


class NcclComm:
    """An NCCL communicator made with the NCCL C API (ctypes), for the C-ABI collective
    hc_shard_sizes_allgather.  The unique id travels over the already initialised torch.distributed
    control group (gloo or nccl).  Uses the NCCL library torch ships so that one copy lives in the process."""

    def __init__(self, rank, world, group=None):
        import ctypes as C
        import glob
        import os
        import hc_b200
        self.C, self.rank, self.world = C, rank, world
        self.L = hc_b200.lib()
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*")) + ["libnccl.so.2"]
        self.nccl = None
        for c in cands:
            try:
                self.nccl = C.CDLL(c, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if self.nccl is None:
            raise RuntimeError("NCCL library not found")
        uid = (C.c_char * 128)()
        if rank == 0:
            self._ck(self.nccl.ncclGetUniqueId(C.byref(uid)))
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0, group=group)
        uid = (C.c_char * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
        self.comm = C.c_void_p()

        class _Uid(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]
        self.nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _Uid, C.c_int]
        u = _Uid()
        C.memmove(C.byref(u), uid, 128)
        self._ck(self.nccl.ncclCommInitRank(C.byref(self.comm), world, u, rank))
        self.ws = None

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL error %d" % rc)

    def gather_sizes(self, local_sizes, nfiles, align=16, stream=None):
        """local_sizes: int64 CUDA tensor of this rank's contiguous shard -> (sizes, offsets, total) in global order."""
        import hc_b200
        dev = local_sizes.device
        need = int(self.L.hc_shard_ws_bytes(nfiles, self.world))
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        sizes = torch.empty(nfiles, dtype=torch.int64, device=dev)
        offs = torch.empty(nfiles, dtype=torch.int64, device=dev)
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        hc_b200.check(self.L.hc_shard_sizes_allgather(self.comm, self.rank, self.world, local_sizes.data_ptr(), nfiles, align,
                                                      sizes.data_ptr(), offs.data_ptr(), total.data_ptr(), self.ws.data_ptr(), st),
                      "hc_shard_sizes_allgather", self.L)
        return sizes, offs, total

    def close(self):
        if self.comm:
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = None


# ---------------------------------------------------------------------------------------------------
# Cost-aware sharding: the FGK stage is serial per stream and its cost is the
# stream's post-RLE symbol count, which is only known after the transform stage.  Every rank transforms
# its contiguous shard, the symbol counts are all-gathered, every rank computes the same
# longest-processing-time assignment, the symbol streams that change owner travel over NVLink
# (all_to_all on the NCCL backend; send/recv pairs on gloo for the CPU tests), and each rank entropy
# codes the streams it owns.  The offsets table is again derived from an all-gather of the sizes.
# ---------------------------------------------------------------------------------------------------
def lpt_assign(costs, world):
    """Longest-processing-time-first: files in order of decreasing cost (ties: lower index first) go to the
    rank with the least load so far (ties: lower rank).  Deterministic, identical on every rank.
    Returns (owner[nfiles] as a list of ranks, load[world])."""
    import heapq
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    owner = [0] * len(costs)
    for i in order:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + int(costs[i]), r))
    load = [0] * world
    for i, r in enumerate(owner):
        load[r] += int(costs[i])
    return owner, load


def contiguous_assign(nfiles, world):
    owner = []
    for r in range(world):
        lo, hi = shard_range(nfiles, r, world)
        owner += [r] * (hi - lo)
    return owner


def exchange_streams(parts, group=None):
    """parts[d] = 1-D uint8 tensor of the bytes this rank sends to rank d (parts[rank] stays local).
    Returns the list of tensors received, indexed by source rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return [parts[0]]
    dev = parts[0].device
    send_counts = torch.tensor([p.numel() for p in parts], dtype=torch.int64, device=dev)
    recv_counts = torch.empty(world, dtype=torch.int64, device=dev)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        rc = recv_counts.tolist()
        sc = send_counts.tolist()
        out = torch.empty(sum(rc), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(out, torch.cat(parts) if sum(sc) else torch.empty(0, dtype=torch.uint8, device=dev),
                               output_split_sizes=rc, input_split_sizes=sc, group=group)
        res, pos = [], 0
        for n in rc:
            res.append(out[pos:pos + n])
            pos += n
        return res
    # gloo (CPU tests): counts by all_gather, payload by send / recv pairs ordered to avoid deadlock
    allc = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, send_counts.cpu(), group=group)
    res = [None] * world
    res[rank] = parts[rank]
    for step in range(1, world):
        dst, src = (rank + step) % world, (rank - step) % world
        rbuf = torch.empty(int(allc[src][rank]), dtype=torch.uint8)
        reqs = []
        if parts[dst].numel():
            reqs.append(dist.isend(parts[dst].contiguous(), dst, group=group))
        if rbuf.numel():
            reqs.append(dist.irecv(rbuf, src, group=group))
        for q in reqs:
            q.wait()
        res[src] = rbuf
    return res


class ShardedCompressor:
    """Batched huffCompress over all ranks with cost-aware placement of the FGK stage, built on the
    stage-level C ABI (hc_diff_apply_batch, hc_adapt_encode_batch / hc_rle_encode_batch, hc_gather_batch,
    hc_fgk_encode_batch).  Device resident; equal-sized files (side x side images)."""

    def __init__(self, L, rank, world, device, side, use_adapt=True, group=None):
        self.L, self.rank, self.world, self.dev, self.side, self.use_adapt, self.group = L, rank, world, device, side, use_adapt, group
        self.FB = side * side

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %d" % (what, rc))

    def transform(self, d_in):
        """diff model + adaptive block RLE (or MNP-5 RLE) of this rank's shard: nf x FB u8 on the device.
        Returns (symbol buffer, offsets, lengths) with capacity-strided, 256-byte aligned streams."""
        L, FB, dev = self.L, self.FB, self.dev
        nf = d_in.shape[0]
        i64 = torch.int64
        st = torch.cuda.current_stream(dev).cuda_stream
        off = torch.arange(nf, dtype=i64, device=dev) * FB
        ln = torch.full((nf,), FB, dtype=i64, device=dev)
        a = torch.empty(nf * FB + 512, dtype=torch.uint8, device=dev)
        self._ck(L.hc_diff_apply_batch(d_in.data_ptr(), off.data_ptr(), a.data_ptr(), off.data_ptr(), ln.data_ptr(), nf, FB, st), "hc_diff_apply_batch")
        bound = int(L.hc_adapt_bound(self.side, self.side)) if self.use_adapt else int(L.hc_rle_bound(FB))
        stride = (bound + 16 + 255) // 256 * 256
        b = torch.empty(nf * stride + 512, dtype=torch.uint8, device=dev)
        boff = torch.arange(nf, dtype=i64, device=dev) * stride
        blen = torch.zeros(nf, dtype=i64, device=dev)
        if self.use_adapt:
            wd = torch.full((nf,), self.side, dtype=i64, device=dev)
            stt = torch.zeros(nf, dtype=torch.int32, device=dev)
            ws = torch.empty(int(L.hc_adapt_encode_ws_bytes(nf, FB)) + 512, dtype=torch.uint8, device=dev)
            self._ck(L.hc_adapt_encode_batch(a.data_ptr(), off.data_ptr(), wd.data_ptr(), wd.data_ptr(), b.data_ptr(), boff.data_ptr(),
                                             blen.data_ptr(), None, stt.data_ptr(), nf, FB, ws.data_ptr(), st), "hc_adapt_encode_batch")
            assert int(stt.abs().sum().item()) == 0
        else:
            self._ck(L.hc_rle_encode_batch(a.data_ptr(), off.data_ptr(), ln.data_ptr(), b.data_ptr(), boff.data_ptr(), blen.data_ptr(), nf, FB, st),
                     "hc_rle_encode_batch")
        return b, boff, blen, stride

    def compress(self, d_in, lo, nfiles, policy="lpt"):
        """d_in: this rank's contiguous shard [lo, lo + nf) of a batch of nfiles.  Returns a dict with the global
        ids this rank entropy-coded, their .out bytes (device buffer + offsets + lengths), the global sizes and
        offsets table, and the per-rank FGK load (symbols)."""
        L, dev, world, rank = self.L, self.dev, self.world, self.rank
        i64 = torch.int64
        st = torch.cuda.current_stream(dev).cuda_stream
        nf = d_in.shape[0]
        b, boff, blen, stride = self.transform(d_in)
        # symbol counts of the whole batch (the path's collective, now carrying costs)
        costs = gather_sizes(blen, nfiles, self.group).cpu().tolist()
        owner = lpt_assign(costs, world)[0] if policy == "lpt" else contiguous_assign(nfiles, world)
        load = [0] * world
        for i, r in enumerate(owner):
            load[r] += costs[i]
        # pack what goes to every rank (streams padded to 256 bytes so that they can be coded in place)
        mine = list(range(lo, lo + nf))
        parts, sent_ids, keep = [], [], []
        for d in range(world):
            ids = [g for g in mine if owner[g] == d]
            sent_ids.append(ids)
            if not ids:
                parts.append(torch.empty(0, dtype=torch.uint8, device=dev))
                continue
            sel = torch.tensor([g - lo for g in ids], dtype=i64, device=dev)
            # (every device array handed to the C ABI stays referenced until after the call: a temporary freed
            # inside the argument list would be recycled by the allocator for the next temporary)
            lens = blen[sel].contiguous()
            src_off = boff[sel].contiguous()
            pad = (lens + 16 + 255) // 256 * 256
            poff = (torch.cumsum(pad, 0) - pad).contiguous()
            max_len = int(lens.max().item())
            buf = torch.zeros(int(pad.sum().item()), dtype=torch.uint8, device=dev)
            self._ck(L.hc_gather_batch(b.data_ptr(), src_off.data_ptr(), lens.data_ptr(), buf.data_ptr(), poff.data_ptr(), len(ids), max_len, st),
                     "hc_gather_batch")
            keep.append((lens, src_off, poff))
            parts.append(buf)
        torch.cuda.current_stream(dev).synchronize()
        recv = exchange_streams(parts, self.group)
        # what arrived, in (source rank, global id) order -- every rank can derive it from `owner`
        got_ids, got_len, got_off, pos = [], [], [], 0
        chunks = []
        for src in range(world):
            slo, shi = shard_range(nfiles, src, world)
            ids = [g for g in range(slo, shi) if owner[g] == rank]
            p = 0
            for g in ids:
                got_ids.append(g)
                got_len.append(costs[g])
                got_off.append(pos + p)
                p += (costs[g] + 16 + 255) // 256 * 256
            assert p == recv[src].numel(), (src, p, recv[src].numel())
            pos += p
            chunks.append(recv[src])
        n_own = len(got_ids)
        sym = torch.cat(chunks + [torch.zeros(512, dtype=torch.uint8, device=dev)])
        cap = (int(L.hc_fgk_bound(max(got_len) if got_len else 0)) + 16 + 255) // 256 * 256
        out = torch.empty(n_own * cap + 512, dtype=torch.uint8, device=dev)
        o_off = torch.arange(n_own, dtype=i64, device=dev) * cap
        o_cap = torch.full((n_own,), cap, dtype=i64, device=dev)
        o_len = torch.zeros(n_own, dtype=i64, device=dev)
        stt = torch.zeros(n_own, dtype=torch.int32, device=dev)
        flags = torch.full((n_own,), 0x80 | (0x40 if self.use_adapt else 0), dtype=torch.uint8, device=dev)
        if n_own:
            s_off = torch.tensor(got_off, dtype=i64, device=dev)
            s_len = torch.tensor(got_len, dtype=i64, device=dev)
            self._ck(L.hc_fgk_encode_batch(sym.data_ptr(), s_off.data_ptr(), s_len.data_ptr(), flags.data_ptr(), out.data_ptr(), o_off.data_ptr(),
                                           o_cap.data_ptr(), o_len.data_ptr(), stt.data_ptr(), n_own, st), "hc_fgk_encode_batch")
            assert int(stt.abs().sum().item()) == 0
        # sizes of the whole batch in global file order -> offsets table
        per = max(1, max(sum(1 for r in owner if r == k) for k in range(world)))
        pad_ids = torch.full((per,), -1, dtype=i64, device=dev)
        pad_sz = torch.zeros(per, dtype=i64, device=dev)
        if n_own:
            pad_ids[:n_own] = torch.tensor(got_ids, dtype=i64, device=dev)
            pad_sz[:n_own] = o_len
        if world > 1:
            all_ids = torch.empty(per * world, dtype=i64, device=dev)
            all_sz = torch.empty(per * world, dtype=i64, device=dev)
            dist.all_gather_into_tensor(all_ids, pad_ids, group=self.group)
            dist.all_gather_into_tensor(all_sz, pad_sz, group=self.group)
        else:
            all_ids, all_sz = pad_ids, pad_sz
        sizes = torch.zeros(nfiles, dtype=i64, device=dev)
        ok = all_ids >= 0
        sizes[all_ids[ok]] = all_sz[ok]
        offs, total = global_offsets(sizes, 16)
        return {"ids": got_ids, "out": out, "out_off": o_off, "out_len": o_len, "sizes": sizes, "offsets": offs, "total": total,
                "load": load, "owner": owner}
