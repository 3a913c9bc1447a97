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
