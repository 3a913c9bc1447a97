"""ctypes binding over the C ABI of libhc_b200.so (include/hc_b200.h).

This is the host-side entry used by tests/ and bench.py.  It loads the in-tree CUDA library
and NOTHING else: there is no CPU implementation to fall back to -- a missing library or a
missing CUDA device raises.  (tests/emu binds the same signatures onto the SIMT-emulated test
double through `bind()`; the package itself never does.)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhc_b200.so")
# Kernel debugging only (e.g. a -DHC_FGK_CHECK build of the same sources): honoured only together with
# HC_B200_DEBUG=1, and never for anything that is not a CUDA build of this library (see lib()).
if os.environ.get("HC_B200_DEBUG") == "1" and os.environ.get("HC_B200_LIB"):
    LIB_PATH = os.environ["HC_B200_LIB"]

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p
u32 = C.c_uint32
u64 = C.c_uint64

HC_ALIGN = 256
HC_E_CAPACITY = 100
KIND_PLAIN, KIND_ADAPT, KIND_DIFF = 1, 2, 4

# name -> (restype, argtypes).  Device pointers are passed as integers (c_void_p).
SIGNATURES = {
    "hc_version": (C.c_char_p, []),
    "hc_device_count": (C.c_int, []),
    "hc_error_string": (C.c_char_p, [C.c_int]),
    "hc_launch_count": (u64, []),
    "hc_rle_bound": (u64, [u64]),
    "hc_adapt_bound": (u64, [u64, u64]),
    "hc_fgk_bound": (u64, [u64]),
    "hc_block_count": (u64, [u64, u64, u64]),
    "hc_diff_apply_batch": (C.c_int, [vp, vp, vp, vp, vp, u32, u64, vp]),
    "hc_diff_revert_batch": (C.c_int, [vp, vp, vp, vp, vp, u32, u64, vp]),
    "hc_rle_encode_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, u32, u64, vp]),
    "hc_rle_decode_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, u32, u64, vp]),
    "hc_adapt_encode_ws_bytes": (u64, [u32, u64]),
    "hc_adapt_encode_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, u64, vp, vp]),
    "hc_adapt_decode_ws_bytes": (u64, [u32, u64]),
    "hc_adapt_decode_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, u32, u64, u64, vp, vp]),
    "hc_fgk_encode_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]),
    "hc_fgk_decode_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]),
    "hc_offsets_from_lens": (C.c_int, [vp, vp, vp, u32, u32, vp]),
    "hc_gather_batch": (C.c_int, [vp, vp, vp, vp, vp, u32, u64, vp]),
    "hc_codec_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "hc_codec_destroy": (None, [vp]),
    "hc_host_alloc": (vp, [C.c_size_t]),
    "hc_host_free": (None, [vp]),
    "hc_compress_batch": (C.c_int, [vp, vp, vp, vp, u32, C.c_int, C.c_int, vp, vp, u64, vp, vp, vp]),
    "hc_decompress_batch": (C.c_int, [vp, vp, vp, vp, u32, vp, u64, vp, vp, vp]),
    "hc_compress_device": (C.c_int, [vp, vp, vp, vp, vp, u32, u64, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "hc_decompress_device": (C.c_int, [vp, vp, vp, vp, u32, u64, u64, C.c_int, vp, vp, vp, vp, vp]),
    "hc_pipeline_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int]),
    "hc_pipeline_destroy": (None, [vp]),
    "hc_pipeline_submit_compress": (C.c_int64, [vp, vp, vp, vp, u32, C.c_int, C.c_int, vp, vp, u64, vp, vp, vp]),
    "hc_pipeline_submit_decompress": (C.c_int64, [vp, vp, vp, vp, u32, vp, u64, vp, vp, vp]),
    "hc_pipeline_wait": (C.c_int, [vp, C.c_int64]),
    "hc_shard_ws_bytes": (u64, [u32, C.c_int]),
    "hc_shard_sizes_allgather": (C.c_int, [vp, C.c_int, C.c_int, vp, u32, u32, vp, vp, vp, vp, vp]),
    "hc_codec_stream": (vp, [vp]),
    "hc_codec_stage_times": (C.c_int, [vp, C.POINTER(C.c_float), C.c_int]),
    "hc_stage_name": (C.c_char_p, [vp, C.c_int]),
    "hc_codec_enable_stage_timing": (None, [vp, C.c_int]),
}


def bind(path):
    """Load a shared library exporting the hc_b200 C ABI and attach the signatures."""
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib():
    """The product library.  Raises if it has not been built (python __graft_entry__.py)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libhc_b200.so is missing (%s); build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " There is no CPU fallback." % LIB_PATH)
        _LIB = bind(LIB_PATH)
        if b"sm_100a" not in _LIB.hc_version():
            raise RuntimeError("%s is not a CUDA build of libhc_b200 (%r)" % (LIB_PATH, _LIB.hc_version()))
    return _LIB


def check(rc, what="hc call", L=None):
    if rc != 0:
        L = L or lib()
        raise RuntimeError("%s failed: %d (%s)" % (what, rc, L.hc_error_string(rc).decode()))


def align_up(v, a=HC_ALIGN):
    return (int(v) + a - 1) // a * a


def _np_ptr(a):
    return a.ctypes.data


class Codec:
    """Batched huffCompress / huffDecompress (reference src/main.cpp:39-128) on one GPU.

    Files are passed as a list of bytes-like / uint8 arrays; results come back as a list of
    numpy uint8 arrays plus the per-file status (reference exit code, 0 = ok)."""

    def __init__(self, device=0, L=None):
        self.L = L or lib()
        h = vp()
        check(self.L.hc_codec_create(C.byref(h), device), "hc_codec_create", self.L)
        self.h = h

    def close(self):
        if self.h:
            self.L.hc_codec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def pack(files, align=16):
        """-> (buffer u8, off u64[nf], len u64[nf]); starts aligned to `align` bytes."""
        lens = np.array([len(f) for f in files], dtype=np.uint64)
        offs = np.zeros(len(files), dtype=np.uint64)
        pos = 0
        for i, n in enumerate(lens):
            offs[i] = pos
            pos += align_up(int(n), align)
        buf = np.zeros(max(pos, 1), dtype=np.uint8)
        for i, f in enumerate(files):
            a = np.frombuffer(bytes(f), np.uint8) if not isinstance(f, np.ndarray) else f.reshape(-1)
            buf[int(offs[i]):int(offs[i]) + a.size] = a
        return buf, offs, lens

    def compress_packed(self, buf, offs, lens, diff=False, adapt=False, widths=None, out=None):
        """Host buffers in, host buffer out (compact, 16-byte aligned starts)."""
        nf = len(lens)
        if out is None:
            bound = 0
            for n in lens:
                m = int(n) + int(n) // 3 + 64 + (int(n) // 8 + 64 if adapt else 0)
                bound += align_up(int(self.L.hc_fgk_bound(m)) + 16, 16)
            out = np.empty(max(bound, 16), dtype=np.uint8)
        out_off = np.zeros(nf, np.uint64)
        out_len = np.zeros(nf, np.uint64)
        status = np.zeros(nf, np.int32)
        w = None
        if widths is not None:
            w = np.ascontiguousarray(np.broadcast_to(np.asarray(widths, dtype=np.uint64), (nf,)))
        rc = self.L.hc_compress_batch(self.h, _np_ptr(buf), _np_ptr(offs), _np_ptr(lens), nf, int(diff), int(adapt),
                                      _np_ptr(w) if w is not None else None, _np_ptr(out), out.size,
                                      _np_ptr(out_off), _np_ptr(out_len), _np_ptr(status))
        check(rc, "hc_compress_batch", self.L)
        return out, out_off, out_len, status

    def compress(self, files, diff=False, adapt=False, width=512):
        buf, offs, lens = self.pack(files)
        out, oo, ol, st = self.compress_packed(buf, offs, lens, diff, adapt, width)
        return [out[int(o):int(o) + int(n)].copy() for o, n in zip(oo, ol)], st

    def decompress_packed(self, buf, offs, lens, out_cap):
        nf = len(lens)
        out = np.empty(max(int(out_cap), 16), dtype=np.uint8)
        out_off = np.zeros(nf, np.uint64)
        out_len = np.zeros(nf, np.uint64)
        status = np.zeros(nf, np.int32)
        rc = self.L.hc_decompress_batch(self.h, _np_ptr(buf), _np_ptr(offs), _np_ptr(lens), nf, _np_ptr(out), out.size,
                                        _np_ptr(out_off), _np_ptr(out_len), _np_ptr(status))
        return rc, out, out_off, out_len, status

    def decompress(self, files, out_cap=None):
        """-> (list of arrays, status).  A file that does not fit the first buffer comes back with status
        HC_E_CAPACITY and the size it needs, so one second call with the exact total settles it; anything
        beyond 255 x the input (no MNP-5 token expands further) stays that file's error."""
        buf, offs, lens = self.pack(files)
        in_total = int(lens.sum())
        cap = out_cap if out_cap is not None else max(1 << 20, 8 * in_total)
        ceiling = 255 * in_total + 16 * len(files) + 4096
        for attempt in range(2):
            rc, out, oo, ol, st = self.decompress_packed(buf, offs, lens, cap)
            check(rc, "hc_decompress_batch", self.L)
            if attempt == 0 and (st == HC_E_CAPACITY).any():
                need = int(sum(align_up(int(n), 16) for n, s in zip(ol, st) if s in (0, HC_E_CAPACITY)))
                if need <= ceiling:
                    cap = need + 4096
                    continue
            break
        res = [out[int(o):int(o) + int(n)].copy() if s == 0 else np.zeros(0, np.uint8) for o, n, s in zip(oo, ol, st)]
        return res, st
