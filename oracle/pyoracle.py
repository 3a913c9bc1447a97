"""ctypes access to the CPU oracle (oracle/_build/libhc_oracle.so) and, when it has been
built, to the unmodified reference (oracle/_ref/libhcref.so, oracle/_ref/huffman-codec).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Nothing under huffman-codec_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libhc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhcref.so")
REF_BIN = os.path.join(HERE, "_ref", "huffman-codec")
REF_BIN_O0 = os.path.join(HERE, "_ref", "huffman-codec-O0")

_u8p = C.POINTER(C.c_uint8)


def build(ref=True):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _arr(a):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a,
                             dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


class Oracle:
    """Plain-C restatement (oracle/hc_oracle.c)."""

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.hco_free.argtypes = [C.c_void_p]
        L.hco_diff_apply.argtypes = [_u8p, C.c_size_t]
        L.hco_diff_revert.argtypes = [_u8p, C.c_size_t]
        L.hco_rle_bound.argtypes = [C.c_size_t]
        L.hco_rle_bound.restype = C.c_size_t
        L.hco_rle_encode.argtypes = [_u8p, C.c_size_t, _u8p]
        L.hco_rle_encode.restype = C.c_size_t
        L.hco_rle_decode.argtypes = [_u8p, C.c_size_t, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.hco_block_count.argtypes = [C.c_uint64] * 3
        L.hco_block_count.restype = C.c_uint64
        L.hco_adapt_encode_bs.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.hco_adapt_encode.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.POINTER(_u8p),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.hco_adapt_decode.argtypes = [_u8p, C.c_size_t, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.hco_fgk_encode.argtypes = [_u8p, C.c_size_t, C.c_int, C.POINTER(_u8p),
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.hco_fgk_decode.argtypes = [_u8p, C.c_size_t, C.c_uint64, C.c_int, _u8p]
        L.hco_compress.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                   C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.hco_decompress.argtypes = [_u8p, C.c_size_t, C.c_int, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.hco_fgk_stats.argtypes = [_u8p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint32)]

    def _take(self, p, n):
        out = np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy() if n else np.zeros(0, np.uint8)
        self.L.hco_free(p)
        return out

    def diff_apply(self, a):
        a = np.array(a, dtype=np.uint8, copy=True)
        self.L.hco_diff_apply(a.ctypes.data_as(_u8p), a.size)
        return a

    def diff_revert(self, a):
        a = np.array(a, dtype=np.uint8, copy=True)
        self.L.hco_diff_revert(a.ctypes.data_as(_u8p), a.size)
        return a

    def rle_encode(self, a):
        a, p = _arr(a)
        out = np.empty(self.L.hco_rle_bound(a.size), np.uint8)
        m = self.L.hco_rle_encode(p, a.size, out.ctypes.data_as(_u8p))
        return out[:m].copy()

    def rle_decode(self, a):
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t()
        rc = self.L.hco_rle_decode(p, a.size, C.byref(o), C.byref(n))
        assert rc == 0
        return self._take(o, n.value)

    def block_count(self, w, h, b):
        return self.L.hco_block_count(w, h, b)

    def adapt_encode_bs(self, a, w, h, b):
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t()
        self.L.hco_adapt_encode_bs(p, w, h, b, C.byref(o), C.byref(n))
        return self._take(o, n.value)

    def adapt_encode(self, a, w, h):
        """-> (rc, bytes, chosen block size)"""
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t(); b = C.c_uint64()
        rc = self.L.hco_adapt_encode(p, w, h, C.byref(o), C.byref(n), C.byref(b))
        if rc:
            return rc, None, 0
        return 0, self._take(o, n.value), b.value

    def adapt_decode(self, a):
        """-> (rc, bytes)"""
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t()
        rc = self.L.hco_adapt_decode(p, a.size, C.byref(o), C.byref(n))
        if rc:
            return rc, None
        return 0, self._take(o, n.value)

    def fgk_encode(self, a, mode=1):
        """-> (packed bytes, raw bit count)"""
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t(); nb = C.c_uint64()
        self.L.hco_fgk_encode(p, a.size, mode, C.byref(o), C.byref(n), C.byref(nb))
        return self._take(o, n.value), nb.value

    def fgk_decode(self, a, count, mode=1):
        """-> (rc, symbols)"""
        a, p = _arr(a)
        out = np.empty(max(count, 1), np.uint8)
        rc = self.L.hco_fgk_decode(p, a.size, count, mode, out.ctypes.data_as(_u8p))
        return rc, (out[:count] if rc == 0 else None)

    def compress(self, a, diff=False, adapt=False, width=512, mode=1):
        """-> (rc, .out bytes)"""
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t()
        rc = self.L.hco_compress(p, a.size, int(diff), int(adapt), width, mode, C.byref(o), C.byref(n))
        if rc:
            return rc, None
        return 0, self._take(o, n.value)

    def decompress(self, a, mode=1):
        a, p = _arr(a)
        o = _u8p(); n = C.c_size_t()
        rc = self.L.hco_decompress(p, a.size, mode, C.byref(o), C.byref(n))
        if rc:
            return rc, None
        return 0, self._take(o, n.value)

    def fgk_stats(self, a):
        a, p = _arr(a)
        lv = C.c_uint64(); sw = C.c_uint64(); md = C.c_uint32()
        self.L.hco_fgk_stats(p, a.size, C.byref(lv), C.byref(sw), C.byref(md))
        return lv.value, sw.value, md.value


class Ref:
    """The unmodified reference compiled into oracle/_ref (stage functions via ref_shim.cpp)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.L = C.CDLL(path)
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_diff_apply.argtypes = [_u8p, C.c_size_t]
        L.ref_diff_revert.argtypes = [_u8p, C.c_size_t]
        for name in ("ref_rle_encode", "ref_rle_decode", "ref_adapt_decode", "ref_fgk_encode"):
            f = getattr(L, name)
            f.argtypes = [_u8p, C.c_size_t, C.POINTER(C.c_size_t)]
            f.restype = _u8p
        L.ref_adapt_encode.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.POINTER(C.c_size_t)]
        L.ref_adapt_encode.restype = _u8p
        L.ref_adapt_encode_bs.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_size_t)]
        L.ref_adapt_encode_bs.restype = _u8p
        L.ref_fgk_decode.argtypes = [_u8p, C.c_size_t, C.c_uint64, _u8p]

    def _take(self, p, n):
        out = np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy() if n else np.zeros(0, np.uint8)
        self.L.ref_free(p)
        return out

    def _call(self, name, a, *extra):
        a, p = _arr(a)
        n = C.c_size_t()
        if extra:
            o = getattr(self.L, name)(p, *extra, C.byref(n))
        else:
            o = getattr(self.L, name)(p, a.size, C.byref(n))
        return self._take(o, n.value)

    def diff_apply(self, a):
        a = np.array(a, dtype=np.uint8, copy=True)
        self.L.ref_diff_apply(a.ctypes.data_as(_u8p), a.size)
        return a

    def diff_revert(self, a):
        a = np.array(a, dtype=np.uint8, copy=True)
        self.L.ref_diff_revert(a.ctypes.data_as(_u8p), a.size)
        return a

    def rle_encode(self, a): return self._call("ref_rle_encode", a)
    def rle_decode(self, a): return self._call("ref_rle_decode", a)
    def adapt_encode(self, a, w, h): return self._call("ref_adapt_encode", a, w, h)
    def adapt_encode_bs(self, a, w, h, b): return self._call("ref_adapt_encode_bs", a, w, h, b)
    def adapt_decode(self, a): return self._call("ref_adapt_decode", a)
    def fgk_encode(self, a): return self._call("ref_fgk_encode", a)

    def fgk_decode(self, a, count):
        a, p = _arr(a)
        out = np.empty(max(count, 1), np.uint8)
        self.L.ref_fgk_decode(p, a.size, count, out.ctypes.data_as(_u8p))
        return out[:count]


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


def ref_cli(args, cwd=None, binary=REF_BIN):
    """Run the reference binary; -> (exit code, stderr text)."""
    r = subprocess.run([binary] + list(args), cwd=cwd, capture_output=True)
    return r.returncode, r.stderr.decode(errors="replace"), r.stdout.decode(errors="replace")
