#!/usr/bin/env python3
"""FGK kernels under the DEBUG build of the SIMT emulator (CPU only): -DHC_EMU_DEBUG validates the whole tree (weight order,
parent / child links, symbol and path tables) after every symbol and checks that all lanes of a warp sit in the same collective.
Built with -O0: at -O1 and above the host compiler duplicates code after lane-dependent branches, one collective then has two call
sites and the same-collective check reports a false positive.
    python tools/emu_debug_fgk.py"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "huffman-codec_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import backend  # noqa: E402
import hc_b200  # noqa: E402
import pyoracle  # noqa: E402

so = os.path.join(ROOT, "tests", "emu", "_build", "libhc_emu_dbg.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.run(["g++", "-std=c++17", "-O0", "-g", "-DHC_EMU", "-DHC_EMU_DEBUG", "-x", "c++", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-pthread",
                "-I", os.path.join(ROOT, "tests", "emu"), "-o", so, os.path.join(ROOT, "huffman-codec_b200", "csrc", "hc_api.cu"), "-ldl"], check=True)


class Dbg(backend.EmuBackend):
    def __init__(self):
        self.L = hc_b200.bind(so)
        self.stream = None


be, orc = Dbg(), pyoracle.Oracle()
rng = np.random.default_rng(3)
files = [rng.integers(0, 256, 6000, dtype=np.uint8), rng.integers(0, 16, 5000, dtype=np.uint8), (np.cumsum(rng.integers(-2, 3, 5000)) & 255).astype(np.uint8),
         np.repeat(rng.integers(0, 256, 400, dtype=np.uint8), rng.integers(1, 30, 400))[:5000].astype(np.uint8), rng.integers(0, 256, 20000, dtype=np.uint8)]
src = backend.Batch(be, [f.size for f in files], files)
enc = backend.Batch(be, [be.L.hc_fgk_bound(f.size) for f in files], fill=0xEE)
flags, st = be.upload(np.zeros(src.nf, np.uint8)), be.upload(np.zeros(src.nf, np.int32))
t = time.time()
rc = be.L.hc_fgk_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, flags.ptr, enc.data.ptr, enc.d_off.ptr, enc.d_cap.ptr, enc.d_len.ptr, st.ptr,
                              src.nf, be.stream)
assert rc == 0 and not be.download(st, src.nf * 4, np.int32).any(), "encode: status (102 = the validator found an inconsistent tree)"
outs = enc.files(enc.lens())
for f, g in zip(files, outs):
    bits, _ = orc.fgk_encode(f)
    assert np.array_equal(g, np.concatenate([np.frombuffer(int(f.size).to_bytes(8, "little") + b"\0", np.uint8), bits]))
src2 = backend.Batch(be, [o.size for o in outs], outs)
dst = backend.Batch(be, [f.size for f in files], fill=0xEE)
fl, st2 = be.upload(np.zeros(src.nf, np.uint8)), be.upload(np.zeros(src.nf, np.int32))
rc = be.L.hc_fgk_decode_batch(src2.data.ptr, src2.d_off.ptr, src2.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr, dst.d_len.ptr, fl.ptr, st2.ptr,
                              src.nf, be.stream)
assert rc == 0 and not be.download(st2, src.nf * 4, np.int32).any(), "decode: status"
for f, g in zip(files, dst.files(dst.lens())):
    assert np.array_equal(f, g)
print("FGK encode + decode under the debug emulator: %d streams, %d symbols, tree valid after every symbol, %.0f s" % (len(files), sum(f.size for f in files), time.time() - t))
