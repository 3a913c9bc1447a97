"""Test backends for the stage-level C ABI.

`cuda` : the product library libhc_b200.so with torch-allocated device memory (needs a GPU).
`emu`  : tests/emu/_build/libhc_emu.so -- the SAME kernel source compiled with -DHC_EMU and
         executed by the fiber-based SIMT emulator (tests/emu/hc_emu.h); "device"
         memory is host memory.  It exists so that kernel logic is checked on the CPU box
         before GPU minutes are spent; it is a test double, never a product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import hc_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu", "_build")
EMU_SO = os.path.join(EMU_DIR, "libhc_emu.so")
CSRC = os.path.join(ROOT, "huffman-codec_b200", "csrc")


def build_emu(force=False):
    emu_src = os.path.join(ROOT, "tests", "emu")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(emu_src, f) for f in os.listdir(emu_src) if f.endswith(".h")]
    if not force and os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(s) for s in srcs):
        return EMU_SO
    os.makedirs(EMU_DIR, exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-DHC_EMU", "-x", "c++", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-pthread", "-I", emu_src, "-o", EMU_SO, os.path.join(CSRC, "hc_api.cu")], check=True)
    return EMU_SO


class Buf:
    def __init__(self, ptr, nbytes, keep):
        self.ptr, self.nbytes, self.keep = ptr, nbytes, keep


class EmuBackend:
    name = "emu"

    def __init__(self):
        self.L = hc_b200.bind(build_emu())
        self.stream = None

    def alloc(self, nbytes, fill=0):
        a = np.full(int(nbytes) + 64, fill, dtype=np.uint8)
        return Buf(a.ctypes.data, int(nbytes), a)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        raw = arr.view(np.uint8).reshape(-1)
        b = self.alloc(raw.size)
        b.keep[:raw.size] = raw
        return b

    def download(self, buf, nbytes=None, dtype=np.uint8, offset=0):
        n = buf.nbytes - offset if nbytes is None else int(nbytes)
        return buf.keep[offset:offset + n].copy().view(dtype)

    def sync(self):
        pass


class CudaBackend:
    name = "cuda"

    def __init__(self):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required")
        self.L = hc_b200.lib()
        self.stream = torch.cuda.current_stream().cuda_stream

    def alloc(self, nbytes, fill=0):
        t = self.torch.full((int(nbytes) + 64,), fill, dtype=self.torch.uint8, device="cuda")
        return Buf(t.data_ptr(), int(nbytes), t)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        raw = arr.view(np.uint8).reshape(-1)
        b = self.alloc(raw.size)
        if raw.size:
            b.keep[:raw.size].copy_(self.torch.from_numpy(raw.copy()))
        return b

    def download(self, buf, nbytes=None, dtype=np.uint8, offset=0):
        n = buf.nbytes - offset if nbytes is None else int(nbytes)
        self.torch.cuda.synchronize()
        return buf.keep[offset:offset + n].cpu().numpy().copy().view(dtype)

    def sync(self):
        self.torch.cuda.synchronize()


class Batch:
    """A batch buffer in the layout the C ABI expects (256-byte aligned, padded regions)."""

    def __init__(self, be, caps, files=None, fill=0):
        self.be = be
        self.nf = len(caps)
        caps = [hc_b200.align_up(int(c) + 16) for c in caps]
        self.caps = np.array(caps, dtype=np.uint64)
        self.offs = np.zeros(self.nf, dtype=np.uint64)
        pos = 0
        for i, c in enumerate(caps):
            self.offs[i] = pos
            pos += c
        self.total = pos
        host = np.full(pos + 64, fill, dtype=np.uint8)
        lens = np.zeros(self.nf, dtype=np.uint64)
        if files is not None:
            for i, f in enumerate(files):
                f = np.asarray(f, dtype=np.uint8).reshape(-1)
                host[int(self.offs[i]):int(self.offs[i]) + f.size] = f
                lens[i] = f.size
        self.data = be.upload(host)
        self.d_off = be.upload(self.offs)
        self.d_cap = be.upload(self.caps)
        self.d_len = be.upload(lens)
        self.max_len = int(lens.max()) if self.nf else 0

    def lens(self):
        return self.be.download(self.d_len, self.nf * 8, np.uint64)

    def files(self, lens=None):
        lens = self.lens() if lens is None else lens
        host = self.be.download(self.data, self.total)
        return [host[int(o):int(o) + int(n)].copy() for o, n in zip(self.offs, lens)]
