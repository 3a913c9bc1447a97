"""The C-ABI library loads and exports every symbol include/hc_b200.h declares (no GPU needed),
and its host-only helpers agree with the oracle."""
import os
import re

import numpy as np

import hc_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(hc_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = hc_b200.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
        assert n in hc_b200.SIGNATURES, "binding missing for " + n
    assert b"sm_100a" in L.hc_version()


def test_no_cpu_fallback_without_device():
    import torch
    L = hc_b200.lib()
    if torch.cuda.is_available():
        return
    # without a CUDA device the product must fail loudly, never compute on the CPU
    assert L.hc_device_count() <= 0
    import ctypes as C
    h = C.c_void_p()
    assert L.hc_codec_create(C.byref(h), 0) != 0


def test_bounds_and_geometry(oracle):
    L = hc_b200.lib()
    for w, h, b in ((512, 512, 8), (512, 517, 64), (9, 8, 8), (4096, 4096, 1024), (17, 33, 16)):
        assert L.hc_block_count(w, h, b) == oracle.block_count(w, h, b)
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 3, 4, 100, 999):
        worst = np.repeat(rng.integers(0, 256, n // 3 + 1), 3)[:n].astype(np.uint8)   # runs of exactly 3: 4/3 expansion
        assert oracle.rle_encode(worst).size <= L.hc_rle_bound(n)
