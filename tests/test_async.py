"""The asynchronous host API (hc_pipeline_*): jobs submitted back to back on a 2-deep pipeline give the
same bytes as the synchronous calls.  Runs on the SIMT emulator (CPU) and on the GPU."""
import ctypes as C

import numpy as np
import pytest

import hc_b200
import synth
from backend import CudaBackend, EmuBackend

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("name", BACKENDS)
def test_pipeline_matches_synchronous_calls(name, oracle):
    be = EmuBackend() if name == "emu" else CudaBackend()
    L = be.L
    n = 24 if name == "emu" else 64
    batches = [[synth.image(synth.CLASSES[(i + b) % 4], n, 100 * b + i, n + 8).reshape(-1) for i in range(5)] for b in range(4)]
    pipe = C.c_void_p()
    hc_b200.check(L.hc_pipeline_create(C.byref(pipe), 0, 2), "hc_pipeline_create", L)
    try:
        jobs = []
        for files in batches:
            buf, offs, lens = hc_b200.Codec.pack(files)
            widths = np.full(len(files), n, np.uint64)
            out = np.zeros(int(sum(L.hc_fgk_bound(int(L.hc_adapt_bound(n, n + 8))) + 16 for _ in files)), np.uint8)
            o_off, o_len, o_st = np.zeros(len(files), np.uint64), np.zeros(len(files), np.uint64), np.zeros(len(files), np.int32)
            t = L.hc_pipeline_submit_compress(pipe, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(files), 1, 1,
                                              widths.ctypes.data, out.ctypes.data, out.size, o_off.ctypes.data, o_len.ctypes.data,
                                              o_st.ctypes.data)
            assert t >= 0
            jobs.append((t, files, (buf, offs, lens, widths), out, o_off, o_len, o_st))
        djobs = []
        for t, files, keep, out, o_off, o_len, o_st in jobs:
            assert L.hc_pipeline_wait(pipe, t) == 0 and not o_st.any()
            for f, o, ln in zip(files, o_off, o_len):
                rc, exp = oracle.compress(f, diff=True, adapt=True, width=n)
                assert rc == 0 and np.array_equal(out[int(o):int(o) + int(ln)], exp)
            dec = np.zeros(sum(f.size + 16 for f in files) + 64, np.uint8)
            r_off, r_len, r_st = np.zeros(len(files), np.uint64), np.zeros(len(files), np.uint64), np.zeros(len(files), np.int32)
            t2 = L.hc_pipeline_submit_decompress(pipe, out.ctypes.data, o_off.ctypes.data, o_len.ctypes.data, len(files), dec.ctypes.data,
                                                 dec.size, r_off.ctypes.data, r_len.ctypes.data, r_st.ctypes.data)
            assert t2 >= 0
            djobs.append((t2, files, dec, r_off, r_len, r_st))
        for t2, files, dec, r_off, r_len, r_st in djobs:
            assert L.hc_pipeline_wait(pipe, t2) == 0 and not r_st.any()
            for f, o, ln in zip(files, r_off, r_len):
                assert np.array_equal(dec[int(o):int(o) + int(ln)], f)
        assert L.hc_pipeline_wait(pipe, 10 ** 6) != 0            # unknown ticket
        assert L.hc_pipeline_wait(pipe, djobs[0][0]) != 0        # a ticket is waited for once: no second answer, no hang
    finally:
        L.hc_pipeline_destroy(pipe)
