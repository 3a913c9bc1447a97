"""Full-size parity cases of BASELINE configs 4 and 5 that only make sense on the GPU (`-m gpu`):
a 4096 x 4096 high-entropy image (one 16.8 M symbol FGK stream, block-size search over 8 candidates
up to 1024, reference src/transform.cpp:294-328) and a Fibonacci-weighted stream whose FGK tree grows
deeper than 32 levels, so that codes no longer fit one 32-bit word."""
import numpy as np
import pytest

import hc_b200
import synth
from backend import Batch, CudaBackend

pytestmark = pytest.mark.gpu


def fib_stream(ns=34):
    """symbol s repeated F(s) times in increasing order (the FGK tree degenerates into a chain of depth
    ns), then every symbol a few more times: the rare ones are sent with codes of up to ns bits."""
    f = [1, 1]
    while len(f) < ns:
        f.append(f[-1] + f[-2])
    tail = np.arange(ns, dtype=np.uint8)
    return np.concatenate([np.full(f[i], i, np.uint8) for i in range(ns)] + [tail, tail, tail, tail[::-1]])


def test_fgk_codes_beyond_32_bits(oracle):
    be = CudaBackend()
    deep = fib_stream(34)                                         # 14.9 M symbols, depth 34
    assert oracle.fgk_stats(deep)[2] > 32
    files = [deep, fib_stream(30)[::-1].copy()]
    src = Batch(be, [f.size for f in files], files)
    enc = Batch(be, [be.L.hc_fgk_bound(f.size) for f in files], fill=0xEE)
    flags = be.upload(np.zeros(src.nf, np.uint8))
    st = be.upload(np.zeros(src.nf, np.int32))
    assert be.L.hc_fgk_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, flags.ptr, enc.data.ptr, enc.d_off.ptr,
                                    enc.d_cap.ptr, enc.d_len.ptr, st.ptr, src.nf, be.stream) == 0
    assert not be.download(st, src.nf * 4, np.int32).any()
    lens = enc.lens()
    outs = enc.files(lens)
    for i, (f, g) in enumerate(zip(files, outs)):
        bits, _ = oracle.fgk_encode(f)
        exp = np.concatenate([np.frombuffer(int(f.size).to_bytes(8, "little") + b"\0", np.uint8), bits])
        assert int(lens[i]) == exp.size and np.array_equal(g, exp), i
    src2 = Batch(be, [o.size for o in outs], outs)
    dst = Batch(be, [f.size for f in files], fill=0xEE)
    fl = be.upload(np.zeros(src.nf, np.uint8))
    st2 = be.upload(np.zeros(src.nf, np.int32))
    assert be.L.hc_fgk_decode_batch(src2.data.ptr, src2.d_off.ptr, src2.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                    dst.d_len.ptr, fl.ptr, st2.ptr, src.nf, be.stream) == 0
    assert not be.download(st2, src.nf * 4, np.int32).any()
    for i, (f, g) in enumerate(zip(files, dst.files(dst.lens()))):
        assert np.array_equal(g, f), i
    # a stream cut inside a long code still fails with the reference's exit code 9
    cut = outs[0][: outs[0].size - 3].copy()
    src3 = Batch(be, [cut.size], [cut])
    dst3 = Batch(be, [files[0].size], fill=0xEE)
    st3 = be.upload(np.zeros(1, np.int32))
    assert be.L.hc_fgk_decode_batch(src3.data.ptr, src3.d_off.ptr, src3.d_len.ptr, dst3.data.ptr, dst3.d_off.ptr, dst3.d_cap.ptr,
                                    dst3.d_len.ptr, fl.ptr, st3.ptr, 1, be.stream) == 0
    assert list(be.download(st3, 4, np.int32)) == [9]


def test_random_image_4096(oracle):
    """BASELINE config 4, class `random` (seed 2000 + 2): 4096 x 4096, -m -a -w 4096."""
    cd = hc_b200.Codec(0)
    img = synth.image("random", 4096, 2002).reshape(-1)
    outs, st = cd.compress([img], diff=True, adapt=True, width=4096)
    assert st[0] == 0
    rc, exp = oracle.compress(img, diff=True, adapt=True, width=4096, mode=1)
    assert rc == 0 and outs[0].size == exp.size and np.array_equal(outs[0], exp)
    back, st = cd.decompress(outs)
    assert st[0] == 0 and np.array_equal(back[0], img)


def test_rle_stage_multi_mib(oracle):
    """Plain MNP-5 (`-m` without `-a`) at the file sizes of BASELINE config 4: one CTA streams hundreds of 16 KiB tiles;
    the decoder of the flat images needs 64 output windows per token tile."""
    be = CudaBackend()
    rng = np.random.default_rng(5)
    n = 4 << 20
    flat = np.full(n, 9, np.uint8)
    flat[rng.integers(0, n, 40)] = 200                                # a few breaks inside runs of many times 258
    files = [oracle.diff_apply(synth.image(k, 2048, 70 + i).reshape(-1)) for i, k in enumerate(("walk", "smooth", "random", "longrun"))]
    files += [flat, np.repeat(rng.integers(0, 4, n // 3, dtype=np.uint8), 3)[: n - 5]]
    src = Batch(be, [f.size for f in files], files)
    enc = Batch(be, [be.L.hc_rle_bound(f.size) for f in files], fill=0xEE)
    assert be.L.hc_rle_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, enc.data.ptr, enc.d_off.ptr, enc.d_len.ptr,
                                    src.nf, src.max_len, be.stream) == 0
    elens = enc.lens()
    outs = enc.files(elens)
    for i, (f, g) in enumerate(zip(files, outs)):
        exp = oracle.rle_encode(f)
        assert int(elens[i]) == exp.size and np.array_equal(g, exp), i
    tsrc = Batch(be, [o.size for o in outs], outs)
    dst = Batch(be, [f.size for f in files], fill=0xEE)
    st = be.upload(np.zeros(src.nf, np.int32))
    assert be.L.hc_rle_decode_batch(tsrc.data.ptr, tsrc.d_off.ptr, tsrc.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                    dst.d_len.ptr, st.ptr, src.nf, tsrc.max_len, be.stream) == 0
    assert not be.download(st, src.nf * 4, np.int32).any()
    for i, (f, g) in enumerate(zip(files, dst.files(dst.lens()))):
        assert np.array_equal(g, f), i
