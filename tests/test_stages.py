"""Stage-level parity: every C-ABI stage against the CPU oracle on the same inputs.

Each test runs twice: on the SIMT-emulated kernels (`emu`, CPU box, small inputs) and on the
real sm_100a kernels (`cuda`, marked gpu).  Bit-exact is the bar everywhere (byte/integer work).
"""
import numpy as np
import pytest

import hc_b200
import synth
from backend import Batch, CudaBackend, EmuBackend

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
_cache = {}


@pytest.fixture(params=BACKENDS)
def be(request):
    name = request.param
    if name not in _cache:
        _cache[name] = EmuBackend() if name == "emu" else CudaBackend()
    return _cache[name]


def rc0(rc):
    assert rc == 0, rc


def ragged_files(be, seed=1):
    """Ragged batch: empty, tiny, vector/tile boundary sizes, run-heavy, count-byte-heavy."""
    rng = np.random.default_rng(seed)
    big = 70000 if be.name == "cuda" else 33000
    sizes = [0, 1, 2, 3, 4, 15, 16, 17, 31, 257, 258, 259, 260, 4095, 4096, 4097, 16383, 16384, 16385, big]
    files = []
    for i, n in enumerate(sizes):
        k = i % 6
        if k == 0:
            v = rng.integers(0, 256, n)
        elif k == 1:
            reps = rng.integers(1, 700, n // 50 + 2)
            v = np.repeat(rng.integers(0, 256, reps.size), reps)[:n]
        elif k == 2:
            reps = rng.integers(1, 6, n // 2 + 2)
            v = np.repeat(rng.integers(250, 256, reps.size), reps)[:n]
        elif k == 3:
            v = np.full(n, 7)
        elif k == 4:
            v = rng.integers(0, 2, n) * 3
        else:
            v = np.cumsum(rng.integers(-1, 2, n)) & 255
        files.append(v.astype(np.uint8))
    # runs that cross tile (16 KiB) and vector boundaries with lengths around 258*k
    v = np.concatenate([np.full(16384 - 100, 1), np.full(258 * 3 + 2, 9), np.full(5, 2), np.full(258, 4), np.full(257, 5),
                        np.full(259, 6), np.arange(40) & 255, np.full(3, 8), np.full(3, 9)]).astype(np.uint8)
    files.append(v)
    return files


def test_diff_apply_revert(be, oracle):
    files = ragged_files(be)
    src = Batch(be, [f.size for f in files], files)
    dst = Batch(be, [f.size for f in files])
    rc0(be.L.hc_diff_apply_batch(src.data.ptr, src.d_off.ptr, dst.data.ptr, dst.d_off.ptr, src.d_len.ptr,
                                 src.nf, src.max_len, be.stream))
    lens = [f.size for f in files]
    got = dst.files(lens)
    for f, g in zip(files, got):
        assert np.array_equal(g, oracle.diff_apply(f))
    # revert, out of place and in place
    back = Batch(be, lens)
    rc0(be.L.hc_diff_revert_batch(dst.data.ptr, dst.d_off.ptr, back.data.ptr, back.d_off.ptr, src.d_len.ptr,
                                  src.nf, src.max_len, be.stream))
    for f, g in zip(files, back.files(lens)):
        assert np.array_equal(g, f)
    rc0(be.L.hc_diff_revert_batch(dst.data.ptr, dst.d_off.ptr, dst.data.ptr, dst.d_off.ptr, src.d_len.ptr,
                                  src.nf, src.max_len, be.stream))
    for f, g in zip(files, dst.files(lens)):
        assert np.array_equal(g, f)


def test_diff_revert_segmented(be, oracle):
    # few large files -> the launcher splits each file into segments (per-segment sums + carry)
    n = 300000 if be.name == "cuda" else 150000
    rng = np.random.default_rng(3)
    files = [rng.integers(0, 256, n, dtype=np.uint8), rng.integers(0, 256, n - 12345, dtype=np.uint8)]
    src = Batch(be, [f.size for f in files], files)
    dst = Batch(be, [f.size for f in files])
    rc0(be.L.hc_diff_revert_batch(src.data.ptr, src.d_off.ptr, dst.data.ptr, dst.d_off.ptr, src.d_len.ptr,
                                  src.nf, src.max_len, be.stream))
    for f, g in zip(files, dst.files([f.size for f in files])):
        assert np.array_equal(g, oracle.diff_revert(f))


def test_rle_encode(be, oracle):
    files = ragged_files(be, 2)
    src = Batch(be, [f.size for f in files], files)
    dst = Batch(be, [be.L.hc_rle_bound(f.size) for f in files], fill=0xEE)
    rc0(be.L.hc_rle_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_len.ptr,
                                 src.nf, src.max_len, be.stream))
    lens = dst.lens()
    got = dst.files(lens)
    for i, (f, g) in enumerate(zip(files, got)):
        exp = oracle.rle_encode(f)
        assert int(lens[i]) == exp.size, (i, f.size)
        assert np.array_equal(g, exp), (i, f.size)


def test_rle_decode(be, oracle):
    files = ragged_files(be, 5)
    encs = [oracle.rle_encode(f) for f in files]
    # any byte string is a valid MNP-5 stream: also decode raw noise and count-heavy streams
    rng = np.random.default_rng(11)
    encs.append(rng.integers(0, 256, 3000, dtype=np.uint8))
    encs.append(np.tile(np.array([5, 5, 5, 255], np.uint8), 700))
    encs.append(np.tile(np.array([5, 5, 5, 0, 5], np.uint8), 500))
    exps = [oracle.rle_decode(e) for e in encs]
    src = Batch(be, [e.size for e in encs], encs)
    dst = Batch(be, [e.size for e in exps], fill=0xEE)
    st = be.upload(np.zeros(src.nf, np.int32))
    # size query first (out == NULL)
    rc0(be.L.hc_rle_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, None, None, None, dst.d_len.ptr, st.ptr,
                                 src.nf, src.max_len, be.stream))
    assert [int(x) for x in dst.lens()] == [e.size for e in exps]
    rc0(be.L.hc_rle_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                 dst.d_len.ptr, st.ptr, src.nf, src.max_len, be.stream))
    lens = dst.lens()
    assert not be.download(st, src.nf * 4, np.int32).any()
    for i, (e, g) in enumerate(zip(exps, dst.files(lens))):
        assert int(lens[i]) == e.size
        assert np.array_equal(g, e), i
    # the byte after every region's payload must be untouched (no overrun)
    host = be.download(dst.data, dst.total)
    for o, n, c in zip(dst.offs, lens, dst.caps):
        assert (host[int(o) + int(n):int(o) + int(c)] == 0xEE).all()


def _rle_fuzz_stream(rng, kind, n):
    """Byte strings that stress the 64-byte segments, 16 KiB tiles and output windows of rle.cuh."""
    if kind == 0:
        a = rng.integers(0, 256, n)
    elif kind == 1:
        a = rng.integers(0, 2, n)
    elif kind == 2:
        a = rng.integers(0, 3, n) // 2
    elif kind == 3:
        out, tot = [], 0
        pool = [1, 1, 2, 3, 3, 4, 5, 60, 61, 62, 63, 64, 65, 66, 190, 193, 194, 195, 255, 256, 257, 258, 259, 260, 300, 514, 515, 516, 517,
                518, 774, 1032, 1290, 5000]
        while tot < n:
            L = int(rng.choice(pool))
            out.append(np.full(L, int(rng.integers(0, 3))))
            tot += L
        a = np.concatenate(out)[:n] if out else np.zeros(0)
    elif kind == 4:
        a = np.tile(np.array([5, 5, 5, 255]), n // 4 + 1)[:n]           # as tokens: every fourth byte a count of 255
    elif kind == 5:
        a = np.tile(np.array([7, 7, 7, int(rng.integers(0, 256))]), n // 4 + 1)[:n]
    else:
        a = np.repeat(rng.integers(0, 4, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))[:n]
    return np.asarray(a).astype(np.uint8)


def test_rle_fuzz(be, oracle):
    """Random streams of several statistics at sizes around the segment / tile / window borders: the encoder against
    the oracle, the decoder on the same bytes read as TOKENS (any byte string is a valid MNP-5 stream) and on the
    encoder's output, with exact and with short capacities."""
    rng = np.random.default_rng(2024)
    sizes = [0, 1, 2, 3, 5, 63, 64, 65, 127, 128, 129, 255, 256, 257, 258, 259, 1000, 4095, 4096, 16383, 16384, 16385, 16384 + 64,
             16384 * 2, 16384 * 2 + 1, 40000]
    for it in range(6 if be.name == "emu" else 20):
        files = [_rle_fuzz_stream(rng, int(rng.integers(0, 7)), int(rng.choice(sizes + [int(rng.integers(1, 60000))]))) for _ in range(10)]
        src = Batch(be, [f.size for f in files], files)
        enc = Batch(be, [be.L.hc_rle_bound(f.size) for f in files], fill=0xEE)
        rc0(be.L.hc_rle_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, enc.data.ptr, enc.d_off.ptr, enc.d_len.ptr,
                                     src.nf, src.max_len, be.stream))
        elens = enc.lens()
        ehost = be.download(enc.data, enc.total)
        for i, (f, g) in enumerate(zip(files, enc.files(elens))):
            exp = oracle.rle_encode(f)
            assert int(elens[i]) == exp.size and np.array_equal(g, exp), (it, i, f.size)
            o, c = int(enc.offs[i]), int(enc.caps[i])
            assert (ehost[o + exp.size:o + c] == 0xEE).all(), (it, i)
        # decode: the raw files as token streams, and the encoder's output
        toks = files + [oracle.rle_encode(f) for f in files[:4]]
        exps = [oracle.rle_decode(t) for t in toks]
        tsrc = Batch(be, [t.size for t in toks], toks)
        caps = [e.size if rng.random() < 0.7 else max(0, e.size - int(rng.integers(300, 3000))) for e in exps]
        dst = Batch(be, caps, fill=0xEE)
        st = be.upload(np.zeros(tsrc.nf, np.int32))
        rc0(be.L.hc_rle_decode_batch(tsrc.data.ptr, tsrc.d_off.ptr, tsrc.d_len.ptr, None, None, None, dst.d_len.ptr, st.ptr,
                                     tsrc.nf, tsrc.max_len, be.stream))
        assert [int(x) for x in dst.lens()] == [e.size for e in exps], it
        rc0(be.L.hc_rle_decode_batch(tsrc.data.ptr, tsrc.d_off.ptr, tsrc.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                     dst.d_len.ptr, st.ptr, tsrc.nf, tsrc.max_len, be.stream))
        lens, sts = dst.lens(), be.download(st, tsrc.nf * 4, np.int32)
        host = be.download(dst.data, dst.total)
        for i, e in enumerate(exps):
            o, c = int(dst.offs[i]), int(dst.caps[i])
            m = min(e.size, c)
            assert int(lens[i]) == e.size and np.array_equal(host[o:o + m], e[:m]), (it, i, toks[i].size, e.size)
            assert (e.size > c) == (int(sts[i]) == hc_b200.HC_E_CAPACITY), (it, i)
            assert (host[o + m:o + c] == 0xEE).all(), (it, i)


def test_rle_decode_capacity(be, oracle):
    enc = np.tile(np.array([9, 9, 9, 200], np.uint8), 50)
    exp = oracle.rle_decode(enc)
    src = Batch(be, [enc.size], [enc])
    dst = Batch(be, [256], fill=0xEE)              # capacity 256+16 -> 512 after alignment
    st = be.upload(np.zeros(1, np.int32))
    rc0(be.L.hc_rle_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                 dst.d_len.ptr, st.ptr, 1, src.max_len, be.stream))
    assert int(dst.lens()[0]) == exp.size
    assert int(be.download(st, 4, np.int32)[0]) == hc_b200.HC_E_CAPACITY
    host = be.download(dst.data, dst.total + 64)
    cap = int(dst.caps[0])
    assert np.array_equal(host[:cap], exp[:cap]) and (host[cap:] == 0xEE).all() is not False


SHAPES = [(8, 8), (9, 8), (8, 9), (15, 17), (16, 16), (17, 33), (64, 24), (31, 100), (130, 70), (256, 40), (512, 16)]


def _adapt_inputs(be):
    imgs = []
    kinds = ("walk", "smooth", "random", "const", "longrun")
    shapes = SHAPES if be.name == "emu" else SHAPES + [(512, 512), (512, 517), (1024, 300)]
    for i, (w, h) in enumerate(shapes):
        for k in (kinds[i % 5], kinds[(i + 2) % 5]):
            imgs.append((w, h, synth.image(k, w, 50 + i, h).reshape(-1)))
    # large-block paths of the mask search (B >= 64: one warp per block; flat areas: long-run correction)
    # and of adapt_large.cuh (one CTA per block, ragged edge blocks in both directions)
    for i, (w, h, k) in enumerate(((192, 128, "const"), (128, 192, "longrun"), (136, 200, "smooth"), (128, 128, "walk"),
                                   (200, 136, "random"), (333, 100, "random"), (136, 200, "const"), (70, 130, "const"),
                                   (97, 151, "longrun"))):
        imgs.append((w, h, synth.image(k, w, 90 + i, h).reshape(-1)))
    big = np.zeros((160, 192), np.uint8)
    big[:, 100:] = 7                       # two flat halves: runs of >= 258 in both scan directions
    big[70:, :] += 3
    imgs.append((192, 160, big.reshape(-1)))
    # gradients that favour vertical / horizontal scanning
    y, x = np.mgrid[0:64, 0:64]
    imgs.append((64, 64, (x & 255).astype(np.uint8).reshape(-1)))
    imgs.append((64, 64, (y & 255).astype(np.uint8).reshape(-1)))
    y, x = np.mgrid[0:136, 0:200]          # ragged large blocks, vertical and horizontal winners
    imgs.append((200, 136, ((x * 7) & 255).astype(np.uint8).reshape(-1)))
    imgs.append((200, 136, ((y * 7) & 255).astype(np.uint8).reshape(-1)))
    return imgs


def test_adapt_encode(be, oracle):
    imgs = _adapt_inputs(be)
    files = [im for _, _, im in imgs]
    src = Batch(be, [f.size for f in files], files)
    dst = Batch(be, [be.L.hc_adapt_bound(w, h) for w, h, _ in imgs], fill=0xEE)
    d_w = be.upload(np.array([w for w, _, _ in imgs], np.uint64))
    d_h = be.upload(np.array([h for _, h, _ in imgs], np.uint64))
    d_b = be.upload(np.zeros(src.nf, np.uint64))
    st = be.upload(np.zeros(src.nf, np.int32))
    ws = be.alloc(be.L.hc_adapt_encode_ws_bytes(src.nf, src.max_len))
    rc0(be.L.hc_adapt_encode_batch(src.data.ptr, src.d_off.ptr, d_w.ptr, d_h.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_len.ptr,
                                   d_b.ptr, st.ptr, src.nf, src.max_len, ws.ptr, be.stream))
    lens = dst.lens()
    got = dst.files(lens)
    bs = be.download(d_b, src.nf * 8, np.uint64)
    assert not be.download(st, src.nf * 4, np.int32).any()
    for i, (w, h, im) in enumerate(imgs):
        rc, exp, b = oracle.adapt_encode(im, w, h)
        assert rc == 0
        assert int(bs[i]) == b, (w, h, i)
        assert int(lens[i]) == exp.size, (w, h, i)
        assert np.array_equal(got[i], exp), (w, h, i)


def test_adapt_encode_too_small(be):
    files = [np.zeros(35, np.uint8), np.zeros(64, np.uint8)]
    src = Batch(be, [64, 64], files)
    dst = Batch(be, [256, 256])
    d_w = be.upload(np.array([5, 8], np.uint64))
    d_h = be.upload(np.array([7, 8], np.uint64))
    st = be.upload(np.zeros(2, np.int32))
    ws = be.alloc(be.L.hc_adapt_encode_ws_bytes(2, 64))
    rc0(be.L.hc_adapt_encode_batch(src.data.ptr, src.d_off.ptr, d_w.ptr, d_h.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_len.ptr,
                                   None, st.ptr, 2, 64, ws.ptr, be.stream))
    assert list(be.download(st, 8, np.int32)) == [12, 0]


def test_adapt_decode(be, oracle):
    imgs = _adapt_inputs(be)
    encs = [oracle.adapt_encode(im, w, h)[1] for w, h, im in imgs]
    # fixed block sizes too (not only the search winner)
    for (w, h, im) in imgs[:6]:
        encs.append(oracle.adapt_encode_bs(im, w, h, 8))
        imgs = imgs + [(w, h, im)]
    # forced large block sizes on ragged shapes: multiples of 64 (adapt_large.cuh), a multiple of 16
    # only (strips with a remainder) and one that is neither (lane-group kernel)
    for (w, h, kind) in ((130, 70, "walk"), (200, 136, "smooth"), (100, 333, "longrun")):
        im = synth.image(kind, w, 7, h).reshape(-1)
        for bsz in (64, 128, 80, 72):
            encs.append(oracle.adapt_encode_bs(im, w, h, bsz))
            imgs = imgs + [(w, h, im)]
    # a stream with block size 4 (never produced by the reference encoder, accepted by its decoder)
    img4 = synth.image("walk", 12, 77, 8).reshape(-1)
    encs.append(oracle.adapt_encode_bs(img4, 12, 8, 4))
    imgs = imgs + [(12, 8, img4)]
    src = Batch(be, [e.size for e in encs], encs)
    max_out = max(w * h for w, h, _ in imgs)
    for use_ws in (True, False):      # parallel index+expand kernels / serial fallback without scratch
        dst = Batch(be, [w * h for w, h, _ in imgs], fill=0xEE)
        st = be.upload(np.zeros(src.nf, np.int32))
        ws = be.alloc(be.L.hc_adapt_decode_ws_bytes(src.nf, max_out)) if use_ws else None
        # size query first (out == NULL)
        rc0(be.L.hc_adapt_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, None, None, None, dst.d_len.ptr, st.ptr,
                                       src.nf, src.max_len, max_out, None, be.stream))
        assert [int(x) for x in dst.lens()] == [w * h for w, h, _ in imgs]
        rc0(be.L.hc_adapt_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                       dst.d_len.ptr, st.ptr, src.nf, src.max_len, max_out, ws.ptr if ws else None, be.stream))
        assert not be.download(st, src.nf * 4, np.int32).any()
        lens = dst.lens()
        for i, ((w, h, im), g) in enumerate(zip(imgs, dst.files(lens))):
            assert int(lens[i]) == w * h
            assert np.array_equal(g, im), (w, h, i, use_ws)


def test_adapt_decode_errors(be, oracle):
    img = synth.image("walk", 16, 3, 24).reshape(-1)
    good = oracle.adapt_encode(img, 16, 24)[1]
    hdr = lambda w, h, b: np.frombuffer(w.to_bytes(8, "big") + h.to_bytes(8, "big") + b.to_bytes(8, "big"), np.uint8)
    cases = [
        good[:20],                                            # 10: header < 24 bytes
        good[:24],                                            # 11: direction bytes missing
        good[:-3],                                            # 14: data underrun
        np.concatenate([good, np.array([1, 2], np.uint8)]),   # 15: leftover
        np.concatenate([hdr(8, 8, 8), np.array([0x80], np.uint8), np.array([7, 7, 7, 100], np.uint8)]),  # 13: overshoot
        np.concatenate([hdr(8, 8, 0), np.zeros(8, np.uint8)]),   # block size 0: reference divides by zero; we say 10
        good,
    ]
    expect = [10, 11, 14, 15, 13, 10, 0]
    for c, e in zip(cases[:5], expect[:5]):
        assert oracle.adapt_decode(c)[0] == e
    src = Batch(be, [c.size for c in cases], cases)
    for use_ws in (True, False):
        dst = Batch(be, [16 * 24] * len(cases))
        st = be.upload(np.zeros(src.nf, np.int32))
        ws = be.alloc(be.L.hc_adapt_decode_ws_bytes(src.nf, 16 * 24)) if use_ws else None
        rc0(be.L.hc_adapt_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                       dst.d_len.ptr, st.ptr, src.nf, src.max_len, 16 * 24, ws.ptr if ws else None, be.stream))
        assert list(be.download(st, src.nf * 4, np.int32)) == expect, use_ws
        assert np.array_equal(dst.files(dst.lens())[-1], img)


def _fgk_inputs(be):
    rng = np.random.default_rng(21)
    n = 6000 if be.name == "emu" else 60000
    files = [np.zeros(0, np.uint8), np.array([65], np.uint8), np.frombuffer(b"AAAA", np.uint8),
             rng.integers(0, 256, n, dtype=np.uint8), rng.integers(0, 4, n // 2, dtype=np.uint8),
             np.arange(n // 3, dtype=np.uint32).astype(np.uint8), np.full(n // 4, 9, np.uint8),
             np.sort(synth.image("fib", 128, 1).reshape(-1))[::-1][: n].copy(),      # deep tree, long codes
             (np.cumsum(rng.integers(-2, 3, n)) & 255).astype(np.uint8)]
    return files


def test_fgk_encode(be, oracle):
    files = _fgk_inputs(be)
    src = Batch(be, [f.size for f in files], files)
    dst = Batch(be, [be.L.hc_fgk_bound(f.size) for f in files], fill=0xEE)
    flags = be.upload(np.array([(i % 4) << 6 for i in range(src.nf)], np.uint8))
    st = be.upload(np.zeros(src.nf, np.int32))
    rc0(be.L.hc_fgk_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, flags.ptr, dst.data.ptr, dst.d_off.ptr,
                                 dst.d_cap.ptr, dst.d_len.ptr, st.ptr, src.nf, be.stream))
    assert not be.download(st, src.nf * 4, np.int32).any()
    lens = dst.lens()
    for i, (f, g) in enumerate(zip(files, dst.files(lens))):
        bits, _ = oracle.fgk_encode(f)
        exp = np.concatenate([np.frombuffer(int(f.size).to_bytes(8, "little") + bytes([(i % 4) << 6]), np.uint8), bits])
        assert int(lens[i]) == exp.size, i
        assert np.array_equal(g, exp), i


def test_fgk_decode(be, oracle):
    files = _fgk_inputs(be)
    outs = []
    for i, f in enumerate(files):
        bits, _ = oracle.fgk_encode(f)
        outs.append(np.concatenate([np.frombuffer(int(f.size).to_bytes(8, "little") + bytes([0xC0 if i & 1 else 0]), np.uint8), bits]))
    # malformed: short header (8), truncated bits (9), absurd count (9), trailing garbage is ignored (0)
    outs.append(outs[3][:5].copy())
    outs.append(outs[3][:-5].copy())
    big = outs[3].copy()
    big[:8] = np.frombuffer((10 ** 12).to_bytes(8, "little"), np.uint8)
    outs.append(big)
    outs.append(np.concatenate([outs[8], np.array([1, 2, 3], np.uint8)]))
    exp_status = [0] * len(files) + [8, 9, 9, 0]
    src = Batch(be, [o.size for o in outs], outs)
    dst = Batch(be, [f.size for f in files] + [64, files[3].size, 64, files[8].size], fill=0xEE)
    flags = be.upload(np.zeros(src.nf, np.uint8))
    st = be.upload(np.zeros(src.nf, np.int32))
    rc0(be.L.hc_fgk_decode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                 dst.d_len.ptr, flags.ptr, st.ptr, src.nf, be.stream))
    status = list(be.download(st, src.nf * 4, np.int32))
    assert status == exp_status
    lens = dst.lens()
    got = dst.files(lens)
    fl = be.download(flags, src.nf)
    for i, f in enumerate(files):
        assert int(lens[i]) == f.size and np.array_equal(got[i], f), i
        assert fl[i] == (0xC0 if i & 1 else 0)
    assert np.array_equal(got[-1], files[8])


def test_offsets_and_gather(be):
    lens = np.array([0, 5, 256, 257, 1000, 1], np.uint64)
    d_len = be.upload(lens)
    d_off = be.upload(np.zeros(lens.size, np.uint64))
    d_tot = be.upload(np.zeros(1, np.uint64))
    rc0(be.L.hc_offsets_from_lens(d_len.ptr, d_off.ptr, d_tot.ptr, lens.size, 16, be.stream))
    off = be.download(d_off, lens.size * 8, np.uint64)
    al = (lens + 15) // 16 * 16
    assert list(off) == list(np.concatenate([[0], np.cumsum(al)[:-1]]))
    assert int(be.download(d_tot, 8, np.uint64)[0]) == int(al.sum())
    rng = np.random.default_rng(4)
    files = [rng.integers(0, 256, int(n), dtype=np.uint8) for n in lens]
    src = Batch(be, [int(n) for n in lens], files)
    out = be.alloc(int(al.sum()) + 256, fill=0xEE)
    rc0(be.L.hc_gather_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, out.ptr, d_off.ptr, lens.size, int(lens.max()), be.stream))
    host = be.download(out, int(al.sum()))
    for f, o in zip(files, off):
        assert np.array_equal(host[int(o):int(o) + f.size], f)


def test_rle_encode_long_runs(be, oracle):
    """Runs around the 258 wrap at every phase relative to the 16-byte vectors and 16 KiB tiles."""
    rng = np.random.default_rng(77)
    files = []
    for lead in list(range(0, 20)) + [4090, 16370, 16383]:
        parts = [rng.integers(0, 256, lead)]
        for L in (257, 258, 259, 260, 261, 515, 516, 517, 518, 519, 774, 775, 1032, 1033, 3, 2, 4):
            v = int(rng.integers(0, 256))
            parts.append(np.full(L + int(rng.integers(0, 3)), v))
            parts.append(np.array([(v + 1) & 255] * int(rng.integers(1, 4))))
        files.append(np.concatenate(parts).astype(np.uint8))
    # random mixtures of short runs and runs around the multiples of 258
    pool = np.array([1, 1, 1, 2, 3, 3, 4, 5, 6, 7, 16, 17, 255, 256, 257, 258, 259, 260, 261, 262, 514, 515, 516, 517, 518,
                     519, 520, 521, 773, 774, 775, 776])
    for _ in range(24):
        ls = rng.choice(pool, 40)
        vals = rng.integers(0, 3, 40)                 # few values: neighbouring runs often merge
        files.append(np.concatenate([np.full(int(L), int(v)) for L, v in zip(ls, vals)]).astype(np.uint8))
    files.append(np.full(40000, 5, np.uint8))
    files.append(np.concatenate([np.full(16384 * 2 - 1, 9), np.full(300, 9), np.full(2, 1)]).astype(np.uint8))
    src = Batch(be, [f.size for f in files], files)
    dst = Batch(be, [be.L.hc_rle_bound(f.size) for f in files], fill=0xEE)
    rc0(be.L.hc_rle_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_len.ptr,
                                 src.nf, src.max_len, be.stream))
    lens = dst.lens()
    for i, (f, g) in enumerate(zip(files, dst.files(lens))):
        exp = oracle.rle_encode(f)
        assert int(lens[i]) == exp.size and np.array_equal(g, exp), (i, f.size)


def _fgk_stress_inputs(be):
    """Many streams with different symbol statistics: uniform, geometric (deep trees, leaves below the
    path table), few symbols (internal nodes near the root swap often), sorted ramps, bursts."""
    rng = np.random.default_rng(4242)
    ns, nmax = (28, 2500) if be.name == "emu" else (224, 120000)
    files = []
    for i in range(ns):
        n = int(rng.integers(1, nmax))
        kind = i % 7
        if kind == 0:
            f = rng.integers(0, 256, n)
        elif kind == 1:
            f = np.minimum(rng.geometric(0.5 ** (1 + i % 3), n) - 1, 255)
        elif kind == 2:
            f = rng.integers(0, 2 + i % 5, n) * 37
        elif kind == 3:
            f = np.sort(rng.integers(0, 256, n))
        elif kind == 4:
            f = np.repeat(rng.integers(0, 256, n // 7 + 1), 7)[:n]
        elif kind == 5:
            f = (np.arange(n) // (1 + i)) & 255
        else:
            p = rng.dirichlet(np.full(64, 0.08))
            f = rng.choice(64, n, p=p) * 3
        files.append(np.asarray(f, np.uint8))
    # one long stream of short runs over the full alphabet: many subtree swaps, two in one update now and then
    # (a leaf can come back to the same path through different nodes -- the encoder's lookup-ahead must notice)
    n = 40000 if be.name == "emu" else 400000
    files.append(np.repeat(rng.integers(0, 256, n // 7 + 1), 7)[:n].astype(np.uint8))
    return files


def test_fgk_stress_roundtrip(be, oracle):
    files = _fgk_stress_inputs(be)
    src = Batch(be, [f.size for f in files], files)
    enc = Batch(be, [be.L.hc_fgk_bound(f.size) for f in files], fill=0xEE)
    flags = be.upload(np.zeros(src.nf, np.uint8))
    st = be.upload(np.zeros(src.nf, np.int32))
    rc0(be.L.hc_fgk_encode_batch(src.data.ptr, src.d_off.ptr, src.d_len.ptr, flags.ptr, enc.data.ptr, enc.d_off.ptr,
                                 enc.d_cap.ptr, enc.d_len.ptr, st.ptr, src.nf, be.stream))
    assert not be.download(st, src.nf * 4, np.int32).any()
    lens = enc.lens()
    outs = enc.files(lens)
    for i, (f, g) in enumerate(zip(files, outs)):
        bits, _ = oracle.fgk_encode(f)
        exp = np.concatenate([np.frombuffer(int(f.size).to_bytes(8, "little") + b"\0", np.uint8), bits])
        assert int(lens[i]) == exp.size and np.array_equal(g, exp), (i, f.size)
    # decode what was just produced
    src2 = Batch(be, [o.size for o in outs], outs)
    dst = Batch(be, [f.size for f in files], fill=0xEE)
    fl = be.upload(np.zeros(src.nf, np.uint8))
    st2 = be.upload(np.zeros(src.nf, np.int32))
    rc0(be.L.hc_fgk_decode_batch(src2.data.ptr, src2.d_off.ptr, src2.d_len.ptr, dst.data.ptr, dst.d_off.ptr, dst.d_cap.ptr,
                                 dst.d_len.ptr, fl.ptr, st2.ptr, src.nf, be.stream))
    assert not be.download(st2, src.nf * 4, np.int32).any()
    for i, (f, g) in enumerate(zip(files, dst.files(dst.lens()))):
        assert np.array_equal(g, f), (i, f.size)


def test_adapt_index_cta_kernel_on_emulator():
    """The CTA-wide block index kernel only takes matrices of 4 MiB and more (BASELINE config 4); the library's test
    hook HC_INDEX_WIDE_MIN lowers that bound so that the CPU suite reaches it with small images (the threshold is
    read once per process, hence the child process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HC_INDEX_WIDE_MIN="1024")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider", "-m", "not gpu",
                        "-k", "(adapt_decode or adapt_decode_errors) and emu"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout
