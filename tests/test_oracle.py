"""CPU tests: pin the oracle (oracle/hc_oracle.c) against the reference.

(1) golden vectors generated from the unmodified reference binary (tests/golden/),
(2) SURVEY A.7 hand vectors, (3) live differential fuzzing against oracle/_ref.
"""
import hashlib

import numpy as np
import pytest

import synth

MODES = {"plain": (False, False), "m": (True, False), "a": (False, True), "ma": (True, True)}


def _sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.mark.parametrize("mode", list(MODES))
def test_samples_match_reference_golden(oracle, golden, samples, mode):
    diff, adapt = MODES[mode]
    for name, data in samples.items():
        g = golden["samples"][name]
        assert _sha(data) == g["in_sha256"]
        rc, out = oracle.compress(data, diff=diff, adapt=adapt, width=512, mode=1)
        assert rc == 0
        assert len(out) == g[mode]["size"], (name, mode)
        assert _sha(out) == g[mode]["sha256"], (name, mode)
        rc, back = oracle.decompress(out)
        assert rc == 0 and np.array_equal(back, data)


def test_faithful_tree_equals_array_tree_on_samples(oracle, samples):
    # mode 0 = pointer tree with recursive findSuccNode; mode 1 = number-indexed arrays
    for name in ("hd01", "df1hvx", "hd09"):
        sym = oracle.rle_encode(oracle.diff_apply(samples[name]))[:60000]
        b0, n0 = oracle.fgk_encode(sym, mode=0)
        b1, n1 = oracle.fgk_encode(sym, mode=1)
        assert n0 == n1 and np.array_equal(b0, b1)
        rc, s0 = oracle.fgk_decode(b0, len(sym), mode=0)
        assert rc == 0 and np.array_equal(s0, sym)
        rc, s1 = oracle.fgk_decode(b0, len(sym), mode=1)
        assert rc == 0 and np.array_equal(s1, sym)


def test_a7_hand_vectors(oracle, golden):
    for v in golden["a7"]:
        data = np.frombuffer(bytes.fromhex(v["in"]), np.uint8)
        f = v["flags"]
        width = int(f[f.index("-w") + 1]) if "-w" in f else 512
        for mode in (0, 1):
            rc, out = oracle.compress(data, diff="-m" in f, adapt="-a" in f, width=width, mode=mode)
            assert rc == 0 and bytes(out).hex() == v["out"], v
    # spot check against the literal table in SURVEY.md A.7
    rc, out = oracle.compress(np.frombuffer(b"AAAA", np.uint8))
    assert bytes(out).hex() == "05" + "00" * 7 + "00" + "41c010"
    rc, out = oracle.compress(np.zeros(0, np.uint8))
    assert bytes(out) == b"\x00" * 9


def test_synthetic_golden(oracle, golden):
    for e in golden["synthetic"]:
        w = e.get("w", e.get("n"))
        h = e.get("h", e.get("n"))
        img = synth.image(e["kind"], w, e["seed"], h).reshape(-1)
        rc, out = oracle.compress(img, diff="m" in e["mode"], adapt="a" in e["mode"], width=w)
        assert rc == 0
        assert (len(out), _sha(out)) == (e["size"], e["sha256"]), e
        if "B" in e:
            rc, sym, b = oracle.adapt_encode(oracle.diff_apply(img) if "m" in e["mode"] else img, w, h)
            assert b == e["B"]


def test_stage_known_answers(oracle, golden):
    for s in golden["stages"]:
        v = np.frombuffer(bytes.fromhex(s["in"]), np.uint8)
        assert bytes(oracle.diff_apply(v)).hex() == s["diff"]
        assert bytes(oracle.rle_encode(v)).hex() == s["rle"]
        assert bytes(oracle.fgk_encode(v, mode=0)[0]).hex() == s["fgk"]
        assert bytes(oracle.fgk_encode(v, mode=1)[0]).hex() == s["fgk"]
        assert np.array_equal(oracle.rle_decode(oracle.rle_encode(v)), v)
        assert np.array_equal(oracle.diff_revert(oracle.diff_apply(v)), v)


def test_error_codes(oracle, golden):
    for c in golden["cli"]:
        if "malformed" in c:
            rc, _ = oracle.decompress(np.frombuffer(bytes.fromhex(c["blob"]), np.uint8))
            assert rc == c["rc"], c
    small = np.arange(30, dtype=np.uint8)
    assert oracle.compress(small, adapt=True, width=7)[0] == 6
    assert oracle.compress(small, adapt=True, width=5)[0] == 12
    assert oracle.compress(np.zeros(0, np.uint8), adapt=True, width=512)[0] == 12


def _fuzz_inputs(rng, count, maxn):
    for i in range(count):
        n = int(rng.integers(0, maxn))
        k = i % 5
        if k == 0:
            v = rng.integers(0, 256, n)
        elif k == 1:
            v = rng.integers(0, 2, n) * 255
        elif k == 2:
            reps = rng.integers(1, 600, n // 100 + 2)
            v = np.repeat(rng.integers(0, 256, reps.size), reps)[:n]
        elif k == 3:
            reps = rng.integers(1, 7, n // 2 + 2)
            v = np.repeat(rng.integers(250, 256, reps.size), reps)[:n]
        else:
            v = np.cumsum(rng.integers(-1, 2, n)) & 255
        yield v.astype(np.uint8)


def test_fuzz_stages_vs_reference(oracle, ref):
    rng = np.random.default_rng(2024)
    for v in _fuzz_inputs(rng, 150, 6000):
        assert np.array_equal(oracle.diff_apply(v), ref.diff_apply(v))
        assert np.array_equal(oracle.diff_revert(v), ref.diff_revert(v))
        e = oracle.rle_encode(v)
        assert np.array_equal(e, ref.rle_encode(v))
        assert np.array_equal(oracle.rle_decode(e), ref.rle_decode(e))
        # decoding arbitrary bytes is also defined (any byte string is a valid RLE stream)
        assert np.array_equal(oracle.rle_decode(v[:400]), ref.rle_decode(v[:400]))
        f = ref.fgk_encode(v)
        assert np.array_equal(oracle.fgk_encode(v, mode=1)[0], f)
        assert np.array_equal(oracle.fgk_encode(v[:1500], mode=0)[0], ref.fgk_encode(v[:1500]))
        assert np.array_equal(ref.fgk_decode(f, v.size), v)
        rc, s = oracle.fgk_decode(f, v.size, mode=1)
        assert rc == 0 and np.array_equal(s, v)


def test_fuzz_adaptive_vs_reference(oracle, ref):
    rng = np.random.default_rng(7)
    shapes = [(8, 8), (9, 8), (8, 9), (15, 17), (16, 16), (17, 33), (64, 24), (31, 100), (130, 70), (256, 40)]
    for i, (w, h) in enumerate(shapes * 3):
        kind = ("walk", "smooth", "random", "const", "longrun")[i % 5]
        img = synth.image(kind, w, 100 + i, h).reshape(-1)
        rc, enc, b = oracle.adapt_encode(img, w, h)
        assert rc == 0
        assert np.array_equal(enc, ref.adapt_encode(img, w, h)), (w, h, kind)
        for bs in (8, 16, 32):
            if bs <= w and bs <= h:
                assert np.array_equal(oracle.adapt_encode_bs(img, w, h, bs), ref.adapt_encode_bs(img, w, h, bs))
        rc, dec = oracle.adapt_decode(enc)
        assert rc == 0 and np.array_equal(dec, img)
        assert np.array_equal(ref.adapt_decode(enc), img)


def test_fgk_deep_tree(oracle, ref):
    # Fibonacci-skewed stream: long codes (depth > 16), multi-byte code words
    img = synth.image("fib", 256, 1).reshape(-1)
    sym = np.sort(img)[::-1].copy()
    b1, n1 = oracle.fgk_encode(sym, mode=1)
    assert np.array_equal(b1, ref.fgk_encode(sym))
    lv, sw, md = oracle.fgk_stats(sym)
    assert md >= 16
    rc, s = oracle.fgk_decode(b1, sym.size)
    assert rc == 0 and np.array_equal(s, sym)
