"""Whole-file parity through the host-level C ABI (hc_compress_batch / hc_decompress_batch
= huffCompress / huffDecompress, reference src/main.cpp:39-128).

Compared against: the CPU oracle, the golden digests produced by the unmodified reference
binary (tests/golden/golden.json) and -- when oracle/_ref travelled with the snapshot -- the
reference binary itself in both directions (ours -> ref -d, ref -c -> ours).
"""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import hc_b200
import pyoracle
import synth
from backend import CudaBackend, EmuBackend

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
_cache = {}
MODES = {"plain": (False, False), "m": (True, False), "a": (False, True), "ma": (True, True)}


@pytest.fixture(params=BACKENDS)
def codec(request):
    name = request.param
    if name not in _cache:
        be = EmuBackend() if name == "emu" else CudaBackend()
        _cache[name] = (hc_b200.Codec(0, be.L), name)
    return _cache[name]


def _sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def test_a7_vectors(codec, golden):
    cd, _ = codec
    for v in golden["a7"]:
        data = np.frombuffer(bytes.fromhex(v["in"]), np.uint8)
        f = v["flags"]
        width = int(f[f.index("-w") + 1]) if "-w" in f else 512
        outs, st = cd.compress([data], diff="-m" in f, adapt="-a" in f, width=width)
        assert st[0] == 0 and bytes(outs[0]).hex() == v["out"], v
        back, st = cd.decompress(outs)
        assert st[0] == 0 and np.array_equal(back[0], data)


@pytest.mark.parametrize("mode", list(MODES))
def test_small_batch_vs_oracle(codec, oracle, mode):
    cd, name = codec
    diff, adapt = MODES[mode]
    n = 32 if name == "emu" else 96
    files = []
    for i, k in enumerate(synth.CLASSES + ("fib", "longrun")):
        files.append(synth.image(k, n, 400 + i, n + 8 * (i % 3)).reshape(-1))
    if not adapt:
        files += [np.zeros(0, np.uint8), np.array([7], np.uint8), np.arange(1000, dtype=np.uint32).astype(np.uint8)[:999]]
    outs, st = cd.compress(files, diff=diff, adapt=adapt, width=n)
    assert not st.any()
    for f, o in zip(files, outs):
        rc, exp = oracle.compress(f, diff=diff, adapt=adapt, width=n)
        assert rc == 0 and np.array_equal(o, exp)
    back, st = cd.decompress(outs)
    assert not st.any()
    for f, b in zip(files, back):
        assert np.array_equal(b, f)


def test_mixed_kinds_in_one_decode_batch(codec, oracle):
    cd, _ = codec
    img = synth.image("walk", 32, 9, 40).reshape(-1)
    outs = [oracle.compress(img, diff=d, adapt=a, width=32)[1] for d in (0, 1) for a in (0, 1)]
    back, st = cd.decompress(outs)
    assert not st.any()
    for b in back:
        assert np.array_equal(b, img)


def test_compress_errors(codec):
    cd, _ = codec
    files = [np.arange(30, dtype=np.uint8), np.arange(64, dtype=np.uint8), np.arange(35, dtype=np.uint8)]
    outs, st = cd.compress(files, adapt=True, width=[7, 8, 5])
    assert list(st) == [6, 0, 12]            # src/main.cpp:54-58, ok, src/transform.cpp:300-304
    assert outs[0].size == 0 and outs[2].size == 0 and outs[1].size > 9
    outs, st = cd.compress([np.zeros(0, np.uint8)], adapt=True, width=512)
    assert list(st) == [12]                  # empty file with -a


def test_decompress_errors(codec, golden, oracle):
    cd, _ = codec
    blobs, expect = [], []
    for c in golden["cli"]:
        if "malformed" in c:
            blobs.append(np.frombuffer(bytes.fromhex(c["blob"]), np.uint8))
            expect.append(c["rc"])
    good = oracle.compress(np.frombuffer(b"hello world, hello world", np.uint8))[1]
    blobs.append(good)
    expect.append(0)
    blobs.append(np.concatenate([good, np.array([1, 2, 3, 4], np.uint8)]))     # trailing bytes are ignored
    expect.append(0)
    back, st = cd.decompress(blobs)
    assert list(st) == expect
    assert bytes(back[-1]) == b"hello world, hello world" and bytes(back[-2]) == bytes(back[-1])


def test_mutation_fuzz_matches_oracle(codec, oracle):
    """Bit flips, truncations and appended bytes in the body of valid .out files (the 9-byte container header and the
    24 bytes of an adaptive header stay intact, so sizes stay bounded): status and bytes must equal the oracle's for
    every file of the batch, and a broken file never disturbs its neighbours."""
    cd, name = codec
    rng = np.random.default_rng(99)
    n = 24 if name == "emu" else 48
    base = []
    for i in range(6):
        img = synth.image(synth.CLASSES[i % 4], n, 600 + i, n).reshape(-1)
        for diff, adapt in ((False, False), (True, True)):
            base.append(oracle.compress(img, diff=diff, adapt=adapt, width=n)[1])
    blobs = []
    for k in range(5 * len(base)):
        b = base[k % len(base)].copy()
        kind = k % 5
        if kind == 0 and b.size > 12:
            for _ in range(1 + k % 3):
                b[int(rng.integers(10, b.size))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1 and b.size > 12:
            b = b[: int(rng.integers(9, b.size))].copy()
        elif kind == 2:
            b = np.concatenate([b, rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8)])
        elif kind == 3 and b.size > 20:
            i = int(rng.integers(10, b.size - 4))
            b[i:i + 4] = rng.integers(0, 256, 4, dtype=np.uint8)
        blobs.append(b)
    back, st = cd.decompress(blobs)
    agree = 0
    for b, got, s in zip(blobs, back, st):
        rc, exp = oracle.decompress(b)
        assert int(s) == rc, (int(s), rc)
        if rc == 0:
            assert np.array_equal(got, exp)
            agree += 1
    assert agree >= len(base)                        # the untouched and the harmlessly extended ones at least


def _container(sym, flags, oracle):
    bits, _ = oracle.fgk_encode(sym)
    return np.concatenate([np.frombuffer(int(sym.size).to_bytes(8, "little") + bytes([flags]), np.uint8), bits])


def test_hostile_adaptive_header_fails_alone(codec, oracle):
    """A crafted adaptive header (w = h = b = 2^31: 2^62 output bytes promised by 13 payload bytes; the reference dies
    in its allocation, src/transform.cpp:340) fails with its own status and does not take the batch with it, whatever
    output capacity the caller offers; a file that merely does not fit reports HC_E_CAPACITY and the size it needs."""
    cd, _ = codec
    be = (2 ** 31).to_bytes(8, "big")
    hostile = _container(np.frombuffer(be * 3 + bytes([0x80]) + bytes(range(13)), np.uint8), 0x40, oracle)
    huge_b = _container(np.frombuffer((16).to_bytes(8, "big") * 2 + (2 ** 63).to_bytes(8, "big") + bytes([0x80, 7, 7, 7, 200, 1]), np.uint8),
                        0x40, oracle)
    good1 = oracle.compress(np.frombuffer(b"hello world, hello world", np.uint8))[1]
    img = np.arange(64 * 64, dtype=np.uint32).astype(np.uint8)
    good2 = oracle.compress(img, diff=True, adapt=True, width=64)[1]
    for cap in (1 << 20, 1 << 26):
        back, st = cd.decompress([good1, hostile, good2, huge_b], out_cap=cap)
        assert st[0] == 0 and st[2] == 0 and st[1] in (13, 14) and st[3] in (13, 14), list(st)
        assert bytes(back[0]) == b"hello world, hello world" and np.array_equal(back[2], img)
    # too small a buffer: per-file capacity status carrying the need, then one exact retry inside Codec.decompress
    buf, offs, lens = hc_b200.Codec.pack([good1, good2])
    rc, out, oo, ol, st = cd.decompress_packed(buf, offs, lens, 64)
    assert rc == 0 and list(st) == [0, hc_b200.HC_E_CAPACITY] and int(ol[1]) == img.size
    back, st = cd.decompress([good1, good2], out_cap=64)
    assert not st.any() and np.array_equal(back[1], img)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(MODES))
def test_samples_golden(golden, samples, mode):
    """BASELINE config 2: data/*.raw, all four flag sets, byte identical to the reference."""
    cd = hc_b200.Codec(0)
    diff, adapt = MODES[mode]
    names = sorted(samples)
    outs, st = cd.compress([samples[n] for n in names], diff=diff, adapt=adapt, width=512)
    assert not st.any()
    for n, o in zip(names, outs):
        g = golden["samples"][n][mode]
        assert (o.size, _sha(o)) == (g["size"], g["sha256"]), (n, mode)
    back, st = cd.decompress(outs)
    assert not st.any()
    for n, b in zip(names, back):
        assert _sha(b) == golden["samples"][n]["in_sha256"]


@pytest.mark.gpu
def test_synthetic_golden(golden):
    cd = hc_b200.Codec(0)
    for e in golden["synthetic"]:
        w, h = e.get("w", e.get("n")), e.get("h", e.get("n"))
        img = synth.image(e["kind"], w, e["seed"], h).reshape(-1)
        outs, st = cd.compress([img], diff="m" in e["mode"], adapt="a" in e["mode"], width=w)
        assert st[0] == 0 and (outs[0].size, _sha(outs[0])) == (e["size"], e["sha256"]), e


@pytest.mark.gpu
def test_cross_decode_with_reference_binary(samples):
    """ours -> reference -d, and reference -c -> ours (only where oracle/_ref is present)."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built")
    cd = hc_b200.Codec(0)
    files = {"hd01": samples["hd01"], "extra": samples["hd01extra"], "walk": synth.image("walk", 200, 5, 120).reshape(-1)}
    widths = {"hd01": 512, "extra": 512, "walk": 200}
    with tempfile.TemporaryDirectory() as tmp:
        for name, data in files.items():
            for flags in ([], ["-m"], ["-a", "-w", str(widths[name])], ["-m", "-a", "-w", str(widths[name])]):
                ours, st = cd.compress([data], diff="-m" in flags, adapt="-a" in flags, width=widths[name])
                assert st[0] == 0
                p_in, p_out, p_dec = (os.path.join(tmp, x) for x in ("in.raw", "o.out", "o.dec"))
                open(p_out, "wb").write(bytes(ours[0]))
                r = subprocess.run([pyoracle.REF_BIN, "-d", "-i", p_out, "-o", p_dec], capture_output=True)
                assert r.returncode == 0 and open(p_dec, "rb").read() == bytes(data)
                open(p_in, "wb").write(bytes(data))
                r = subprocess.run([pyoracle.REF_BIN, "-c"] + flags + ["-i", p_in, "-o", p_out], capture_output=True)
                assert r.returncode == 0
                ref_out = np.frombuffer(open(p_out, "rb").read(), np.uint8)
                assert np.array_equal(ref_out, ours[0])
                back, st = cd.decompress([ref_out])
                assert st[0] == 0 and np.array_equal(back[0], data)


@pytest.mark.gpu
def test_large_images_4096(oracle):
    """BASELINE config 4 geometry: 4096x4096, -m -a -w 4096 (block sizes up to 1024), one image per
    low/medium entropy class plus a 1024x1024 random image; byte identical to the oracle."""
    cd = hc_b200.Codec(0)
    files, widths = [], []
    for i, kind in enumerate(("smooth", "const", "walk")):
        files.append(synth.image(kind, 4096, 2000 + i).reshape(-1))
        widths.append(4096)
    files.append(synth.image("random", 1024, 2003).reshape(-1))
    widths.append(1024)
    outs, st = cd.compress(files, diff=True, adapt=True, width=widths)
    assert not st.any()
    for f, w, o in zip(files, widths, outs):
        rc, exp = oracle.compress(f, diff=True, adapt=True, width=w, mode=1)
        assert rc == 0 and o.size == exp.size and np.array_equal(o, exp), w
    back, st = cd.decompress(outs)
    assert not st.any()
    for f, b in zip(files, back):
        assert np.array_equal(f, b)
    # plain RLE path on a 16 MiB file (segmented diff kernels, long CTA streams)
    outs, st = cd.compress(files[:1], diff=True, adapt=False)
    rc, exp = oracle.compress(files[0], diff=True, adapt=False, mode=1)
    assert st[0] == 0 and np.array_equal(outs[0], exp)
    back, st = cd.decompress(outs)
    assert st[0] == 0 and np.array_equal(back[0], files[0])


@pytest.mark.gpu
def test_cli_matches_reference_cli(samples, golden):
    """The huffman-codec binary of this repo against the reference CLI contract."""
    cli = os.path.join(os.path.dirname(hc_b200.HERE), "huffman-codec_b200", "huffman-codec")
    if not os.path.exists(cli):
        pytest.skip("CLI not built")
    with tempfile.TemporaryDirectory() as tmp:
        p_in, p_out, p_dec = (os.path.join(tmp, x) for x in ("hd01.raw", "hd01.out", "hd01.dec"))
        open(p_in, "wb").write(bytes(samples["hd01"]))
        for mode, flags in (("plain", []), ("ma", ["-m", "-a", "-w", "512"])):
            r = subprocess.run([cli, "-c"] + flags + ["-i", p_in, "-o", p_out], capture_output=True)
            out = open(p_out, "rb").read()
            g = golden["samples"]["hd01"][mode]
            assert r.returncode == 0 and (len(out), _sha(out)) == (g["size"], g["sha256"])
            assert r.stderr.decode() == "writing %d bytes to %s\n" % (len(out), p_out)
            r = subprocess.run([cli, "-d", "-i", p_out, "-o", p_dec], capture_output=True)
            assert r.returncode == 0 and open(p_dec, "rb").read() == bytes(samples["hd01"])
        # error behaviour: same exit codes and stderr text as the reference binary
        for c in golden["cli"]:
            if "malformed" in c:
                bad = os.path.join(tmp, "bad.out")
                open(bad, "wb").write(bytes.fromhex(c["blob"]))
                r = subprocess.run([cli, "-d", "-i", bad, "-o", os.path.join(tmp, "o")], capture_output=True)
            else:
                small = os.path.join(tmp, "small.raw")
                open(small, "wb").write(bytes(range(30)))
                r = subprocess.run([cli] + [a.replace("$TMP", tmp) for a in c["args"]], capture_output=True, cwd=tmp)
                assert r.stdout.decode() == c["stdout"]
            assert r.returncode == c["rc"], c
            assert r.stderr.decode() == c["stderr"].replace("$TMP", tmp), c


def test_many_files_cross_group_pipeline(codec, oracle, monkeypatch):
    """The host-level calls split large batches into several group pipelines (own stream and buffers
    each; by default one per 4096 files, forced to 3 here); offsets, lengths and statuses must still
    line up with the file order."""
    cd, name = codec
    monkeypatch.setenv("HC_GROUPS", "3")
    rng = np.random.default_rng(5)
    nfile = 200
    files, widths = [], []
    for i in range(nfile):
        w = int(rng.integers(8, 24))
        h = int(rng.integers(8, 20))
        img = synth.image(("walk", "smooth", "random", "const")[i % 4], w, 1000 + i, h).reshape(-1)
        if i % 37 == 5:
            img = img[:-3]                     # size no longer a multiple of the width -> status 6
        files.append(img)
        widths.append(w)
    outs, st = cd.compress(files, diff=True, adapt=True, width=widths)
    good = []
    for i, (f, w, o) in enumerate(zip(files, widths, outs)):
        rc, exp = oracle.compress(f, diff=True, adapt=True, width=w)
        assert st[i] == rc, i
        if rc == 0:
            assert np.array_equal(o, exp), i
            good.append(i)
        else:
            assert o.size == 0
    back, st2 = cd.decompress([outs[i] for i in good])
    assert not st2.any()
    for i, b in zip(good, back):
        assert np.array_equal(b, files[i]), i


@pytest.mark.gpu
def test_mixed_entropy_stress_all_modes(oracle):
    """Mini C5: random / smooth / const thirds plus walk, long-run and Fibonacci-skewed
    streams, random shapes up to 300 x 300 (ragged edge blocks, every block-size path of the adaptive
    coder), all four flag combinations, each output compared with the oracle and decoded back."""
    cd = hc_b200.Codec(0)
    rng = np.random.default_rng(31337)
    kinds = ("random", "smooth", "const", "walk", "longrun", "fib")
    files, widths = [], []
    for i in range(180):
        w = int(rng.integers(8, 300))
        h = int(rng.integers(8, 300))
        files.append(synth.image(kinds[i % 6], w, 5000 + i, h).reshape(-1))
        widths.append(w)
    for diff in (False, True):
        for adapt in (False, True):
            outs, st = cd.compress(files, diff=diff, adapt=adapt, width=widths)
            assert not st.any()
            for i, (f, w, o) in enumerate(zip(files, widths, outs)):
                rc, exp = oracle.compress(f, diff=diff, adapt=adapt, width=w)
                assert rc == 0 and np.array_equal(o, exp), (i, diff, adapt, w, f.size // w)
            back, st2 = cd.decompress(outs)
            assert not st2.any()
            for i, (f, b) in enumerate(zip(files, back)):
                assert np.array_equal(b, f), (i, diff, adapt)
