#!/usr/bin/env python3
"""Regenerate tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run in the build container only:

    make -C oracle all && python tests/golden/make_golden.py

Writes
  tests/golden/samples.npz  the reference's 12 sample inputs (data/*.raw), zlib-packed
  tests/golden/golden.json  size / sha256 / M / flags / chosen block size of every
                            reference .out for samples x {plain,-m,-a,-m -a}; the
                            hand-checkable vectors of SURVEY A.7; synthetic-image pins;
                            small stage-level known answers; CLI exit codes.
The GPU box has no /root/reference: tests read only these two files (and, when it
travelled with the snapshot, oracle/_ref for live differential checks).
"""
import glob
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "huffman-codec_b200"))
import pyoracle  # noqa: E402
import synth  # noqa: E402

REF_DATA = "/root/reference/data"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))
MODES = {"plain": [], "m": ["-m"], "a": ["-a", "-w", "512"], "ma": ["-m", "-a", "-w", "512"]}


def run_ref(data, flags, tmp):
    """compress with the reference binary -> (.out bytes); also checks its own round trip."""
    inp = os.path.join(tmp, "in.raw")
    outp = os.path.join(tmp, "x.out")
    dec = os.path.join(tmp, "x.dec")
    with open(inp, "wb") as f:
        f.write(bytes(data))
    r = subprocess.run([pyoracle.REF_BIN, "-c"] + flags + ["-i", inp, "-o", outp], capture_output=True)
    if r.returncode != 0:
        return r.returncode, None
    out = open(outp, "rb").read()
    r = subprocess.run([pyoracle.REF_BIN, "-d", "-i", outp, "-o", dec], capture_output=True)
    assert r.returncode == 0 and open(dec, "rb").read() == bytes(data), "reference round trip failed"
    return 0, out


def describe(out, adapt, ora):
    d = {"size": len(out), "sha256": hashlib.sha256(out).hexdigest(),
         "M": int.from_bytes(out[:8], "little"), "flags": out[8]}
    if adapt:
        rc, sym = ora.fgk_decode(np.frombuffer(out[9:], np.uint8), d["M"])
        assert rc == 0
        d["B"] = int.from_bytes(bytes(sym[16:24]), "big")
    return d


def main():
    ora = pyoracle.Oracle()
    ref = pyoracle.Ref()
    gold = {"samples": {}, "a7": [], "synthetic": [], "stages": [], "cli": []}
    samples = {}
    with tempfile.TemporaryDirectory() as tmp:
        for path in sorted(glob.glob(os.path.join(REF_DATA, "*.raw"))):
            name = os.path.basename(path)[:-4]
            data = open(path, "rb").read()
            samples[name] = np.frombuffer(data, np.uint8)
            gold["samples"][name] = {"in_size": len(data), "in_sha256": hashlib.sha256(data).hexdigest()}
            for mode, flags in MODES.items():
                rc, out = run_ref(data, flags, tmp)
                assert rc == 0
                gold["samples"][name][mode] = describe(out, "-a" in flags, ora)
                print(name, mode, gold["samples"][name][mode]["size"], flush=True)

        # SURVEY A.7 hand-checkable vectors, regenerated rather than trusted
        a7 = [(b"", []), (b"A", []), (b"AAAA", []), (b"AAAAA", []), (b"A" * 258 + b"B", []),
              (b"A" * 259, []), (b"A" * 600, []), (b"ABAB", []), (b"ABAB", ["-m"]),
              (bytes([x for y in range(8) for x in range(8)]), ["-a", "-w", "8"]),
              (bytes([y for y in range(8) for x in range(8)]), ["-a", "-w", "8"]),
              (bytes(range(256)) * 3, ["-m"]), (bytes(range(256)) * 3, [])]
        for data, flags in a7:
            rc, out = run_ref(data, flags, tmp)
            assert rc == 0
            gold["a7"].append({"in": data.hex(), "flags": flags, "out": out.hex()})

        # synthetic classes, one image per class and size
        for n, seed in ((64, 7), (512, 1234)):
            for kind in synth.CLASSES + ("fib", "longrun"):
                img = synth.image(kind, n, seed).reshape(-1)
                for mode, flags in (("m", ["-m"]), ("ma", ["-m", "-a", "-w", str(n)]), ("a", ["-a", "-w", str(n)])):
                    rc, out = run_ref(img, flags, tmp)
                    assert rc == 0
                    e = describe(out, "-a" in flags, ora)
                    e.update({"kind": kind, "n": n, "seed": seed, "mode": mode})
                    gold["synthetic"].append(e)
        # odd shapes: width x height not multiples of 8, W != H
        for (w, h, seed) in ((8, 8, 1), (9, 8, 2), (8, 9, 3), (17, 15, 4), (24, 40, 5), (100, 36, 6), (513, 9, 7), (33, 257, 8)):
            for kind in ("walk", "smooth", "const"):
                img = synth.image(kind, w, seed, h).reshape(-1)
                rc, out = run_ref(img, ["-m", "-a", "-w", str(w)], tmp)
                assert rc == 0
                e = describe(out, True, ora)
                e.update({"kind": kind, "w": w, "h": h, "seed": seed, "mode": "ma"})
                gold["synthetic"].append(e)

        # CLI exit codes, from the reference binary itself
        bad = os.path.join(tmp, "bad.out")
        small = os.path.join(tmp, "small.raw")
        open(small, "wb").write(bytes(range(30)))
        cases = [(["-h"], None), ([], None), (["-i"], None), (["-x"], None), (["-c", "-w", "0", "-i", small], None),
                 (["-i", os.path.join(tmp, "nonexistent")], None),
                 (["-c", "-a", "-w", "7", "-i", small], None),     # 30 % 7 != 0 -> 6
                 (["-c", "-a", "-w", "5", "-i", small], None),     # W < 8 -> 12
                 ]
        for args, _ in cases:
            r = subprocess.run([pyoracle.REF_BIN] + args, capture_output=True, cwd=tmp)
            gold["cli"].append({"args": [a.replace(tmp, "$TMP") for a in args], "rc": r.returncode,
                                "stderr": r.stderr.decode().replace(tmp, "$TMP"),
                                "stdout": r.stdout.decode()})
        # malformed .out files for the decoder
        rc, good = run_ref(bytes(range(64)) * 2, ["-a", "-w", "8"], tmp)
        rc, plain = run_ref(b"hello world, hello world", [], tmp)
        malformed = {
            "short_header": plain[:5],                                       # 8
            "bit_underrun": plain[:-2],                                      # 9
            "count_too_big": (len(plain) * 100).to_bytes(8, "little") + plain[8:],  # 9
            "adapt_flag_on_plain": plain[:8] + bytes([0x40]) + plain[9:],    # 10 (stream < 24 B)
        }
        for name, blob in malformed.items():
            open(bad, "wb").write(blob)
            r = subprocess.run([pyoracle.REF_BIN, "-d", "-i", bad, "-o", os.path.join(tmp, "o")], capture_output=True)
            gold["cli"].append({"malformed": name, "blob": blob.hex(), "rc": r.returncode,
                                "stderr": r.stderr.decode()})

    # stage-level known answers from the reference's own stage functions (libhcref.so)
    rng = np.random.default_rng(99)
    for i in range(24):
        n = int(rng.integers(1, 700))
        kind = i % 4
        if kind == 0:
            v = rng.integers(0, 256, n)
        elif kind == 1:
            v = rng.integers(0, 3, n)
        elif kind == 2:
            v = np.repeat(rng.integers(0, 256, n // 40 + 1), rng.integers(1, 300, n // 40 + 1))[:n]
        else:
            v = np.repeat(rng.integers(250, 256, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))[:n]
        v = v.astype(np.uint8)
        rle = ref.rle_encode(v)
        gold["stages"].append({"in": bytes(v).hex(), "diff": bytes(ref.diff_apply(v)).hex(),
                               "rle": bytes(rle).hex(), "fgk": bytes(ref.fgk_encode(v)).hex()})

    np.savez_compressed(os.path.join(OUT_DIR, "samples.npz"), **samples)
    with open(os.path.join(OUT_DIR, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", OUT_DIR)


if __name__ == "__main__":
    main()
