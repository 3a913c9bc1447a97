"""CLI contract that needs no GPU: option parsing, help text and exit codes 0-5 are decided
before the codec is touched (reference src/main.cpp:152-207; golden values from the reference binary)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "huffman-codec_b200", "huffman-codec")


def test_cli_argument_errors_match_reference(golden):
    if not os.path.exists(CLI):
        import __graft_entry__
        __graft_entry__.build()
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "small.raw"), "wb").write(bytes(range(30)))
        n = 0
        for c in golden["cli"]:
            if "malformed" in c or c["rc"] > 5:
                continue
            r = subprocess.run([CLI] + [a.replace("$TMP", tmp) for a in c["args"]], capture_output=True, cwd=tmp)
            assert r.returncode == c["rc"], c
            assert r.stdout.decode() == c["stdout"] and r.stderr.decode() == c["stderr"].replace("$TMP", tmp), c
            n += 1
        assert n >= 6


def test_cli_extensions_without_gpu():
    """-w with a non-number ends like an invalid width (the reference aborts on an uncaught exception,
    src/main.cpp:176); a missing list file is a missing input."""
    if not os.path.exists(CLI):
        import __graft_entry__
        __graft_entry__.build()
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "small.raw"), "wb").write(bytes(range(30)))
        r = subprocess.run([CLI, "-c", "-a", "-w", "abc", "-i", "small.raw"], capture_output=True, cwd=tmp)
        assert r.returncode == 4 and b"invalid 2D data width" in r.stderr
        r = subprocess.run([CLI, "-c", "-L", "nolist.txt"], capture_output=True, cwd=tmp)
        assert r.returncode == 5
