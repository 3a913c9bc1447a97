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
