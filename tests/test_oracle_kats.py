"""Pins the oracle (CPU restatement) against the hand-checkable KATs and the survey-derived fixture digests.

The reference ships no golden vectors for this path and cannot be built here (no OCaml): parity is therefore
"unpinned" in the sense of the task statement; these are the strongest pins available.
"""
import hashlib
import json
import os

import pytest

from conftest import GOLDEN, run_cli
from kats import K9, KATS


def materialise(tmp_path, files, argv):
    paths = {}
    for name, data in files.items():
        p = tmp_path / name
        p.write_bytes(data)
        paths[name] = str(p)
    out = []
    for a in argv:
        if a.startswith("{") and a.endswith("}"):
            out.append(paths[a[1:-1]])
        else:
            out.append(a)
    return out


@pytest.mark.parametrize("kat", KATS, ids=[k[0] for k in KATS])
def test_oracle_kat(oracle_bin, tmp_path, kat):
    name, files, argv, expected, code = kat
    rc, out, err = run_cli(oracle_bin, materialise(tmp_path, files, argv))
    assert rc == code, err.decode(errors="replace")
    assert out == expected


def test_oracle_k9_bucket_order(oracle_bin, tmp_path):
    name, files, argv, head, code = K9
    rc, out, err = run_cli(oracle_bin, materialise(tmp_path, files, argv))
    assert rc == code
    assert out.split(b"\n")[: len(head)] == head


def test_oracle_output_file_naming(oracle_bin, tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">s\nACGTNACG\n")
    rc, out, _ = run_cli(oracle_bin, ["-k", "3", "-l", "x", "-f", str(fa), "-o", str(tmp_path / "pre")])
    assert rc == 0 and out == b""
    assert (tmp_path / "pre.KPopSpectra.txt").read_bytes() == b"\tx\n06\t3\n"
    rc, out, _ = run_cli(oracle_bin, ["-k", "3", "-l", "x", "-f", str(fa), "-o", "/dev/stdout"])
    assert rc == 0 and out == b"\tx\n06\t3\n"


REF_TEST = "/root/reference/test"


@pytest.mark.skipif(not os.path.isdir(REF_TEST), reason="reference fixtures only exist in the build container")
def test_oracle_fixture_digests(oracle_bin):
    """md5 / line / byte counts of SURVEY.md 8c, derived there by an independent Python restatement."""
    with open(os.path.join(GOLDEN, "fixture_digests.json")) as f:
        cases = json.load(f)
    for c in cases:
        argv = [a.replace("{REF_TEST}", REF_TEST) for a in c["argv"]]
        rc, out, _ = run_cli(oracle_bin, argv)
        assert rc == 0
        assert len(out) == c["bytes"], c["argv"]
        assert out.count(b"\n") == c["lines"], c["argv"]
        assert hashlib.md5(out).hexdigest() == c["md5"], c["argv"]
