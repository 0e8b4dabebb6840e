"""When a real (OCaml) KPopCount of the reference is available (baseline/_ref/, oracle/_ref/ or PATH: see
oracle/refprobe.py), the C++ oracle is pinned to it: every KAT, K9 (bucket order for k > 12) and the golden inputs
under tests/golden/inputs.  The image has no OCaml toolchain, so without such a binary these tests skip and say so --
parity then rests on the survey-derived digests (tests/test_oracle_kats.py): "parity unpinned"."""
import gzip
import os
import sys

import pytest

from conftest import GOLDEN, ROOT, run_cli
from kats import K9, KATS
from test_oracle_kats import materialise

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from refprobe import find_reference_kpopcount  # noqa: E402

REF = find_reference_kpopcount()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference KPopCount binary under baseline/_ref, oracle/_ref or on PATH: "
                                                   "the oracle stays pinned to the survey-derived digests only")


def test_probe_never_returns_the_repos_own_front_ends():
    assert REF is None or not os.path.realpath(REF).startswith(os.path.realpath(os.path.join(ROOT, "kpop_b200")))


@needs_ref
@pytest.mark.parametrize("kat", KATS + [K9], ids=[k[0] for k in KATS + [K9]])
def test_oracle_equals_reference_on_kats(oracle_bin, tmp_path, kat):
    name, files, argv, _expected, _code = kat
    args = materialise(tmp_path, files, argv)
    rc_r, out_r, _ = run_cli(REF, args)
    rc_o, out_o, _ = run_cli(oracle_bin, args)
    assert (rc_o, out_o) == (rc_r, out_r), name


@needs_ref
@pytest.mark.parametrize("argv", [["-k", "5", "-L"], ["-k", "5", "-l", "all"], ["-k", "12", "-l", "x"], ["-k", "13", "-l", "x"],
                                  ["-k", "30", "-l", "x"], ["-k", "12", "-M", "1000", "-l", "x"], ["-k", "15", "-M", "5000", "-l", "x"]],
                         ids=lambda a: "_".join(a))
def test_oracle_equals_reference_on_golden_inputs(oracle_bin, tmp_path, argv):
    for name in ("wuhan.fasta", "clusters-small.fasta"):
        if name.startswith("clusters") and "-L" not in argv and argv[1] != "5":
            continue
        p = tmp_path / name
        with gzip.open(os.path.join(GOLDEN, "inputs", name + ".gz"), "rb") as f:
            p.write_bytes(f.read())
        rc_r, out_r, _ = run_cli(REF, argv + ["-f", str(p)])
        rc_o, out_o, _ = run_cli(oracle_bin, argv + ["-f", str(p)])
        assert (rc_o, out_o) == (rc_r, out_r), (name, argv)
