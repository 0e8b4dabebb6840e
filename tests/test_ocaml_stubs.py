"""The OCaml side of the boundary (ocaml/): there is no OCaml toolchain in the image, so the C stubs are type-checked by
gcc against a mock of <caml/*.h> (tests/mock_caml: documented signatures only, no behaviour), and the `external`
declarations of Kpc_gpu.ml are checked against the stubs and against the symbols the library exports."""
import os
import re
import subprocess

from conftest import ROOT

OCAML = os.path.join(ROOT, "ocaml")


def test_stubs_typecheck_against_mock_runtime():
    p = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "tests", "mock_caml"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(OCAML, "kpc_stubs.c")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()


def test_every_external_has_a_stub_with_the_same_arity():
    ml = open(os.path.join(OCAML, "Kpc_gpu.ml")).read()
    c = open(os.path.join(OCAML, "kpc_stubs.c")).read()
    ext = re.findall(r"external\s+(\w+)\s*:\s*(.*?)\s*=\s*\"(\w+)\"", ml, flags=re.S)
    assert len(ext) >= 14
    for name, sig, sym in ext:
        m = re.search(r"^value %s\((.*?)\)\s*\{" % sym, c, flags=re.M)
        assert m, f"no stub for external {name} = {sym}"
        n_c = len([a for a in m.group(1).split(",") if a.strip()])
        n_ml = sig.count("->")  # arguments of the OCaml type (all first order here)
        assert n_c == n_ml, f"{sym}: {n_c} C parameters for {n_ml} OCaml arguments"
        assert n_c <= 5, "more than five arguments would need a bytecode stub as well"


def test_stubs_only_call_exported_entry_points():
    c = open(os.path.join(OCAML, "kpc_stubs.c")).read()
    header = open(os.path.join(ROOT, "include", "kpopcount.h")).read()
    called = set(re.findall(r"\b(kpc_[a-z_]+)\(", c)) - set(re.findall(r"value (kpc_ml_\w+)\(", c))
    called = {f for f in called if not f.startswith("kpc_ml_")}
    declared = set(re.findall(r"\b(kpc_[a-z_]+)\(", header))
    assert called and called <= declared, called - declared


def test_compute_iterates_the_reference_input_list():
    """The new compute is written against what exists in the reference: a Files.Type.t list (bin/KPopCount.ml:217-238)."""
    ml = open(os.path.join(OCAML, "KPopCount_gpu.ml")).read()
    assert "Files.Type.t list" in ml and "List.iter" in ml and "iter_files" not in ml
    for ctor in ("Files.Type.FASTA", "SingleEndFASTQ", "PairedEndFASTQ"):
        assert ctor in ml


def test_every_stub_is_bound_and_every_identifier_the_driver_uses_is_declared():
    ml = open(os.path.join(OCAML, "Kpc_gpu.ml")).read()
    c = open(os.path.join(OCAML, "kpc_stubs.c")).read()
    drv = open(os.path.join(OCAML, "KPopCount_gpu.ml")).read()
    bound = set(re.findall(r"=\s*\"(kpc_ml_\w+)\"", ml))
    stubs = set(re.findall(r"^value (kpc_ml_\w+)\(", c, flags=re.M))
    assert stubs == bound, (stubs - bound, bound - stubs)
    # values, types and constructors Kpc_gpu.ml offers
    declared = set(re.findall(r"external\s+(\w+)\s*:", ml)) | set(re.findall(r"^type\s+(\w+)", ml, flags=re.M))
    for rhs in re.findall(r"^type\s+\w+\s*=\s*([A-Z][\w\s|]*)$", ml, flags=re.M):
        declared |= {x.strip() for x in rhs.split("|")}
    used = set(re.findall(r"\bKpc_gpu\.(\w+)", drv))
    assert used and used <= declared, used - declared
