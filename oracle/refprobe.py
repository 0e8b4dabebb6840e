"""oracle/refprobe.py -- TEST INFRASTRUCTURE ONLY.

Looks for a real (OCaml) `KPopCount` of the reference: the image has no OCaml toolchain, so one can only be there if
somebody put it under baseline/_ref/ or on PATH (bioconda package `kpop`, README.md:41-44 of the reference).  When found,
the harnesses prefer it: the tests pin the C++ oracle to it on every fixture, bench.py times it as the CPU baseline and
says `"kind": "reference"`.  The repo's own front-ends (kpop_b200/bin/KPopCount, tests/emul/_build/KPopCount_emul) are
never taken for it."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_OWN = (os.path.join(ROOT, "kpop_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"))


def _is_own(path):
    rp = os.path.realpath(path)
    return any(rp.startswith(os.path.realpath(d) + os.sep) for d in _OWN)


def _looks_like_kpopcount(path):
    try:
        p = subprocess.run([path, "-V"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=20)
    except (OSError, subprocess.TimeoutExpired):
        return False
    return p.returncode == 0 and p.stdout.strip().isdigit()


def find_reference_kpopcount():
    """Path of a reference KPopCount binary, or None."""
    cands = [os.environ.get("KPC_REFERENCE_KPOPCOUNT", ""),
             os.path.join(ROOT, "baseline", "_ref", "KPopCount"),
             os.path.join(ROOT, "baseline", "_ref", "bin", "KPopCount"),
             os.path.join(ROOT, "oracle", "_ref", "KPopCount")]
    w = shutil.which("KPopCount")
    if w:
        cands.append(w)
    for c in cands:
        if c and os.path.isfile(c) and os.access(c, os.X_OK) and not (_is_own(c) and "_ref" not in c):
            if _looks_like_kpopcount(c):
                return c
    return None


def cpu_reference():
    """(binary, kind, description) of the CPU implementation harnesses should time / check against."""
    ref = find_reference_kpopcount()
    if ref:
        return ref, "reference", f"reference KPopCount found at {ref}"
    oracle = os.path.join(ROOT, "oracle", "_build", "kpopcount_oracle")
    return oracle, "port", "oracle/kpopcount_oracle: C++ restatement of the reference (no OCaml toolchain, no reference binary found)"


if __name__ == "__main__":
    print(cpu_reference())
