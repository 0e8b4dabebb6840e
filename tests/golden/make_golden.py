"""How the files under tests/golden/ were made (run in the build container, where /root/reference exists).

* inputs/*.fasta.gz     gzip -9 of the reference's own test inputs (test/wuhan.fasta, test/refTB.fasta,
                        test/clusters-small.fasta): DATA files, needed because /root/reference does not exist on the
                        GPU box.  No reference source code is copied.
* fixture_digests.json  line / byte counts and md5 of the spectra the reference semantics give for those inputs.  The
                        reference itself cannot be run here (pure OCaml, no toolchain), so the digests come from two
                        independent restatements that agree: the survey's Python one and
                        oracle/kpopcount_oracle.cpp (this script re-derives them with the latter and checks).
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_TEST = "/root/reference/test"


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    oracle = os.path.join(ROOT, "oracle", "_build", "kpopcount_oracle")
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    for name in ("wuhan", "refTB", "clusters-small"):
        with open(os.path.join(REF_TEST, name + ".fasta"), "rb") as src, \
                gzip.GzipFile(os.path.join(HERE, "inputs", name + ".fasta.gz"), "wb", 9, mtime=0) as dst:
            shutil.copyfileobj(src, dst)
    with open(os.path.join(HERE, "fixture_digests.json")) as f:
        cases = json.load(f)
    for c in cases:
        argv = [a.replace("{REF_TEST}", REF_TEST) for a in c["argv"]]
        out = subprocess.run([oracle] + argv, stdout=subprocess.PIPE, check=True).stdout
        got = {"lines": out.count(b"\n"), "bytes": len(out), "md5": hashlib.md5(out).hexdigest()}
        ok = all(got[k] == c[k] for k in got)
        print(("ok   " if ok else "DIFF ") + " ".join(c["argv"]), got)
        if not ok:
            sys.exit(1)


if __name__ == "__main__":
    main()
