"""Shared pytest plumbing.

* registers the ``gpu`` marker (the driver runs ``-m "not gpu"`` on a CPU box and ``-m gpu`` on a B200);
* builds the oracle (test infrastructure, see oracle/*.cpp headers) once per session;
* exposes the small case tables (KATs) used both by the oracle pin tests and by the GPU parity tests.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_BIN = os.path.join(ORACLE_DIR, "_build", "kpopcount_oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_bin():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    assert os.path.exists(ORACLE_BIN)
    return ORACLE_BIN


def run_cli(binary, args, stdin=None, env=None):
    """Run a KPopCount-compatible executable; returns (exit code, stdout bytes, stderr bytes)."""
    p = subprocess.run([binary] + list(args), input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return p.returncode, p.stdout, p.stderr
