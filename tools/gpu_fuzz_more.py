#!/usr/bin/env python
"""Extra fuzz seeds beyond the test-suite's (same generators): CLI vs oracle on small adversarial cases, and the fast
FASTQ pipeline vs oracle on large ones.  usage: gpu_fuzz_more.py [first_seed] [n_seeds]"""
import os, random, sys, tempfile, pathlib, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import subprocess
subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
from conftest import run_cli, ORACLE_BIN
from test_gpu_parity import one_case, env_chunk, GPU_BIN
from test_gpu_fastq_fast import big_fastq, count_with_binding, SHAPES
first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bad = 0
cases = 0
with tempfile.TemporaryDirectory() as td:
    td = pathlib.Path(td)
    for seed in range(first, first + n):
        rng = random.Random(5000 + seed)
        for idx in range(25):
            argv, fmt = one_case(rng, td, idx)
            chunk = rng.choice([None, 4096, 70_000])
            rc_o, out_o, _ = run_cli(ORACLE_BIN, argv)
            rc_g, out_g, err_g = run_cli(GPU_BIN, argv, env=env_chunk(chunk))
            if rc_g == 2 and b"code -9" in err_g:
                continue
            cases += 1
            if (rc_g, out_g) != (rc_o, out_o):
                bad += 1
                print("MISMATCH small", seed, idx, chunk, " ".join(argv), flush=True)
        shape = SHAPES[seed % len(SHAPES)]
        rng = random.Random(zlib.crc32(shape.encode()) + seed)
        data = big_fastq(rng, shape)
        p = td / f"big{seed}.fq"; p.write_bytes(data)
        k = rng.choice([12, 12, 11, 9, 8, 5, 4]); content = rng.choice(["DNA-ds", "DNA-ds", "DNA-ss"])
        if content == "DNA-ss" and k == 12: k = 11
        rc_o, out_o, _ = run_cli(ORACLE_BIN, ["-k", str(k), "-C", content, "-l", "x", "-s", str(p)])
        for chunk, dev, launch in [(None, False, None), (None, True, 65536), (None, True, None)]:
            rc_f, out_f = count_with_binding(data, k, content, "x", chunk, True, dev, launch)
            if rc_f == -9: continue
            cases += 1
            if (rc_f, out_f) != (0, out_o):
                bad += 1
                print("MISMATCH big", seed, shape, k, content, chunk, dev, launch, flush=True)
print(f"fuzz: {cases} cases, {bad} mismatches (seeds {first}..{first + n - 1})")
