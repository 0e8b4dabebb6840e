#!/usr/bin/env python
"""The other configurations of BASELINE.json / SURVEY 8d beside the bench.py line: C1, C2 (real files, golden digests)
and a C4 sample (synthetic ~5 Mb genomes, k = 30, hash-table path), each timed on the GPU through the Python binding
(one context, device work + text to the host) with the oracle (C++ restatement of the reference, 1 core) beside it.
Prints one JSON line per configuration.  Run on a B200:  python tools/bench_configs.py  [n_genomes]"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle", "_build", "kpopcount_oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


_TEXT = None


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def synth_genome(g, seed=4):
    """header '>G<g> synthetic', one contig of 4,750,000 + (splitmix64(seed, g) mod 500,001) i.i.d. uniform
    ACGT bases, 70 per line (numpy's generator seeded by (seed, g) stands in for the per-word splitmix stream)."""
    import numpy as np
    n = 4_750_000 + splitmix64((seed << 32) ^ g) % 500_001
    rng = np.random.default_rng([seed, g])
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]
    full = n // 70 * 70
    body = np.concatenate([seq[:full].reshape(-1, 70), np.full((full // 70, 1), 10, dtype=np.uint8)], axis=1).tobytes()
    tail = seq[full:].tobytes() + (b"\n" if n > full else b"")
    return b">G%d synthetic\n" % g + body + tail, n


def gpu_run(k, label, per_record, kind, data, repeat=3):
    import torch
    from kpop_b200 import KMerCounter
    global _TEXT
    if _TEXT is None:
        _TEXT = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)  # caller-owned sink buffer (kpc_set_sink_buffer)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    best, text = None, None
    with KMerCounter(k=k, label="" if per_record else label) as kc:
        kc.set_text_buffer(_TEXT.data_ptr(), _TEXT.numel())
        for _ in range(repeat):
            kc.reset()
            t0 = time.perf_counter()
            kc.begin(kind)
            kc.feed_pointer(src.data_ptr(), len(data), eof=True)
            kc.end()
            kc.finish()
            n = kc.text_buffer_used()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        text = bytes(_TEXT[:n].numpy())
    return best, text


def oracle_run(args, path):
    t0 = time.perf_counter()
    p = subprocess.run([ORACLE] + args + [path], stdout=subprocess.PIPE, check=True)
    return time.perf_counter() - t0, p.stdout


def kmers_of(text):
    return sum(int(line.split(b"\t")[1]) for line in text.split(b"\n") if line and not line.startswith(b"\t"))


def main():
    n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    digests = {tuple(d["argv"][:-1]): d for d in json.load(open(os.path.join(GOLDEN, "fixture_digests.json")))}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        for name, fname, k, args, per_record in (("C1", "clusters-small.fasta", 5, ["-k", "5", "-L", "-f"], True),
                                                 ("C2", "refTB.fasta", 12, ["-k", "12", "-l", "refTB", "-f"], False)):
            data = gzip.open(os.path.join(GOLDEN, "inputs", fname + ".gz")).read()
            path = os.path.join(td, fname)
            open(path, "wb").write(data)
            t_cpu, want = oracle_run(args, path)
            t_gpu, got = gpu_run(k, "refTB", per_record, "fasta", data)
            gold = digests[tuple(args)]
            km = kmers_of(got)
            print(json.dumps({"config": name, "argv": "KPopCount " + " ".join(args) + " " + fname, "input_bytes": len(data),
                              "kmers": km, "text_bytes": len(got), "identical_to_oracle": got == want,
                              "md5": hashlib.md5(got).hexdigest(), "md5_matches_survey_digest": hashlib.md5(got).hexdigest() == gold["md5"],
                              "gpu_s": t_gpu, "gpu_kmers_per_s": km / t_gpu, "cpu_oracle_s_1core": t_cpu,
                              "cpu_kmers_per_s": km / t_cpu, "note": "pinned host bytes in, text in a pinned host buffer out; best of 3; context creation excluded"}), flush=True)
        # C4 sample: k = 30 (the reference's maximum, KMers.ml:264-267), hash-table path, one spectrum per genome
        tot_gpu = tot_cpu = 0.0
        tot_km = tot_bytes = 0
        all_same = True
        for g in range(n_genomes):
            data, n = synth_genome(g)
            t_gpu, got = gpu_run(30, "G%d" % g, False, "fasta", data, repeat=4)
            tot_gpu += t_gpu
            tot_km += kmers_of(got)
            tot_bytes += len(data)
            if g < 2:  # the oracle takes a few seconds per genome
                path = os.path.join(td, "g.fa")
                open(path, "wb").write(data)
                t_cpu, want = oracle_run(["-k", "30", "-l", "G%d" % g, "-f"], path)
                tot_cpu += t_cpu
                all_same = all_same and got == want
        print(json.dumps({"config": "C4 sample", "argv": "KPopCount -k 30 -l G<g> -f <synthetic genome g>", "genomes": n_genomes,
                          "input_bytes": tot_bytes, "kmers": tot_km, "identical_to_oracle_first_2": all_same,
                          "gpu_s_per_genome": tot_gpu / n_genomes, "gpu_kmers_per_s": tot_km / tot_gpu,
                          "cpu_oracle_s_per_genome_1core": tot_cpu / min(2, n_genomes),
                          "note": "hash-table path (generic tile kernel + atomicCAS insert + device-side ordering); samples are "
                                  "independent: C4 shards by sample with no collective"}), flush=True)


if __name__ == "__main__":
    main()
