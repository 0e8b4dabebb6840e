"""bench_c4.py -- `bench.py --workload c4`: BASELINE.json configs[3] (SURVEY.md 8d "C4").

Synthetic bacterial genomes (one contig of 4,750,000 + splitmix64(seed 4, g) mod 500,001 i.i.d. ACGT bases, 70 per line,
header ">G<g> synthetic"), one spectrum per genome at the reference's maximum k (30: KMers.ml:264-267), i.e. what
`Parallel ... KPopCount -k 30 -l G<g> -f G<g>.fa` does with one process per sample (README.md:579,1020).  Samples are
independent: rank r takes genomes g = r, r + N, ... ("sharded by sample"); there is no collective on the data path.

A step = `--genomes-per-gpu` genomes per rank, each one begin / feed / end / finish through the C ABI on ONE context
(kpc_reset_label between samples).
  value   : genome bytes resident in HBM (kpc_feed_device), spectra formatted on the device and left there; CUDA events
            on the library's stream, max over ranks.
  e2e     : genome bytes in pinned host memory (kpc_feed), spectrum text copied back into a pinned host buffer.
  roofline: algorithmic bytes per genome = N_in + 2 * distinct * 12 over the kernel time of the step.
  parity  : on rank 0 the text of >= 16 genomes (e2e path) is byte-compared with the CPU reference run beside it
            (`nproc` concurrent single-sample processes: the reference's own scaling idiom); that run is the cpu_baseline.
"""
import concurrent.futures
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
K = 30
SEED = 4


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def synth_genome(g, seed=SEED):
    """One C4 genome as FASTA bytes (numpy's generator seeded by (seed, g) stands in for the per-word splitmix stream)."""
    import numpy as np
    n = 4_750_000 + splitmix64((seed << 32) ^ g) % 500_001
    rng = np.random.default_rng([seed, g])
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]
    full = n // 70 * 70
    body = np.concatenate([seq[:full].reshape(-1, 70), np.full((full // 70, 1), 10, dtype=np.uint8)], axis=1).tobytes()
    tail = seq[full:].tobytes() + (b"\n" if n > full else b"")
    return b">G%d synthetic\n" % g + body + tail, n


def _cpu_one(args):
    binary, path, label = args
    t0 = time.perf_counter()
    out = subprocess.run([binary, "-k", str(K), "-l", label, "-f", path], stdout=subprocess.PIPE, check=True).stdout
    return hashlib.md5(out).hexdigest(), len(out), out.count(b"\n") - 1, time.perf_counter() - t0


def cpu_leg(genomes, labels, workers):
    """The reference's scaling idiom: `workers` concurrent single-sample processes.  Returns (per-genome results, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from refprobe import cpu_reference
    binary, kind, desc = cpu_reference()
    if not os.path.exists(binary):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        jobs = []
        for (data, _n), lab in zip(genomes, labels):
            p = os.path.join(td, lab + ".fa")
            with open(p, "wb") as f:
                f.write(data)
            jobs.append((binary, p, lab))
        t0 = time.perf_counter()
        with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as ex:
            res = list(ex.map(_cpu_one, jobs))
        dt = time.perf_counter() - t0
    return res, dt, kind, desc


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def main(args, rank, world, local_rank, n_gpus):
    if args.impl == "reference":
        return reference_arm(args, rank, n_gpus)
    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from kpop_b200 import KMerCounter

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    G = args.genomes_per_gpu
    ids = [rank + world * i for i in range(G)]
    genomes = [synth_genome(g) for g in ids]
    labels = ["G%d" % g for g in ids]
    # C library contexts per GPU (--contexts): samples are independent, so several small samples can be in flight at once
    # (each context has its own streams and is driven by its own host thread; ctypes calls release the GIL)
    from concurrent.futures import ThreadPoolExecutor
    C = max(1, min(int(getattr(args, "contexts", 1) or 1), G))
    kcs = [KMerCounter(k=K, label=labels[0], device=local_rank) for _ in range(C)]
    kc = kcs[0]
    streams = [torch.cuda.ExternalStream(x.stream_handle()) for x in kcs]
    stream = streams[0]
    pool = ThreadPoolExecutor(C) if C > 1 else None
    dev = []
    for data, _n in genomes:
        t = torch.empty(len(data) + 64, dtype=torch.uint8, device="cuda")
        t[: len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        dev.append(t)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ---------------------------------------------------------------------
    for x in kcs:
        x.discard_text(True)

    def run_samples(c):
        torch.cuda.set_device(local_rank)
        x = kcs[c]
        for i in range(c, G, C):
            x.reset_label(labels[i])
            x.begin("fasta")
            x.feed_device(dev[i].data_ptr(), len(genomes[i][0]), eof=True)
            x.end()
            x.finish()

    def step_device():
        if pool is None:
            run_samples(0)
        else:
            list(pool.map(run_samples, range(C)))

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(C)]
    launches0 = sum(x.kernel_launches() for x in kcs)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)  # the device is idle here (barrier above): every context's work starts after this point
    for _ in range(args.steps):
        step_device()
    for c in range(C):
        ev1[c].record(streams[c])
    barrier()
    wall = time.perf_counter() - t0
    launches = sum(x.kernel_launches() for x in kcs) - launches0
    clocks = sampler.stop() if rank == 0 else None
    text_bytes_step = kc.text_bytes()  # of the last sample; per-sample sizes come from the e2e arm below
    ms = torch.tensor([max(ev0.elapsed_time(e) for e in ev1), wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms[0]) / args.steps
    wall_ms_step = float(ms[1]) / args.steps

    # ---------------- end-to-end arm: pinned host bytes in, spectrum text in a pinned host buffer out ------------
    kc.discard_text(False)
    text_host = torch.empty(192 << 20, dtype=torch.uint8, pin_memory=True)
    kc.set_text_buffer(text_host.data_ptr(), text_host.numel())
    pinned = []
    for data, _n in genomes:
        t = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        pinned.append(t)
    digests, d2h = [None] * G, [0]

    def step_e2e(keep):
        d2h[0] = 0
        for i in range(G):
            kc.reset_label(labels[i])
            kc.begin("fasta")
            kc.feed_pointer(pinned[i].data_ptr(), len(genomes[i][0]), eof=True)
            kc.end()
            kc.finish()
            n = kc.text_buffer_used()
            d2h[0] += n
            assert n > 0 and int(text_host[n - 1]) == 10
            if keep:
                buf = bytes(text_host[:n].numpy())
                digests[i] = (hashlib.md5(buf).hexdigest(), n, buf.count(b"\n") - 1)

    step_e2e(True)
    n_e2e = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e(False)
    barrier()
    dt = (time.perf_counter() - t0) / n_e2e
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])

    # k-mers of the step (all ranks): every window of a genome is valid (no N): n - k + 1 per genome
    kmers = torch.tensor([sum(n - K + 1 for _d, n in genomes)], dtype=torch.float64, device="cuda")
    in_bytes = torch.tensor([sum(len(d) for d, _n in genomes)], dtype=torch.float64, device="cuda")
    distinct = torch.tensor([sum(x[2] for x in digests)], dtype=torch.float64, device="cuda")
    if world > 1:
        for t in (kmers, in_bytes, distinct):
            dist.all_reduce(t)
    kmers, in_bytes, distinct = float(kmers[0]), float(in_bytes[0]), float(distinct[0])

    # ---------------- CPU reference beside it + full-text parity (rank 0, N = 1 only) ---------------------------
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        workers = max(1, min(host_cores(), G))
        res, secs, kind, desc = cpu_leg(genomes, labels, workers)
        same = sum(1 for (m, n, _d), r in zip(digests, res) if (m, n) == (r[0], r[1]))
        parity = {"genomes_compared": G, "identical": same}
        km = sum(n - K + 1 for _d, n in genomes)
        cpu = {"value": km / secs, "unit": "k-mers/s", "cores": workers, "kind": kind,
               "sample": f"{G} genomes ({sum(len(d) for d, _ in genomes)} B) as {workers} concurrent single-sample processes "
                         f"(-k 30 -l G<g> -f), {secs:.1f} s wall, {sum(r[3] for r in res) / G:.2f} s per genome on one core; "
                         f"{desc}; host has {host_cores()} cores"}
        assert same == G, f"spectra differ from the CPU reference: {same} of {G} identical"

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            with open(pk) as f:
                peaks = json.load(f)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg = in_bytes + 2.0 * distinct * 12.0          # per step, all ranks
        achieved = alg / world / (ms_step * 1e-3) / 1e9  # per GPU
        out = {
            "metric": "k-mers counted/sec (bit-exact) at k=30, samples sharded across GPUs", "value": kmers / (ms_step * 1e-3),
            "unit": "k-mers/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"C4: synthetic bacterial genomes (4.75-5.25 Mb, seed 4), {G} per GPU per step, genome g on GPU "
                                   f"g mod {world}, KPopCount -k 30 -l G<g> -f per sample (sort path: no hash table)",
                       "k": K, "genomes_per_gpu": G, "contexts_per_gpu": C, "bytes_per_step": int(in_bytes), "kmers_per_step": int(kmers),
                       "distinct_per_step": int(distinct), "ms_per_genome": ms_step / G, "wall_ms_per_genome": wall_ms_step / G,
                       "l2": "every genome (5 MB) fits L2; 16 different genomes + their 16 MiB bucket tables and ~170 MB of "
                             "entries per sample cycle through it between repeats"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
                         "kernel": "whole step (framing + staging pass, scan, streaming scatter, bucket merge, formatter): see profiles/ for the launch list",
                         "algorithmic_bytes_per_genome": alg / (G * world)},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": kmers / dt, "unit": "k-mers/s", "h2d_bytes_per_step": int(in_bytes),
                    "d2h_bytes_per_step": int(d2h[0]) * world, "ms_per_step": dt * 1e3, "ms_per_genome": dt * 1e3 / G},
            "cpu_baseline": cpu, "parity": parity,
        }
        print(json.dumps(out), flush=True)
    for x in kcs:
        x.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reference_arm(args, rank, n_gpus):
    """The reference's own CPU path for C4: `nproc` concurrent single-sample KPopCount processes."""
    if rank != 0:
        return 0
    G = max(host_cores(), 8)
    genomes = [synth_genome(g) for g in range(min(G, 64))]
    labels = ["G%d" % g for g in range(len(genomes))]
    workers = max(1, min(host_cores(), len(genomes)))
    rates, secs_all = [], 0.0
    for i in range(args.warmup + args.steps):
        res, secs, kind, desc = cpu_leg(genomes, labels, workers)
        if i >= args.warmup:
            rates.append(sum(n - K + 1 for _d, n in genomes) / secs)
            secs_all += secs
    value = sum(rates) / len(rates)
    out = {"impl": "reference", "metric": "k-mers counted/sec (bit-exact) at k=30, samples sharded across GPUs", "value": value,
           "unit": "k-mers/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": secs_all / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int63", "data": "synthetic",
           "config": {"workload": f"C4: {len(genomes)} synthetic bacterial genomes per step as {workers} concurrent single-sample "
                                  "processes (-k 30 -l G<g> -f)", "k": K},
           "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": workers, "kind": kind, "sample": desc},
           "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0
