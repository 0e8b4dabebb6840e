#!/usr/bin/env python
"""bench.py -- k-mers counted per second (bit-exact KPopCount tables) at k = 12 on N B200s, with the HBM roofline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
Under torchrun (N > 1) every rank owns one GPU and one shard of the reads; the 4^12 tables are summed with an NCCL
all-reduce over NVLink before rank 0 formats the spectrum.

Workload (BASELINE.json configs[2] / configs[4] shape, SURVEY.md 8d "C3"): synthetic single-end 150 bp FASTQ,
record i = "@S<i>\\n" + 150 bases + "\\n+\\n" + 150 x 'I' + "\\n", bases i.i.d. uniform ACGT with N at p = 2^-10, seed 3,
the first whole records that fit in 10^10 bytes (31,781,305 reads = 9,999,999,965 B) PER GPU; rank r counts records
[r * R, (r + 1) * R) of the same stream ("scaling": "weak").  k = 12, DNA-ds, one label: the dense 4^12 table path.

A step = reset the table + frame / lint / roll / count every byte of the shard + (N > 1: all-reduce) + compact and
format the table on the device.
  value  : device-resident input, CUDA events on the library's stream, max over ranks.
  e2e    : the same call sequence through kpc_feed() from PINNED HOST memory, text copied back to the host and
           handed to a Python sink, wall clock between device synchronisations, max over ranks.
  roofline: the framing/counting kernel alone: algorithmic bytes (every input byte once) / its CUDA-event time.
  cpu_baseline / --impl reference: the oracle (C++ restatement of the reference; the reference is OCaml and
           cannot be built in this image) on a bounded prefix of the same stream, 1 thread -- the reference is
           single-threaded per sample (README.md:593).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3_RECORDS = 31_781_305          # first whole records <= 10^10 bytes
C5_RECORDS = 316_807_313         # first whole records <= 10^11 bytes
SEED = 3
K = 12
ORACLE = os.path.join(ROOT, "oracle", "_build", "kpopcount_oracle")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from refprobe import cpu_reference  # noqa: E402  (a real KPopCount under baseline/_ref or on PATH wins over the C++ port)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5"],
                    help="c3: 10 GB of reads per GPU, k = 12 (the headline, weak scaling); c5: ONE 100 GB stream split across the "
                         "GPUs (strong scaling); c4: ~5 Mb genomes at k = 30, sharded by sample")
    ap.add_argument("--records-per-gpu", type=int, default=None)
    ap.add_argument("--genomes-per-gpu", type=int, default=16)
    ap.add_argument("--contexts", type=int, default=2,
                    help="c4: library contexts per GPU, each fed by its own host thread (samples are independent)")
    ap.add_argument("--cpu-sample-records", type=int, default=1_500_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in a thread every 5 ms
    (nvidia-smi -lms as a fallback when the NVML binding is missing)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reason_bits = [], [], 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            try:
                self.mx = [self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)]
            except Exception:
                self.mx = []
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1)
            nv, bits = self.nv, self.reason_bits
            names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                     ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                     ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
            reasons = [n for n, attr in names if bits & int(getattr(nv, attr, 0))]
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx[0] if self.mx else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a bounded prefix of the same stream
# ------------------------------------------------------------------------------------------------------------
SYNTH = os.path.join(ROOT, "oracle", "_build", "synth_fastq")


def synth_prefix_to_file(n_records, first_record, path):
    """Host generator (oracle/synth_fastq.cpp; same definition as the device generator, kpc_synth.h)."""
    if not os.path.exists(SYNTH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    with open(path, "wb") as f:
        subprocess.run([SYNTH, str(first_record), str(n_records), str(SEED)], stdout=f, check=True)
    return os.path.getsize(path)


def run_cpu_baseline(sample_path, sample_records):
    """Times the CPU reference (a real KPopCount if one is found, else the oracle port; 1 thread: the reference is
    single-threaded per sample) on the sample; returns (kmers/s, kmers, seconds)."""
    if not os.path.exists(ORACLE):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    binary, _kind, _desc = cpu_reference()
    t0 = time.perf_counter()
    p = subprocess.run([binary, "-k", str(K), "-l", "S3", "-s", sample_path], stdout=subprocess.PIPE, check=True)
    dt = time.perf_counter() - t0
    kmers = 0
    for line in p.stdout.split(b"\n")[1:]:
        if line:
            kmers += int(line.split(b"\t")[1])
    return kmers / dt, kmers, dt


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned buffer is allocated: pinned
    pages are then local to the PCIe root the copies go through (at N = 8 the ranks otherwise share whatever node the
    allocations happen to land on).  Returns a small record for the JSON line; harmless when /sys has no topology."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"gpu": bdf, "node": node, "bound": False}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"gpu": bdf, "node": node, "cpus": len(allowed), "bound": bool(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": str(e)[:80]}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)

    if args.workload == "c4":
        import bench_c4
        return bench_c4.main(args, rank, world, local_rank, n_gpus)
    if args.impl == "reference":
        return reference_arm(args, rank, world, n_gpus)

    import torch
    import torch.distributed as dist
    from kpop_b200 import KMerCounter

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if os.environ.get("KPC_BENCH_NUMA", "1") != "0" else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # c3: every rank counts its own 10 GB of the seed-3 stream (weak); c5: ONE seed-5 stream of 100 GB cut into `world`
    # contiguous record ranges (strong: read-chunk sharding)
    c5 = args.workload == "c5"
    seed = 5 if c5 else SEED
    if c5:
        total = args.records_per_gpu * world if args.records_per_gpu else C5_RECORDS
        first = rank * total // world
        R = (rank + 1) * total // world - first
    else:
        R = args.records_per_gpu or C3_RECORDS
        first = rank * R
    label = "S5" if c5 else "S3"
    kc = KMerCounter(k=K, label=label, device=local_rank)
    nbytes = kc.synth_offset(first + R) - kc.synth_offset(first)
    data = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    assert data.data_ptr() % 16 == 0
    kc.synth_fastq(data.data_ptr(), first, R, seed)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(kc.stream_handle())
    lo_ptr, _hi, nbins = kc.dense_table()

    class _Dev:  # zero-copy view of the library's table for torch.distributed
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

    table32 = torch.as_tensor(_Dev(lo_ptr, nbins, "<i4"), device="cuda")
    table_bytes = nbins * 4
    scalar = torch.zeros(1, dtype=torch.int64, device="cuda")

    def reduce_tables():
        """Sum the per-rank 4^k tables into every rank (NCCL over NVLink).  u32 unless a bin could wrap."""
        if world == 1:
            return
        cur = torch.cuda.current_stream()
        mx = kc.dense_max()
        scalar.fill_(mx)
        dist.all_reduce(scalar)
        if int(scalar.item()) < (1 << 32):
            cur.wait_stream(stream)
            dist.all_reduce(table32)
            stream.wait_stream(cur)
        else:
            kc.dense_promote()
            _lo, hi_ptr, _n = kc.dense_table()
            t64 = torch.as_tensor(_Dev(hi_ptr, nbins, "<i8"), device="cuda")
            cur.wait_stream(stream)
            dist.all_reduce(t64)
            stream.wait_stream(cur)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm -------------------------------------------------------------------
    kc.discard_text(True)

    def step_device(ev_k0=None, ev_k1=None):
        kc.reset()
        kc.begin("single-end")
        if ev_k0 is not None:
            ev_k0.record(stream)
        kc.feed_device(data.data_ptr(), nbytes, eof=True)
        if ev_k1 is not None:
            ev_k1.record(stream)
        kc.end()
        reduce_tables()
        if rank == 0:
            kc.finish()

    for _ in range(args.warmup):
        step_device()
    barrier()
    kmers_rank = kc.kmers_counted() if world == 1 else None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = kc.kernel_launches()
    kc.profile_enable(True)   # CUDA events around every partition / count launch, on the library's stream
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step_device(*kev[i])
    ev1.record(stream)
    barrier()
    launches = kc.kernel_launches() - launches0
    part_ms, count_ms, fq_launches, fq_bytes = kc.profile_read()
    kc.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    ms_kernel = float(t[1])

    # k-mers counted by the whole job: sum of the (reduced) table on rank 0
    if world > 1:
        kc.reset(); kc.begin("single-end"); kc.feed_device(data.data_ptr(), nbytes, eof=True); kc.end()
        reduce_tables()
    total_kmers = kc.kmers_counted()
    text_bytes = 0
    if rank == 0:
        kc.finish()
        text_bytes = kc.text_bytes()

    # ---------------- end-to-end arm (pinned host -> spectra text on the host) --------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.copy_(data[:nbytes])
        torch.cuda.synchronize()
        kc.discard_text(False)
        got = [0]
        # the spectrum text lands in a pinned host buffer owned by the caller (kpc_set_sink_buffer): what an embedding
        # host (OCaml Bigarray, C) would do; rank 0 reads its length and touches the last line
        text_host = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True) if rank == 0 else None
        if rank == 0:
            kc.set_text_buffer(text_host.data_ptr(), text_host.numel())

        def step_e2e():
            kc.reset()
            kc.begin("single-end")
            kc.feed_pointer(host.data_ptr(), nbytes, eof=True)
            kc.end()
            reduce_tables()
            if rank == 0:
                kc.finish()
                got[0] = kc.text_buffer_used()
                assert got[0] > 0 and int(text_host[got[0] - 1]) == 10  # the text ends with a line feed

        step_e2e()
        n_e2e = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        # what the link gives: a plain pinned -> device copy of 1 GiB (explains e2e: the path is H2D bound)
        probe_n = min(int(nbytes), 1 << 30)
        h2d_gbs = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            tp0 = time.perf_counter()
            data[:probe_n].copy_(host[:probe_n], non_blocking=True)
            torch.cuda.synchronize()
            h2d_gbs = max(h2d_gbs, probe_n / (time.perf_counter() - tp0) / 1e9)
        e2e = {"value": total_kmers / dt, "unit": "k-mers/s", "h2d_bytes_per_step": int(nbytes) * world,
               "d2h_bytes_per_step": int(got[0]), "ms_per_step": dt * 1e3, "steps": n_e2e,
               "h2d_gbs_achieved": nbytes / dt / 1e9, "h2d_gbs_plain_copy": h2d_gbs, "numa": numa}
        if rank == 0:
            kc.set_text_buffer(0, 0)
        del host

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            sp = os.path.join(td, "sample.fq")
            sbytes = synth_prefix_to_file(args.cpu_sample_records, 0, sp)
            rate, km, secs = run_cpu_baseline(sp, args.cpu_sample_records)
        _bin, kind, desc = cpu_reference()
        cpu = {"value": rate, "unit": "k-mers/s", "cores": 1, "kind": kind,
               "sample": f"first {args.cpu_sample_records} reads ({sbytes} B, {km} k-mers) of the same stream, "
                         f"-k 12 -l S3 -s, {secs:.1f} s, host has {host_cores()} cores; {desc}; "
                         "the reference is single-threaded per sample (README.md:593)"}

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            with open(pk) as f:
                peaks = json.load(f)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: fq_partition_kernel (framing + linting + canonical k-mers + bucket appends); every input byte
        # is algorithmic for it, its time is the sum of its launches' CUDA-event times over the timed steps
        launch_ms = part_ms / max(1, fq_launches)
        launch_bytes = fq_bytes / max(1, fq_launches)
        achieved = launch_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
        step_alg = (nbytes + 2 * table_bytes) / (ms_step * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "fq_partition_traffic.json")
        if os.path.exists(tp):  # dram read + write of one launch from the committed ncu --set full capture
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_launch"] * (launch_bytes / tj["launch_bytes"])
        out = {
            "metric": "k-mers counted/sec (bit-exact) at k=12", "value": total_kmers / (ms_step * 1e-3),
            "unit": "k-mers/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if c5 else "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": ("C5: ONE synthetic 150bp SE FASTQ stream, seed 5, %d reads cut into %d contiguous record ranges "
                                    "(%d reads, %d B on rank 0), KPopCount -k 12 -l S5 -s, dense 4^12 table%s"
                                    % (R * world if args.records_per_gpu else C5_RECORDS, world, R, nbytes,
                                       ", NCCL all-reduce of the tables" if world > 1 else "")) if c5 else
                                   "%s: synthetic 150bp SE FASTQ, seed 3, %d reads (%d B) per GPU, KPopCount -k 12 -l S3 -s, "
                                   "dense 4^12 table%s" % ("C3" if R == C3_RECORDS else "C3 shape, other size", R, nbytes,
                                                           ", NCCL all-reduce of the tables" if world > 1 else ""),
                       "k": K, "records_per_gpu": R, "bytes_per_gpu": int(nbytes), "kmers_total": int(total_kmers),
                       "l2": "input (%.1f GB per GPU) is far larger than L2; no flush needed" % (nbytes / 1e9),
                       "spectrum_text_bytes": int(text_bytes)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
                         "kernel": "fq_partition_kernel<DNA-ds, k=12>", "algorithmic_bytes_per_launch": int(launch_bytes),
                         "launch_ms": launch_ms, "launches_per_step": fq_launches / max(1, args.steps),
                         "partition_ms_per_step": part_ms / max(1, args.steps),
                         "count_ms_per_step": count_ms / max(1, args.steps),
                         "feed_ms_per_step": ms_kernel, "step_algorithmic_gbs": step_alg,
                         "traffic_source": "profiles/fq_partition_traffic.json (ncu dram read+write of one launch, scaled)" if traffic else None},
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu,
        }
        if kmers_rank is not None:
            out["config"]["kmers_per_gpu"] = int(kmers_rank)
        print(json.dumps(out), flush=True)
    kc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reference_arm(args, rank, world, n_gpus):
    """The reference's own CPU path (oracle port, 1 thread) on a bounded sample of the same workload."""
    if rank != 0:
        return 0
    rates = []
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        sp = os.path.join(td, "sample.fq")
        # bounded: the whole run stays near four minutes (about 1e5 reads/s on one core), and a step is never smaller than
        # 1 M reads, so that the fixed costs of one process (the 2^24-bucket table of Hashtbl.create, the final bucket
        # walk) stay below a few percent of a step, as they are on the full-size input
        per_step = max(1_000_000, min(args.cpu_sample_records, 24_000_000 // max(1, args.steps + args.warmup)))
        sbytes = synth_prefix_to_file(per_step, 0, sp)
        km = 0
        t_all = 0.0
        for i in range(args.warmup + args.steps):
            rate, km, secs = run_cpu_baseline(sp, per_step)
            if i >= args.warmup:
                rates.append(rate)
                t_all += secs
    value = sum(rates) / len(rates)
    _bin, kind, desc = cpu_reference()
    out = {"impl": "reference", "metric": "k-mers counted/sec (bit-exact) at k=12", "value": value, "unit": "k-mers/s",
           "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int63", "data": "synthetic",
           "config": {"workload": "C3: synthetic 150bp SE FASTQ, seed 3, KPopCount -k 12 -l S3 -s, dense 4^12 table; each step = "
                                  f"the first {per_step} reads ({sbytes} B, {km} k-mers) of the stream (bounded sample: the rate "
                                  "does not depend on the size once the table is warm)",
                      "k": K, "sample_reads": per_step, "sample_bytes": sbytes, "sample_kmers": km},
           "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": 1, "kind": kind,
                            "sample": f"{per_step} reads per step; {desc}; single-threaded per sample (README.md:593); "
                                      f"host has {host_cores()} cores"},
           "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
