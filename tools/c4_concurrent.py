#!/usr/bin/env python
"""C4 idiom on one GPU: samples are independent, so T host threads each drive their own context (own CUDA stream) and
the GPU overlaps their latency-bound kernels.  Prints genomes/s for T = 1, 2, 4, 8.  usage: c4_concurrent.py [genomes]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from bench_configs import synth_genome
from kpop_b200 import KMerCounter

G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
genomes = []
for g in range(4):
    data, n = synth_genome(g)
    genomes.append((torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory(), len(data)))

def worker(tid, T, out):
    text = torch.empty(128 << 20, dtype=torch.uint8, pin_memory=True)
    with KMerCounter(k=30, label="G") as kc:
        kc.set_text_buffer(text.data_ptr(), text.numel())
        for it in range(-2, G // T):          # two warm-up samples per context
            if it == 0:
                barrier.wait(); out[tid] = time.perf_counter()
            src, n = genomes[(tid + max(it, 0)) % len(genomes)]
            kc.reset(); kc.begin("fasta"); kc.feed_pointer(src.data_ptr(), n, eof=True); kc.end(); kc.finish()
        assert kc.text_buffer_used() > 80_000_000

for T in (1, 2, 4, 8):
    barrier = threading.Barrier(T)
    starts = [0.0] * T
    th = [threading.Thread(target=worker, args=(i, T, starts)) for i in range(T)]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - min(starts)
    n = (G // T) * T
    print(f"threads {T}: {n} genomes in {dt*1e3:.1f} ms = {n/dt:.1f} genomes/s ({dt/n*1e3:.2f} ms per genome)", flush=True)
