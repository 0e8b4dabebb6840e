#!/usr/bin/env python
"""One C4 genome (k = 30, hash path) through the binding: for ncu launch lists.  usage: c4_one.py [repeats]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from bench_configs import synth_genome
from kpop_b200 import KMerCounter
data, n = synth_genome(0)
src = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
text = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
with KMerCounter(k=30, label="G0") as kc:
    kc.set_text_buffer(text.data_ptr(), text.numel())
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        kc.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        kc.begin("fasta"); kc.feed_pointer(src.data_ptr(), len(data), eof=True); t1 = time.perf_counter(); kc.end(); t2 = time.perf_counter(); kc.finish()
        t3 = time.perf_counter()
        print(f"iter {it}: feed {1e3*(t1-t0):.2f} end {1e3*(t2-t1):.2f} finish {1e3*(t3-t2):.2f} ms, text {kc.text_buffer_used()} B", flush=True)
