#!/usr/bin/env python
"""Where the end-to-end time of one KPopCount call goes (pinned host FASTQ -> spectrum text on the host)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kpop_b200 import KMerCounter
R = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
kc = KMerCounter(k=12, label="S3", device=0)
nbytes = kc.synth_offset(R)
data = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
kc.synth_fastq(data.data_ptr(), 0, R, 3)
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.copy_(data[:nbytes]); torch.cuda.synchronize()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); kc.reset(); t1 = T(); kc.begin("single-end"); kc.feed_pointer(host.data_ptr(), nbytes, eof=True); t2 = T()
    kc.end(); t3 = T(); kc.finish(); t4 = T(); n = len(kc.take_text()); t5 = T()
    print(f"iter {it}: reset {1e3*(t1-t0):.1f} feed {1e3*(t2-t1):.1f} ({nbytes/(t2-t1)/1e9:.1f} GB/s) end {1e3*(t3-t2):.1f} finish {1e3*(t4-t3):.1f} take_text {1e3*(t5-t4):.1f} ms  text {n} B", flush=True)
