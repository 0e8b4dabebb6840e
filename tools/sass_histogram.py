#!/usr/bin/env python
"""Static SASS opcode histogram of the hot kernels of libkpopcount_gpu.so (cuobjdump -sass on stdin).
usage: cuobjdump -sass kpop_b200/libkpopcount_gpu.so | python tools/sass_histogram.py > profiles/<name>.txt"""
import re
import sys
from collections import Counter, OrderedDict

funcs = OrderedDict()
cur = None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
print("# SASS opcode histogram (static instruction counts) of libkpopcount_gpu.so, sm_100a cubins; cuobjdump -sass")
print("# bulk-copy / mbarrier tile staging shows as UBLKCP (cp.async.bulk), SYNCS (mbarrier), UBLKPF (bulk prefetch to L2)")
for f, c in funcs.items():
    if not any(k in f for k in ("fq_partition", "fq_count", "bucket_finalize", "BucketCount", "BucketScatter")):
        continue
    tot = sum(c.values())
    print(f"\n{f[:160]}\n  total {tot}: " + ", ".join(f"{k} {v}" for k, v in c.most_common(30)))
    print("  memory / sync ops: " + ", ".join(f"{k} {c[k]}" for k in ("UBLKCP", "SYNCS", "UBLKPF", "ATOMS", "ATOMG", "REDG", "LDS", "STS", "LDG", "STG", "BAR") if c[k]))
