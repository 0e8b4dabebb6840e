#!/usr/bin/env python
"""Stall samples and executed instructions per SASS opcode from an `ncu --page source --csv --print-source cuda,sass`
export (every SASS address counted once).  usage: ncu_opcodes.py export.csv"""
import csv, collections, re, sys
def f(x):
    try: return float(x)
    except ValueError: return 0.0
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; agg = collections.Counter(); inst = collections.Counter(); seen = set()
st = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) < 60: continue
    addr = r[2]
    if addr in ("-", "...", "") or addr in seen: continue
    seen.add(addr)
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3].strip())
    if not m: continue
    op = m.group(2); base = op.split('.')[0]
    key = op[:12] if base in ("ATOMS", "STS", "LDS", "BAR", "LDG", "STG", "ATOMG", "SHFL", "REDG", "IDP") else base
    agg[key] += f(r[6]); inst[key] += f(r[7])
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h: st[key][h[6:]] += f(r[j])
tot = sum(agg.values()); ti = sum(inst.values())
print(f"total samples {tot:.0f}, warp instructions {ti:.4g}")
for op, s in agg.most_common(26):
    top = " ".join(f"{k}={int(v)}" for k, v in st[op].most_common(3))
    print(f"{op:14} samples {100*s/tot:5.1f}%  instr {100*inst[op]/ti:5.1f}%   {top}")
