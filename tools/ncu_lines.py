#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
share of executed warp instructions and of stall samples per CUDA source line (lines with '-' address are the
CUDA lines; the SASS rows that follow them are skipped).  usage: ncu_lines.py export.csv [min_pct]"""
import csv, sys
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; out = []
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) < len(hdr) - 2 or r[2] != "-": continue
    ii = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
    st = {h: float(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    out.append((cur_file, r[0], r[1], float(r[ii] or 0), float(r[sa] or 0), st))
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total warp instructions {ti:.4g}, samples {ts:.0f}")
for f, ln, src, i, s, st in out:
    if 100 * i / ti >= thr or 100 * s / ts >= thr:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        tops = " ".join(f"{k[6:]}={v:.0f}" for k, v in top if v)
        print(f"{f[:14]:14}:{ln:>4} {100*i/ti:5.1f}%i {100*s/ts:5.1f}%s  {src.strip()[:90]:90} {tops}")
