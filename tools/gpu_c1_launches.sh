#!/bin/bash
# launch list of C1 (KPopCount -k 5 -L on clusters-small.fasta) and C2 through the CLI
mkdir -p gpurun_out
f1=$(ls tests/golden/inputs/*clusters-small* 2>/dev/null | head -n 1)
if [[ "$f1" == *.gz ]]; then gunzip -c "$f1" > /tmp/c1.fa; else cp "$f1" /tmp/c1.fa; fi
for i in 1 2 3; do kpop_b200/bin/KPopCount -k 5 -L -f /tmp/c1.fa > /tmp/c1.out; done
md5sum /tmp/c1.out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c1_launches.csv kpop_b200/bin/KPopCount -k 5 -L -f /tmp/c1.fa > /dev/null 2> gpurun_out/c1_ncu.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/c1_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i+1; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[start:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki][:80]][0]+=1; agg[r[ki][:80]][1]+=v
tot=sum(v[1] for v in agg.values())
print("kernels total us", tot/1e3, "launches", sum(v[0] for v in agg.values()))
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print(f"{v[1]/1e3:9.1f} us {v[0]:4d}x  {k}")
PY
