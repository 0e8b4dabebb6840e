#!/bin/bash
# turns gpurun_out/fq_full.ncu-rep (tools/gpu_iter.sh) into the summaries kept under profiles/:  usage: ncu_summarise.sh <tag>
tag=${1:-r2_final}
rep=gpurun_out/fq_full.ncu-rep
ncu -i $rep --page raw --csv 2>/dev/null > /tmp/ncu_raw.csv
( echo "# ncu --set full --clock-control none --import-source on, bench.py --records-per-gpu 4000000 (first launch = 1 GiB); $tag kernels"; python tools/ncu_raw.py /tmp/ncu_raw.csv ) > profiles/${tag}_ncu_fq_summary.txt
ncu -i $rep --page source --csv --print-source cuda,sass --kernel-name regex:fq_partition 2>/dev/null > /tmp/ncu_src.csv
python tools/ncu_lines.py /tmp/ncu_src.csv 0.7 > profiles/${tag}_ncu_partition_lines.txt
python - "$tag" <<'PY'
import csv, json, sys
rows = list(csv.reader(open('/tmp/ncu_raw.csv')))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    if 'fq_partition' in d['Kernel Name']:
        def gb(x, u):
            v = float(x.replace(',', ''))
            return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]
        u = dict(zip(h, rows[1]))
        rd = gb(d['dram__bytes_read.sum'], u['dram__bytes_read.sum'])
        wr = gb(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
        json.dump({"kernel": "fq_partition_kernel<DNA-ds, k=12>", "launch_bytes": 1073741824, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_bytes_per_launch": rd + wr, "source": "ncu --set full, profiles/%s_ncu_fq_summary.txt" % sys.argv[1]},
                  open('profiles/fq_partition_traffic.json', 'w'))
        break
PY
cat profiles/fq_partition_traffic.json
