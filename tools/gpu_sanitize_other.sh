mkdir -p gpurun_out; rm -f gpurun_out/sanitizer_other_paths.log
gunzip -c tests/golden/inputs/clusters-small.fasta.gz | head -c 400000 > /tmp/c1.fa
python - <<'PY'
import sys
sys.path.insert(0, 'tools')
from bench_configs import synth_genome
open('/tmp/g0.fa', 'wb').write(synth_genome(0)[0][:600000])
PY
for spec in "memcheck -k 30 -l G0 -f /tmp/g0.fa" "memcheck -k 5 -L -f /tmp/c1.fa" "racecheck -k 30 -l G0 -f /tmp/g0.fa" "memcheck -k 12 -l x -f /tmp/g0.fa"; do
  set -- $spec; tool=$1; shift
  timeout 600 compute-sanitizer --tool $tool --log-file /tmp/san.log kpop_b200/bin/KPopCount "$@" > /dev/null 2>&1
  tail -n 1 /tmp/san.log | sed "s|^|[$tool $*] |" | tee -a gpurun_out/sanitizer_other_paths.log
done
