#!/bin/bash
# Wall time of the drop-in CLI on config 2 (FASTA file -> rows): python scratch/cli_time.sh <tag>
set -e
python - <<'PY'
from pangenome_b200.synth import pangenome
open("/tmp/cfg2.fa", "wb").write(pangenome(10, 5_000_000))
PY
for i in 1 2 3; do
  rm -f /tmp/cfg2.fa_*
  /usr/bin/time -f "cli_wall_s %e" python kmer_b200.py -m -i /tmp/cfg2.fa -k 27 --no-dump-db > /tmp/cli_out.tab 2> /tmp/cli_err.txt || true
  tail -1 /tmp/cli_err.txt; grep -c . /tmp/cli_out.tab
done
grep "^#" /tmp/cli_out.tab
