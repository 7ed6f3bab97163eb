#!/bin/bash
# Wall time of the drop-in CLI on config 2 (FASTA file -> rows on stdout, .xyz + .mcl side files): bash scratch/cli_time.sh
python - <<'PY'
from pangenome_b200.synth import pangenome
open("/tmp/cfg2.fa", "wb").write(pangenome(10, 5_000_000))
PY
python - <<'PY'
# a fresh process per run: python start-up + torch import + CUDA context + the pipeline
import glob, os, subprocess, sys, time
for i in range(3):
    for f in glob.glob("/tmp/cfg2.fa_*"): os.remove(f)
    t0 = time.time()
    with open("/tmp/cli_out.tab", "w") as out:
        rc = subprocess.call([sys.executable, "kmer_b200.py", "-m", "-i", "/tmp/cfg2.fa", "-k", "27", "--no-dump-db"], stdout=out)
    print("cli_wall_s %.3f rc=%d rows=%d" % (time.time() - t0, rc, sum(1 for l in open("/tmp/cli_out.tab") if not l.startswith("#"))))
PY
grep "^#" /tmp/cli_out.tab
python - <<'PY'
# the same inside one process (python start-up, torch import and CUDA context excluded): what a long-running host pays per file
import io, sys, time
sys.argv = ["kmer_b200.py"]
import torch
from pangenome_b200 import cli
torch.cuda.init(); torch.zeros(1, device="cuda")
for i in range(3):
    import glob, os
    for f in glob.glob("/tmp/cfg2.fa_*"): os.remove(f)
    out = io.StringIO()
    t0 = time.time()
    cli.entry_point(["kmer_b200.py", "-m", "-i", "/tmp/cfg2.fa", "-k", "27", "--no-dump-db"], out=out)
    torch.cuda.synchronize()
    print("cli_in_process_s %.3f rows=%d" % (time.time() - t0, sum(1 for l in out.getvalue().splitlines() if not l.startswith("#"))))
    print("   ", [l for l in out.getvalue().splitlines() if l.startswith("# finished")])
PY
