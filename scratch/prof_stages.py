"""One pass of every stage on config 2 for an ncu capture of the kernels after the dBG build (K4..K8) and K1:
    ncu --set full --clock-control none --import-source on -k regex:'k1_|k4_|k5_|k6_|k7_|k8_' -o gpurun_out/prof_stages python scratch/prof_stages.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangenome_b200 import engine, _lib, builder as pgbuilder, graph
from pangenome_b200.synth import pangenome
torch.cuda.set_device(0)
k = 27
data = pangenome(10, 5_000_000)
packed = engine.PackedSeqs(engine.to_device_bytes(data))
table, n_rec, b = pgbuilder.build_table(packed, k)
res = graph.seq2graph_device(packed, table.select_rdbg(), k, Ns=2 ** 62, rc=False)
rows = res.rows(packed, data)
torch.cuda.synchronize()
print("ok", table.n_keys(), len(rows))
