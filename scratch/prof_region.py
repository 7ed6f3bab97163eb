"""Config-2 streaming build (K1, K2a, K2c, K3s) a few times, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'k2c_|k3s_region|k2a_' --launch-skip 6 -c 3 -o gpurun_out/prof_region python scratch/prof_region.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangenome_b200 import engine, _lib, builder as pgbuilder
from pangenome_b200.synth import pangenome
torch.cuda.set_device(0)
k = 27
data = pangenome(10, 5_000_000)
d = engine.to_device_bytes(data)
bld = pgbuilder.RoundBuilder(k, _lib.PG_MODE_CANONICAL, len(data))
for i in range(4):
    bld.begin()
    t = bld.build_async(engine.PackedSeqs(d, lazy=True))
    torch.cuda.synchronize()
    bld.verify()
print("ok", t.n_keys(), t.capacity)
