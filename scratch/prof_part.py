import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from pangenome_b200 import engine, _lib
from pangenome_b200.synth import pangenome
data = pangenome(10, 5_000_000)
p = engine.PackedSeqs(engine.to_device_bytes(data))
for i in range(3):
    t, n_rec, b = engine.build_dbg_partitioned(p, 27, capacity=1 << 25)
torch.cuda.synchronize()
print("ok", t.count())
