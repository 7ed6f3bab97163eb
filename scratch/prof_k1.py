import sys, torch
sys.path.insert(0, '.')
from pangenome_b200 import engine
from pangenome_b200.synth import pangenome
d = engine.to_device_bytes(pangenome(10, 5_000_000))
ev = lambda: torch.cuda.Event(enable_timing=True)
for i in range(5):
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    p = engine.PackedSeqs(d, lazy=True)
    b.record(); torch.cuda.synchronize()
    print("K1 total %.3f ms" % a.elapsed_time(b), flush=True)
print(p.n_rec, p.n_bases)
