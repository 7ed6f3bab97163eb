import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from pangenome_b200 import engine, _lib
from pangenome_b200.engine import _ptr, _stream
from pangenome_b200.synth import pangenome
L = _lib.load()
data = pangenome(10, 5_000_000)
p = engine.PackedSeqs(engine.to_device_bytes(data))
k = 27; n_rec = p.n_rec; npos = p.n_positions(k)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(f, reps=7):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); a, b = ev(), ev(); a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
import os
out = []
for cap, sub in ((1 << 25, 32 << 20), (1 << 25, 8 << 20), (1 << 27, 128 << 20), (1 << 27, 8 << 20)):
    t = engine.DbgTable(cap, k, 2)
    clear_ms = timeit(t.clear)
    sb = engine.sub_bits_for(cap, sub)
    bk = engine.partition_kmers(p, k, 2, n_rec, 0, sb)
    def both():
        t.clear(); engine.check(L.pg_insert_records(ctypes.byref(t.c), _ptr(bk.records), _ptr(bk.seg_off), _ptr(bk.counts), bk.n_parts, 1, bk.part_cap, _stream()), "ins")
    ms = timeit(both) - clear_ms
    out.append("cap2^%d/%dparts %.3f" % (int(np.log2(cap)), bk.n_parts, ms))
print("ILP", os.environ.get("PG_K3_ILP"), "EVICT", os.environ.get("PG_K3_EVICT"), "GRID", os.environ.get("PG_K3_GRID"), " | ".join(out), flush=True)
