import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from pangenome_b200 import engine, _lib
from pangenome_b200.synth import pangenome
def run(name, data, k=27, cap=None, reps=5):
    d = engine.to_device_bytes(data)
    p = engine.PackedSeqs(d)
    t, n_rec = engine.build_dbg(p, k, capacity=cap)
    used, ent = t.count()
    cap = t.capacity
    ev = lambda: torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(reps):
        t.clear()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record(); t.insert(p, n_rec); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts)//2]
    npos = p.n_positions(k)
    print("%-28s positions %9d keys %9d table %7.1f MB load %.2f  insert %.3f ms  %.1f G pos/s" % (name, npos, used, cap*16/1e6, used/cap, ms, npos/ms/1e6), flush=True)
run("cfg2 (2GB table)", pangenome(10, 5_000_000))
run("cfg2 tight (512MB)", pangenome(10, 5_000_000), cap=1<<25)
run("40x0.5M snp1e-4 (32MB)", pangenome(40, 500_000, snp=0.0001), cap=1<<21)
run("40x0.5M snp1e-4 (16MB)", pangenome(40, 500_000, snp=0.0001), cap=1<<20)
run("20x1M snp1e-3 (64MB)", pangenome(20, 1_000_000, snp=0.001), cap=1<<22)
run("4x1M snp1e-2 (64MB)", pangenome(4, 1_000_000, snp=0.01), cap=1<<22)
run("1x4M unique (128MB)", pangenome(1, 4_000_000), cap=1<<23)
run("1x4M unique (2GB)", pangenome(1, 4_000_000), cap=1<<27)
