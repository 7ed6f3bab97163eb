import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from pangenome_b200 import engine, _lib
from pangenome_b200.engine import _ptr, _stream
from pangenome_b200.synth import pangenome
L = _lib.load()
data = pangenome(10, 5_000_000)
d = engine.to_device_bytes(data)
p = engine.PackedSeqs(d)
k = 27; n_rec = p.n_rec; npos = p.n_positions(k)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(f, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); a, b = ev(), ev(); a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
for cap in (1 << 27, 1 << 25):
    t = engine.DbgTable(cap, k, 2)
    clear_ms = timeit(t.clear)
    def fused():
        t.clear(); t.insert(p, n_rec)
    f_ms = timeit(fused) - clear_ms
    print("cap 2^%d  clear %.3f ms  fused insert %.3f ms (%.1f G pos/s)" % (int(np.log2(cap)), clear_ms, f_ms, npos / f_ms / 1e6), flush=True)
    for sub_bytes in (128 << 20, 64 << 20, 32 << 20, 16 << 20, 8 << 20, 2 << 20):
        sb = engine.sub_bits_for(cap, sub_bytes)
        bk = engine.partition_kmers(p, k, 2, n_rec, 0, sb)
        part_ms = timeit(lambda: engine.partition_kmers(p, k, 2, n_rec, 0, sb, buckets=bk))
        def ins():
            engine.check(L.pg_insert_records(ctypes.byref(t.c), _ptr(bk.records), _ptr(bk.seg_off), _ptr(bk.counts), bk.n_parts, 1, bk.part_cap, _stream()), "ins")
        def both():
            t.clear(); ins()
        ins_ms = timeit(both) - clear_ms
        print("   sub_bytes %4d MB  parts %4d  partition %.3f ms  insert_records %.3f ms  total %.3f ms (%.1f G pos/s)" % (
            sub_bytes >> 20, bk.n_parts, part_ms, ins_ms, part_ms + ins_ms, npos / (part_ms + ins_ms) / 1e6), flush=True)
        assert not t.overflowed()
