"""Where should the table clear go?  serial / at begin (overlaps K1) / after K1 (overlaps K2a) /
one build ahead overlapping K3 (two tables)."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from pangenome_b200 import engine, _lib
from pangenome_b200.engine import _ptr, _stream, check, PgTable
import bench as B

data, _wl = B.workload("cfg2")
k = 27
d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda(); torch.cuda.synchronize()
packed = engine.PackedSeqs(d)
bd = engine.TwoPhaseBuilder(k, _lib.PG_MODE_CANONICAL, packed.n_positions(k), estimate=False, double_buffer=True)
ref = engine.build_dbg(packed, k)[0].checksum()
L = bd.L
st = torch.cuda.current_stream(); side = bd.side
E = lambda: torch.cuda.Event(enable_timing=True)
state = {"cur": 0}

def k2a(p, b):
    desc = PgTable(None, 2, None, bd.mode, bd.k)
    check(L.pg_kmer_partition_dev(ctypes.byref(desc), _ptr(p.pk2), _ptr(p.amb), _ptr(p.d_seq_off), _ptr(p.d_counts), p.cap_records, p.nbytes,
                                  bd.owner_bits, bd.sub_bits, _ptr(b.records), b.part_cap, _ptr(b.counts), _stream()), "k2a")
def k3(t, p, b):
    check(L.pg_count_short_dev(ctypes.byref(t.c), _ptr(p.d_seq_off), _ptr(p.d_counts), p.cap_records, _stream()), "cs")
    check(L.pg_insert_records(ctypes.byref(t.c), _ptr(b.records), _ptr(b.seg_off), _ptr(b.counts), b.n_parts, 1, b.part_cap, _stream()), "k3")

def step(mode, ev=None):
    b = bd.buckets
    marks = [E() for _ in range(4)] if ev is not None else None
    rec = (lambda i: marks[i].record(st)) if ev is not None else (lambda i: None)
    rec(0)
    if mode == "serial":
        t = bd.tables[0]; t.clear()
        p = engine.PackedSeqs(d, lazy=True); rec(1); k2a(p, b); rec(2); k3(t, p, b)
    elif mode == "begin":
        t = bd.tables[0]
        side.wait_stream(st)
        with torch.cuda.stream(side): t.clear()
        p = engine.PackedSeqs(d, lazy=True); rec(1); k2a(p, b); rec(2); st.wait_stream(side); k3(t, p, b)
    elif mode == "after_k1":
        t = bd.tables[0]
        p = engine.PackedSeqs(d, lazy=True); rec(1)
        side.wait_stream(st)
        with torch.cuda.stream(side): t.clear()
        k2a(p, b); rec(2); st.wait_stream(side); k3(t, p, b)
    elif mode == "under_k3":
        cur = state["cur"] = state["cur"] ^ 1
        t, other = bd.tables[cur], bd.tables[cur ^ 1]
        p = engine.PackedSeqs(d, lazy=True); rec(1); k2a(p, b); rec(2)
        st.wait_stream(side)                 # the clear of t, enqueued during the previous step
        side.wait_stream(st)                 # after K2a, and after everything that read `other`
        with torch.cuda.stream(side): other.clear()
        k3(t, p, b)
    rec(3)
    if ev is not None: ev.append(marks)
    return t

for mode in ("serial", "begin", "after_k1", "under_k3"):
    for tt in bd.tables: tt.clear()
    torch.cuda.synchronize()
    for _ in range(4): t = step(mode)
    torch.cuda.synchronize()
    assert t.checksum() == ref, mode
    ev = []; e0, e1 = E(), E(); torch.cuda.synchronize(); e0.record(st)
    for _ in range(40): t = step(mode, ev)
    e1.record(st); torch.cuda.synchronize()
    assert t.checksum() == ref, mode
    n = len(ev)
    seg = [sum(x[i].elapsed_time(x[i + 1]) for x in ev) / n for i in range(3)]
    print("%-9s step %.3f ms   [clear+]K1 %.3f  K2a %.3f  K3(+wait) %.3f" % (mode, e0.elapsed_time(e1) / 40, *seg))
