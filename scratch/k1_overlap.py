"""Does K1 of the next input hide under K3 of the current build (second stream)?"""
import sys, torch
sys.path.insert(0, ".")
from pangenome_b200 import engine, _lib
import bench as B

data, _wl = B.workload("cfg2")
k = 27
d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda(); torch.cuda.synchronize()
packed = engine.PackedSeqs(d)
n_ins = packed.n_insertions(k)
ref = engine.build_dbg(packed, k)[0].checksum()
bd = engine.TwoPhaseBuilder(k, _lib.PG_MODE_CANONICAL, packed.n_positions(k), estimate=False)
main = torch.cuda.current_stream()
side = torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)

def step_serial():
    bd.begin()
    return bd.build_async(engine.PackedSeqs(d, lazy=True))

def step_overlap():
    side.wait_stream(main) if False else None
    with torch.cuda.stream(side):
        p = engine.PackedSeqs(d, lazy=True)
        done = torch.cuda.Event(); done.record(side)
    for t in (p.pk2, p.amb, p._idx, p._ws):
        t.record_stream(main)
    main.wait_event(done)
    bd.begin()
    return bd.build_async(p)

for name, fn in (("serial", step_serial), ("K1 on a second stream", step_overlap), ("serial", step_serial), ("K1 on a second stream", step_overlap)):
    for _ in range(5): t = fn()
    torch.cuda.synchronize()
    assert t.checksum() == ref
    e0, e1 = E(), E(); e0.record(main)
    for _ in range(60): t = fn()
    e1.record(main); torch.cuda.synchronize()
    bd.verify()
    assert t.checksum() == ref
    ms = e0.elapsed_time(e1) / 60
    print("%-24s %.3f ms/step  %.1f G k-mers/s" % (name, ms, n_ins / ms / 1e6))
