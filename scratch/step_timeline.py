"""Where does a bench step spend its time?  Events between the stages + host enqueue time."""
import sys, time, torch
sys.path.insert(0, ".")
from pangenome_b200 import engine, synth, _lib
import bench as B

data, _wl = B.workload("cfg2")
k = 27
host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
d = host.to("cuda"); torch.cuda.synchronize()
packed = engine.PackedSeqs(d)
builder = engine.TwoPhaseBuilder(k, _lib.PG_MODE_CANONICAL, packed.n_positions(k), estimate=False, double_buffer="--double" in sys.argv)
st = torch.cuda.current_stream()
E = lambda: torch.cuda.Event(enable_timing=True)
def step(ev=None):
    if ev is not None: a = E(); a.record(st)
    builder.begin()
    if ev is not None: b = E(); b.record(st)
    p = engine.PackedSeqs(d, lazy=True)
    if ev is not None: c = E(); c.record(st)
    t = builder.build_async(p)
    if ev is not None: e = E(); e.record(st); ev.append((a, b, c, e))
    return t
for _ in range(5): step()
torch.cuda.synchronize()
for label, rec in (("no events", False), ("events", True)):
    ev = [] if rec else None
    e0, e1 = E(), E()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record(st)
    for _ in range(50): step(ev)
    e1.record(st); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(label, "gpu ms/step %.3f  host enqueue ms/step %.3f  wall %.3f" % (e0.elapsed_time(e1) / 50, (t1 - t0) * 1e3 / 50, (t2 - t0) * 1e3 / 50))
    if rec:
        n = len(ev)
        print("  begin %.3f  K1 %.3f  build %.3f" % tuple(sum(x[i].elapsed_time(x[i + 1]) for x in ev) / n for i in range(3)))
        print("  step-to-step gap %.3f" % (sum(ev[i][3].elapsed_time(ev[i + 1][0]) for i in range(n - 1)) / (n - 1)))
# K1 alone
e0, e1 = E(), E(); torch.cuda.synchronize(); e0.record(st)
for _ in range(50): p = engine.PackedSeqs(d, lazy=True)
e1.record(st); torch.cuda.synchronize()
print("K1 alone ms %.3f" % (e0.elapsed_time(e1) / 50))
# clear alone
e0, e1 = E(), E(); torch.cuda.synchronize(); e0.record(st)
for _ in range(20): builder.table.clear()
e1.record(st); torch.cuda.synchronize()
print("clear alone ms %.3f  (%.2f GB)" % (e0.elapsed_time(e1) / 20, builder.table.capacity * 16 / 1e9))
