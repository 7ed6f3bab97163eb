"""How long the HOST takes to enqueue one step (no synchronisation) against the GPU time of the step."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangenome_b200 import _lib, builder, engine, synth
data = synth.pangenome(10, 5_000_000)
d = engine.to_device_bytes(data)
b = builder.RoundBuilder(27, _lib.PG_MODE_CANONICAL, len(data))
def step():
    b.begin(); p = engine.PackedSeqs(d, lazy=True); return b.build_async(p)
for _ in range(3):
    step()
torch.cuda.synchronize(); b.verify()
for _ in range(3):
    step()
torch.cuda.synchronize()
n = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    step()
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue per step %.3f ms, gpu per step %.3f ms, wall per step %.3f ms" % (1e3 * (t1 - t0) / n, e0.elapsed_time(e1) / n, 1e3 * (t2 - t0) / n))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
