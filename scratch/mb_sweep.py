"""Random-slot ceiling sweep (profiles/r2*_microbench.jsonl): pg_microbench_slots over table size, region size,
threads per SM and independent operations in flight per thread."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangenome_b200 import measure
torch.cuda.set_device(0)
n_ops = 1 << 25
for cap_bits in (27,):
    for region_mb in (8, 32):
        for mode, name in ((0, "load"), (3, "mix"), (11, "mix+records")):
            for ctas, ilp in ((5, 1), (8, 1), (4, 2), (6, 2), (4, 4), (2, 8), (4, 8)):
                g, ms = measure.slot_ceiling(1 << cap_bits, (region_mb << 20) // 16, n_ops, mode, ctas, ilp)
                print(json.dumps({"table_slots": 1 << cap_bits, "region_mb": region_mb, "mode": name, "ctas_per_sm": ctas, "ilp": ilp,
                                  "G_ops_per_s": round(g, 2), "ms": round(ms, 4)}), flush=True)
