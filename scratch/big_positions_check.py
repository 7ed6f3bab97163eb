"""More than 2^31 k-mer positions on ONE GPU: 32-bit overflow hunt.

440 x 5 Mbp genomes (2.2 Gbp, 1 % SNP from one ancestor) -> K1, then the two-phase build (key-space
estimator sizes the table) and the fused single-launch build must agree on the table checksum, and
the distinct-key count must match what the ancestor + SNP model predicts to a few percent.
Needs ~90 GB of HBM and ~8 GB of host RAM; run by hand:  python scratch/big_positions_check.py
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pangenome_b200 import engine, synth, _lib

n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 440
k = 27
t0 = time.time()
data = synth.pangenome(n_genomes, 5_000_000, seed=3)
print("generated %.2f GB in %.0f s" % (len(data) / 1e9, time.time() - t0), flush=True)
d = engine.to_device_bytes(data)
del data
packed = engine.PackedSeqs(d)
npos = packed.n_positions(k)
print("records %d  bases %d  positions %d (2^31 = %d)" % (packed.n_rec, packed.n_bases, npos, 2 ** 31), flush=True)
assert npos > 2 ** 31
torch.cuda.synchronize(); t0 = time.time()
t, n_rec, b = engine.build_dbg_partitioned(packed, k)
torch.cuda.synchronize()
print("two-phase build: %.3f s, capacity 2^%d, keys %d, load %.3f" % (time.time() - t0, int(np.log2(t.capacity)), t.n_keys(), t.n_keys() / t.capacity), flush=True)
assert int(b.counts.sum().item()) == npos
cs2 = t.checksum()
used, entries = t.count()
assert used == t.n_keys()
cap = t.capacity
del t, b
torch.cuda.empty_cache()
t0 = time.time()
tf = engine.DbgTable(cap, k, _lib.PG_MODE_CANONICAL)
tf.insert(packed)
torch.cuda.synchronize()
print("fused build: %.3f s" % (time.time() - t0), flush=True)
assert not tf.overflowed()
assert tf.checksum() == cs2, (tf.checksum(), cs2)
print("OK: both builds agree on", cs2, "entries", entries)
