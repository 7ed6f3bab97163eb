"""Run under torchrun (one rank per GPU): distributed build == oracle on the concatenated input.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mg_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pangenome_b200 import engine, multigpu, synth
    import oracle
    k = 21
    anc = np.random.default_rng(3).integers(0, 4, 60000, dtype=np.uint8)
    shards = []
    for r in range(world):
        recs = []
        for g in range(3):
            gid = r * 3 + g
            recs.append((b"g%d" % gid, synth._ACGT[synth._snp_copy(np.random.default_rng(50 + gid), anc, 0.02)]))
        if r == world - 1:
            recs.append((b"short", b"ACGT"))          # short-record sentinel on one rank only
            recs.append((b"polyA", b"A" * 700))       # a hot key: every occurrence must reach one owner; count clamps at 255
        shards.append(synth.fasta_bytes(recs, width=70))
    packed = engine.PackedSeqs(engine.to_device_bytes(shards[rank]))
    for sub_bytes, cls in ((8 << 20, multigpu.DistributedBuilder), (1 << 14, multigpu.DistributedBuilder),
                           (8 << 20, multigpu.PeerBuilder), (1 << 14, multigpu.PeerBuilder)):
        builder = cls(k, packed.n_positions(k), world, rank, sub_bytes=sub_bytes)
        for _ in range(3):                      # repeated builds: buffer reuse / double buffering
            t = builder.build(packed, packed.n_rec)
        if hasattr(builder, "build_async"):     # device-side bounds, nothing read back between K1 and K3
            for _ in range(3):
                lazy = engine.PackedSeqs(engine.to_device_bytes(shards[rank]), lazy=True)
                t = builder.build_async(lazy)
        torch.cuda.synchronize()
        builder.verify()
        merged = multigpu.gather_export(t, world, rank)
        if rank == 0:
            ref = oracle.run(b"".join(shards), k, stages=1)
            ks, vs, cs = merged
            assert np.array_equal(ks, ref["dbg"][0]), "keys differ"
            assert np.array_equal(vs, ref["dbg"][1]), "masks differ"
            assert np.array_equal(cs, ref["dbg"][2]), "counts differ"
            print("mg_check ok: %s world %d, sub_bytes %d, %d entries, max count %d" % (cls.__name__, world, sub_bytes, ks.size, int(cs.max())), flush=True)
        if hasattr(builder, "close"):
            builder.close()
        if cls is multigpu.PeerBuilder and sub_bytes == 8 << 20:
            # stages 2-5 distributed: rdBG all-gather, hits gathered to rank 0, K6-K8 there
            for c_flag in (2, 3):
                res, rows = multigpu.seq2graph_distributed(packed, t, k, world, rank, shards[rank], rc=bool(c_flag & 1))
                if rank == 0:
                    full = oracle.run(b"".join(shards), k, c=c_flag)
                    rk, _ = res.rdbg.rdbg_export()
                    assert np.array_equal(rk, full["rdbg"]), "rdBG differs"
                    assert res.xyz_lines() == full["xyz"], "xyz differs"
                    assert rows == full["rows"], "rows differ"
                    print("mg_check ok: distributed graph -c %d: %d rdBG nodes, %d edges, %d rows" %
                          (c_flag, rk.size, len(full["xyz"]), len(rows)), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
