"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: splitting ONE file into record-aligned byte ranges, the
``-n`` record prefix that continues across ranks, the merged table checksum, the padded variable-length gather.
No kernels involved (stub objects stand in for the device tables)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _spawn(fn, world=2, port=29571):
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(fn, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world and all(out[r] is True for r in range(world)), dict(out)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


FASTA = b"junk before the first header\n>r0 first\nACGTACGT\nACG\n>r1\n" + b"ACGT" * 40 + b"\n>r2 x\n\n>r3\nGG\n>r4\n" + b"T" * 90 + b"\n"


def _range_worker(rank, world, port, out):
    import tempfile
    _init(rank, world, port)
    from pangenome_b200 import shard
    path = [None]
    if rank == 0:
        d = tempfile.mkdtemp(prefix="pg_shard_")
        path[0] = os.path.join(d, "in.fa")
        with open(path[0], "wb") as f:
            f.write(FASTA)
    dist.broadcast_object_list(path, src=0)
    data, (a, b), size = shard.read_rank_range(path[0], world, rank)
    parts = [None] * world
    dist.all_gather_object(parts, (a, b, bytes(data)))
    ok = size == len(FASTA) and b"".join(p[2] for p in parts) == FASTA          # the ranges tile the file
    ok = ok and all(p[2][:1] == b">" or i == 0 or p[0] == p[1] for i, p in enumerate(parts))     # every range starts at a record
    ok = ok and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_byte_ranges_tile_the_file_gloo():
    _spawn(_range_worker, 2, 29571)


def test_cut_points_properties():
    from pangenome_b200 import shard
    starts = [i for i in range(len(FASTA)) if FASTA[i:i + 1] == b">" and (i == 0 or FASTA[i - 1:i] == b"\n")]
    for world in (1, 2, 3, 4, 8, 16):
        cuts = shard.cut_points(FASTA, world)
        assert cuts == shard.cut_points_from_starts(starts, len(FASTA), world)
        assert cuts[0] == 0 and cuts[-1] == len(FASTA) and cuts == sorted(cuts)
        assert all(c in starts or c == len(FASTA) for c in cuts[1:-1])
    assert shard.cut_points(b"", 4) == [0, 0, 0, 0, 0]
    assert shard.cut_points(b">only\nACGT\n", 4) == [0, 11, 11, 11, 11]          # one record: rank 0 takes it, the others are empty


class _StubPacked:
    def __init__(self, lens):
        self.seq_lengths = np.asarray(lens, dtype=np.int64)
        self.pk2 = torch.zeros(1)

    def record_prefix(self, Ns, strands):
        lens = self.seq_lengths * strands
        cum = np.cumsum(lens)
        over = np.nonzero(cum > Ns)[0]
        return int(over[0]) + 1 if over.size else int(lens.size)


def _prefix_worker(rank, world, port, out):
    _init(rank, world, port)
    from pangenome_b200 import builder
    lens = [[100, 50, 70], [30, 500, 20]][rank]
    p = _StubPacked(lens)
    whole = _StubPacked([100, 50, 70, 30, 500, 20])
    ok = True
    for strands in (1, 2):
        for Ns in (0, 99, 100, 219, 220, 249, 250, 251, 749, 750, 10 ** 9):
            mine = builder.global_record_prefix(p, Ns, strands, world)
            want_total = whole.record_prefix(Ns, strands)
            want = min(max(want_total - (0 if rank == 0 else 3), 0), 3)
            ok = ok and mine == want
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_record_prefix_continues_across_ranks_gloo():
    _spawn(_prefix_worker, 2, 29573)


class _StubTable:
    def __init__(self, entries, short):
        self.entries, self.short = entries, short
        self.slots = torch.zeros(1)

    def checksum(self):
        from pangenome_b200 import measure
        ent = list(self.entries)
        if self.short > 0:
            ent.append(((1 << 64) - 1, 32, min(self.short, 255)))
        n, s, x = 0, 0, 0
        for k, v, c in ent:
            h = measure.entry_mix(k, v, c)
            n, s, x = n + 1, (s + h) & ((1 << 64) - 1), x ^ h
        return n, s, x

    def stats_host(self):
        a = np.zeros(8, np.int64)
        a[1] = self.short
        return a


def _checksum_worker(rank, world, port, out):
    _init(rank, world, port)
    from pangenome_b200 import measure
    rng = np.random.default_rng(5)
    ent = [(int(rng.integers(0, 2 ** 62)), int(rng.integers(0, 4096)), int(rng.integers(1, 256))) for _ in range(200)]
    ok = True
    for shorts in ((0, 0), (3, 0), (200, 100), (0, 2)):
        mine = _StubTable(ent[rank::world], shorts[rank])
        whole = _StubTable(ent, sum(shorts))
        ok = ok and measure.merged_checksum(mine, world) == whole.checksum()
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_merged_checksum_gloo():
    """per-rank checksums combine into the checksum of the whole table; the short-record sentinel is summed first and
    clamped after (200 + 100 -> 255, not 200 and 100 separately)."""
    _spawn(_checksum_worker, 2, 29575)


def test_entry_mix_matches_the_oracle_checksum():
    import oracle
    from pangenome_b200 import measure
    keys = np.array([0, 5, 2 ** 63 - 1, 2 ** 64 - 1], dtype=np.uint64)
    vals = np.array([1007, 32, 4095, 32], dtype=np.uint16)
    cnts = np.array([1, 255, 7, 2], dtype=np.uint8)
    n, s, x = 0, 0, 0
    for k, v, c in zip(keys.tolist(), vals.tolist(), cnts.tolist()):
        h = measure.entry_mix(k, v, c)
        n, s, x = n + 1, (s + h) & ((1 << 64) - 1), x ^ h
    assert (n, s, x) == oracle.table_checksum(keys, vals, cnts)


def test_log2_exact():
    from pangenome_b200 import multigpu
    assert [multigpu.log2_exact(n) for n in (1, 2, 4, 8)] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        multigpu.log2_exact(3)


def _gv_worker(rank, world, port, out):
    _init(rank, world, port)
    from pangenome_b200 import multigpu
    t = torch.arange(rank * 10, rank * 10 + rank + 1, dtype=torch.int64)       # rank r contributes r+1 items
    cat, sizes = multigpu.gather_varlen(t, world)
    ok = cat.tolist() == [0, 10, 11] and sizes == [1, 2]
    e, _ = multigpu.gather_varlen(torch.zeros(0, dtype=torch.int64), world)
    out[rank] = bool(ok and e.numel() == 0)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_varlen_gloo():
    _spawn(_gv_worker, 2, 29577)


def test_round_planning():
    from pangenome_b200 import builder
    assert builder.plan_rounds(50_000_000, 4) == (4, 12500992)
    assert builder.plan_rounds(100, 3) == (1, 8192)
    n, r = builder.plan_rounds(2_000_000_000, None, 1 << 27)
    assert n * r >= 2_000_000_000 and r % 8192 == 0 and n == 15
    assert builder.table_capacity_for(50_000_000, 180e9) == 1 << 27
    assert builder.table_capacity_for(4_000_000_000, 170e9) == 1 << 32          # capped by the free HBM


def _count_exchange_worker(rank, world, port, out):
    _init(rank, world, port)
    from pangenome_b200 import builder
    # rank r "stored" 100 * r + o compact and 7000 + 10 * r + o wide records into owner o's receive buffers
    send = torch.tensor([100 * rank + o for o in range(world)] + [7000 + 10 * rank + o for o in range(world)], dtype=torch.int64)
    recv = torch.zeros((world, 2), dtype=torch.int64)
    c, w = builder.exchange_record_counts(send, recv, world)
    ok = c.tolist() == [100 * s + rank for s in range(world)] and w.tolist() == [7000 + 10 * s + rank for s in range(world)]
    ok = ok and c.is_contiguous() and w.is_contiguous()
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_compact_count_exchange_gloo():
    """the all-to-all of the compact build: per owner (compact, wide) counts in, per source (compact, wide) counts out"""
    _spawn(_count_exchange_worker, 2, 29577)


def test_plan_levels():
    from pangenome_b200 import builder
    for world in (1, 2, 8):
        for region_log in range(0, 19):
            lv = builder.plan_levels(region_log, world)
            assert sum(lv) == region_log and len(lv) <= 3, (region_log, world, lv)
            assert lv[0] <= 10 and all(1 <= x <= 8 for x in lv[1:]), (region_log, world, lv)
    assert builder.plan_levels(13, 1) == [8, 5] and builder.plan_levels(18, 2) == [10, 8]
    assert builder.plan_levels(17, 1) == [8, 4, 5] and builder.plan_levels(17, 1, 9) == [8, 9] and builder.plan_levels(18, 1, 9) == [8, 5, 5]
