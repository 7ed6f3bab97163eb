"""GPU parity of the streaming builder (pangenome_b200/builder.py): rounds, spill buckets, the receiver-side
split K2b, against the oracle and the reference's golden vectors.  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle
from conftest import load_small_cases
from pangenome_b200.synth import pangenome, plant_like, survey_4x1m

pytestmark = pytest.mark.gpu
CASES = load_small_cases()


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import builder, engine
    return engine, builder


def _triples(t):
    ks, vs, cs = t.export()
    return [[int(a), int(b), int(d)] for a, b, d in zip(ks, vs, cs)]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_round_builder_golden(mods, case):
    eng, bld = mods
    data = case["input_latin1"].encode("latin-1")
    k, c, Ns = case["k"], case["c"], case.get("Ns", 2 ** 63)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    rc0 = bool((c >> 1) & 1)
    for mode in ((2, 1) if rc0 else (0,)):
        for rounds in (1, 3):
            for rb in (0, 8):       # L2-atomic K3 / shared-memory regions of 256 slots
                t, n_rec, b = bld.build_table(packed, k, rc=rc0, Ns=Ns, mode=mode, rounds=rounds, region_bits=rb)
                assert b.region_bits == rb or b.table.capacity < 256
                assert _triples(t) == case["dbg"], "mode %d rounds %d region_bits %d" % (mode, rounds, rb)
            if mode == 2:           # compact 8-byte records: 4 regions of 4096 slots, later rounds reload the regions
                t, n_rec, b = bld.build_table(packed, k, rc=rc0, Ns=Ns, mode=mode, rounds=rounds, region_bits=12, capacity=1 << 14, sample=False)
                assert b.compact and t.c.hash_kind == 1 and b.region_bits == 12
                assert _triples(t) == case["dbg"], "compact, rounds %d" % rounds


def test_rounds_device_bounds_and_rebuild(mods):
    """All records with bounds read on the device (lazy K1), several rounds, the same builder twice (buffer parity and the
    epoch bump carry over from build to build)."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    data = survey_4x1m()
    want = oracle.table_checksum(*oracle.run(data, 27, stages=1)["dbg"])
    d = eng.to_device_bytes(data)
    for rounds in (1, 2, 5):
        b = bld.RoundBuilder(27, _lib.PG_MODE_CANONICAL, len(data), rounds=rounds)
        for _ in range(3):
            b.begin()
            t = b.build_async(eng.PackedSeqs(d, lazy=True))
            torch.cuda.synchronize()
            b.verify()
            assert t.checksum() == want, rounds
        assert b.n_rounds == rounds


def test_spill_absorbs_hash_skew(mods):
    """One k-mer (poly-A) making up a large share of the input overflows its bucket; the spill takes the surplus and the
    table is still exact.  Without enough spill the build reports LostRecords instead of a wrong table."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    rng = np.random.default_rng(11)
    recs = [b">r%d\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 3000)) + b"A" * 4000 + b"\n" for i in range(12)]
    data = b"".join(recs)
    k = 21
    ref = oracle.run(data, k, stages=1)
    want = oracle.table_checksum(*ref["dbg"])
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    # tight buckets (slack 1.0 -> about 1/n_sub of the records each): the poly-A bucket must spill
    b = bld.RoundBuilder(k, _lib.PG_MODE_CANONICAL, len(data), capacity=1 << 18, sub_bytes=1 << 14, slack=1.0, spill_frac=1.0, compact=False)
    assert b.sets[0].n_parts >= 64
    b.sets[0].c.part_cap = b.sets[0].part_cap = 2048          # far below the 4000 x 12 poly-A records that hash to one bucket
    b.sets[0].seg_off.copy_(torch.arange(b.sets[0].n_parts + 1, device="cuda") * 2048)
    b.begin()
    t = b.build_async(packed, packed.n_rec)
    torch.cuda.synchronize()
    b.verify()
    assert int(b.sets[0].counts[-1].item()) > 10000            # the spill really was used
    assert t.checksum() == want
    # no room in the spill either: the flag trips, build_table() recovers
    b2 = bld.RoundBuilder(k, _lib.PG_MODE_CANONICAL, len(data), capacity=1 << 18, sub_bytes=1 << 14, compact=False)
    b2.sets[0].c.part_cap = b2.sets[0].part_cap = 2048
    b2.sets[0].c.spill_cap = b2.sets[0].spill_cap = 16
    b2.sets[0].seg_off.copy_(torch.arange(b2.sets[0].n_parts + 1, device="cuda") * 2048)
    b2.begin()
    b2.build_async(packed, packed.n_rec)
    torch.cuda.synchronize()
    with pytest.raises(bld.LostRecords):
        b2.verify()
    t3, _, _ = bld.build_table(packed, k)
    assert t3.checksum() == want


def test_k2b_split_path_on_one_gpu(mods):
    """The multi-GPU data flow on one device: K2a into 2 coarse buckets, K2b (pg_records_split) re-sorts the two segments into
    hash-prefix regions (+ spill), plan, K3."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    L = _lib.load()
    data = plant_like(n_genomes=3, length=200_000, n_chrom=2, n_families=40)
    k = 27
    want = oracle.table_checksum(*oracle.run(data, k, stages=1)["dbg"])
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    n_rec = packed.n_rec
    g0, g1 = int(packed.seq_off[0]), int(packed.seq_off[n_rec])
    coarse = bld.LocalBuckets(1, int((g1 - g0) * 0.6) + 4096, 0, "cuda")
    for sub_bits, part_cap in ((6, None), (9, None), (4, 3000)):
        n = g1 - g0
        fine = bld.LocalBuckets(sub_bits, part_cap or int(n / (1 << sub_bits) * 1.3) + 2048, n, "cuda")
        t = eng.DbgTable(4 * n, k, _lib.PG_MODE_CANONICAL)
        desc = _lib.PgTable(None, 2, None, _lib.PG_MODE_CANONICAL, k, 1, 0, 0)
        P, S = eng._ptr, eng._stream
        eng.check(L.pg_kmer_partition_to(ctypes.byref(desc), P(packed.pk2), P(packed.amb), P(packed.d_seq_off), n_rec, g0, g1, None, 0, 0,
                                         ctypes.byref(coarse.c), None, 0, None, S()), "pg_kmer_partition_to")
        eng.check(L.pg_count_short(ctypes.byref(t.c), P(packed.d_seq_off), n_rec, g0, g1, S()), "pg_count_short")
        eng.check(L.pg_records_split(P(coarse.records), P(coarse.seg_off), P(coarse.counts), 2, coarse.part_cap, ctypes.byref(fine.c),
                                     P(t.stats), S()), "pg_records_split")
        eng.check(L.pg_buckets_plan(ctypes.byref(fine.c), P(fine.seg_cnt), P(t.stats), S()), "pg_buckets_plan")
        eng.check(L.pg_insert_records(ctypes.byref(t.c), P(fine.records), P(fine.seg_off), P(fine.seg_cnt), fine.n_parts + 1, 1, 0, S()),
                  "pg_insert_records")
        torch.cuda.synchronize()
        st = t.stats_host()
        assert st[_lib.PG_STAT_LOST] == 0 and st[_lib.PG_STAT_OVERFLOW] == 0
        cnt = fine.counts.cpu().numpy()
        assert int(cnt[:-1].sum()) == packed.n_positions(k)
        if part_cap:
            assert cnt[-1] > 0                               # tight buckets: the spill took the surplus
        assert t.checksum() == want, (sub_bits, part_cap)


def test_plant_like_single_gpu_parity(mods):
    """BASELINE configs 4/5 in miniature (repeat families, poly-A / microsatellite tracts, 5 chromosomes per genome):
    dBG, rdBG, .xyz in file order and rows against the oracle."""
    eng, bld = mods
    from pangenome_b200 import graph
    data = plant_like(n_genomes=4, length=400_000, n_chrom=5, n_families=60)
    k = 27
    ref = oracle.run(data, k, stages=4)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    assert packed.n_rec == 20
    t, n_rec, b = bld.build_table(packed, k, rounds=3)
    ks, vs, cs = t.export()
    assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])
    rd = t.select_rdbg()
    rk, _ = rd.rdbg_export()
    assert np.array_equal(rk, ref["rdbg"])
    res = graph.seq2graph_device(packed, rd, k)
    assert res.xyz_lines() == ref["xyz"]
    assert res.rows(packed, data) == ref["rows"]


@pytest.mark.parametrize("region_bits,rounds,compact", [(12, 1, True), (12, 3, True), (12, 1, False), (12, 3, False), (8, 1, False), (8, 2, False)])
def test_region_build_refine_and_retune(mods, region_bits, rounds, compact):
    """K2a -> K2c (coarse buckets refined to one bucket per region) -> K3s (regions built in shared memory) -> spill, with the
    bounds read on the device; the same builder three times: verify() retunes the capacity to the distinct keys found
    (load <= 0.5), which changes the number of regions and K2c's fan-out between builds."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    data = survey_4x1m()
    want = oracle.table_checksum(*oracle.run(data, 27, stages=1)["dbg"])
    d = eng.to_device_bytes(data)
    b = bld.RoundBuilder(27, _lib.PG_MODE_CANONICAL, len(data), rounds=rounds, region_bits=region_bits, compact=compact)
    assert b.region_bits == region_bits and b.adaptive and b.compact == compact
    caps = []
    for _ in range(3):
        b.begin()
        t = b.build_async(eng.PackedSeqs(d, lazy=True))
        torch.cuda.synchronize()
        caps.append(t.capacity)
        assert (t.capacity >> region_bits) > (1 << b.sub_bits), "this test must go through K2c"
        b.verify()
        assert t.capacity == caps[-1]                          # retuning applies to the NEXT build
        assert t.checksum() == want, (region_bits, rounds, caps)
        n_keys = t.n_keys()
        used, entries = t.count()
        assert used == n_keys == b._last_used
    assert caps[1] < caps[0] and caps[2] == caps[1]            # retuned once, then stable
    assert 0.25 < b._last_used / caps[-1] <= 0.5


def test_region_build_sampled_capacity_and_overflow_recovery(mods):
    """build_table on one GPU: the table is sized from K2a's key-space sample (load <= 0.5); a table forced too small
    fills a region, which is reported (TableFull) and rebuilt larger - never a wrong table."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    data = survey_4x1m()
    ref = oracle.run(data, 27, stages=1)
    want = oracle.table_checksum(*ref["dbg"])
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, n_rec, b = bld.build_table(packed, 27)
    assert b.region_bits == 12 and b.sampler is not None
    used = t.n_keys()
    assert abs(b.last_estimate - used) < 0.02 * used
    assert 0.2 < used / t.capacity <= 0.5
    assert t.checksum() == want
    ks, vs, cs = t.export()
    assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])
    # the stages after the build read the region table like any other
    rd = t.select_rdbg()
    assert rd.n_members == len(oracle.run(data, 27, stages=2)["rdbg"])
    # 2^20 slots for ~1.95 M keys: regions fill up
    small = bld.RoundBuilder(27, _lib.PG_MODE_CANONICAL, len(data), capacity=1 << 20)
    small.begin()
    small.build_async(packed, packed.n_rec)
    torch.cuda.synchronize()
    with pytest.raises(bld.TableFull):
        small.verify()
    t2, _, b2 = bld.build_table(packed, 27, capacity=1 << 20)
    assert t2.capacity > 1 << 20 and t2.checksum() == want


def test_region_table_takes_later_upserts(mods):
    """A table built in shared-memory regions probes inside a region; the fused insert kernel (pg_kmer_insert) follows the
    same rule, so adding the same records again only doubles the counts."""
    eng, bld = mods
    data = pangenome(3, 60_000, seed=3)
    ref = oracle.run(data + data.replace(b">", b">x"), 21, stages=1)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, n_rec, b = bld.build_table(packed, 21, region_bits=8)
    assert b.region_bits == 8 and t.c.region_bits == 8
    t.insert(packed, n_rec)
    ks, vs, cs = t.export()
    assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])


def test_compact_wide_spill_takes_hash_skew_and_edges(mods):
    """Compact records: a poly-A k-mer overflows its 8-byte bucket, the surplus travels as wide records; record edges and N
    runs are wide from the start.  Exact table; k even, so palindromes take the fold bit."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    rng = np.random.default_rng(12)
    recs = []
    for i in range(12):
        body = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 3000)) + b"A" * 4000 + b"NNNNNNNNNN" + b"ACGT" * 40 + b"TTTTAAAA" * 20
        recs.append(b">r%d\n" % i + body + b"\n")
    data = b"".join(recs)
    for k in (20, 27):
        ref = oracle.run(data, k, stages=1)
        want = oracle.table_checksum(*ref["dbg"])
        packed = eng.PackedSeqs(eng.to_device_bytes(data))
        b = bld.RoundBuilder(k, _lib.PG_MODE_CANONICAL, len(data), capacity=1 << 18, slack=1.0, spill_frac=1.0)
        assert b.compact and b.sets[0].n_parts == 64
        b.sets[0].c.part_cap = b.sets[0].part_cap = 2048          # far below the 4000 x 12 poly-A records of one bucket
        b.begin()
        t = b.build_async(packed, packed.n_rec)
        torch.cuda.synchronize()
        b.verify()
        assert int(b.sets[0].counts.max().item()) > 10000 and int(b.sets[0].wide_count.item()) > 10000
        assert t.checksum() == want, k
        ks, vs, cs = t.export()
        assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])


def test_compact_falls_back_on_short_records(mods):
    """An input the compact form does not suit (thousands of 60-base records: every position is a record edge) overflows
    the wide spill: verify() reports it and puts the builder back on 16-byte records; build_table() recovers by itself."""
    import torch
    eng, bld = mods
    from pangenome_b200 import _lib
    rng = np.random.default_rng(13)
    data = b"".join(b">s%d\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 60)) + b"\n" for i in range(40000))
    k = 27
    want = oracle.table_checksum(*oracle.run(data, k, stages=1)["dbg"])
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    b = bld.RoundBuilder(k, _lib.PG_MODE_CANONICAL, len(data), capacity=1 << 22, spill_frac=1.0 / 64)
    assert b.compact
    b.begin()
    b.build_async(packed, packed.n_rec)
    torch.cuda.synchronize()
    with pytest.raises(bld.LostRecords):
        b.verify()
    b.begin()
    assert not b.compact and b.table.c.hash_kind == 0
    t = b.build_async(packed, packed.n_rec)
    torch.cuda.synchronize()
    b.verify()
    assert t.checksum() == want
    t2, _, b2 = bld.build_table(packed, k)
    assert t2.checksum() == want


def test_compact_table_takes_later_upserts_and_stages(mods):
    """A table placed by the hash of the 2-bit code (hash_kind 1) takes later generic upserts (the fused insert kernel
    converts the base-5 key for the placement hash) and feeds the stages after it like any other."""
    eng, bld = mods
    from pangenome_b200 import graph
    data = pangenome(3, 60_000, seed=3) + b">n\n" + b"ACGTNNACGTTGCA" * 30 + b"\n"
    ref2 = oracle.run(data + data.replace(b">", b">x"), 21, stages=1)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, n_rec, b = bld.build_table(packed, 21, capacity=1 << 19)
    assert b.compact and t.c.hash_kind == 1
    ref = oracle.run(data, 21, stages=4)
    ks, vs, cs = t.export()
    assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])
    rd = t.select_rdbg()
    assert np.array_equal(rd.rdbg_export()[0], ref["rdbg"])
    res = graph.seq2graph_device(packed, rd, 21)
    assert res.xyz_lines() == ref["xyz"]
    assert res.rows(packed, data) == ref["rows"]
    t.insert(packed, n_rec)
    ks, vs, cs = t.export()
    assert np.array_equal(ks, ref2["dbg"][0]) and np.array_equal(vs, ref2["dbg"][1]) and np.array_equal(cs, ref2["dbg"][2])
