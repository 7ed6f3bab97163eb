"""GPU parity of the two-phase build (K2a pg_kmer_partition + K3 pg_insert_records)."""
import numpy as np
import pytest

import oracle
from conftest import load_small_cases
from pangenome_b200.synth import pangenome, survey_4x1m

pytestmark = pytest.mark.gpu
CASES = load_small_cases()


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import engine
    return engine


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_partitioned_golden(eng, case):
    data = case["input_latin1"].encode("latin-1")
    k, c, Ns = case["k"], case["c"], case.get("Ns", 2 ** 63)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    rc0 = bool((c >> 1) & 1)
    for mode in ((2, 1) if rc0 else (0,)):
        for sub_bytes in (32 << 20, 4096):        # one bucket / many buckets
            t, _, _ = eng.build_dbg_partitioned(packed, k, rc=rc0, Ns=Ns, mode=mode, sub_bytes=sub_bytes)
            ks, vs, cs = t.export()
            got = [[int(a), int(b), int(d)] for a, b, d in zip(ks, vs, cs)]
            assert got == case["dbg"], "mode %d sub_bytes %d" % (mode, sub_bytes)


def test_partitioned_big_and_buckets(eng, big_facts):
    data = survey_4x1m()
    ref = oracle.run(data, 27, stages=1)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    for sub_bytes in (32 << 20, 1 << 20, 1 << 16):
        t, n_rec, b = eng.build_dbg_partitioned(packed, 27, sub_bytes=sub_bytes)
        assert t.checksum() == oracle.table_checksum(*ref["dbg"]), sub_bytes
        counts = b.counts.cpu().numpy()
        assert int(counts.sum()) == packed.n_positions(27)
        assert counts.max() < 1.2 * counts.mean() + 4096          # hash-uniform buckets
    # every record of bucket b hashes into table region b
    t, n_rec, b = eng.build_dbg_partitioned(packed, 27, sub_bytes=1 << 20)
    import torch
    rec = b.records.view(-1, 2)
    cnt = b.counts.cpu().numpy()
    bits = int(np.log2(b.n_parts))
    for p in (0, b.n_parts // 2, b.n_parts - 1):
        keys = rec[p * b.part_cap:p * b.part_cap + int(cnt[p]), 0].cpu().numpy().view(np.uint64)
        h = keys.copy()
        with np.errstate(over="ignore"):
            h ^= h >> np.uint64(33); h *= np.uint64(0xff51afd7ed558ccd); h ^= h >> np.uint64(33)
            h *= np.uint64(0xc4ceb9fe1a85ec53); h ^= h >> np.uint64(33)
        assert np.all((h >> np.uint64(64 - bits)) == p)


def test_partitioned_skew_falls_back(eng):
    """poly-A: one key dominates -> its bucket overflows -> fused kernel takes over, result exact"""
    data = b">a\n" + b"A" * 300000 + b"\n>b\n" + pangenome(1, 20000)[len(b">g0 synthetic\n"):]
    ref = oracle.run(data, 21, stages=1)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, _, _ = eng.build_dbg_partitioned(packed, 21, sub_bytes=1 << 12)
    assert t.checksum() == oracle.table_checksum(*ref["dbg"])


def test_inserts_count_distinct_keys(eng):
    """PG_STAT_USED is kept by the insert kernels themselves (fused and two-phase)."""
    data = pangenome(6, 200_000)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t1, _ = eng.build_dbg(packed, 27)
    kept = t1.n_keys()
    used, _ = t1.count()
    assert kept == used
    t2, _, _ = eng.build_dbg_partitioned(packed, 27)
    assert t2.n_keys() == used
    from pangenome_b200 import _lib
    b = eng.TwoPhaseBuilder(27, _lib.PG_MODE_CANONICAL, packed.n_positions(27))
    for _ in range(3):                      # reused buffers, clear overlapped on a side stream
        t3 = b.build(packed, packed.n_rec)
        assert t3.n_keys() == used
        assert t3.checksum() == t1.checksum()
    b.verify()


def test_table_sized_from_key_sample(eng):
    """K2a's 1/256 key-space sample estimates the distinct keys within a few percent; a table sized
    from it gives the same result as the upper-bound table."""
    from pangenome_b200 import _lib
    data = pangenome(8, 400_000)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    ref_t, _ = eng.build_dbg(packed, 27)
    b = eng.TwoPhaseBuilder(27, _lib.PG_MODE_CANONICAL, packed.n_positions(27), estimate=True)
    t = b.build(packed, packed.n_rec)
    b.verify()
    assert abs(b.last_estimate - ref_t.n_keys()) < 0.05 * ref_t.n_keys()
    assert t.capacity < b.cap_max and 0.17 < t.n_keys() / t.capacity <= 0.36
    assert t.checksum() == ref_t.checksum()


def test_async_build_device_args(eng):
    """pack -> partition -> insert enqueued without reading K1's record index back (device-side bounds)."""
    from pangenome_b200 import _lib
    for data in (pangenome(5, 300_000), b">a\nACGTACGTTGCAAGGCTTAACCGGATAGGCTTACGATCGGATCTTAGGA\n>s\nACG\n>e\n\n>b\nTTGACGGTCATTGCAGGCATTACGGATCGATCGGCTAGCTAGGCTAGG\n"):
        ref_p = eng.PackedSeqs(eng.to_device_bytes(data))
        ref_t, _ = eng.build_dbg(ref_p, 21)
        b = eng.TwoPhaseBuilder(21, _lib.PG_MODE_CANONICAL, max(ref_p.n_positions(21), 64), estimate=False)
        for _ in range(2):
            b.begin()
            p = eng.PackedSeqs(eng.to_device_bytes(data), lazy=True)
            t = b.build_async(p)
        b.verify()
        assert t.checksum() == ref_t.checksum()
        assert t.export()[0].tolist() == ref_t.export()[0].tolist()
        assert p.seq_off.tolist() == ref_p.seq_off.tolist()          # the lazy index is still available afterwards
    # a record index that does not fit cap_records is reported, not silently truncated
    many = b"".join(b">r%d\nACGTTGCAAGGCTTAACCGGATAGG\n" % i for i in range(40))
    p = eng.PackedSeqs(eng.to_device_bytes(many), cap_records=8, lazy=True)
    b = eng.TwoPhaseBuilder(11, _lib.PG_MODE_CANONICAL, 4096, estimate=False)
    b.build_async(p)
    import pytest as _pt
    with _pt.raises(_lib.PgError):
        b.verify()
    assert p.n_rec == 40                                             # fetching re-packs with a larger index


def test_builder_reuse_alternating_inputs(eng):
    """One TwoPhaseBuilder, two different inputs back to back without synchronising in between: the table is
    emptied by an epoch bump only, so every slot still holds the previous build's key - each result must be its
    own input's table (a stale slot read as live, or a live one as stale, changes the checksum)."""
    from pangenome_b200 import _lib
    ds, refs = [], []
    for seed in (11, 12):
        d = eng.to_device_bytes(pangenome(3, 150_000, seed=seed))
        ds.append(d)
        refs.append(eng.build_dbg(eng.PackedSeqs(d), 19)[0].checksum())
    assert refs[0] != refs[1]
    b = eng.TwoPhaseBuilder(19, _lib.PG_MODE_CANONICAL, 3 * 150_000 + 64, estimate=False)
    for i in (0, 0, 1, 0, 1, 1, 0, 1):
        b.begin()
        t = b.build_async(eng.PackedSeqs(ds[i], lazy=True))
        assert t.checksum() == refs[i]
        assert t.n_keys() == t.count()[0]
    p = eng.PackedSeqs(ds[1])
    assert b.build(p, p.n_rec).checksum() == refs[1]       # the synchronous entry on the same builder
    b.verify()
