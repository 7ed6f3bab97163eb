"""GPU regression tests for defects found in review."""
import ctypes

import numpy as np
import pytest

import oracle
from pangenome_b200.synth import pangenome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import engine
    return engine


def test_epoch_wrap_clears_the_whole_allocation(eng):
    """The 10-bit generation tag wraps while set_capacity() selects a prefix of the buffer: slots beyond the
    prefix must not keep tags of the previous cycle, or a later, larger build sees ghost keys."""
    import torch
    from pangenome_b200 import _lib
    k = 15
    big = pangenome(2, 40_000, seed=3)
    small = pangenome(1, 3_000, seed=4)
    p_big = eng.PackedSeqs(eng.to_device_bytes(big))
    p_small = eng.PackedSeqs(eng.to_device_bytes(small))
    want_big = oracle.table_checksum(*oracle.run(big, k, stages=1)["dbg"])
    want_small = oracle.table_checksum(*oracle.run(small, k, stages=1)["dbg"])
    cap_big = eng.next_pow2(4 * p_big.n_positions(k))
    cap_small = eng.next_pow2(4 * p_small.n_positions(k))
    t = eng.DbgTable(cap_big, k, _lib.PG_MODE_CANONICAL)

    def build(packed, cap):
        t.set_capacity(cap)
        t.clear()
        t.insert(packed)
        torch.cuda.synchronize()
        assert not t.overflowed()
        return t.checksum()

    # fill the whole allocation under a tag that is about to come round again after the wrap
    t.c.epoch = 2
    t.insert(p_big)
    assert t.checksum() == want_big
    t.c.epoch = _lib_epoch_max() - 1
    assert build(p_small, cap_small) == want_small          # epoch 1023, small prefix only
    assert build(p_small, cap_small) == want_small          # wraps: epoch 1 + full clear of ALL allocated slots
    assert t.c.epoch == 1
    t.c.epoch = 1                                            # next clear() -> epoch 2: the tag the big fill was written under
    assert build(p_big, cap_big) == want_big                 # ghost keys beyond the small prefix would break this
    # the ghost scenario proper: epoch 2 is live again, table emptied, capacity grown - nothing may be visible
    t.set_capacity(cap_small)
    t.c.epoch = _lib_epoch_max()
    t.clear()                                                # wrap with the SMALL capacity selected
    t.set_capacity(cap_big)
    for e in (1, 2, 3, 1022, 1023):
        t.c.epoch = e
        used, entries = t.count()
        assert (used, entries) == (0, 0), "epoch %d sees %d stale slots" % (e, used)


def _lib_epoch_max():
    return 1023


def test_count_short_ranges_are_half_open(eng):
    """Consecutive C-ABI ranges [a,b) [b,c) over one record index: a shorter-than-k record that starts exactly at b
    belongs to the second range only; the range that reaches the end of the stream owns trailing empty records."""
    import torch
    from pangenome_b200 import _lib
    L = _lib.load()
    k = 5
    data = b">a\nACGTACGTAC\n>b\nAC\n>c\nGGGTTTAAAC\n>d\n\n"          # b: 2 bases (< k) at offset 10; d: empty at the stream end
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    assert packed.seq_off.tolist() == [0, 10, 12, 22, 22]
    ref = oracle.run(data, k, stages=1)
    want = oracle.table_checksum(*ref["dbg"])
    for cuts in ([0, 22], [0, 10, 22], [0, 10, 12, 22], [0, 5, 10, 11, 22], [0, 12, 22]):
        t = eng.DbgTable(1024, k, _lib.PG_MODE_CANONICAL)
        for a, b in zip(cuts[:-1], cuts[1:]):
            eng.check(L.pg_kmer_insert(ctypes.byref(t.c), eng._ptr(packed.pk2), eng._ptr(packed.amb), eng._ptr(packed.d_seq_off),
                                       packed.n_rec, a, b, eng._stream()), "pg_kmer_insert")
        torch.cuda.synchronize()
        assert int(t.stats_host()[_lib.PG_STAT_SHORT]) == 4, cuts        # records b and d, two strands each
        assert t.checksum() == want, cuts
