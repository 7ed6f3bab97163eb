"""GPU parity: the CUDA path (through the C-ABI of libpgdbg.so) against the
oracle and the reference's golden vectors.  Bit-exact (integer work)."""
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


def unpack_syms(packed):
    pk2 = packed.pk2.cpu().numpy().view(np.uint32)
    amb = packed.amb.cpu().numpy().view(np.uint32)
    i = np.arange(packed.n_bases)
    d = (pk2[i >> 4] >> (2 * (i & 15)).astype(np.uint32)) & 3
    a = (amb[i >> 5] >> (i & 31).astype(np.uint32)) & 1
    return np.where(a == 1, 4 + d, d).astype(np.uint8)


def syms_of_bytes(seq):
    lut = np.full(256, 5, np.uint8)
    for ch, v in ((b"Aa", 0), (b"Gg", 1), (b"Cc", 2), (b"Tt", 3), (b"Nn", 4)):
        for c in ch:
            lut[c] = v
    return lut[seq]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_pack_dbg_rdbg_golden(eng, case):
    data = case["input_latin1"].encode("latin-1")
    k, c, Ns = case["k"], case["c"], case.get("Ns", 2 ** 63)
    ref = oracle.run(data, k, c=c, Ns=Ns, stages=2)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    assert packed.seq_off.tolist() == ref["seq_off"].tolist()
    assert packed.hdr_off.tolist() == ref["hdr_off"].tolist()
    junk = int(packed.seq_off[0]) if packed.n_rec else 0
    assert unpack_syms(packed)[junk:].tolist() == syms_of_bytes(ref["seq"]).tolist()
    rc0 = bool((c >> 1) & 1)
    for mode in ((2, 1) if rc0 else (0,)):
        t, n_used = eng.build_dbg(packed, k, rc=rc0, Ns=Ns, mode=mode)
        ks, vs, cs = t.export()
        got = [[int(a), int(b), int(d)] for a, b, d in zip(ks, vs, cs)]
        assert got == case["dbg"], "mode %d" % mode
        assert t.checksum() == oracle.table_checksum(*ref["dbg"])
        rd = t.select_rdbg()
        rk, rv = rd.rdbg_export()
        assert rk.tolist() == case["rdbg"], "mode %d" % mode
        assert rv.tolist() == ref["rdbg_vals"].tolist()


def test_pack_edge_cases(eng):
    for data, n_rec, n_bases in ((b"", 0, 0), (b">x ACGT", 0, 0), (b">x\n", 1, 0), (b"\n", 0, 0),
                                 (b"junk\n>s1 desc\nACGT\nAC\n>s2\n\n>s3\nGGGTT", 3, 14)):
        p = eng.PackedSeqs(eng.to_device_bytes(data))
        ref = oracle.run(data, 3, stages=1)
        assert (p.n_rec, p.n_bases) == (n_rec, n_bases), data
        # bases before the first header stay in the stream ([0, seq_off[0])) but belong to no record
        assert (p.seq_off - p.seq_off[0]).tolist() == ref["seq_off"].tolist()


def test_pack_ragged_tiles(eng):
    """Line widths and sizes that put newlines, headers and the end of file at
    every position relative to the 16-byte chunks and 16 KB tiles."""
    rng = np.random.default_rng(7)
    for width, nrec, L in ((1, 3, 700), (15, 5, 5000), (16, 4, 16384), (17, 3, 40000), (61, 7, 33000),
                           (100000, 2, 70000)):
        recs = []
        for r in range(nrec):
            s = rng.choice(np.frombuffer(b"ACGTNacgtRY", dtype=np.uint8), size=L + r, p=[.23, .23, .23, .23, .02, .01, .01, .01, .01, .01, .01])
            body = b"\n".join(bytes(s[i:i + width]) for i in range(0, s.size, width))
            recs.append(b">r%d some text\n" % r + body + b"\n")
        data = b"".join(recs)
        for cut in (0, 1, 5):
            d = data[:len(data) - cut]
            ref = oracle.run(d, 5, stages=1)
            p = eng.PackedSeqs(eng.to_device_bytes(d))
            assert p.seq_off.tolist() == ref["seq_off"].tolist()
            assert p.hdr_off.tolist() == ref["hdr_off"].tolist()
            assert np.array_equal(unpack_syms(p)[int(p.seq_off[0]):], syms_of_bytes(ref["seq"]))


def test_big_4x1m(eng, big_facts):
    data = survey_4x1m()
    ref = oracle.run(data, 27, stages=2)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    assert packed.n_rec == 4 and packed.n_bases == 4_000_000
    assert packed.n_insertions(27) == 7999792
    for mode in (2, 1):
        t, _ = eng.build_dbg(packed, 27, mode=mode)
        assert t.checksum() == oracle.table_checksum(*ref["dbg"])
        ks, vs, cs = t.export()
        assert ks.size == big_facts["dbg_entries"]
        assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])
        rd = t.select_rdbg()
        rk, rv = rd.rdbg_export()
        assert rk.size == big_facts["rdbg_entries"]
        assert np.array_equal(rk, ref["rdbg"])


def test_count_saturation_and_hot_keys(eng):
    """poly-A and a 2-mer satellite: > 255 occurrences of one k-mer, heavy
    same-address atomic contention; counts clamp at 255 like the uint8 upstream."""
    data = b">a\n" + b"A" * 5000 + b"\n>b\n" + b"AC" * 3000 + b"\n>c\n" + b"A" * 300 + b"G" + b"A" * 300 + b"\n"
    ref = oracle.run(data, 11, stages=2)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    for mode in (2, 1):
        t, _ = eng.build_dbg(packed, 11, mode=mode)
        ks, vs, cs = t.export()
        assert np.array_equal(ks, ref["dbg"][0]) and np.array_equal(vs, ref["dbg"][1]) and np.array_equal(cs, ref["dbg"][2])
        assert int(cs.max()) == 255


def test_table_growth_on_overflow(eng):
    data = pangenome(2, 20000)
    ref = oracle.run(data, 15, stages=1)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, _ = eng.build_dbg(packed, 15, capacity=1024)     # far too small: must grow and still be exact
    assert t.capacity > 1024
    assert t.checksum() == oracle.table_checksum(*ref["dbg"])


def test_config2_scaled_checksum(eng):
    """Config-2 shape at 1/10 scale (10 x 500 kbp, 1 % SNP), k sweep 15/21/27."""
    data = pangenome(10, 500_000)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    for k in (15, 21, 27):
        ref = oracle.run(data, k, stages=1)
        t, _ = eng.build_dbg(packed, k)
        assert t.checksum() == oracle.table_checksum(*ref["dbg"]), k


def test_single_pass_k1_matches(eng):
    """PG_K1_SINGLE_PASS=1 (one launch, decoupled look-back) packs exactly like the default three launches.
    The switch is read once per process, so the variant runs in a subprocess."""
    import os, subprocess, sys, textwrap
    from conftest import ROOT
    code = textwrap.dedent("""
        import sys, hashlib, numpy as np
        sys.path.insert(0, %r)
        from pangenome_b200 import engine
        from pangenome_b200.synth import pangenome
        out = []
        for data in (pangenome(3, 300_000), b">a desc\\nACGTNNacgtRY\\nAC\\n\\n>b\\n>c\\nGGGTT", b"junk\\n>x\\n" + b"ACGT" * 9000, b""):
            p = engine.PackedSeqs(engine.to_device_bytes(data))
            nw = p.n_bases // 16 + 1
            out.append((p.n_rec, p.n_bases, p.seq_off.tolist(), p.hdr_off.tolist(),
                        hashlib.sha256(p.pk2[:nw].cpu().numpy().tobytes() + p.amb[:nw // 2 + 1].cpu().numpy().tobytes()).hexdigest()))
        print(repr(out))
    """ % ROOT)
    res = {}
    for flag in ("0", "1"):
        env = dict(os.environ, PG_K1_SINGLE_PASS=flag)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[flag] = r.stdout.strip().splitlines()[-1]
    assert res["0"] == res["1"]


def test_table_generations_and_wrap(eng):
    """DbgTable.clear() is an epoch bump (pg_table_reset): slots written by earlier builds must read as
    free in every later one - also across the wrap of the 10-bit generation tag, where the slots are
    rewritten - and in all three table modes.  Two different inputs alternate so that a stale slot
    leaking through would change the table."""
    from pangenome_b200 import _lib
    inputs = [b">a\nACGTTGCAAGGCTTAACCGGATAGGCTTACGATCGGANCTTAGGA\n>b\nTTGACGGTCATTG\n",
              b">x\nGGGGGGGGGGGGGGGGGGGACGTACGTACGTTTTTTTTTTTTTTTTTTT\n>s\nAC\n"]
    for mode, rc in ((_lib.PG_MODE_CANONICAL, True), (_lib.PG_MODE_LITERAL_RC, True), (_lib.PG_MODE_LITERAL, False)):
        packs = [eng.PackedSeqs(eng.to_device_bytes(d)) for d in inputs]
        want = []
        for d in inputs:
            r = oracle.run(d, 7, c=3 if rc else 0, stages=1)
            want.append(oracle.table_checksum(*r["dbg"]))
        t = eng.DbgTable(256, 7, mode)
        assert t.c.epoch == 1
        seen_wrap = False
        for it in range(1100):
            i = (it * 7 + it // 3) & 1
            t.clear()
            seen_wrap |= t.c.epoch == 1
            t.insert(packs[i])
            if it % 97 == 0 or it > 1015:
                assert t.checksum() == want[i], (mode, it, t.c.epoch)
                assert t.n_keys() == t.count()[0]
        assert seen_wrap and 1 <= t.c.epoch <= 1023


def test_bad_epoch_is_rejected(eng):
    import ctypes
    from pangenome_b200 import _lib
    t = eng.DbgTable(64, 5, _lib.PG_MODE_CANONICAL)
    for bad in (0, 1024, -3):
        c = _lib.PgTable(t.c.d_slots, t.c.capacity, t.c.d_stats, t.c.mode, t.c.k, bad, 0)
        assert _lib.load().pg_table_count(ctypes.byref(c), None) != 0
        assert b"epoch" in _lib.load().pg_last_error()
