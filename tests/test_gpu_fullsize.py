"""BASELINE config 2 at FULL size (10 x 5 Mbp, k = 27) on the GPU against facts computed once by the
C oracle (tests/golden/cfg2_oracle_facts.json; the oracle is itself pinned to the reference), plus
size-independent properties: fused == two-phase, rebuild idempotence, strand symmetry of the table."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from pangenome_b200.synth import pangenome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import engine, graph
    facts = json.load(open(os.path.join(GOLDEN, "cfg2_oracle_facts.json")))
    data = pangenome(10, 5_000_000)
    assert hashlib.sha256(data).hexdigest() == facts["file_sha256"]
    packed = engine.PackedSeqs(engine.to_device_bytes(data))
    return engine, graph, facts, data, packed


def test_cfg2_dbg_checksum_both_builds(ctx):
    engine, graph, facts, data, packed = ctx
    assert packed.n_insertions(27) == facts["n_inserts"]
    t_fused, _ = engine.build_dbg(packed, 27)
    assert list(t_fused.checksum()) == facts["dbg_checksum"]
    t_two, _, _ = engine.build_dbg_partitioned(packed, 27)
    assert list(t_two.checksum()) == facts["dbg_checksum"]
    assert t_two.n_keys() == t_fused.n_keys() == facts["dbg_entries"] // 2     # no palindromes at odd k without N
    # idempotence: building again into a fresh table gives the same table
    t_again, _, _ = engine.build_dbg_partitioned(packed, 27)
    assert t_again.checksum() == t_two.checksum()
    # literal two-inserts-per-position mode agrees with the canonical pairing
    t_lit, _ = engine.build_dbg(packed, 27, mode=1)
    assert list(t_lit.checksum()) == facts["dbg_checksum"]


def test_cfg2_strand_symmetry(ctx):
    """every key has its reverse complement with the same count (both strands are inserted)"""
    engine, graph, facts, data, packed = ctx
    t, _, _ = engine.build_dbg_partitioned(packed, 27)
    ks, vs, cs = t.export()
    assert ks.size == facts["dbg_entries"]
    k = 27
    x = ks.copy()
    rc = np.zeros_like(x)
    for _ in range(k):
        d = x % np.uint64(5)
        x //= np.uint64(5)
        rc = rc * np.uint64(5) + np.where(d == 4, d, np.uint64(3) - d)
    pos = np.searchsorted(ks, rc)
    assert np.array_equal(ks[pos], rc)
    assert np.array_equal(cs[pos], cs)


def test_cfg2_rdbg_graph_rows(ctx):
    engine, graph, facts, data, packed = ctx
    t, _, _ = engine.build_dbg_partitioned(packed, 27)
    rd = t.select_rdbg()
    rk, _ = rd.rdbg_export()
    assert rk.size == facts["rdbg_entries"]
    assert hashlib.sha256("\n".join("%d" % v for v in rk.tolist()).encode()).hexdigest() == facts["rdbg_sha256"]
    res = graph.seq2graph_device(packed, rd, 27)
    lines = res.xyz_lines()
    assert len(lines) == facts["xyz_edges"]
    assert hashlib.sha256("\n".join(lines).encode()).hexdigest() == facts["xyz_fileorder_sha256"]
    assert res.graph.n_components == facts["n_components"]
    assert res.rows(packed, data) == [tuple(r) for r in facts["rows"]]


@pytest.mark.parametrize("k", [15, 21, 27])
def test_cfg3_full_size_checksum(k):
    """BASELINE config 3 (200 x 5 Mbp = 1 Gbp, k sweep 15/21/27): 2.0 G insertions through the streaming builder; the table
    checksum must equal the C oracle's (tests/golden/cfg3_oracle_facts.json, 12-15 CPU-minutes per k to produce).  First the
    one-shot product build (upper-bound table of 2^31 slots: too many regions, L2-atomic K3), then the serving loop: the
    second build of an adaptive builder runs at load <= 0.5 in shared-memory regions behind a multi-level partition.
    Needs ~60 GB of HBM."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    from pangenome_b200 import builder, engine
    facts = json.load(open(os.path.join(GOLDEN, "cfg3_oracle_facts.json")))
    data = _cfg3_data()
    assert len(data) == facts["file_bytes"]
    packed = engine.PackedSeqs(engine.to_device_bytes(data))
    f = facts["k"][str(k)]
    assert packed.n_insertions(k) == f["n_inserts"]
    t, _, b = builder.build_table(packed, k)
    assert list(t.checksum()) == f["dbg_checksum"]
    assert t.n_keys() * 2 == f["dbg_entries"]          # no palindromes at odd k without N
    b.close()
    del t, b
    torch.cuda.empty_cache()
    from pangenome_b200 import _lib
    b2 = builder.RoundBuilder(k, _lib.PG_MODE_CANONICAL, len(data))
    for i in range(2):
        b2.begin()
        t2 = b2.build_async(packed, packed.n_rec)
        torch.cuda.synchronize()
        b2.verify()
        assert list(t2.checksum()) == f["dbg_checksum"], (i, b2.describe())
    assert b2.region_bits == 12 and len(b2.levels) >= 2, b2.describe()
    assert f["dbg_entries"] / 2 / t2.capacity <= 0.7
    b2.close()


_CFG3 = {}


def _cfg3_data():
    if "d" not in _CFG3:
        d = pangenome(200, 5_000_000)
        facts = json.load(open(os.path.join(GOLDEN, "cfg3_oracle_facts.json")))
        assert hashlib.sha256(d).hexdigest() == facts["file_sha256"]
        _CFG3["d"] = d
    return _CFG3["d"]
