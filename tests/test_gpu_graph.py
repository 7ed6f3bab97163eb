"""GPU parity of the path stages (K5-K8): .xyz edge list incl. file order and
weights, and the region table, against the golden vectors / the oracle."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import load_small_cases
from pangenome_b200.synth import pangenome, survey_4x1m

pytestmark = pytest.mark.gpu
CASES = load_small_cases()


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import engine, graph
    return engine, graph


def full(mods, data, k, c=2, Ns=2 ** 63, mode=None):
    eng, graph = mods
    rc0, rc1 = bool((c >> 1) & 1), bool(c & 1)
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, _ = eng.build_dbg(packed, k, rc=rc0, Ns=Ns, mode=mode)
    rd = t.select_rdbg()
    res = graph.seq2graph_device(packed, rd, k, Ns=Ns, rc=rc1)
    return packed, res


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_graph_golden(mods, case):
    data = case["input_latin1"].encode("latin-1")
    k, c, Ns = case["k"], case["c"], case.get("Ns", 2 ** 63)
    modes = (2, 1) if (c >> 1) & 1 else (0,)
    for mode in modes:
        packed, res = full(mods, data, k, c, Ns, mode)
        assert res.xyz_lines() == case["xyz"], "mode %d" % mode        # same edges, weights AND file order
        assert res.rows(packed, data) == [tuple(r) for r in case["rows"]], "mode %d" % mode
        if case["mcl"]:
            assert sorted(res.mcl_lines()) == sorted(case["mcl"])


def test_graph_big_4x1m(mods, big_facts):
    data = survey_4x1m()
    packed, res = full(mods, data, 27)
    lines = res.xyz_lines()
    assert len(lines) == big_facts["xyz_edges"] == 142203
    assert hashlib.sha256("\n".join(lines).encode()).hexdigest() == big_facts["xyz_fileorder_sha256"]
    assert res.rows(packed, data) == [tuple(r) for r in big_facts["rows"]]


def test_graph_many_components_and_rc(mods):
    """unrelated genomes -> several components; -c 3 walks both strands"""
    rng = np.random.default_rng(11)
    from pangenome_b200.synth import fasta_bytes, _ACGT
    recs = []
    for fam in range(5):
        anc = rng.integers(0, 4, 3000 + 500 * fam, dtype=np.uint8)
        for g in range(2 + fam % 3):
            s = anc.copy()
            m = rng.random(anc.size) < 0.02
            s[m] = (s[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) % 4
            recs.append((b"fam%d_g%d" % (fam, g), _ACGT[s]))
    data = fasta_bytes(recs, width=70)
    for c in (2, 3, 1, 0):
        for k in (9, 21):
            ref = oracle.run(data, k, c=c)
            packed, res = full(mods, data, k, c)
            assert res.xyz_lines() == ref["xyz"], (c, k)
            assert res.rows(packed, data) == ref["rows"], (c, k)
            assert len(set(res.nodes[3].tolist())) == ref["n_components"]


def test_graph_min_weight_filter(mods):
    """weak-edge filtering: components over edges of weight >= W only"""
    data = pangenome(4, 30000, snp=0.02, seed=9)
    eng, graph = mods
    packed = eng.PackedSeqs(eng.to_device_bytes(data))
    t, _ = eng.build_dbg(packed, 15)
    rd = t.select_rdbg()
    r1 = graph.seq2graph_device(packed, rd, 15, min_weight=1)
    r3 = graph.seq2graph_device(packed, rd, 15, min_weight=3)
    c0, v0, c1, v1, w = r3.edges
    # reference for W=3: union-find over the strong edges on the host
    names = {}
    for a, b in zip(zip(c0.tolist(), v0.tolist()), zip(c1.tolist(), v1.tolist())):
        names.setdefault(a, len(names)); names.setdefault(b, len(names))
    par = list(range(len(names)))
    def find(x):
        while par[x] != x:
            par[x] = par[par[x]]; x = par[x]
        return x
    for a, b, ww in zip(zip(c0.tolist(), v0.tolist()), zip(c1.tolist(), v1.tolist()), w.tolist()):
        if ww >= 3:
            ra, rb = find(names[a]), find(names[b])
            if ra != rb:
                par[ra] = rb
    n_comp = len({find(i) for i in range(len(names))})
    assert len(set(r3.nodes[3].tolist())) == n_comp
    assert len(set(r1.nodes[3].tolist())) <= n_comp
