"""`_db.npz` interop.  CPU part: the oakht image written by pangenome_b200.npz is
what the reference's own loader expects - checked by probing like `oakht.pointer` (always) and by
running the reference's load_on_disk + dbg2rdbg on it (only where oracle/_ref exists).  GPU part:
`-d file.npz` produced by the reference-layout writer round-trips through the GPU table."""
import os

import numpy as np
import pytest

import oracle
from conftest import load_small_cases
from oracle import refrun

CASES = {c["name"]: c for c in load_small_cases()}


def _triples(case):
    k = np.array([r[0] for r in case["dbg"]], dtype=np.uint64)
    v = np.array([r[1] for r in case["dbg"]], dtype=np.uint16)
    c = np.array([r[2] for r in case["dbg"]], dtype=np.uint8)
    return k, v, c


def _fnv4(v):
    a = 0xcbf29ce484222325
    for _ in range(4):
        a ^= v & 0xff
        a = (a * 0x100000001b3) & (2 ** 64 - 1)
        v >>= 8
    return a


@pytest.mark.parametrize("name", ["test_fsa_k5", "test_fsa_k27", "nasty_k11_lf", "rand05_k15_c2"])
def test_oakht_image_is_probeable_like_the_reference(name):
    from pangenome_b200 import npz
    case = CASES[name]
    keys, vals, cnts = _triples(case)
    params, ok, ov, oc = npz.oakht_image(keys, vals, cnts)
    ref = oracle.run(case["input_latin1"].encode("latin-1"), case["k"], c=case["c"], stages=1, image=True)
    assert int(params[0]) == ref["dbg_capacity"] and int(params[2]) == keys.size     # same capacity / size as upstream
    assert params.tolist()[1:] == [750000000, keys.size, 1, 1, 0]
    M = int(params[0])
    for k_, v_, c_ in zip(keys.tolist(), vals.tolist(), cnts.tolist()):
        j = _fnv4(k_) % M
        j0, t = j, 0
        while not (int(ok[j]) == k_ or oc[j] == 0):      # oakht.pointer (:521-538)
            j = (j0 + t * t) % M
            t += 1
        assert int(ok[j]) == k_ and ov[j] == v_ and oc[j] == c_


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref (patched reference) not built on this machine")
def test_reference_loads_our_image(tmp_path):
    """The reference's load_on_disk + dbg2rdbg on an image we wrote -> the golden rdBG."""
    from pangenome_b200 import npz
    mod = refrun.load()
    case = CASES["nasty_k11_lf"]
    keys, vals, cnts = _triples(case)
    params, ok, ov, oc = npz.oakht_image(keys, vals, cnts)
    fn = str(tmp_path / "x_db")
    np.savez_compressed(fn, parameters=params, keys=ok, values=ov, counts=oc)
    offset, kd = mod.load_on_disk(fn + ".npz")
    assert kd.size == keys.size
    rd = mod.dbg2rdbg(kd)
    rk, _, _ = refrun.table_triples(rd)
    assert rk.tolist() == case["rdbg"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_fsa_k5", "nasty_k11_lf", "rand05_k15_c2", "rand21_k15_c0"])
def test_gpu_dump_and_load_roundtrip(name, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import engine, graph, npz
    case = CASES[name]
    data = case["input_latin1"].encode("latin-1")
    k, c = case["k"], case["c"]
    rc0, rc1 = bool((c >> 1) & 1), bool(c & 1)
    packed = engine.PackedSeqs(engine.to_device_bytes(data))
    t, _ = engine.build_dbg(packed, k, rc=rc0)
    fn = npz.dump(t, str(tmp_path / "in.fa_db"))
    z = np.load(fn)
    live = z["counts"] > 0
    o = np.argsort(z["keys"][live])
    got = [[int(a), int(b), int(d)] for a, b, d in zip(z["keys"][live][o], z["values"][live][o], z["counts"][live][o])]
    assert got == case["dbg"]
    # -d: load the file into a (literal-key) GPU table and run the remaining stages on it
    t2 = npz.load(fn, k)
    ks, vs, cs = t2.export()
    assert [[int(a), int(b), int(d)] for a, b, d in zip(ks, vs, cs)] == case["dbg"]
    rd = t2.select_rdbg()
    rk, _ = rd.rdbg_export()
    assert rk.tolist() == case["rdbg"]
    res = graph.seq2graph_device(packed, rd, k, rc=rc1)
    assert res.xyz_lines() == case["xyz"]
    assert res.rows(packed, data) == [tuple(r) for r in case["rows"]]
