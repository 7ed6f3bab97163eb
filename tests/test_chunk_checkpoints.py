"""Chunk checkpoints of the dBG stage (kmer_numba.py:1252-1266, -r).

The golden cases were produced by the reference's own ``seq2rdbg`` with a small ``chunk``
(tests/golden/make_golden.py chunk): final table, offset and live entries of the last
``<qry>_db_brkpt.npz``, that file itself, and what resuming from it gives.
CPU: the planning logic against the reference's offsets (records parsed by the oracle).
GPU: the chunked build, the checkpoint written, and resuming from the REFERENCE's image.
"""
import base64
import io
import json
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "chunk_cases.json")))
IDS = [c["name"] for c in CASES]


def _expected_offset(case):
    """Offset of the last checkpoint according to plan_chunks + the oracle's record parser."""
    from pangenome_b200.stages import plan_chunks, _last_line_start
    data = case["input_latin1"].encode("latin-1")
    r = oracle.run(data, case["k"], c=case["c"], stages=1)
    strands = 2 if (case["c"] >> 1) & 1 else 1
    lens = np.diff(r["seq_off"]).astype(np.int64) * strands
    off, tail = None, False
    for r0, r1, ckpt, is_tail in plan_chunks(lens, case["chunk"], case["Ns"]):
        if is_tail:
            tail = True
            break
        if ckpt:
            off = int(r["hdr_off"][r1]) if r1 < len(lens) else _last_line_start(data)
    return off, tail


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_plan_matches_reference_offsets(case):
    off, tail = _expected_offset(case)
    assert off == case["offset"]
    # the only way a chunked build differs from the one-shot build of the same records: the extra empty
    # record upstream parses after a checkpoint on the last record (2 sentinel hits when rc)
    data = case["input_latin1"].encode("latin-1")
    if case["Ns"] == 2 ** 63:
        ref = oracle.run(data, case["k"], c=case["c"], stages=1)
        want = {int(a): (int(b), int(d)) for a, b, d in zip(*ref["dbg"])}
        got = {a: (b, d) for a, b, d in case["dbg"]}
        sent = 0xFFFFFFFFFFFFFFFF
        strands = 2 if (case["c"] >> 1) & 1 else 1
        quirk = tail and data[:1] == b">" and data.find(b"\n", case["offset"]) >= 0
        if quirk:
            v, n = want.get(sent, (32, 0))
            want[sent] = (32, min(255, n + strands))
        assert got == want


def test_last_line_start():
    from pangenome_b200.stages import _last_line_start
    assert _last_line_start(b">a\nAC\n>b\nGG\n") == 9
    assert _last_line_start(b">a\nAC\n>b\nGG") == 9
    assert _last_line_start(b">a\n") == 0
    assert _last_line_start(b">a") == 0
    assert _last_line_start(b"") == 0


@pytest.fixture(scope="module")
def stages():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import stages
    return stages


def _triples(table):
    ks, vs, cs = table.export()
    return [[int(a), int(b), int(d)] for a, b, d in zip(ks, vs, cs)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_chunked_build_and_resume(stages, case, tmp_path):
    data = case["input_latin1"].encode("latin-1")
    qry = str(tmp_path / "in.fa")
    open(qry, "wb").write(data)
    rc0 = bool((case["c"] >> 1) & 1)
    h = stages.seq2rdbg(qry, case["k"], 5, case["Ns"], chunk=case["chunk"], brkpt="", rc=rc0)
    assert _triples(h.table) == case["dbg"]
    ck = qry + "_db_brkpt.npz"
    assert os.path.isfile(ck) == (case["offset"] is not None)
    if case["offset"] is None:
        return
    z = np.load(ck)
    assert int(z["parameters"][5]) == case["offset"]
    live = z["counts"] > 0
    o = np.argsort(z["keys"][live], kind="stable")
    got = [[int(a), int(b), int(d)] for a, b, d in
           zip(z["keys"][live][o].tolist(), z["values"][live][o].tolist(), z["counts"][live][o].tolist())]
    assert got == case["brkpt_live"]
    # -r with the REFERENCE's own checkpoint image
    ref_ck = str(tmp_path / "ref_brkpt.npz")
    open(ref_ck, "wb").write(base64.b64decode(case["brkpt_b64"]))
    h2 = stages.seq2rdbg(qry, case["k"], 5, case["Ns"], chunk=2 ** 33, brkpt=ref_ck, rc=rc0)
    assert _triples(h2.table) == case["resume_dbg"]
    # ... and with ours
    h3 = stages.seq2rdbg(qry, case["k"], 5, case["Ns"], chunk=2 ** 33, brkpt=ck, rc=rc0)
    assert _triples(h3.table) == case["resume_dbg"]
    # later stages run on the resumed (literal-key) table like on a one-shot build; the empty record of the
    # last-record quirk can add the short-record sentinel (value 32: an rdBG member) where there was none
    if case["Ns"] == 2 ** 63:
        want = oracle.run(data, case["k"], c=case["c"], stages=2)["rdbg"].tolist()
        sent = 0xFFFFFFFFFFFFFFFF
        if any(e[0] == sent for e in case["resume_dbg"]) and sent not in want:
            want.append(sent)
        assert stages.dbg2rdbg(h2).table.rdbg_export()[0].tolist() == want
