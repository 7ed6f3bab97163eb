"""Randomised parity: many tiny FASTA files with hostile line structure (blank lines, '>' inside
sequence lines, CR/LF mixes, junk before the first header, IUPAC/lowercase/N, records of length
0..k+3 except k+1) through the whole GPU pipeline vs the oracle.  Seeded, bit-exact."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ALPH = np.frombuffer(b"ACGTACGTACGTACGTNnacgtRYKM>", dtype=np.uint8)


def make_case(rng, k):
    parts = []
    if rng.random() < 0.3:
        parts.append(bytes(rng.choice(ALPH[:16], size=int(rng.integers(1, 30))).tolist()) + b"\n")     # junk before the first header
    for r in range(int(rng.integers(1, 7))):
        n = int(rng.choice([0, 1, k - 1, k, k + 2, k + 3, int(rng.integers(k + 2, 6 * k + 40))]))
        n = max(0, n)
        seq = bytes(rng.choice(ALPH[:-1] if rng.random() < 0.8 else ALPH, size=n).tolist())
        if rng.random() < 0.3 and n > 2 * k:
            a = int(rng.integers(0, n - k))
            seq = seq[:a] + seq[a:a + k + 3] + seq[a:]                 # a repeat
        hdr = b">r%d %s" % (r, bytes(rng.choice(ALPH[:8], size=int(rng.integers(0, 12))).tolist()))
        width = int(rng.choice([1, 7, 16, 31, 60, 1000]))
        nl = b"\r\n" if rng.random() < 0.15 else b"\n"
        lines = [seq[i:i + width] for i in range(0, len(seq), width)] or ([b""] if rng.random() < 0.5 else [])
        body = b"".join(l + nl for l in lines)
        if rng.random() < 0.2:
            body = body.replace(nl, nl + b"\n", 1)                     # a blank line inside the record
        parts.append(hdr + nl + body)
    data = b"".join(parts)
    if rng.random() < 0.3 and data.endswith(b"\n"):
        data = data[:-1]                                               # no final newline (Q8)
    return data


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_whole_pipeline(seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import engine, graph
    rng = np.random.default_rng(1000 + seed)
    n_checked = 0
    for it in range(40):
        k = int(rng.choice([3, 4, 5, 7, 11, 16, 21, 27]))
        c = int(rng.choice([2, 2, 2, 3, 0, 1]))
        data = make_case(rng, k)
        if data[:1] == b"\n":
            continue                       # the reference's path stage mis-parses a leading blank line (:1833); out of parity scope
        ref = oracle.run(data, k, c=c)
        if ref["ub_count"]:
            continue                       # a record of length k+1 (after parsing): undefined upstream (Q2)
        rc0, rc1 = bool((c >> 1) & 1), bool(c & 1)
        packed = engine.PackedSeqs(engine.to_device_bytes(data))
        assert packed.seq_off.tolist() == (ref["seq_off"] + (packed.seq_off[0] if packed.n_rec else 0)).tolist(), data
        assert packed.hdr_off.tolist() == ref["hdr_off"].tolist(), data
        t, _ = engine.build_dbg(packed, k, rc=rc0)
        ks, vs, cs = t.export()
        assert ks.tolist() == ref["dbg"][0].tolist() and vs.tolist() == ref["dbg"][1].tolist() and cs.tolist() == ref["dbg"][2].tolist(), data
        t2, _, _ = engine.build_dbg_partitioned(packed, k, rc=rc0)
        assert t2.checksum() == t.checksum(), data
        if rc0:      # compact 8-byte records (here nearly every position is an edge or ambiguous: the wide path)
            from pangenome_b200 import builder
            t3, _, b3 = builder.build_table(packed, k, rc=True, capacity=1 << 13, region_bits=12, sample=False, rounds=1 + it % 2)
            assert b3.compact and t3.checksum() == t.checksum(), data
            b3.close()
        rd = t.select_rdbg()
        assert rd.rdbg_export()[0].tolist() == ref["rdbg"].tolist(), data
        res = graph.seq2graph_device(packed, rd, k, rc=rc1)
        assert res.xyz_lines() == ref["xyz"], data
        assert res.rows(packed, data) == ref["rows"], data
        n_checked += 1
    assert n_checked >= 20
