#!/usr/bin/env python
"""Generate tests/golden/*.json by running the REFERENCE ITSELF.

Run in the build container (needs /root/reference):

    python oracle/make_ref.py && python tests/golden/make_golden.py

Each case stores the input FASTA bytes and what the patched reference
(oracle/_ref, see oracle/make_ref.py for the five patches) produced for it:
sorted dBG (key,val,count) triples, sorted rdBG keys, the ``.xyz`` edge lines in
file order, the cluster file written by the connected-components stand-in for
``mcl`` and the region table rows.  ``big_4x1M.json`` stores only sizes and
sha256 digests for the SURVEY App. C 4 x 1 Mbp set.

The reference exhibits undefined behaviour for records of length k+1 (Q2);
the generators below never emit such records.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refrun  # noqa: E402


def fasta(records, width=60, crlf=False, final_newline=True):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for hdr, seq in records:
        out.append(b">" + hdr + nl)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width] + nl)
    data = b"".join(out)
    if not final_newline and data.endswith(nl):
        data = data[:-len(nl)]
    return data


def mutate(rng, s, rate):
    s = bytearray(s)
    for i in range(len(s)):
        if rng.random() < rate:
            s[i] = rng.choice(list(b"ACGT"))
    return bytes(s)


def rand_case(rng, k):
    """Multi-record input with shared ancestry, repeats, N runs, lowercase,
    IUPAC, poly-A, records of length k, < k and 0; never length k+1."""
    L = int(rng.integers(40, 400))
    anc = bytes(rng.choice(list(b"ACGT"), size=L).tolist())
    if rng.random() < 0.5:   # inject a repeat
        a = int(rng.integers(0, L - 10))
        rep = anc[a:a + int(rng.integers(5, 40))]
        b = int(rng.integers(0, L))
        anc = anc[:b] + rep + anc[b:]
    recs = []
    for g in range(int(rng.integers(1, 6))):
        s = bytearray(mutate(rng, anc, 0.03))
        r = rng.random()
        if r < 0.15:
            p = int(rng.integers(0, len(s)))
            s[p:p + int(rng.integers(1, 6))] = b"N" * int(rng.integers(1, 6))
        elif r < 0.3:
            p = int(rng.integers(0, len(s)))
            s[p:p + 8] = bytes(s[p:p + 8]).lower()
        elif r < 0.4:
            p = int(rng.integers(0, len(s)))
            s[p:p + 1] = bytes([rng.choice(list(b"RYKMSWn"))])
        elif r < 0.5:
            p = int(rng.integers(0, len(s)))
            s[p:p] = b"A" * (k + int(rng.integers(0, 6)))
        recs.append((b"g%d synthetic %d" % (g, k), bytes(s)))
    r = rng.random()
    if r < 0.2:
        recs.insert(int(rng.integers(0, len(recs) + 1)), (b"exactk", anc[:k]))
    elif r < 0.4:
        recs.insert(int(rng.integers(0, len(recs) + 1)), (b"short", anc[:max(0, k - 1 - int(rng.integers(0, 3)))]))
    elif r < 0.5:
        recs.insert(int(rng.integers(0, len(recs) + 1)), (b"empty", b""))
    width = int(rng.choice([17, 60, 80, 1000]))
    final = rng.random() > 0.2
    return no_ub(lambda rr: fasta(rr, width=width, final_newline=final), recs, k)


def nasty(k):
    recs = [
        (b"polyA desc text", b"A" * (k + 9) + b"CGT" + b"A" * (k + 2)),
        (b"nrun", b"ACGTACGTTGCA" * 3 + b"NNNNN" + b"ACGGTCATTGCA" * 3 + b"N" + b"GATTACA" * 4),
        (b"lower", b"acgtacgttgcaACGTTGCATTGACGGTCAttgacca" * 2),
        (b"iupac", b"ACGTRYKMACGTTGCASWBDHVACGTTGACCATG" * 2),
        (b"short", b"ACG"[:max(0, min(3, k - 1))]),
        (b"exactk", (b"GATTACAGATTACAGATTACAGATTACAGATTACA")[:k]),
        (b"empty", b""),
        (b"rep1", b"TTGACGGTCATTGCAGGCATTACGGATCGATCGGCTAGCTAGGCTA" * 3),
        (b"rep2", b"TTGACGGTCATTGCAGGCATTACGGATCGTTCGGCTAGCTAGGCTA" * 2 + b"G"),
    ]
    recs = [(h, s if len(s) != k + 1 else s + b"C") for h, s in recs]
    return recs


def no_ub(build, recs, k):
    """Re-build until no record parses to length k+1 (Q2: undefined upstream).
    Lengths are taken after parsing (CRLF adds a base per line, a missing final
    newline drops one), using the C oracle's record parser."""
    import oracle
    recs = list(recs)
    for _ in range(8):
        data = build(recs)
        r = oracle.run(data, k, stages=1)
        lens = np.diff(r["seq_off"])
        bad = [i for i, n in enumerate(lens.tolist()) if n == k + 1]
        if not bad:
            return data
        for i in bad:
            recs[i] = (recs[i][0], recs[i][1] + b"C")
    raise RuntimeError("could not avoid n == k+1")


def digest(lines):
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def pack(name, data, k, c, res):
    ks, vs, cs = res["dbg"]
    return {
        "name": name, "k": k, "c": c, "input_latin1": data.decode("latin-1"),
        "dbg": [[int(a), int(b), int(d)] for a, b, d in zip(ks, vs, cs)],
        "rdbg": [int(x) for x in res["rdbg"]],
        "xyz": res["xyz"], "mcl": res["mcl"],
        "rows": [list(r) for r in res["rows"]],
        "reference_error": res["error"],
    }


def survey_4x1m():
    """SURVEY.md App. C generator, draw order exactly as specified."""
    rng = np.random.default_rng(1234)
    L = 1_000_000
    anc = rng.integers(0, 4, L, dtype=np.uint8)
    recs = []
    for g in range(4):
        s = anc.copy()
        m = rng.random(L) < 0.01
        s[m] = (s[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) % 4
        recs.append((b"g%d synthetic" % g, np.frombuffer(b"ACGT", dtype=np.uint8)[s].tobytes()))
    return fasta(recs, width=80)


def main_small():
    small = []
    test_fsa = open("/root/reference/test/test.fsa", "rb").read()
    for k in (5, 27):
        small.append(pack("test_fsa_k%d" % k, test_fsa, k, 2, refrun.run(test_fsa, k)))
        print("test.fsa k=%d done" % k, flush=True)
    for k in (5, 11, 27):
        recs = nasty(k)
        for variant, kw in (("lf", {}), ("crlf", {"crlf": True}), ("nofinal", {"final_newline": False})):
            data = no_ub(lambda rr: fasta(rr, width=25, **kw), recs, k)
            small.append(pack("nasty_k%d_%s" % (k, variant), data, k, 2, refrun.run(data, k)))
    print("nasty done", flush=True)
    rng = np.random.default_rng(20261018)
    ks = [3, 4, 5, 7, 11, 15, 21, 27]
    for i in range(32):
        k = ks[i % len(ks)]
        data = rand_case(rng, k)
        c = 2 if i < 20 else [0, 1, 3][i % 3]
        small.append(pack("rand%02d_k%d_c%d" % (i, k, c), data, k, c, refrun.run(data, k, c=c)))
    # -n cap: the record prefix each stage sees (kmer_numba.py:1227,1820,1847)
    data = rand_case(rng, 7)
    res = refrun.run(data, 7, Ns=300)
    d = pack("ncap_k7", data, 7, 2, res)
    d["Ns"] = 300
    small.append(d)
    with open(os.path.join(HERE, "small_cases.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": small}, f, separators=(",", ":"))
    print("small cases: %d" % len(small), flush=True)


def main_big():
    # separate process from main_small(): numba dispatches the python int 2**63 (the CLI default
    # for -n) onto an int64 signature compiled earlier for Ns=300 and raises OverflowError
    big = survey_4x1m()
    assert len(big) == 4050056, len(big)
    res = refrun.run(big, 27)
    ks_, vs_, cs_ = res["dbg"]
    facts = {
        "name": "survey_4x1M_k27", "k": 27, "c": 2, "generator": "survey_4x1m() in tests/golden/make_golden.py",
        "file_bytes": len(big), "file_sha256": hashlib.sha256(big).hexdigest(),
        "dbg_entries": int(ks_.size), "dbg_count_sum": int(cs_.astype(np.int64).sum()),
        "dbg_sha256": digest(["%d\t%d\t%d" % (a, b, d) for a, b, d in zip(ks_.tolist(), vs_.tolist(), cs_.tolist())]),
        "rdbg_entries": int(res["rdbg"].size), "rdbg_sha256": digest(["%d" % x for x in res["rdbg"].tolist()]),
        "xyz_edges": len(res["xyz"]), "xyz_sorted_sha256": digest(sorted(res["xyz"])),
        "xyz_fileorder_sha256": digest(res["xyz"]),
        "mcl_clusters": len(res["mcl"]), "rows": [list(r) for r in res["rows"]],
    }
    with open(os.path.join(HERE, "big_4x1M.json"), "w") as f:
        json.dump(facts, f, indent=1)
    print(json.dumps({k: v for k, v in facts.items() if "sha" not in k}))


def main_chunk():
    """Chunk checkpoints of the dBG stage (kmer_numba.py:1252-1266): the reference's
    ``seq2rdbg`` with a small ``chunk``.  Stored per case: the final dBG, the offset and the live
    entries of the last ``<qry>_db_brkpt.npz`` it left behind, and that file itself (base64) so the
    GPU path can resume from the reference's own image."""
    import base64
    import io
    rng = np.random.default_rng(20260)
    cases = []
    tiny = (b">a\nACGTTGCAAGGCTTAACCGGATAGGCTTACGATCGGA\nGGCTTAACCAGT\n>b\nTTGACGGTCATTGACCAGTA\n>c\nAC\n"
            b">d\nGGGATTTACCCAGATTTAGGACCA\n")
    # (a file without a final newline is left out: when the checkpoint falls on its last record the reference
    # re-parses from the last line, finds no newline at all and segfaults)
    inputs = [("tiny", tiny, 5),
              ("nasty_k7", no_ub(lambda r: fasta(r, width=30), nasty(7), 7), 7),
              ("nasty_crlf_k11", no_ub(lambda r: fasta(r, width=25, crlf=True), nasty(11), 11), 11)]
    for i in range(4):
        k = int(rng.choice([5, 9, 15, 27]))
        import oracle
        for _ in range(20):       # needs a final newline (see above) and still no record of length k+1 (Q2)
            data = rand_case(rng, k)
            if not data.endswith(b"\n"):
                data += b"\n"
            if (k + 1) not in np.diff(oracle.run(data, k, stages=1)["seq_off"]).tolist():
                break
        else:
            raise RuntimeError("no usable random case")
        inputs.append(("rand%d_k%d" % (i, k), data, k))
    # every -n 2**63 case first: numba dispatches the python int 2**63 onto an int64 signature compiled for a
    # small -n and raises OverflowError (same reason main_small/main_big are separate processes)
    todo = []
    for name, data, k in inputs:
        total = 2 * sum(len(x) for x in data.split(b"\n") if not x.startswith(b">"))
        for c in (2, 0):
            for chunk in sorted({10, max(20, total // 5), max(30, total // 2), total + 1000}):
                for Ns in ((2 ** 63, 150) if (name.startswith("tiny") and chunk < 100) else (2 ** 63,)):
                    todo.append((Ns != 2 ** 63, len(todo), name, data, k, c, chunk, Ns))
    for _, _, name, data, k, c, chunk, Ns in sorted(todo):
        if True:
            if True:
                if True:
                    r = refrun.run_chunked(data, k, c=c, Ns=Ns, chunk=chunk)
                    case = {"name": "%s_c%d_chunk%d%s" % (name, c, chunk, "" if Ns == 2 ** 63 else "_n%d" % Ns),
                            "input_latin1": data.decode("latin-1"), "k": k, "c": c, "chunk": chunk, "Ns": Ns,
                            "dbg": [[int(a), int(b), int(d)] for a, b, d in zip(*[x.tolist() for x in r["dbg"]])],
                            "offset": r["offset"], "brkpt_b64": None, "brkpt_live": None}
                    if r["brkpt"] is not None:
                        z = np.load(io.BytesIO(r["brkpt"]))
                        live = z["counts"] > 0
                        o = np.argsort(z["keys"][live], kind="stable")
                        case["brkpt_live"] = [[int(a), int(b), int(d)] for a, b, d in
                                              zip(z["keys"][live][o].tolist(), z["values"][live][o].tolist(), z["counts"][live][o].tolist())]
                        case["brkpt_b64"] = base64.b64encode(r["brkpt"]).decode()
                        # what ``-r <that file>`` gives (same -n, the CLI's chunk); without a -n cap it must be
                        # the final table again
                        rr = refrun.run_chunked(data, k, c=c, Ns=Ns, brkpt_bytes=r["brkpt"])
                        case["resume_dbg"] = [[int(a), int(b), int(d)] for a, b, d in zip(*[x.tolist() for x in rr["dbg"]])]
                        if Ns == 2 ** 63:
                            assert case["resume_dbg"] == case["dbg"], case["name"]
                    cases.append(case)
                    print(case["name"], len(case["dbg"]), case["offset"], flush=True)
    with open(os.path.join(HERE, "chunk_cases.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases), "chunk cases")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        import subprocess
        for w in ("small", "big", "chunk"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), w])
    elif which == "small":
        main_small()
    elif which == "chunk":
        main_chunk()
    else:
        main_big()
