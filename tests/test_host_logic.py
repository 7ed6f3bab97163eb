"""CPU tests of the host side: CLI parsing, C-ABI symbols, loud failure without CUDA."""
import ctypes
import io
import os
import re

import pytest

from conftest import ROOT


def test_cli_parse_like_reference():
    from pangenome_b200 import cli
    a, extra, flags = cli.parse_args(["prog", "-m", "-i", "x.fa", "-k27", "-n", "5e8", "-c", "3", "--min-edge-weight", "2", "-z"])
    assert a["-i"] == "x.fa" and a["-k"] == "27" and a["-c"] == "3" and a["-n"] == "5e8"
    assert extra["--min-edge-weight"] == "2"
    assert cli._eval_n("2**63") == 2 ** 63 and cli._eval_n("5e8") == 500000000
    out = io.StringIO()
    with pytest.raises(SystemExit):
        cli.entry_point(["prog"], out=out)
    assert out.getvalue().startswith("Usage:") and len(out.getvalue().strip().split("\n")) == 11


def test_abi_exports_every_declared_symbol():
    """libpgdbg.so loads and exports each function include/pgdbg.h declares; the ctypes table
    covers exactly that set."""
    from pangenome_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "pgdbg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", hdr))
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().pg_version() >= 100
    assert _lib.load().pg_pack_words(1600) == 132


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pangenome_b200 import engine, _lib
    with pytest.raises(_lib.PgError):
        engine.to_device_bytes(b">a\nACGT\n")
    with pytest.raises(_lib.PgError):
        engine.DbgTable(1024, 5, 2)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pangenome_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
    src = open(os.path.join(ROOT, "kmer_b200.py")).read()
    assert "oracle" not in src


def test_record_prefix_ns():
    """-n semantics: records are consumed until the running count exceeds Ns (the crossing record is kept)."""
    import numpy as np
    from pangenome_b200.engine import PackedSeqs
    p = PackedSeqs.__new__(PackedSeqs)
    p._host = dict(seq_off=np.array([0, 100, 250, 400, 1000]), n_rec=4)
    assert p.record_prefix(2 ** 63, 2) == 4
    assert p.record_prefix(300, 1) == 3      # 100, 250, 400 > 300
    assert p.record_prefix(300, 2) == 2      # 200, 500 > 300
    assert p.record_prefix(0, 1) == 1


def test_record_ids_and_rows_vectorised():
    """seqid extraction from a byte view of the file (headers longer than the search window, no final newline, Q8) and
    the print order of the rows (per record: forward rows, then mirrored rc rows)."""
    import numpy as np
    from pangenome_b200 import graph
    long_hdr = b"L" * 400
    data = b">a desc\nACGT\n>" + long_hdr + b"\nAC\n>c\nGGGT\n>last no newline"
    offs = [i for i in range(len(data)) if data[i:i + 1] == b">" and (i == 0 or data[i - 1:i] == b"\n")]

    class P:
        hdr_off = np.array(offs)
        seq_lengths = np.array([4, 2, 4, 0])
    want = ["a desc", long_hdr.decode(), "c", "last no newlin"]
    assert graph.record_ids(P, data) == want
    assert graph.record_ids(P, np.frombuffer(data, dtype=np.uint8)) == want
    res = graph.GraphResult()
    res.rows_raw = [(np.array([2, 0, 0]), np.array([0, 0, 2]), np.array([3, 2, 4]), 1, np.array([7, 5, 6])),
                    (np.array([0]), np.array([1]), np.array([3]), -1, np.array([9]))]
    assert res.rows(P, data) == [("a desc", 0, 2, "+", 5), ("a desc", 2, 4, "+", 6), ("a desc", 1, 3, "-", 9), ("c", 0, 3, "+", 7)]


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The ctypes structures in pangenome_b200/_lib.py must have the size and field offsets gcc gives the structs of
    include/pgdbg.h (a silent mismatch would hand the kernels garbage)."""
    import ctypes
    import subprocess
    from pangenome_b200 import _lib
    structs = {"pg_table": _lib.PgTable, "pg_bucket_set": _lib.PgBucketSet, "pg_cbuckets": _lib.PgCBuckets, "pg_graph": _lib.PgGraph}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pgdbg.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_edge_stage_checkpoint_flag_is_rejected_loudly():
    """-R (kmer_numba.py:1876-1890: resume the edge-weight stage from <qry>_rdb_brkpt.npz) is deliberately not built: the
    reference only ever writes that file after 2^33 bases of one stage call, and its dump/reload cycle (Dict.popitem while
    dumping, re-insertion on load) reorders every later line of the .xyz file, so a faithful resume would have to
    replicate numba's typed-dict iteration order.  The CLI must say so instead of silently ignoring the flag."""
    from pangenome_b200 import cli
    with pytest.raises(SystemExit) as e:
        cli.entry_point(["prog", "-i", "x.fa", "-k", "27", "-R", "x.fa_rdb_brkpt.npz"], out=io.StringIO())
    assert "-R" in str(e.value) and "not supported" in str(e.value)
