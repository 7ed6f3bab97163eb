"""The reference's stage API on the B200 path: ``seq2rdbg``, ``dbg2rdbg``,
``seq2graph`` (kmer_numba.py:1234-1268, 1313-1321, 1853-1951) with the same
argument meaning.  ``seq2graph`` prints the table rows on stdout and writes
``<qry>_rdbg_weight.xyz`` (+ ``.mcl``) beside the input, like the reference.
"""
import os
import sys

import numpy as np

from . import _lib, engine, graph


class DbgHandle:
    """What seq2rdbg returns: the device table plus the packed sequences it was
    built from (later stages reuse them instead of re-reading the file)."""

    def __init__(self, table, packed, data, n_rec):
        self.table, self.packed, self.data, self.n_rec = table, packed, data, n_rec


def _read(qry):
    return np.fromfile(qry, dtype=np.uint8)


def seq_chk(qry):
    """kmer_numba.py:873-886: peek 40 bytes, '>' anywhere at a line start -> fasta."""
    with open(qry, "rb") as f:
        s = f.read(40)
    if s[:1] == b">" or b"\n>" in s:
        return "fasta"
    if s[:1] == b"@" or b"\n@" in s:
        return "fastq"
    return None


def seq2rdbg(qry, kmer=13, bits=5, Ns=1e6, chunk=2 ** 32, brkpt="./breakpoint", saved="dBG_disk", hashfunc=None,
             jit=True, rc=True, data=None):
    """Stage 1: build the dBG (both strands when ``rc``)."""
    if bits != 5:
        raise ValueError("only the reference's base-5 code (bits=5) is supported")
    kmer = min(max(1, int(kmer)), 27)
    if data is None:
        if seq_chk(qry) != "fasta":
            raise SystemExit("pangenome_b200: %s is not FASTA (the reference's fastq branch is broken upstream)" % qry)
        data = _read(qry)
    raw = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
    packed = engine.PackedSeqs(engine.to_device_bytes(data))
    table, n_rec = engine.build_dbg(packed, kmer, rc=bool(rc), Ns=Ns)
    return DbgHandle(table, packed, raw, n_rec)


def dump(kmer_dict, fn):
    """kmer_numba.py dump (:243-261): ``<fn>.npz`` in the reference's oakht layout."""
    from . import npz
    return npz.dump(kmer_dict.table, fn)


def load_dbg(qry, fn, kmer):
    """``-d`` / ``-D``: a table saved by the reference (or by ``dump``) + the packed input."""
    from . import npz
    kmer = min(max(1, int(kmer)), 27)
    data = _read(qry)
    packed = engine.PackedSeqs(engine.to_device_bytes(data))
    table = npz.load(fn, kmer, device=packed.pk2.device)
    return DbgHandle(table, packed, data.tobytes(), packed.n_rec)


def dbg2rdbg(kmer_dict):
    """Stage 2: keep nodes with indegree != 1 or outdegree != 1."""
    rd = kmer_dict.table.select_rdbg()
    return DbgHandle(rd, kmer_dict.packed, kmer_dict.data, kmer_dict.n_rec)


def seq2graph(qry, kmer=13, bits=5, Ns=1e6, brkpt="./breakpoint_rdbg.npz", rdbg_dict=None, saved=None, hashfunc=None,
              jit=True, chunk=2 ** 33, rc=False, cluster=True, min_weight=1, out=None, write_mcl=True):
    """Stages 3-5: edge weights -> ``<qry>_rdbg_weight.xyz``; components (or an existing
    ``.xyz.mcl`` cluster file) -> labels; region table on ``out`` (stdout)."""
    out = out or sys.stdout
    kmer = min(max(1, int(kmer)), 27)
    packed, data = rdbg_dict.packed, rdbg_dict.data
    oname = qry + "_rdbg_weight.xyz"
    mcl_lines = None
    if cluster and os.path.isfile(oname + ".mcl"):
        print("# the mcl has been ran", file=out)
        with open(oname + ".mcl") as f:
            mcl_lines = f.read().split("\n")
            if mcl_lines and mcl_lines[-1] == "":
                mcl_lines.pop()
    res = graph.seq2graph_device(packed, rdbg_dict.table, kmer, Ns=Ns, rc=bool(rc), min_weight=min_weight, mcl_lines=mcl_lines)
    res.write_xyz(oname)
    if mcl_lines is None and write_mcl:
        res.write_mcl(oname + ".mcl")
    for seqid, st, ed, strand, lab in res.rows(packed, data):
        out.write("%s\t%d\t%d\t%s\t%d\n" % (seqid, st, ed, strand, lab))
    return LabelDict(res.nodes[1], res.nodes[2], res.nodes[3])


class LabelDict(dict):
    """The reference returns label_dct {(code, v5): label}; building millions of Python tuples is the
    slowest thing left in the CLI, so the dict is only materialised when somebody looks at it."""

    def __init__(self, code, v5, label):
        super().__init__()
        self._arrays, self._built = (code, v5, label), False

    def _build(self):
        if not self._built:
            self._built = True
            code, v5, label = self._arrays
            super().update({(int(c), int(v)): int(l) for c, v, l in zip(code.tolist(), v5.tolist(), label.tolist())})

    def __getitem__(self, k):
        self._build()
        return super().__getitem__(k)

    def __contains__(self, k):
        self._build()
        return super().__contains__(k)

    def __len__(self):
        return int(self._arrays[0].size)

    def __iter__(self):
        self._build()
        return super().__iter__()

    def items(self):
        self._build()
        return super().items()
