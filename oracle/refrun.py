"""Run the patched reference (``oracle/_ref``) stage by stage and dump every
compared artefact in canonical form.

TEST INFRASTRUCTURE ONLY (golden-vector generation, oracle validation and the
``--impl reference`` CPU arm of bench.py).  Never imported by the product.

Stage calls mirror the reference CLI ``entry_point`` (kmer_numba.py:2103-2144):
``seq2rdbg`` -> ``dbg2rdbg`` -> ``seq2graph``; the ``.npz`` dump/reload between
stage 1 and 2 does not change the result and is skipped here.
"""
import contextlib
import importlib.util
import io
import os
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_mod = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, "ref_patched.py"))


def load(cache_jitclass=True):
    """Import oracle/_ref/ref_patched.py.  With ``cache_jitclass`` the
    reference's ``init_dict`` (kmer_numba.py:1097-1122), which builds a fresh
    jitclass *type* on every call and so re-JITs every stage, is wrapped to
    reuse one type per (ktype, vtype) - same code, compile once (SURVEY.md
    App. B "timing trick")."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("oracle/_ref not built: run oracle/make_ref.py where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    os.environ["PATH"] = os.path.join(REF_DIR, "bin") + os.pathsep + os.environ.get("PATH", "")
    spec = importlib.util.spec_from_file_location("ref_patched", os.path.join(REF_DIR, "ref_patched.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_patched"] = mod
    spec.loader.exec_module(mod)
    if cache_jitclass:
        nb = mod.nb
        cache = {}

        def init_dict(hashfunc=mod.oakht, capacity=2 ** 20, ksize=1, ktype=nb.uint64, vsize=1,
                      vtype=nb.uint32, jit=True):
            key = (str(ktype), str(vtype))
            if key not in cache:
                spec_ = {'capacity': nb.int64, 'load': nb.float32, 'size': nb.int64, 'ksize': nb.int64,
                         'vsize': nb.int64, 'keys': ktype[:], 'values': vtype[:], 'counts': nb.uint8[:]}
                cache[key] = mod.nb_jitclass(spec_)(hashfunc)
            return cache[key](capacity=2 ** 20, ksize=ksize, ktype=ktype, vsize=vsize, vtype=vtype)

        mod.init_dict = init_dict
    _mod = mod
    return mod


def table_triples(ht):
    """Sorted (key:u64, val:u16, cnt:u8) of an ``oakht`` (live slots: counts>0,
    as ``iteritems`` kmer_numba.py:623-631)."""
    import numpy as np
    live = ht.counts > 0
    keys = np.asarray(ht.keys)[live].astype(np.uint64)
    vals = np.asarray(ht.values)[live].astype(np.uint16)
    cnts = np.asarray(ht.counts)[live].astype(np.uint8)
    order = np.argsort(keys, kind="stable")
    return keys[order], vals[order], cnts[order]


def run(fasta_bytes, k, c=2, Ns=2 ** 63, stages="all", timings=None):
    """Run the reference on ``fasta_bytes`` (a scratch copy is written; the
    reference memmaps its input 'r+' and drops side files beside it).

    Returns dict(dbg=(keys,vals,cnts), rdbg=keys, xyz=[lines], mcl=[lines],
    rows=[(seqid,start,end,strand,label)], error=str|None).
    """
    import numpy as np
    mod = load()
    tmp = tempfile.mkdtemp(prefix="pgref_")
    out = {"error": None}
    try:
        qry = os.path.join(tmp, "in.fa")
        with open(qry, "wb") as f:
            f.write(fasta_bytes)
        rc0 = ((c >> 1) == 1)
        rc1 = ((c & 1) == 1)
        t0 = time.perf_counter()
        kd = mod.seq2rdbg(qry, k, 5, Ns, brkpt='', chunk=2 ** 33, rc=rc0)
        t1 = time.perf_counter()
        out["dbg"] = table_triples(kd)
        if stages == "dbg":
            if timings is not None:
                timings["dbg"] = t1 - t0
            return out
        t2 = time.perf_counter()
        rd = mod.dbg2rdbg(kd)
        t3 = time.perf_counter()
        rk, rv, _ = table_triples(rd)
        out["rdbg"] = rk
        out["rdbg_vals"] = rv
        buf = io.StringIO()
        t4 = time.perf_counter()
        try:
            with contextlib.redirect_stdout(buf):
                mod.seq2graph(qry, kmer=k, bits=5, Ns=Ns, rdbg_dict=rd, hashfunc=mod.oakht,
                              chunk=2 ** 33, brkpt='', rc=rc1)
        except Exception as e:  # F9: empty edge set -> untyped-dict TypeError
            out["error"] = "%s: %s" % (type(e).__name__, str(e).split("\n")[0])
        t5 = time.perf_counter()
        xyz = qry + "_rdbg_weight.xyz"
        out["xyz"] = open(xyz).read().split("\n")[:-1] if os.path.isfile(xyz) else []
        out["mcl"] = open(xyz + ".mcl").read().split("\n")[:-1] if os.path.isfile(xyz + ".mcl") else []
        rows = []
        for line in buf.getvalue().split("\n"):
            if line.startswith("#"):
                continue
            f = line.split("\t")
            if len(f) != 5:
                continue
            rows.append((f[0], int(f[1]), int(f[2]), f[3], int(f[4])))
        out["rows"] = rows
        if timings is not None:
            timings.update(dbg=t1 - t0, rdbg=t3 - t2, graph=t5 - t4)
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_chunked(fasta_bytes, k, c=2, Ns=2 ** 63, chunk=2 ** 33, brkpt_bytes=None):
    """Stage 1 only, with the reference's chunk checkpoints (kmer_numba.py:1252-1266): returns
    dict(dbg=(keys,vals,cnts), brkpt=bytes|None, offset=int|None) where ``brkpt`` is the last
    ``<qry>_db_brkpt.npz`` the run left behind.  ``brkpt_bytes``: resume from that image (``-r``)."""
    import numpy as np
    mod = load()
    tmp = tempfile.mkdtemp(prefix="pgref_")
    try:
        qry = os.path.join(tmp, "in.fa")
        with open(qry, "wb") as f:
            f.write(fasta_bytes)
        brk = ''
        if brkpt_bytes is not None:
            brk = os.path.join(tmp, "resume.npz")
            with open(brk, "wb") as f:
                f.write(brkpt_bytes)
        kd = mod.seq2rdbg(qry, k, 5, Ns, brkpt=brk, chunk=chunk, rc=((c >> 1) == 1))
        out = {"dbg": table_triples(kd), "brkpt": None, "offset": None}
        ck = qry + "_db_brkpt.npz"
        if os.path.isfile(ck):
            out["brkpt"] = open(ck, "rb").read()
            out["offset"] = int(np.load(ck)["parameters"][5])
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
