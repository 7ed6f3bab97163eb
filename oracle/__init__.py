"""CPU oracle for the dBG hot path - TEST INFRASTRUCTURE ONLY.

``oracle.run(fasta_bytes, k, ...)`` drives ``oracle/pgoracle.c`` (a plain-C
restatement of /root/reference/kmer_numba.py's algorithm, App. A of SURVEY.md)
through ctypes and returns every compared artefact as numpy arrays.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``pangenome_b200``) never does; it fails loudly without its CUDA library.

Parity status: pinned against the reference's own numba code (fixtures in
``tests/golden``; generator ``tests/golden/make_golden.py``); cluster
membership is unpinned (third-party ``mcl``), see pgoracle.c header.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "libpgoracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "pgoracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-shared", "-fPIC", "-o", _LIB_PATH, src])
    return _LIB_PATH


_PTRS = {
    "hdr_off": ctypes.c_int64, "hdr_len": ctypes.c_int64, "seq_off": ctypes.c_int64, "seq": ctypes.c_uint8,
    "dbg_keys": ctypes.c_uint64, "dbg_vals": ctypes.c_uint16, "dbg_cnts": ctypes.c_uint8,
    "dbg_slot_keys": ctypes.c_uint64, "dbg_slot_vals": ctypes.c_uint16, "dbg_slot_cnts": ctypes.c_uint8,
    "rdbg_keys": ctypes.c_uint64, "rdbg_vals": ctypes.c_uint16,
    "edge_c0": ctypes.c_uint64, "edge_v0": ctypes.c_uint64, "edge_c1": ctypes.c_uint64, "edge_v1": ctypes.c_uint64,
    "edge_w": ctypes.c_int64,
    "node_code": ctypes.c_uint64, "node_v": ctypes.c_uint64, "node_label": ctypes.c_int64,
    "row_rec": ctypes.c_int64, "row_start": ctypes.c_int64, "row_end": ctypes.c_int64,
    "row_strand": ctypes.c_int32, "row_label": ctypes.c_int64,
}
_INTS = ["n_records", "n_inserts", "ub_count", "dbg_size", "dbg_capacity", "rdbg_size", "n_edges", "n_nodes",
         "n_components", "n_rows"]


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.pgo_run.restype = ctypes.c_void_p
        L.pgo_run.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                              ctypes.c_int, ctypes.c_int]
        L.pgo_free.argtypes = [ctypes.c_void_p]
        L.pgo_time.restype = ctypes.c_double
        L.pgo_time.argtypes = [ctypes.c_void_p, ctypes.c_int]
        for name, ct in _PTRS.items():
            f = getattr(L, "pgo_" + name)
            f.restype = ctypes.POINTER(ct)
            f.argtypes = [ctypes.c_void_p]
        for name in _INTS:
            f = getattr(L, "pgo_" + name)
            f.restype = ctypes.c_int64
            f.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def _arr(L, h, name, n):
    if n <= 0:
        return np.zeros(0, dtype=np.dtype(_PTRS[name]))
    p = getattr(L, "pgo_" + name)(h)
    return np.ctypeslib.as_array(p, shape=(int(n),)).copy()


def run(fasta, k, c=2, Ns=2 ** 63, stages=4, edge_offbit=5, image=False):
    """Returns a dict of artefacts in the reference's conventions:

    dbg = (keys u64 ascending, vals u16, cnts u8); rdbg = keys u64 ascending;
    edges = (c0, v0, c1, v1, w) in the reference's file (first-insertion) order;
    nodes = (code, v5, label); rows = list of (seqid, start, end, strand, label);
    xyz = the lines of ``<in>_rdbg_weight.xyz``.
    """
    L = lib()
    buf = np.frombuffer(bytes(fasta), dtype=np.uint8)
    h = L.pgo_run(buf.ctypes.data if buf.size else None, buf.size, int(k), int(c), int(min(Ns, 2 ** 64 - 1)),
                  int(stages), int(edge_offbit))
    try:
        g = {n: int(getattr(L, "pgo_" + n)(h)) for n in _INTS}
        out = dict(g)
        nrec = g["n_records"]
        hdr_off = _arr(L, h, "hdr_off", nrec)
        hdr_len = _arr(L, h, "hdr_len", nrec)
        out["hdr_off"] = hdr_off
        out["seq_off"] = _arr(L, h, "seq_off", nrec + 1)
        out["seq"] = _arr(L, h, "seq", int(out["seq_off"][-1]) if nrec else 0)
        raw = bytes(fasta)
        out["headers"] = [raw[int(o):int(o) + int(l)] for o, l in zip(hdr_off, hdr_len)]
        out["dbg"] = (_arr(L, h, "dbg_keys", g["dbg_size"]), _arr(L, h, "dbg_vals", g["dbg_size"]),
                      _arr(L, h, "dbg_cnts", g["dbg_size"]))
        cap = g["dbg_capacity"]
        if image:   # raw oakht slot image, as the reference's `_db.npz` stores it (kmer_numba.py:243-261)
            out["dbg_image"] = (_arr(L, h, "dbg_slot_keys", cap), _arr(L, h, "dbg_slot_vals", cap),
                                _arr(L, h, "dbg_slot_cnts", cap))
        if stages >= 2:
            out["rdbg"] = _arr(L, h, "rdbg_keys", g["rdbg_size"])
            out["rdbg_vals"] = _arr(L, h, "rdbg_vals", g["rdbg_size"])
        if stages >= 3:
            ne = g["n_edges"]
            out["edges"] = tuple(_arr(L, h, "edge_" + n, ne) for n in ("c0", "v0", "c1", "v1", "w"))
            c0, v0, c1, v1, w = out["edges"]
            out["xyz"] = ["%d_%d\t%d_%d\t%d" % (a, b, cc, d, e) for a, b, cc, d, e in
                          zip(c0.tolist(), v0.tolist(), c1.tolist(), v1.tolist(), w.tolist())]
        if stages >= 4:
            nn = g["n_nodes"]
            out["nodes"] = (_arr(L, h, "node_code", nn), _arr(L, h, "node_v", nn), _arr(L, h, "node_label", nn))
            nr = g["n_rows"]
            rec = _arr(L, h, "row_rec", nr)
            st = _arr(L, h, "row_start", nr)
            ed = _arr(L, h, "row_end", nr)
            sd = _arr(L, h, "row_strand", nr)
            lb = _arr(L, h, "row_label", nr)
            ids = [hd[1:].decode("latin-1") for hd in out["headers"]]
            out["rows"] = [(ids[r], int(a), int(b), "+" if s == 1 else "-", int(l))
                           for r, a, b, s, l in zip(rec.tolist(), st.tolist(), ed.tolist(), sd.tolist(), lb.tolist())]
        out["times"] = {s: L.pgo_time(h, i) for i, s in enumerate(("dbg", "rdbg", "edge", "label"), 1)}
        return out
    finally:
        L.pgo_free(h)


def table_checksum(keys, vals, cnts):
    """Order-independent checksum of a (key, val, cnt) table: (n, sum, xor) of a
    64-bit mix of each triple.  The CUDA export computes the same on device."""
    keys = np.asarray(keys, dtype=np.uint64)
    w = (np.asarray(vals, dtype=np.uint64) << np.uint64(8)) | np.asarray(cnts, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = keys ^ (w * np.uint64(0x9E3779B97F4A7C15))
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
        s = int(x.sum(dtype=np.uint64)) if x.size else 0
    xo = int(np.bitwise_xor.reduce(x)) if x.size else 0
    return int(keys.size), s, xo
