"""`_db.npz` interop with the reference: ``dump`` writes a GPU-built dBG in the
layout of kmer_numba.py's ``dump`` (:243-261) so that ``kmer_numba.py -d`` can load it; ``load``
reads such a file (written by either side) into a GPU table (``load_on_disk`` :289-335)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check


def oakht_image(keys, vals, cnts, offset=0):
    """(parameters, keys[cap], values[cap], counts[cap]) - the arrays the reference saves."""
    L = _lib.load()
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    vals = np.ascontiguousarray(vals, dtype=np.uint16)
    cnts = np.ascontiguousarray(cnts, dtype=np.uint8)
    n = int(keys.size)
    cap = int(L.pg_host_oakht_capacity(n))
    okeys = np.empty(cap, np.uint64)
    ovals = np.empty(cap, np.uint16)
    ocnts = np.empty(cap, np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    check(L.pg_host_build_oakht(P(keys), P(vals), P(cnts), n, cap, P(okeys), P(ovals), P(ocnts)), "pg_host_build_oakht")
    # parameters = [capacity, load_factor * 1e9, size, ksize, vsize, offset]   (:252-258)
    params = np.asarray([cap, int(0.75 * 1e9), n, 1, 1, int(offset)], dtype="uint64")
    return params, okeys, ovals, ocnts


def dump(table, fn, compressed=True, offset=0):
    """Write ``fn`` (``.npz`` appended like the reference does) from a DbgTable.  ``offset``: the byte
    offset a chunk checkpoint resumes from (parameters[5], :258)."""
    k, v, c = table.export(sort=False)
    return dump_entries(k, v, c, fn, compressed, offset)


def dump_entries(k, v, c, fn, compressed=True, offset=0):
    """The same from (key, val, count) entries in the reference's convention (e.g. the merged export of a
    hash-partitioned table)."""
    fn = fn[:-4] if fn.endswith(".npz") else fn
    params, okeys, ovals, ocnts = oakht_image(k, v, c, offset)
    (np.savez_compressed if compressed else np.savez)(fn, parameters=params, keys=okeys, values=ovals, counts=ocnts)
    return fn + ".npz"


def load(fn, k, device="cuda", mode=_lib.PG_MODE_LITERAL, min_capacity=0, with_offset=False):
    """Read a reference-layout dBG (or rdBG) file into a literal-key GPU table (``mode``:
    PG_MODE_LITERAL, or PG_MODE_LITERAL_RC when more records will be inserted on both strands).
    ``with_offset``: also return parameters[5], the resume offset of a chunk checkpoint."""
    import torch
    from . import engine
    if mode == _lib.PG_MODE_CANONICAL:
        raise ValueError("a table image holds literal keys: load it as PG_MODE_LITERAL or PG_MODE_LITERAL_RC")
    z = np.load(fn)
    counts = z["counts"]
    live = counts > 0
    keys = z["keys"][live].astype(np.uint64)
    vals = z["values"][live].astype(np.uint64)
    cnts = counts[live].astype(np.uint64)
    sent = keys == np.uint64(0xFFFFFFFFFFFFFFFF)          # the short-record sentinel lives outside the GPU table
    short = int(cnts[sent][0]) if sent.any() else 0
    keys, vals, cnts = keys[~sent], vals[~sent], cnts[~sent]
    t = engine.DbgTable(max(1024, 2 * int(keys.size) + 2, int(min_capacity)), k, mode, device=device)
    if keys.size:
        d_k = torch.from_numpy(keys.view(np.int64)).to(device)
        d_v = torch.from_numpy((vals | (cnts << np.uint64(32))).view(np.int64)).to(device)
        check(t.L.pg_table_insert_raw(ctypes.byref(t.c), engine._ptr(d_k), engine._ptr(d_v), int(keys.size), engine._stream()),
              "pg_table_insert_raw")
    t.stats[_lib.PG_STAT_SHORT] = short
    t.stats[_lib.PG_STAT_USED] = int(keys.size)
    if t.overflowed():
        raise _lib.PgError("table overflow while loading %s" % fn)
    if with_offset:
        return int(z["parameters"][5]), t
    return t
