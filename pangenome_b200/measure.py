"""Measurement helpers shared by bench.py, the tests and profiles/: merged table checksums across ranks, the random-slot
micro-benchmark ceiling for K3, and CUDA-event timings of the stages after the dBG build (K1, K4..K8)
with their algorithmic bytes.  Nothing here is on the hot path."""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine

_M = (1 << 64) - 1


def mix64(x):
    x &= _M
    x ^= x >> 33
    x = (x * 0xff51afd7ed558ccd) & _M
    x ^= x >> 33
    x = (x * 0xc4ceb9fe1a85ec53) & _M
    x ^= x >> 33
    return x


def entry_mix(key, val, cnt):
    """Host twin of pg_entry_mix (csrc/common.cuh): the per-entry term of the order-independent table checksum."""
    w = ((val << 8) | cnt) & _M
    return mix64(key ^ ((w * 0x9E3779B97F4A7C15) & _M))


def merged_checksum(table, world=1):
    """(entries, sum, xor) of the whole hash-partitioned dBG, identical on every rank: the per-rank checksums without
    their short-record sentinel, combined, plus ONE sentinel entry carrying the summed (then clamped) count - count
    saturation is applied after the cross-GPU merge (SURVEY 8e "result invariance")."""
    n, s, x = table.checksum()
    short = int(table.stats_host()[_lib.PG_STAT_SHORT])
    if short > 0:
        h = entry_mix(_M, 32, min(short, 255))
        n, s, x = n - 1, (s - h) & _M, x ^ h
    if world > 1:
        mine = torch.tensor([n, s - (1 << 64) if s >= (1 << 63) else s, x - (1 << 64) if x >= (1 << 63) else x, short],
                            dtype=torch.int64, device=table.slots.device)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        n = s = x = short = 0
        for p in parts:
            v = [int(q) & _M for q in p.cpu().tolist()]
            n, s, x, short = n + v[0], (s + v[1]) & _M, x ^ v[2], short + v[3]
    if short > 0:
        h = entry_mix(_M, 32, min(short, 255))
        n, s, x = n + 1, (s + h) & _M, x ^ h
    return n, s, x


def slot_ceiling(capacity, region_slots, n_ops, mode, ctas_per_sm=5, ilp=1, reps=3, device="cuda"):
    """G operations/s of pg_microbench_slots (best of ``reps``)."""
    L = _lib.load()
    slots = torch.zeros(2 * capacity, dtype=torch.int64, device=device)
    recs = torch.zeros(2 * n_ops, dtype=torch.int64, device=device) if (mode & 8) else None
    sink = torch.zeros(1, dtype=torch.int64, device=device)
    best = None
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.check(L.pg_microbench_slots(engine._ptr(slots), capacity, region_slots, n_ops, mode, ctas_per_sm, ilp,
                                           engine._ptr(recs) if recs is not None else None, engine._ptr(sink), engine._stream()),
                     "pg_microbench_slots")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return n_ops / (best * 1e-3) / 1e9, best


def _timed(fn, reps=3):
    """best-of CUDA-event time (ms) of fn() on the current stream; returns (ms, last result)."""
    best, out = None, None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return best, out


def stage_rooflines(packed, table, k, peak_gbs, reps=3):
    """CUDA-event time of every kernel group the north star names outside the dBG insert - K1 pack, K4 rdBG selection,
    K5 path hits, K6 edge weights, K7 components, K8 breakpoint labelling - on the given input, with the algorithmic
    bytes of SURVEY.md 8(d) and the fraction of the measured HBM peak they amount to."""
    from . import graph
    L = _lib.load()
    P, S, byref = engine._ptr, engine._stream, ctypes.byref
    out = {}
    n_bases, n_rec = int(packed.n_bases), int(packed.n_rec)
    n_pos = packed.n_positions(k)

    def entry(ms, nbytes, convention, **extra):
        gbs = nbytes / (ms * 1e-3) / 1e9
        d = {"ms_per_launch": ms, "algorithmic_bytes_per_launch": float(nbytes), "achieved": gbs, "frac": gbs / peak_gbs, "convention": convention}
        d.update(extra)
        return d

    # K1: 1 B/byte read + 0.375 B/base written (2-bit digits + 1-bit ambiguity plane)
    ms, _ = _timed(lambda: packed._launch(packed.cap_records), reps)
    out["k1_fasta_scan_pack"] = entry(ms, packed.nbytes + 0.375 * n_bases, "1 B/byte read + 0.375 B/base written (SURVEY 8d: 1.375 B/base)")
    # K4: 16 B/slot scanned + 16 B per kept slot written; count pass and select pass timed together
    cnt = torch.zeros(2, dtype=torch.int64, device=table.slots.device)
    ms_c, _ = _timed(lambda: engine.check(L.pg_rdbg_count(byref(table.c), P(cnt), S()), "pg_rdbg_count"), reps)
    n_slots, n_members = [int(v) for v in cnt.cpu().tolist()]
    rd = engine.DbgTable(max(1024, int(n_slots / 0.5) + 1), table.k, table.mode, device=table.slots.device)

    def sel():
        rd.clear()
        engine.check(L.pg_rdbg_select(byref(table.c), byref(rd.c), S()), "pg_rdbg_select")
    ms_s, _ = _timed(sel, reps)
    rd.n_members, rd.n_slots_used = n_members, n_slots
    out["k4_rdbg_count"] = entry(ms_c, 16.0 * table.capacity, "16 B/slot scanned", slots=table.capacity)
    out["k4_rdbg_select"] = entry(ms_s, 16.0 * table.capacity + 16.0 * n_slots, "16 B/slot scanned + 16 B/kept slot (SURVEY 8d)", kept_slots=n_slots)
    # K5: packed stream + one rdBG probe per position (L2-resident table) + 22 B per hit written
    ms, hits = _timed(lambda: graph.path_hits(packed, rd, n_rec, 0), reps)
    out["k5_path_hits"] = entry(ms, 0.375 * n_bases + 16.0 * n_pos + 22.0 * hits.n,
                                "2 passes timed together incl. one 8-byte D2H: 0.375 B/base + 16 B/position probe + 22 B/hit", hits=hits.n,
                                positions_per_s=n_pos / (ms * 1e-3))
    # K6: node/edge/visit tables, all random single-word CAS: 8 B key + 3 x 8 B CAS + 4 B slot per hit
    g = graph.RdbgGraph(hits.n, packed.pk2.device)

    def k6():
        engine.check(L.pg_graph_clear(byref(g.c), S()), "pg_graph_clear")
        g.add_hits(hits, 1)
    ms, _ = _timed(k6, reps)
    out["k6_edge_weights"] = entry(ms, 44.0 * hits.n + 44.0 * g.cap, "graph clear (44 B/slot) + 44 B/hit of single-word CAS traffic", hits=hits.n)
    # K7: union-find over the edge slots
    ms, _ = _timed(lambda: engine.check(L.pg_graph_components(byref(g.c), 1, S()), "pg_graph_components"), 1)
    nn, ne = g.counts()
    out["k7_components"] = entry(ms, 12.0 * g.cap + 16.0 * ne + 12.0 * g.cap, "edge-slot scan 12 B + 16 B/edge of parent traffic + flatten 12 B/slot", edges=ne)
    g.components(rd, 1, to_host=False)
    # K8: 22 B/hit read + one node-table probe + 12 B/row written
    ms, rows = _timed(lambda: g.regions(hits, packed, k), reps)
    out["k8_label_regions"] = entry(ms, 30.0 * hits.n + 12.0 * len(rows[0]), "22 B/hit + 8 B node probe + 12 B/row (whole call incl. its D2H of the matched count)",
                                    rows=int(len(rows[0])))
    out["breakpoint_labelling_k5_k8"] = entry(out["k5_path_hits"]["ms_per_launch"] + ms, 0.25 * n_bases + 16.0 * hits.n + 12.0 * len(rows[0]),
                                              "SURVEY 8d: 0.25 B/base read + 16 B per hit lookup + 12 B/row written, K5 + K8 time")
    return out
