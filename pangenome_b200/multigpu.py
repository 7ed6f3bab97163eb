"""Hash-partitioned dBG build across the GPUs of one box (one process per GPU,
``torch.distributed``; NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no distributed path; this is the split SURVEY.md 8(e)
prescribes.  Per rank:

  1. K1 + K2a on the rank's own records: every position emits one 16-byte
     update record into bucket (owner, sub) - owner = low bits of mix64(key)
     (which rank's table holds the key), sub = top bits (which L2-sized region
     of that table);
  2. exchange: a W x S count matrix, then the buckets themselves - one
     all-to-all of equal-sized (padded) blocks, the only collective on the data
     path;
  3. K3 on what arrived, region by region (all sources of sub 0, then sub 1, ..).

Every key lives on exactly one rank, so tables never need merging; the merged
export is the concatenation of the ranks' exports (the short-record sentinel is
summed).  Count saturation happens at export, after all occurrences of a key -
from every rank - were added on its owner (SURVEY 8e "result invariance").
"""
import ctypes
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def log2_exact(n):
    b = n.bit_length() - 1
    if n < 1 or (1 << b) != n:
        raise ValueError("world size must be a power of two (1, 2, 4, 8), got %d" % n)
    return b


def exchange_blocks(send, world):
    """send: tensor [world, block] (row d goes to rank d) -> recv [world, block]
    (row s came from rank s).  NCCL: all_to_all_single; gloo (CPU tests) has no
    all-to-all, so it is emulated with all_gather."""
    if world == 1:
        return send.clone()
    recv = torch.empty_like(send)
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv.view(-1), send.view(-1))
    else:
        rank = dist.get_rank()
        gathered = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(gathered, send)
        for s in range(world):
            recv[s] = gathered[s][rank]
    return recv


def segment_plan(recv_counts, part_cap):
    """recv_counts: int64 [world, n_sub] (records rank s sent me for region b).
    Returns (seg_off, seg_cnt) in region-major order so K3 sweeps one table
    region at a time: segment (b, s) starts at record (s * n_sub + b) * part_cap."""
    world, n_sub = recv_counts.shape
    src = torch.arange(world, dtype=torch.int64, device=recv_counts.device).view(1, world)
    sub = torch.arange(n_sub, dtype=torch.int64, device=recv_counts.device).view(n_sub, 1)
    seg_off = ((src * n_sub + sub) * part_cap).reshape(-1).contiguous()
    seg_cnt = recv_counts.t().reshape(-1).contiguous()
    return seg_off, seg_cnt


class DistributedBuilder:
    """Reusable buffers for the distributed build of same-sized shards.

    The rank's positions are cut into ``chunks`` ranges: K2a(c) runs on the compute stream while
    the all-to-all of chunk c-1 runs on a communication stream; one K3 launch then sweeps the
    table region by region over everything that arrived (world x chunks segments per region)."""

    def __init__(self, k, n_positions_local, world, rank, device="cuda", sub_bytes=8 << 20, chunks=4):
        from . import engine
        self.engine = engine
        self.L = _lib.load()
        self.k, self.world, self.rank = int(min(max(1, k), 27)), world, rank
        self.owner_bits = log2_exact(world)
        self.chunks = max(1, int(chunks)) if world > 1 else 1
        # every rank sizes for the global worst case: all positions distinct, spread evenly
        n_glob = torch.tensor([n_positions_local], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(n_glob, op=dist.ReduceOp.SUM)
        self.n_positions_global = int(n_glob.item())
        per_rank = (self.n_positions_global + world - 1) // world
        cap = engine.next_pow2(max(1024, int(per_rank / 0.5) + 1))
        self.table = engine.DbgTable(cap, self.k, _lib.PG_MODE_CANONICAL, device=device)
        self.sub_bits = min(engine.sub_bits_for(cap, sub_bytes), 10 - self.owner_bits)      # K2a handles <= 1024 buckets
        self.n_sub = n_sub = 1 << self.sub_bits
        n_parts = world * n_sub
        # the largest shard decides the block size every rank uses (blocks must be equal-sized)
        m = torch.tensor([n_positions_local], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        per_chunk = (int(m.item()) + self.chunks - 1) // self.chunks + 4096
        self.part_cap = pc = int(per_chunk / n_parts * 1.25) + 2048
        C = self.chunks
        self.send = torch.empty(C, n_parts * pc * 2, dtype=torch.int64, device=device)      # [chunk][owner][sub][part_cap] records
        self.send_counts = torch.zeros(C, n_parts, dtype=torch.int64, device=device)
        self.recv = torch.empty_like(self.send) if world > 1 else self.send                  # [chunk][source][sub][part_cap]
        self.recv_counts = torch.zeros_like(self.send_counts) if world > 1 else self.send_counts
        # static segment offsets, region-major: segment (b, c, s) = records of source s, chunk c, for table region b
        c_i = torch.arange(C, dtype=torch.int64, device=device).view(1, C, 1)
        s_i = torch.arange(world, dtype=torch.int64, device=device).view(1, 1, world)
        b_i = torch.arange(n_sub, dtype=torch.int64, device=device).view(n_sub, 1, 1)
        self.seg_off = (((c_i * world + s_i) * n_sub + b_i) * pc).reshape(-1).contiguous()
        self.seg_cnt = torch.zeros(n_sub * C * world, dtype=torch.int64, device=device)
        self.comm = torch.cuda.Stream(device=device)
        self.launches_per_build = 2 + self.chunks      # count_short, chunks x K2a, K3 (the table reset is a 64-byte memset)

    def build(self, packed, n_rec, ev=None):
        eng, L, t = self.engine, self.L, self.table
        C, W, n_sub = self.chunks, self.world, self.n_sub
        st = torch.cuda.current_stream()
        self.comm.wait_stream(st)
        t.clear()                            # epoch bump: nothing to overlap
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if ev is not None else None
        if e:
            e[0].record(st)
        g_begin = int(packed.seq_off[0]) if n_rec > 0 else 0
        g_end = int(packed.seq_off[n_rec]) if n_rec > 0 else 0
        span = g_end - g_begin
        step = ((span + C - 1) // C + 2047) // 2048 * 2048 if span > 0 else 0
        desc = _lib.PgTable(None, 2, None, _lib.PG_MODE_CANONICAL, self.k, 1, 0)
        for c in range(C):
            lo = min(g_begin + c * step, g_end)
            hi = min(lo + step, g_end)
            # NB every rank runs every chunk (empty ranges produce empty buckets): the collectives must match
            eng.check(L.pg_kmer_partition(ctypes.byref(desc), eng._ptr(packed.pk2), eng._ptr(packed.amb),
                                          eng._ptr(packed.d_seq_off), n_rec, lo, hi, self.owner_bits, self.sub_bits,
                                          eng._ptr(self.send[c]), self.part_cap, eng._ptr(self.send_counts[c]), None, 0, None,
                                          eng._stream()),
                      "pg_kmer_partition")
            if W > 1:
                done = torch.cuda.Event()
                done.record(st)
                self.comm.wait_event(done)
                with torch.cuda.stream(self.comm):       # exchange of chunk c overlaps K2a of chunk c+1
                    if dist.get_backend() == "nccl":
                        dist.all_to_all_single(self.recv_counts[c], self.send_counts[c])
                        dist.all_to_all_single(self.recv[c], self.send[c])
                    else:
                        self.recv_counts[c].copy_(exchange_blocks(self.send_counts[c].view(W, n_sub), W).view(-1))
                        self.recv[c].copy_(exchange_blocks(self.send[c].view(W, -1), W).view(-1))
        if e:
            e_part = torch.cuda.Event(enable_timing=True); e_part.record(st)
            e_comm = torch.cuda.Event(enable_timing=True); e_comm.record(self.comm)
        st.wait_stream(self.comm)
        if e:
            e_k3 = torch.cuda.Event(enable_timing=True); e_k3.record(st)
        if n_rec > 0:
            eng.check(L.pg_count_short(ctypes.byref(t.c), eng._ptr(packed.d_seq_off), n_rec, g_begin, g_end, eng._stream()),
                      "pg_count_short")
        # counts [chunk][source][sub] -> region-major [sub][chunk][source]
        self.seg_cnt.copy_(self.recv_counts.view(C, W, n_sub).permute(2, 0, 1).reshape(-1))
        eng.check(L.pg_insert_records(ctypes.byref(t.c), eng._ptr(self.recv), eng._ptr(self.seg_off), eng._ptr(self.seg_cnt),
                                      n_sub, C * W, self.part_cap, eng._stream()), "pg_insert_records")
        if e:
            e[1].record(st)
            ev.setdefault("build", []).append((e[0], e[1]))
            ev.setdefault("k2a_all_chunks", []).append((e[0], e_part))
            ev.setdefault("until_exchange_done", []).append((e[0], e_comm))
            ev.setdefault("k3", []).append((e_k3, e[1]))
        return t

    def verify(self):
        if int(self.send_counts.max().item()) > self.part_cap:
            raise _lib.PgError("record bucket overflow or truncated record index on rank %d" % self.rank)
        if self.table.overflowed():
            raise _lib.PgError("dBG table overflow on rank %d" % self.rank)


class PeerBuilder:
    """Distributed build with the exchange FUSED into K2a: every rank owns a receive buffer
    [source rank][sub][part_cap] that its peers map through CUDA IPC; K2a's write-out stores each
    bucket straight into its owner's buffer over NVLink (pg_kmer_partition_p2p), so there is no
    staging copy and no separate all-to-all of the records.  The only collective left is the
    all-to-all of the small count matrix, which doubles as the barrier that orders the peer stores
    before K3.  Two receive buffers alternate between builds: a rank that races ahead into the
    next build cannot overwrite what a slower peer's K3 is still reading."""

    def __init__(self, k, n_positions_local, world, rank, device="cuda", sub_bytes=8 << 20):
        from . import engine
        self.engine = engine
        self.L = L = _lib.load()
        self.k, self.world, self.rank = int(min(max(1, k), 27)), world, rank
        self.owner_bits = log2_exact(world)
        n_glob = torch.tensor([n_positions_local], dtype=torch.int64, device=device)
        m = torch.tensor([n_positions_local], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(n_glob, op=dist.ReduceOp.SUM)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        per_rank = (int(n_glob.item()) + world - 1) // world
        cap = engine.next_pow2(max(1024, int(per_rank / 0.5) + 1))
        self.table = engine.DbgTable(cap, self.k, _lib.PG_MODE_CANONICAL, device=device)
        # owner x sub buckets are written in runs over NVLink: keep their number <= 256 so a 4096-position
        # tile still gives ~16-record (256-byte) runs; fewer table regions per rank is the price
        self.sub_bits = min(engine.sub_bits_for(cap, sub_bytes), max(0, int(os.environ.get("PG_MG_PARTBITS", "8")) - self.owner_bits))
        self.n_sub = n_sub = 1 << self.sub_bits
        n_parts = world * n_sub
        self.part_cap = pc = int(int(m.item()) / n_parts * 1.25) + 4096
        self.buf_bytes = world * n_sub * pc * 16
        self.own, self.peer_tables, self._opened = [], [], []
        for _ in range(2):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.check(L.pg_peer_alloc(self.buf_bytes, ctypes.byref(ptr), handle), "pg_peer_alloc")
            handles = [None] * world
            if world > 1:
                dist.all_gather_object(handles, handle.raw)
            addrs = []
            for r in range(world):
                if r == rank:
                    addrs.append(ptr.value)
                else:
                    q = ctypes.c_void_p()
                    _lib.check(L.pg_peer_open(handles[r], ctypes.byref(q)), "pg_peer_open")
                    self._opened.append(q)
                    addrs.append(q.value)
            self.own.append(ptr)
            self.peer_tables.append(torch.tensor(addrs, dtype=torch.int64, device=device))
        self.send_counts = torch.zeros(n_parts, dtype=torch.int64, device=device)
        self.recv_counts = torch.zeros(n_parts, dtype=torch.int64, device=device)
        s_i = torch.arange(world, dtype=torch.int64, device=device).view(1, world)
        b_i = torch.arange(n_sub, dtype=torch.int64, device=device).view(n_sub, 1)
        self.seg_off = ((s_i * n_sub + b_i) * pc).reshape(-1).contiguous()        # region-major over sources
        self.seg_cnt = torch.zeros(n_sub * world, dtype=torch.int64, device=device)
        self.parity = 0
        self.chunks = 1
        self.launches_per_build = 3          # count_short, k2a (fused exchange), k3 (the table reset is a 64-byte memset)
        if world > 1:
            dist.barrier()

    def begin(self):
        """Empty the table for the next build (epoch bump, DbgTable.clear)."""
        self.table.clear()
        self._begun = True

    def build(self, packed, n_rec, ev=None):
        eng, L, t = self.engine, self.L, self.table
        W, n_sub = self.world, self.n_sub
        st = torch.cuda.current_stream()
        if not getattr(self, "_begun", False):
            self.begin()
        self._begun = False
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if ev is not None else None
        if e:
            e[0].record(st)
        g_begin = int(packed.seq_off[0]) if n_rec > 0 else 0
        g_end = int(packed.seq_off[n_rec]) if n_rec > 0 else 0
        buf = self.parity
        self.parity ^= 1
        desc = _lib.PgTable(None, 2, None, _lib.PG_MODE_CANONICAL, self.k, 1, 0)
        eng.check(L.pg_kmer_partition_p2p(ctypes.byref(desc), eng._ptr(packed.pk2), eng._ptr(packed.amb),
                                          eng._ptr(packed.d_seq_off), n_rec, g_begin, g_end, self.owner_bits, self.sub_bits,
                                          eng._ptr(self.peer_tables[buf]), self.rank, self.part_cap,
                                          eng._ptr(self.send_counts), eng._stream()), "pg_kmer_partition_p2p")
        if e:
            e[1].record(st)
        if W > 1:       # counts row d -> rank d; completes only after every peer's K2a (= all stores into my buffer) finished
            dist.all_to_all_single(self.recv_counts, self.send_counts)
        else:
            self.recv_counts.copy_(self.send_counts)
        if e:
            e[2].record(st)
        if n_rec > 0:
            eng.check(L.pg_count_short(ctypes.byref(t.c), eng._ptr(packed.d_seq_off), n_rec, g_begin, g_end, eng._stream()),
                      "pg_count_short")
        self.seg_cnt.copy_(self.recv_counts.view(W, n_sub).t().reshape(-1))
        eng.check(L.pg_insert_records(ctypes.byref(t.c), self.own[buf], eng._ptr(self.seg_off), eng._ptr(self.seg_cnt),
                                      n_sub, W, self.part_cap, eng._stream()), "pg_insert_records")
        if e:
            e[3].record(st)
            ev.setdefault("build", []).append((e[0], e[3]))
            ev.setdefault("k2a_all_chunks", []).append((e[0], e[1]))
            ev.setdefault("until_exchange_done", []).append((e[0], e[2]))
            ev.setdefault("k3", []).append((e[2], e[3]))
        return t

    def build_async(self, packed, ev=None):
        """build() over ALL records of ``packed`` without reading K1's record index back: K2a and the
        short-record count take their bounds from the device (PackedSeqs(lazy=True)), so a step is
        K1 -> K2a+exchange -> counts all-to-all -> K3 enqueued back to back.  verify() reports a
        truncated record index or an overflow afterwards."""
        eng, L, t = self.engine, self.L, self.table
        W, n_sub = self.world, self.n_sub
        st = torch.cuda.current_stream()
        if not getattr(self, "_begun", False):
            self.begin()
        self._begun = False
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if ev is not None else None
        if e:
            e[0].record(st)
        buf = self.parity
        self.parity ^= 1
        desc = _lib.PgTable(None, 2, None, _lib.PG_MODE_CANONICAL, self.k, 1, 0)
        eng.check(L.pg_kmer_partition_p2p_dev(ctypes.byref(desc), eng._ptr(packed.pk2), eng._ptr(packed.amb),
                                              eng._ptr(packed.d_seq_off), eng._ptr(packed.d_counts), packed.cap_records, packed.nbytes,
                                              self.owner_bits, self.sub_bits, eng._ptr(self.peer_tables[buf]), self.rank,
                                              self.part_cap, eng._ptr(self.send_counts), eng._stream()), "pg_kmer_partition_p2p_dev")
        if e:
            e[1].record(st)
        if W > 1:
            dist.all_to_all_single(self.recv_counts, self.send_counts)
        else:
            self.recv_counts.copy_(self.send_counts)
        if e:
            e[2].record(st)
        eng.check(L.pg_count_short_dev(ctypes.byref(t.c), eng._ptr(packed.d_seq_off), eng._ptr(packed.d_counts), packed.cap_records,
                                       eng._stream()), "pg_count_short_dev")
        self.seg_cnt.copy_(self.recv_counts.view(W, n_sub).t().reshape(-1))
        eng.check(L.pg_insert_records(ctypes.byref(t.c), self.own[buf], eng._ptr(self.seg_off), eng._ptr(self.seg_cnt),
                                      n_sub, W, self.part_cap, eng._stream()), "pg_insert_records")
        if e:
            e[3].record(st)
            ev.setdefault("build", []).append((e[0], e[3]))
            ev.setdefault("k2a_all_chunks", []).append((e[0], e[1]))
            ev.setdefault("until_exchange_done", []).append((e[0], e[2]))
            ev.setdefault("k3", []).append((e[2], e[3]))
        return t

    def verify(self):
        if int(self.send_counts.max().item()) > self.part_cap:
            raise _lib.PgError("record bucket overflow or truncated record index on rank %d" % self.rank)
        if self.table.overflowed():
            raise _lib.PgError("dBG table overflow on rank %d" % self.rank)

    def close(self):
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        for q in self._opened:
            self.L.pg_peer_close(q)
        for p_ in self.own:
            self.L.pg_peer_free(p_)
        self._opened, self.own = [], []


def gather_export(table, world, rank):
    """Merged (keys, vals, cnts) of the distributed dBG on rank 0 (test / small outputs only)."""
    k, v, c = table.export(sort=False)
    short = int(table.stats_host()[_lib.PG_STAT_SHORT])
    sent = np.uint64(0xFFFFFFFFFFFFFFFF)
    keep = k != sent                               # the sentinel is merged separately
    k, v, c = k[keep], v[keep], c[keep]
    if world == 1:
        parts, shorts = [(k, v, c)], [short]
    else:
        parts, shorts = [None] * world, [None] * world
        dist.all_gather_object(parts, (k, v, c))
        dist.all_gather_object(shorts, short)
    if rank != 0:
        return None
    K = np.concatenate([p[0] for p in parts])
    V = np.concatenate([p[1] for p in parts])
    C = np.concatenate([p[2] for p in parts])
    tot_short = sum(shorts)
    if tot_short > 0:
        K = np.append(K, sent)
        V = np.append(V, np.uint16(32))
        C = np.append(C, np.uint8(min(tot_short, 255)))
    o = np.argsort(K, kind="stable")
    return K[o], V[o], C[o]


def gather_varlen(t, world):
    """Concatenation over ranks of 1-D tensors of different lengths (padded all_gather).
    Returns (concatenated tensor, per-rank lengths)."""
    if world == 1:
        return t, [int(t.numel())]
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:sz] for p, sz in zip(parts, sizes)]), sizes


class _GlobalIndex:
    """What graph.RdbgGraph.regions / GraphResult.rows need from a PackedSeqs, for the records of all ranks."""

    def __init__(self, seq_off, hdr_ids, device):
        self.seq_off = seq_off
        self.d_seq_off = torch.from_numpy(seq_off).to(device)
        self.ids = hdr_ids

    @property
    def seq_lengths(self):
        return np.diff(self.seq_off)


def seq2graph_distributed(packed, table, k, world, rank, data, rc=False, min_weight=1):
    """Stages 2-5 after a distributed dBG build: every rank selects the rdBG members
    among ITS keys, the small rdBG tables are all-gathered into one full table per rank, each rank
    walks its own records (K5), the hits (a few % of the positions) are gathered in rank-independent
    form (record, position, literal code, v5, v6) and rank 0 runs K6-K8 on them.
    Returns (GraphResult, rows) on rank 0 and (None, None) elsewhere."""
    from . import engine, graph
    L = _lib.load()
    dev = table.slots.device
    chk, P, S = engine.check, engine._ptr, engine._stream
    # ---- rdBG: local select, all-gather of the raw slots, full table on every rank
    rd_local = table.select_rdbg()
    n_loc = rd_local.n_slots_used
    keys = torch.empty(max(n_loc, 1), dtype=torch.int64, device=dev)
    vals = torch.empty(max(n_loc, 1), dtype=torch.int64, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    chk(L.pg_table_export_raw(ctypes.byref(rd_local.c), P(keys), P(vals), n_loc, P(d_n), S()), "pg_table_export_raw")
    assert int(d_n.item()) == n_loc
    K_all, _ = gather_varlen(keys[:n_loc], world)
    V_all, _ = gather_varlen(vals[:n_loc], world)
    rd = engine.DbgTable(max(1024, 2 * int(K_all.numel()) + 2), table.k, table.mode, device=dev)
    chk(L.pg_table_insert_raw(ctypes.byref(rd.c), P(K_all), P(V_all), int(K_all.numel()), S()), "pg_table_insert_raw")
    short = table.stats[_lib.PG_STAT_SHORT:_lib.PG_STAT_SHORT + 1].clone()
    if world > 1:
        dist.all_reduce(short, op=dist.ReduceOp.SUM)
    rd.stats[_lib.PG_STAT_SHORT] = short[0]
    members = torch.tensor([rd_local.n_members - (1 if int(table.stats_host()[_lib.PG_STAT_SHORT]) > 0 else 0)],
                           dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(members, op=dist.ReduceOp.SUM)
    rd.n_members = int(members.item()) + (1 if int(short.item()) > 0 else 0)
    if rd.overflowed():
        raise _lib.PgError("rdBG table overflow")
    # ---- record index of all ranks (host objects: small)
    lens_all, ids_all = [None] * world, [None] * world
    my_ids = graph.record_ids(packed, data)
    if world > 1:
        dist.all_gather_object(lens_all, packed.seq_lengths.tolist())
        dist.all_gather_object(ids_all, my_ids)
    else:
        lens_all, ids_all = [packed.seq_lengths.tolist()], [my_ids]
    rec_base = sum(len(x) for x in lens_all[:rank])
    g_seq_off = np.concatenate([[0], np.cumsum(np.concatenate([np.asarray(x, dtype=np.int64) for x in lens_all]))]).astype(np.int64)
    gidx = _GlobalIndex(g_seq_off, [i for ids in ids_all for i in ids], dev)
    # ---- K5 on the local records, hits to rank-independent form, gather
    n_strands = 2 if rc else 1
    all_hits = []
    for strand in range(n_strands):
        h = graph.path_hits(packed, rd, packed.n_rec, strand)
        n = h.n
        code = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        chk(L.pg_hits_decode(ctypes.byref(rd.c), P(h.node), n, P(code), S()), "pg_hits_decode")
        rec = h.rec[:n].long()
        pos = h.g[:n] - packed.d_seq_off[:packed.n_rec + 1][rec] if n else h.g[:0]
        v5 = (h.node[:n] & 2047).to(torch.int32)
        g_code, _ = gather_varlen(code[:n], world)
        g_pos, _ = gather_varlen(pos, world)
        g_rec, _ = gather_varlen(rec + rec_base, world)
        g_v5, _ = gather_varlen(v5, world)
        g_v6, _ = gather_varlen(h.v6[:n].to(torch.int32), world)      # NCCL has no int16
        all_hits.append((g_code, g_pos, g_rec, g_v5, g_v6))
    if rank != 0:
        return None, None
    # ---- rank 0: re-key against its own rdBG table and run K6-K8 exactly like the single-GPU path
    hits = []
    for strand, (g_code, g_pos, g_rec, g_v5, g_v6) in enumerate(all_hits):
        n = int(g_code.numel())
        node = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        chk(L.pg_hits_rekey(ctypes.byref(rd.c), P(g_code), P(g_v5.contiguous()), n, P(node), S()), "pg_hits_rekey")
        gg = gidx.d_seq_off[g_rec] + g_pos if n else g_pos
        hits.append(graph.Hits(gg.contiguous(), node, g_rec.to(torch.int32).contiguous(), g_v6.to(torch.int16).contiguous(), n, strand))
    total = sum(h.n for h in hits)
    g = graph.RdbgGraph(total, dev)
    for h in hits:
        g.add_hits(h, n_strands)
    res = graph.GraphResult()
    res.edges = g.edges(rd)
    res.nodes = g.components(rd, min_weight)
    for h in hits:
        rec, start, end, lab = g.regions(h, gidx, table.k)
        res.rows_raw.append((rec, start, end, 1 if h.strand == 0 else -1, lab))
    res.graph = g
    res.rdbg = rd
    rows = res.rows(gidx, None)
    return res, rows


def bench(args, world, rank, local, ClockSampler=None):
    """bench.py --gpus N (N > 1): weak scaling, every rank builds from its own 10 x 5 Mbp shard
    (same ancestor, rank-specific genomes), one all-to-all per step."""
    import time
    from . import engine, synth
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    k = args.k
    genomes, length = (10, 5_000_000) if args.workload == "cfg2" else (4, 1_000_000)
    anc = np.random.default_rng(1).integers(0, 4, length, dtype=np.uint8)
    recs = []
    for g in range(genomes):
        gid = rank * genomes + g
        recs.append((b"g%d synthetic" % gid, synth._ACGT[synth._snp_copy(np.random.default_rng(100 + gid), anc, 0.01)]))
    data = synth.fasta_bytes(recs)
    host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    d_fasta = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()
    packed = engine.PackedSeqs(d_fasta)
    n_rec = packed.n_rec
    n_ins_local = packed.n_insertions(k)
    mode = os.environ.get("PG_EXCHANGE", "p2p")
    builder = PeerBuilder(k, packed.n_positions(k), world, rank) if mode == "p2p" else \
        DistributedBuilder(k, packed.n_positions(k), world, rank)
    stream = torch.cuda.current_stream()
    kev = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(record=False):
        if hasattr(builder, "begin"):
            builder.begin()                 # empty the table (epoch bump)
        if hasattr(builder, "build_async"):   # no host read-back inside a step: bounds stay on the device
            return builder.build_async(engine.PackedSeqs(d_fasta, lazy=True), ev=kev if record else None)
        p = engine.PackedSeqs(d_fasta)
        return builder.build(p, n_rec, ev=kev if record else None)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    builder.verify()
    clocks = ClockSampler(local) if (ClockSampler and rank == 0) else None
    if clocks:
        clocks.start()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(args.steps):
        t = step(record=True)
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop() if clocks else None
    builder.verify()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([n_ins_local], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    used, entries = t.count()
    ent = torch.tensor([used], dtype=torch.int64, device="cuda")
    dist.all_reduce(ent, op=dist.ReduceOp.SUM)
    avg = lambda name: sum(a.elapsed_time(b) for a, b in kev[name]) / len(kev[name])
    stage = torch.tensor([avg("build"), avg("k2a_all_chunks"), avg("until_exchange_done"), avg("k3")], dtype=torch.float64, device="cuda")
    dist.all_reduce(stage, op=dist.ReduceOp.MAX)

    # end to end: every step uploads its own copy of the shard from pinned host memory (double-buffered on a
    # copy stream: the upload of step i+1 is issued before step i's result is awaited), builds, and reads the
    # table statistics back
    copy_stream = torch.cuda.Stream()
    dev_in = [torch.empty_like(d_fasta) for _ in range(2)]

    k1_done = [None, None]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            if k1_done[i % 2] is not None:
                copy_stream.wait_event(k1_done[i % 2])       # K1 of the step that last read this buffer
            dev_in[i % 2].copy_(host, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        return done

    def run_e2e(n):
        # as in bench.py: input i+1 uploads and result i-1 (table statistics) is read while step i runs
        nxt = upload(0)
        pending = None
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = upload(i + 1)
            if hasattr(builder, "begin"):
                builder.begin()
            stream.wait_event(cur)
            if hasattr(builder, "build_async"):
                p = engine.PackedSeqs(dev_in[i % 2], lazy=True)
                k1_done[i % 2] = torch.cuda.Event()
                k1_done[i % 2].record(stream)
                tt = builder.build_async(p)
            else:
                p = engine.PackedSeqs(dev_in[i % 2])
                k1_done[i % 2] = torch.cuda.Event()
                k1_done[i % 2].record(stream)
                tt = builder.build(p, n_rec)
            fut = tt.stats_async(i)
            if pending is not None:
                pending.wait()
            pending = fut
        pending.wait()
    run_e2e(2)
    dist.barrier()
    torch.cuda.synchronize()
    g0, g1 = ev(), ev()
    g0.record(stream)
    run_e2e(args.steps)
    g1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ems = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = float(ms.item()) / args.steps
        n_ins = int(tot.item())
        e2e_ms = float(ems.item()) / args.steps
        peak = 6552.3
        pp = os.path.join(sys_path_root, "MEASURED_PEAKS.json")
        if os.path.isfile(pp):
            peak = float(json.load(open(pp))["hbm_gbs"])
        ins_ms = float(stage[0].item())
        alg = 16.0 * n_ins / world
        block_bytes = builder.n_sub * builder.part_cap * 16 * builder.chunks
        sent = torch.zeros(1)
        line = {
            "metric": "dbg_build_kmers_per_s", "value": n_ins / (ms_step * 1e-3) / 1e9, "unit": "G k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "%d x (%d x %d bp genomes, 1%% SNP, shared ancestor), hash-partitioned table, NCCL all-to-all"
                       % (world, genomes, length), "k": k, "rc": True, "insertions_per_step": n_ins,
                       "table_slots_per_gpu": builder.table.capacity, "distinct_canonical_keys": int(ent.item()),
                       "l2": "every step clears and updates a table larger than L2 on every rank"},
            "e2e": {"value": n_ins / (e2e_ms * 1e-3) / 1e9, "unit": "G k-mers/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(host.numel()) * world, "d2h_bytes_per_step": 8 * _lib.PG_STAT_WORDS * world},
            "gpu_launches": (3 + builder.launches_per_build) * args.steps * world,
            "clocks": clk,
            "exchange": "fused into K2a: peer stores over NVLink (CUDA IPC)" if mode == "p2p" else "NCCL all_to_all_single, %d chunks" % builder.chunks,
            "stages_ms": {"build": ins_ms,
                          "k2a_all_chunks": float(stage[1].item()), "until_exchange_done": float(stage[2].item()),
                          "k3": float(stage[3].item())},
            "exchange_bytes_per_gpu_per_step": block_bytes * (world - 1),
            "roofline": {"kernel": "k2a_partition(+exchange) + k3_insert_records", "bound": "hbm",
                         "achieved": alg / (ins_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ins_ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "convention": "per GPU: 16 B per insertion owned by the rank (SURVEY 8d) over the whole pipelined build"},
        }
        print(json.dumps(line), flush=True)
    if hasattr(builder, "close"):
        builder.close()
    dist.barrier()
    dist.destroy_process_group()
