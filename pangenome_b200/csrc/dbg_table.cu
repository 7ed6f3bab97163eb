// dbg_table.cu - K2+K3 (k-mer extraction fused with hash-table insertion), table read-out and
// K4 (reduced-dBG selection).
//
// Table = open addressing, linear probing, power-of-two capacity, 16-byte slots
//   { u64 key | u32 masks | count:22 tag:10 }
// so that one probe touches one 32-byte sector.  A slot is live while its tag equals the table's epoch
// (common.cuh), so a rebuild starts with epoch += 1 instead of rewriting the table.  A free slot is claimed
// with ONE 128-bit CAS (key + first masks + first count + tag) against the contents just loaded; later
// occurrences merge masks with red.or.b32 and count with red.add.u32 (no return value needed:
// "fire and forget" reductions resolved in L2).  Both are skipped when the 16-byte slot load already
// shows the bits set / the count saturated (>= 255; the reference clamps there, kmer_numba.py:551) -
// repeats therefore cost one sector read and no atomic.
// Replaces oakht.push/has_key/get (:521-603), add_kmer (:1036-1047), build_dbg (:1052-1093),
// seq2dbg_jit_ (:1202-1230), build_rdbg_jit_ (:1292-1309).
#include <stdlib.h>
#include "kmer_core.cuh"
#include "table_dev.cuh"

namespace {

template <int MODE>
__global__ void __launch_bounds__(K2_THREADS)
k2_kmer_insert(TableView t, const uint64_t *__restrict__ pk2, const uint32_t *__restrict__ amb, int64_t n_words,
               const int64_t *__restrict__ seq_off, int64_t n_rec, int64_t g_begin, int64_t g_end, int k,
               uint64_t pow5km1, int64_t w_first, int64_t n_tiles) {
    __shared__ __align__(16) uint64_t s_pk[K2_TILE_WORDS + 4];
    __shared__ __align__(16) uint32_t s_am[K2_TILE_WORDS + 8];
    __shared__ uint32_t s_vlut[16];
    __shared__ uint16_t s_lut5[PG_LUT5_SIZE];
    if (threadIdx.x < 16) s_vlut[threadIdx.x] = pg_vlut_entry(threadIdx.x);      // visible after the tile loop's first barrier
    for (int i = threadIdx.x; i < PG_LUT5_SIZE; i += K2_THREADS) s_lut5[i] = (uint16_t)pg_lut5_entry(i);
    uint32_t n_claimed = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t w0 = w_first + tile * K2_TILE_WORDS;
        __syncthreads();
        stage_tile(pk2, amb, w0, n_words, s_pk, s_am);
        __syncthreads();
        const int64_t g0 = (w0 + threadIdx.x) * 32;
        if (g0 >= g_end || g0 + 32 <= g_begin) continue;
        PgWindow w;
        w.prv = s_pk[threadIdx.x + 1]; w.cur = s_pk[threadIdx.x + 2]; w.nxt = s_pk[threadIdx.x + 3];
        w.aprv = s_am[threadIdx.x + 3]; w.acur = s_am[threadIdx.x + 4]; w.anxt = s_am[threadIdx.x + 5];
        int64_t r = find_record(seq_off, n_rec, g0);
        int64_t rs = r >= 0 ? __ldg(seq_off + r) : 0, re = __ldg(seq_off + r + 1);
        auto upsert_pos = [&](int, uint64_t F, uint64_t R, uint32_t vw) {      // vw = vf | vr << 16
            if (MODE == PG_MODE_CANONICAL) {
                PgUpdate u = pg_canonical_update_w(F, R, vw);
                table_upsert(t, u.key, u.masks, u.inc, n_claimed);
            } else {
                table_upsert(t, F, vw & 0xFFFFu, 1, n_claimed);
                if (MODE == PG_MODE_LITERAL_RC) table_upsert(t, R, vw >> 16, 1, n_claimed);
            }
        };
        if (pg_is_interior(w, g0, 32, k, rs, re, r >= 0, g_begin, g_end)) {
            pg_interior_visit<32>(w, 0, k, pow5km1, s_vlut, s_lut5, upsert_pos);      // fast path: no record edge, no ambiguity
            continue;
        }
        uint64_t F, R;
        pg_codes_init(w, 0, k, F, R);
#pragma unroll 1
        for (int j = 0; j < 32; j++) {
            const int64_t g = g0 + j;
            if (g >= g_end) break;
            while (r + 1 < n_rec && g >= re) { r++; rs = re; re = __ldg(seq_off + r + 1); }
            if (g >= g_begin && r >= 0 && g + k <= re) {
                uint32_t vf, vr;
                pg_occ_vals(w, j, g - rs, re - rs, k, vf, vr);
                upsert_pos(j, F, R, vf | (vr << 16));
            }
            pg_codes_roll(w, j, k, pow5km1, F, R);
        }
    }
    publish_claims(t, n_claimed);
}

// records shorter than k insert the sentinel key 2^64-1 once per strand (Q5, build_dbg :1089-1090)
__global__ void k2_count_short(const int64_t *__restrict__ seq_off, int64_t n_rec, int64_t g_begin, int64_t g_end,
                               int k, int strands, int64_t *stats) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t mine = 0;
    const int64_t g_total = seq_off[n_rec];      // true end of the stream: only the range that reaches it owns offset == g_end
    if (i < n_rec) {
        int64_t a = seq_off[i], b = seq_off[i + 1];
        bool owned = a >= g_begin && (a < g_end || (g_end >= g_total && a <= g_end));
        if (owned && b - a < k) mine = strands;
    }
    mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(reinterpret_cast<unsigned long long *>(stats + PG_STAT_SHORT), (unsigned long long)mine);
}

// device-args variant: all records, n_rec read from K1's count block
__global__ void k2_count_short_dev(const int64_t *__restrict__ seq_off, const int64_t *__restrict__ counts, int64_t cap_records,
                                   int k, int strands, int64_t *stats) {
    int64_t n_rec = counts[0];
    if (n_rec > cap_records) return;
    int64_t mine = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rec; i += (int64_t)gridDim.x * blockDim.x)
        if (seq_off[i + 1] - seq_off[i] < k) mine += strands;
    mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(reinterpret_cast<unsigned long long *>(stats + PG_STAT_SHORT), (unsigned long long)mine);
}

__global__ void k_table_clear(uint4 *slots, int64_t n_slots, int64_t *stats) {
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_slots; i += (int64_t)gridDim.x * blockDim.x)
        slots[i] = e;
    if (blockIdx.x == 0 && threadIdx.x < PG_STAT_WORDS) stats[threadIdx.x] = 0;
}

// ---- read-out ----------------------------------------------------------------------------------
__global__ void k_table_count(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, int mode, int k, uint64_t pow5_mid, int64_t *stats) {
    unsigned long long used = 0, ents = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key, v; pg_ld_slot(slots + 2 * i, tag, key, v);
        if (key == PG_EMPTY) continue;
        int n = 1;
        if (mode == PG_MODE_CANONICAL) n = (pg_maybe_palindrome(key, k, pow5_mid) && pg_rc_code(key, k) == key) ? 1 : 2;
        used += 1; ents += n;
    }
    for (int o = 16; o; o >>= 1) { used += __shfl_down_sync(0xffffffffu, used, o); ents += __shfl_down_sync(0xffffffffu, ents, o); }
    if ((threadIdx.x & 31) == 0) {
        if (used) atomicAdd(reinterpret_cast<unsigned long long *>(stats + PG_STAT_USED), used);
        if (ents) atomicAdd(reinterpret_cast<unsigned long long *>(stats + PG_STAT_ENTRIES), ents);
    }
}
__global__ void k_count_finish(int64_t *stats) {   // the short-record sentinel is one more entry
    if (stats[PG_STAT_SHORT] > 0) stats[PG_STAT_ENTRIES] += 1;
}

__global__ void k_table_export(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, int mode, int k,
                               const int64_t *__restrict__ stats, uint64_t *keys, uint16_t *vals, uint8_t *cnts,
                               int64_t out_cap, unsigned long long *n_out) {
    const int lane = threadIdx.x & 31;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t base = i0 - lane; base < cap; base += stride) {   // whole warps iterate together
        int64_t i = base + lane;
        PgEntry e[2]; int n = 0;
        if (i < cap) { uint64_t key, v; pg_ld_slot(slots + 2 * i, tag, key, v); n = pg_slot_entries(key, v, mode, k, e); }
        // warp-aggregated append: one atomicAdd per warp
        unsigned tot = n;
        unsigned pre = n;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += y; }
        tot = __shfl_sync(0xffffffffu, pre, 31);
        pre -= n;
        unsigned long long at = 0;
        if (lane == 0 && tot) at = atomicAdd(n_out, (unsigned long long)tot);
        at = __shfl_sync(0xffffffffu, at, 0);
        for (int q = 0; q < n; q++) {
            int64_t o = (int64_t)at + pre + q;
            if (o < out_cap) { keys[o] = e[q].key; if (vals) vals[o] = (uint16_t)e[q].val; if (cnts) cnts[o] = (uint8_t)e[q].cnt; }
        }
    }
    if (i0 == 0 && stats[PG_STAT_SHORT] > 0) {
        unsigned long long o = atomicAdd(n_out, 1ull);
        if ((int64_t)o < out_cap) {
            keys[o] = PG_EMPTY; if (vals) vals[o] = 32;
            if (cnts) cnts[o] = (uint8_t)(stats[PG_STAT_SHORT] < 255 ? stats[PG_STAT_SHORT] : 255);
        }
    }
}

__global__ void k_table_checksum(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, int mode, int k,
                                 const int64_t *__restrict__ stats, unsigned long long *out) {
    unsigned long long n = 0, sum = 0, x = 0;
    int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key, v; pg_ld_slot(slots + 2 * i, tag, key, v);
        PgEntry e[2];
        int m = pg_slot_entries(key, v, mode, k, e);
        for (int q = 0; q < m; q++) { uint64_t h = pg_entry_mix(e[q].key, e[q].val, e[q].cnt); n++; sum += h; x ^= h; }
    }
    if (i0 == 0 && stats[PG_STAT_SHORT] > 0) {
        uint64_t h = pg_entry_mix(PG_EMPTY, 32, (uint32_t)(stats[PG_STAT_SHORT] < 255 ? stats[PG_STAT_SHORT] : 255));
        n++; sum += h; x ^= h;
    }
    for (int o = 16; o; o >>= 1) {
        n += __shfl_down_sync(0xffffffffu, n, o); sum += __shfl_down_sync(0xffffffffu, sum, o); x ^= __shfl_down_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, n); atomicAdd(out + 1, sum); atomicXor(out + 2, x); }
}

// ---- K4: reduced dBG -----------------------------------------------------------------------------
__global__ void k4_rdbg_count(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, int mode, int k, uint64_t pow5_mid,
                              const int64_t *__restrict__ stats, unsigned long long *out) {
    unsigned long long ns = 0, nm = 0;
    int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key, v; pg_ld_slot(slots + 2 * i, tag, key, v);
        uint32_t f = pg_rdbg_flags_fast(key, v, mode, k, pow5_mid);
        ns += f != 0; nm += (f & 1u) + ((f >> 1) & 1u);
    }
    if (i0 == 0 && stats[PG_STAT_SHORT] > 0) nm += 1;   // sentinel: val 32 -> in-popcount 0 -> member
    for (int o = 16; o; o >>= 1) { ns += __shfl_down_sync(0xffffffffu, ns, o); nm += __shfl_down_sync(0xffffffffu, nm, o); }
    if ((threadIdx.x & 31) == 0) { if (ns) atomicAdd(out, ns); if (nm) atomicAdd(out + 1, nm); }
}

__global__ void k4_rdbg_select(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, int mode, int k, uint64_t pow5_mid,
                               const int64_t *__restrict__ stats, TableView rd) {
    int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key, v; pg_ld_slot(slots + 2 * i, tag, key, v);
        uint32_t f = pg_rdbg_flags_fast(key, v, mode, k, pow5_mid);
        if (!f) continue;
        table_put_or(rd, key, (uint64_t)(uint32_t)v | ((uint64_t)f << 32));     // every key arrives exactly once
    }
    if (i0 == 0) rd.stats[PG_STAT_SHORT] = stats[PG_STAT_SHORT];
}

__global__ void k4_rdbg_export(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, int mode, int k,
                               const int64_t *__restrict__ stats, uint64_t *keys, uint16_t *vals, int64_t out_cap,
                               unsigned long long *n_out) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t base = i0 - lane; base < cap; base += stride) {       // whole warps iterate together
        const int64_t i = base + lane;
        uint64_t key = PG_EMPTY, v = 0;
        if (i < cap) pg_ld_slot(slots + 2 * i, tag, key, v);
        const uint32_t masks = (uint32_t)v, f = key == PG_EMPTY ? 0u : (uint32_t)(v >> 32);
        const unsigned n = (f & 1u) + ((f >> 1) & 1u);
        // warp-ballot stream compaction: one atomicAdd per warp reserves the run, lanes write at their prefix
        unsigned pre = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += y; }
        const unsigned tot = __shfl_sync(0xffffffffu, pre, 31);
        if (tot == 0) continue;
        pre -= n;
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(n_out, (unsigned long long)tot);
        at = __shfl_sync(0xffffffffu, at, 0) + pre;
        if (f & 1u) { if ((int64_t)at < out_cap) { keys[at] = key; if (vals) vals[at] = (uint16_t)(masks & 0xFFFu); } at++; }
        if (f & 2u) { if ((int64_t)at < out_cap) { keys[at] = pg_rc_code(key, k); if (vals) vals[at] = (uint16_t)((masks >> 16) & 0xFFFu); } }
    }
    if (i0 == 0 && stats[PG_STAT_SHORT] > 0) {
        unsigned long long at = atomicAdd(n_out, 1ull);
        if ((int64_t)at < out_cap) { keys[at] = PG_EMPTY; if (vals) vals[at] = 32; }
    }
}

int check_table(const pg_table *t, const char *who) {
    if (!t || !t->d_slots || !t->d_stats) return pg_fail(PG_ERR_INVALID, "%s: null table", who);
    if (t->capacity < 2 || (t->capacity & (t->capacity - 1))) return pg_fail(PG_ERR_INVALID, "%s: capacity %lld is not a power of two >= 2", who, (long long)t->capacity);
    if (t->k < 1 || t->k > 27) return pg_fail(PG_ERR_INVALID, "%s: k=%d outside 1..27", who, t->k);
    if (t->mode < 0 || t->mode > 2) return pg_fail(PG_ERR_INVALID, "%s: bad mode %d", who, t->mode);
    if (t->epoch < 1 || t->epoch > PG_EPOCH_MAX)
        return pg_fail(PG_ERR_INVALID, "%s: epoch %d outside 1..%d (set 1 and call pg_table_clear once, then pg_table_reset)", who, t->epoch, PG_EPOCH_MAX);
    return PG_OK;
}

int scan_grid(int64_t n, int threads) {
    int64_t blocks = (n + threads - 1) / threads;
    int64_t maxb = (int64_t)pg_num_sms() * 8;
    return (int)(blocks < 1 ? 1 : (blocks < maxb ? blocks : maxb));
}

}  // namespace

extern "C" int64_t pg_table_bytes(int64_t capacity) { return capacity * 16; }

extern "C" int pg_table_clear(const pg_table *t, pg_stream_t stream_) {
    int rc = check_table(t, "pg_table_clear"); if (rc) return rc;
    pg_tune_once();
    cudaStream_t stream = (cudaStream_t)stream_;
    // Full clear: 16 B/slot of DRAM writes (tag 0 everywhere = free under every epoch).  Needed once for a fresh
    // allocation and when the epoch wraps; between builds pg_table_reset does the same job without touching HBM.
    // 2 CTAs per SM: stores are fire-and-forget, 16 warps per SM keep HBM busy.
    static int per_sm = -1;
    if (per_sm < 0) { const char *e = getenv("PG_CLEAR_CTAS_PER_SM"); per_sm = e ? atoi(e) : 2; if (per_sm < 1) per_sm = 1; }
    int64_t want = (t->capacity + 255) / 256, cap_grid = (int64_t)pg_num_sms() * per_sm;
    int grid = (int)(want < cap_grid ? want : cap_grid);
    k_table_clear<<<grid, 256, 0, stream>>>(reinterpret_cast<uint4 *>(t->d_slots), t->capacity, t->d_stats);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_table_reset(pg_table *t, pg_stream_t stream_) {
    int rc = check_table(t, "pg_table_reset"); if (rc) return rc;
    if (t->epoch >= PG_EPOCH_MAX) {           // tags would repeat: rewrite the slots and start over
        t->epoch = 1;
        // ALL allocated slots, not only the prefix selected by `capacity`: a slot beyond it that kept a tag of
        // the previous cycle would read as live once the capacity grows again
        pg_table all = *t;
        if (all.alloc_capacity > all.capacity) all.capacity = all.alloc_capacity;
        return pg_table_clear(&all, stream_);
    }
    t->epoch += 1;                            // every slot written so far now reads as free
    PG_CUDA(cudaMemsetAsync(t->d_stats, 0, PG_STAT_WORDS * sizeof(int64_t), (cudaStream_t)stream_));
    return PG_OK;
}

extern "C" int pg_count_short(const pg_table *t, const int64_t *d_seq_off, int64_t n_rec, int64_t g_begin, int64_t g_end,
                              pg_stream_t stream_) {
    int rc = check_table(t, "pg_count_short"); if (rc) return rc;
    if (!d_seq_off || n_rec < 0) return pg_fail(PG_ERR_INVALID, "pg_count_short: bad arguments");
    if (n_rec == 0) return PG_OK;
    int strands = t->mode == PG_MODE_LITERAL ? 1 : 2;
    // a record is owned by the half-open range holding its offset; the range that reaches the end of the stream
    // (d_seq_off[n_rec], read on the device) also owns the trailing empty records at that offset
    k2_count_short<<<(unsigned)((n_rec + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(d_seq_off, n_rec, g_begin, g_end, t->k, strands, t->d_stats);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_count_short_dev(const pg_table *t, const int64_t *d_seq_off, const int64_t *d_counts, int64_t cap_records,
                                  pg_stream_t stream_) {
    int rc = check_table(t, "pg_count_short_dev"); if (rc) return rc;
    if (!d_seq_off || !d_counts || cap_records < 0) return pg_fail(PG_ERR_INVALID, "pg_count_short_dev: bad arguments");
    int strands = t->mode == PG_MODE_LITERAL ? 1 : 2;
    k2_count_short_dev<<<32, 256, 0, (cudaStream_t)stream_>>>(d_seq_off, d_counts, cap_records, t->k, strands, t->d_stats);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_kmer_insert(const pg_table *t, const uint32_t *d_pk2, const uint32_t *d_amb, const int64_t *d_seq_off,
                              int64_t n_rec, int64_t g_begin, int64_t g_end, pg_stream_t stream_) {
    int rc = check_table(t, "pg_kmer_insert"); if (rc) return rc;
    if (!d_pk2 || !d_amb || !d_seq_off || n_rec < 0 || g_begin < 0 || g_end < g_begin)
        return pg_fail(PG_ERR_INVALID, "pg_kmer_insert: bad arguments");
    if ((reinterpret_cast<uintptr_t>(d_pk2) & 15) || (reinterpret_cast<uintptr_t>(d_amb) & 15))
        return pg_fail(PG_ERR_INVALID, "pg_kmer_insert: packed buffers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    TableView tv = make_view(t);
    rc = pg_count_short(t, d_seq_off, n_rec, g_begin, g_end, stream_); if (rc) return rc;
    if (g_end > g_begin && n_rec > 0) {
        int64_t w_first = (g_begin >> 5) & ~(int64_t)3;                 // multiple of 4 words -> 16-byte aligned staging
        int64_t w_last = (g_end + 31) >> 5;
        int64_t n_tiles = (w_last - w_first + K2_TILE_WORDS - 1) / K2_TILE_WORDS;
        int64_t n_words = ((g_end + 31) >> 5) + 4;                      // padding words exist (pg_pack_words)
        int grid = (int)(n_tiles < (int64_t)pg_num_sms() * 8 ? n_tiles : (int64_t)pg_num_sms() * 8);
        const uint64_t *pk = reinterpret_cast<const uint64_t *>(d_pk2);
        uint64_t p5 = pg_pow5(t->k - 1);
        switch (t->mode) {
        case PG_MODE_LITERAL:
            k2_kmer_insert<PG_MODE_LITERAL><<<grid, K2_THREADS, 0, stream>>>(tv, pk, d_amb, n_words, d_seq_off, n_rec, g_begin, g_end, t->k, p5, w_first, n_tiles); break;
        case PG_MODE_LITERAL_RC:
            k2_kmer_insert<PG_MODE_LITERAL_RC><<<grid, K2_THREADS, 0, stream>>>(tv, pk, d_amb, n_words, d_seq_off, n_rec, g_begin, g_end, t->k, p5, w_first, n_tiles); break;
        default:
            k2_kmer_insert<PG_MODE_CANONICAL><<<grid, K2_THREADS, 0, stream>>>(tv, pk, d_amb, n_words, d_seq_off, n_rec, g_begin, g_end, t->k, p5, w_first, n_tiles); break;
        }
    }
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_table_count(const pg_table *t, pg_stream_t stream_) {
    int rc = check_table(t, "pg_table_count"); if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(t->d_stats + PG_STAT_USED, 0, 2 * sizeof(int64_t), stream));
    k_table_count<<<scan_grid(t->capacity, 256), 256, 0, stream>>>(t->d_slots, t->capacity, pg_tag(t), t->mode, t->k, pg_pow5(t->k / 2), t->d_stats);
    k_count_finish<<<1, 1, 0, stream>>>(t->d_stats);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_table_export(const pg_table *t, uint64_t *d_keys, uint16_t *d_vals, uint8_t *d_cnts, int64_t cap,
                               int64_t *d_n, pg_stream_t stream_) {
    int rc = check_table(t, "pg_table_export"); if (rc) return rc;
    if (!d_keys || !d_n || cap < 0) return pg_fail(PG_ERR_INVALID, "pg_table_export: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int64_t), stream));
    k_table_export<<<scan_grid(t->capacity, 256), 256, 0, stream>>>(t->d_slots, t->capacity, pg_tag(t), t->mode, t->k, t->d_stats, d_keys, d_vals, d_cnts, cap,
                                                                  reinterpret_cast<unsigned long long *>(d_n));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_table_checksum(const pg_table *t, uint64_t *d_out, pg_stream_t stream_) {
    int rc = check_table(t, "pg_table_checksum"); if (rc) return rc;
    if (!d_out) return pg_fail(PG_ERR_INVALID, "pg_table_checksum: null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_out, 0, 3 * sizeof(uint64_t), stream));
    k_table_checksum<<<scan_grid(t->capacity, 256), 256, 0, stream>>>(t->d_slots, t->capacity, pg_tag(t), t->mode, t->k, t->d_stats,
                                                                    reinterpret_cast<unsigned long long *>(d_out));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_rdbg_count(const pg_table *dbg, int64_t *d_out, pg_stream_t stream_) {
    int rc = check_table(dbg, "pg_rdbg_count"); if (rc) return rc;
    if (!d_out) return pg_fail(PG_ERR_INVALID, "pg_rdbg_count: null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(int64_t), stream));
    k4_rdbg_count<<<scan_grid(dbg->capacity, 256), 256, 0, stream>>>(dbg->d_slots, dbg->capacity, pg_tag(dbg), dbg->mode, dbg->k, pg_pow5(dbg->k / 2), dbg->d_stats,
                                                                   reinterpret_cast<unsigned long long *>(d_out));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_rdbg_select(const pg_table *dbg, const pg_table *rdbg, pg_stream_t stream_) {
    int rc = check_table(dbg, "pg_rdbg_select"); if (rc) return rc;
    rc = check_table(rdbg, "pg_rdbg_select"); if (rc) return rc;
    if (dbg->mode != rdbg->mode || dbg->k != rdbg->k) return pg_fail(PG_ERR_INVALID, "pg_rdbg_select: mode/k mismatch");
    cudaStream_t stream = (cudaStream_t)stream_;
    TableView rd = make_view(rdbg);
    k4_rdbg_select<<<scan_grid(dbg->capacity, 256), 256, 0, stream>>>(dbg->d_slots, dbg->capacity, pg_tag(dbg), dbg->mode, dbg->k, pg_pow5(dbg->k / 2), dbg->d_stats, rd);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_rdbg_export(const pg_table *rdbg, uint64_t *d_keys, uint16_t *d_vals, int64_t cap, int64_t *d_n,
                              pg_stream_t stream_) {
    int rc = check_table(rdbg, "pg_rdbg_export"); if (rc) return rc;
    if (!d_keys || !d_n || cap < 0) return pg_fail(PG_ERR_INVALID, "pg_rdbg_export: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int64_t), stream));
    k4_rdbg_export<<<scan_grid(rdbg->capacity, 256), 256, 0, stream>>>(rdbg->d_slots, rdbg->capacity, pg_tag(rdbg), rdbg->mode, rdbg->k, rdbg->d_stats, d_keys, d_vals, cap,
                                                                     reinterpret_cast<unsigned long long *>(d_n));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
