// partition.cu - the two-phase dBG build: K2a (k-mer extraction -> hash-partitioned update records)
// and K3 (record insertion, partition by partition).
//
// Why two phases: one fused extract+insert launch (dbg_table.cu) probes a multi-GB table at random -
// ncu shows it DRAM-activate bound (16 G probes/s, ~3.4 DRAM sectors read + 1 written per probe).
// Here K2a streams the packed sequence once and emits one 16-byte record {key, masks, inc} per
// position, bucketed by the high bits of the slot index the key hashes to (and, across GPUs, by the
// owner rank).  K3 then walks the buckets in order, so at any moment the CTAs of the grid hammer one
// table region that fits the 126 MB L2: atomicCAS / red.or / red.add resolve in L2 and every table
// line is written back to HBM once.  The bucketed records are also exactly what the multi-GPU path
// exchanges (owner = low hash bits, disjoint from the slot bits, which are the top ones).
//
// Record emission inside a CTA tile is a counting sort in shared memory (histogram, one global
// atomicAdd per bucket per tile to reserve space, reorder through an index permutation) so global
// stores are coalesced 16-byte runs per bucket.
#include "tile_sort.cuh"

namespace {

struct PartArgs {
    const uint64_t *pk2; const uint32_t *amb; int64_t n_words;
    const int64_t *seq_off; int64_t n_rec, g_begin, g_end; int k; uint64_t pow5km1;
    int64_t t_first, n_tiles;
    int owner_bits;
    BucketOut out;
    // distinct-key estimator: keys whose mix has bits 8..15 == 0 (a 1/256 sample of the KEY space, so every
    // occurrence of a sampled key is sampled) go into a small CAS set; 256 x its size estimates the table
    uint64_t *sample_keys; uint64_t sample_mask; unsigned long long *sample_count;
    // device-side arguments: when set, n_rec / the stream range come from K1's count block and record
    // index ON THE DEVICE, so the host can enqueue K1 -> K2a -> K3 without waiting for K1's result;
    // g_begin / g_end are then offsets RELATIVE to the first record (one round of a multi-round build)
    const int64_t *d_counts; int64_t cap_records;
};

__device__ __forceinline__ void sample_key(const PartArgs &a, uint64_t key, uint64_t h) {
    uint64_t s = (h >> 16) & a.sample_mask;
    for (uint64_t probe = 0; probe <= a.sample_mask; probe++) {
        uint64_t ck = a.sample_keys[s];
        if (ck == key) return;
        if (ck == PG_EMPTY) {
            uint64_t old = atomicCAS(reinterpret_cast<unsigned long long *>(a.sample_keys + s), (unsigned long long)PG_EMPTY, (unsigned long long)key);
            if (old == PG_EMPTY) { atomicAdd(a.sample_count, 1ull); return; }
            if (old == key) return;
        }
        s = (s + 1) & a.sample_mask;
    }
    atomicAdd(a.sample_count, 1ull << 40);     // set full: poison the estimate so the host falls back to the upper bound
}

// GENERAL = key-space sampling and/or an owner rank in the bucket id; the plain single-GPU build needs neither
template <bool GENERAL>
__device__ __forceinline__ uint32_t part_of(const PartArgs &a, uint64_t key) {
    uint64_t h = pg_mix64(key);
    uint32_t sub = a.out.sub_bits ? (uint32_t)(h >> (64 - a.out.sub_bits)) : 0u;      // hash prefix = table region (tv_home)
    if (!GENERAL) return sub;
    if (a.sample_keys && ((h >> 8) & 0xFFu) == 0) sample_key(a, key, h);
    uint32_t owner = a.owner_bits ? (uint32_t)(h & ((1u << a.owner_bits) - 1u)) : 0u;   // low bits: disjoint from the slot bits
    return (owner << a.out.sub_bits) | sub;
}


template <int MODE, int KP_THREADS, bool GENERAL>
__global__ void __launch_bounds__(KP_THREADS, KP_THREADS == 256 ? 3 : 1)
k2a_partition(PartArgs a) {
    constexpr int KP_TILE = KP_THREADS * KP_G;
    if (a.d_counts) {      // all records of the packed stream, bounds read from the device
        const int64_t n_rec = a.d_counts[0];
        if (n_rec > a.cap_records) {          // the record index was truncated: poison bucket 0, the host falls back
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                a.out.part_counts[0] = 1ull << 62;
                if (a.out.spill_cap > 0) a.out.part_counts[a.out.n_parts] = 1ull << 62;      // a spill would absorb bucket 0's "overflow"
            }
            return;
        }
        a.n_rec = n_rec;
        const int64_t s0 = n_rec > 0 ? a.seq_off[0] : 0, s1 = n_rec > 0 ? a.seq_off[n_rec] : 0;
        const int64_t lo = s0 + a.g_begin, hi = s0 + a.g_end;          // this round's slice of [s0, s1)
        a.g_begin = lo < s1 ? lo : s1;
        a.g_end = (a.g_end < 0 || hi > s1) ? s1 : hi;                  // g_end < 0: to the end of the stream
        a.n_words = ((s1 + 31) >> 5) + 4;
        a.t_first = a.g_begin / KP_TILE;
        a.n_tiles = a.g_end > a.g_begin ? (a.g_end + KP_TILE - 1) / KP_TILE - a.t_first : 0;
    }
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int RPP = (MODE == PG_MODE_LITERAL_RC) ? 2 : 1;               // records per position
    constexpr int MAXR = KP_TILE * RPP;
    TileSort<MAXR> ts;
    ts.carve(smem, a.out.n_parts);
    __shared__ uint32_t s_nrec;
    __shared__ uint32_t s_chunk[32];
    const uint64_t pol = pg_policy_evict_first();
    __shared__ uint32_t s_vlut[16];
    __shared__ uint16_t s_lut5[PG_LUT5_SIZE];
    if (threadIdx.x < 16) s_vlut[threadIdx.x] = pg_vlut_entry(threadIdx.x);      // visible after the tile loop's first barrier
    for (int i = threadIdx.x; i < PG_LUT5_SIZE; i += KP_THREADS) s_lut5[i] = (uint16_t)pg_lut5_entry(i);

    auto emit = [&](int slot, uint64_t key, uint32_t masks, uint32_t inc) { ts.emit(slot, part_of<GENERAL>(a, key), key, masks, inc); };
    // staging slot of this thread's q-th position: consecutive lanes take consecutive slots, so the 16-byte
    // record stores of a warp fall into distinct banks (a [thread][q] layout costs 8 wavefronts per store)
    auto emit_pos = [&](int q, uint64_t F, uint64_t R, uint32_t vw) {      // vw = vf | vr << 16
        const int slot = q * KP_THREADS + threadIdx.x;
        if (MODE == PG_MODE_CANONICAL) {
            PgUpdate u = pg_canonical_update_w(F, R, vw);
            emit(slot, u.key, u.masks, u.inc);
        } else {
            emit(slot, F, vw & 0xFFFFu, 1u);
            if (MODE == PG_MODE_LITERAL_RC) emit(slot + KP_TILE, R, vw >> 16, 1u);
        }
    };

    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        __syncthreads();
        ts.reset(a.out.n_parts, KP_THREADS);
        __syncthreads();
        // ---- 1. compute this thread's records ------------------------------------------------
        const int64_t g0 = (a.t_first + tile) * KP_TILE + (int64_t)threadIdx.x * KP_G;
        bool done = false;
        if (g0 < a.g_end && g0 + KP_G > a.g_begin) {
            const int64_t wi = g0 >> 5;
            const int j0 = (int)(g0 & 31);
            PgWindow w;
            w.prv = wi > 0 ? __ldg(a.pk2 + wi - 1) : 0; w.cur = __ldg(a.pk2 + wi); w.nxt = wi + 1 < a.n_words ? __ldg(a.pk2 + wi + 1) : 0;
            w.aprv = wi > 0 ? __ldg(a.amb + wi - 1) : 0; w.acur = __ldg(a.amb + wi); w.anxt = wi + 1 < a.n_words ? __ldg(a.amb + wi + 1) : 0;
            int64_t r = find_record(a.seq_off, a.n_rec, g0);
            int64_t rs = r >= 0 ? __ldg(a.seq_off + r) : 0, re = __ldg(a.seq_off + r + 1);
            const int k = a.k;
            if (pg_is_interior(w, g0, KP_G, k, rs, re, r >= 0, a.g_begin, a.g_end)) {
                // ---- fast path: 16 ACGT positions strictly inside one record (kmer_core.cuh)
                pg_interior_visit<KP_G>(w, j0, k, a.pow5km1, s_vlut, s_lut5, emit_pos);
                done = true;
            } else {
                // ---- generic path: record edges, ambiguity codes, range ends (all the quirks) ----
                uint64_t F, R;
                pg_codes_init(w, j0, k, F, R);
#pragma unroll 1
                for (int q = 0; q < KP_G; q++) {
                    const int64_t g = g0 + q;
                    const int j = j0 + q;
                    bool ok = false;
                    if (g < a.g_end) {
                        while (r + 1 < a.n_rec && g >= re) { r++; rs = re; re = __ldg(a.seq_off + r + 1); }
                        ok = g >= a.g_begin && r >= 0 && g + k <= re;
                    }
                    if (ok) {
                        uint32_t vf, vr;
                        pg_occ_vals(w, j, g - rs, re - rs, k, vf, vr);
                        emit_pos(q, F, R, vf | (vr << 16));
                    } else {
                        for (int e = 0; e < RPP; e++) ts.s_pid[e * KP_TILE + q * KP_THREADS + threadIdx.x] = NOREC;
                    }
                    pg_codes_roll(w, j, k, a.pow5km1, F, R);
                }
                done = true;
            }
        }
        if (!done)
            for (int e = 0; e < KP_G * RPP; e++) ts.s_pid[e * KP_THREADS + threadIdx.x] = NOREC;
        __syncthreads();
        ts.template sort_write<KP_THREADS>(a.out, s_chunk, &s_nrec, pol);
    }
}

// K2b: split coarse buckets into table regions.  On the multi-GPU path records cross NVLink bucketed by OWNER
// only (long runs: a 8192-position tile gives 1024-record = 16 KB runs per peer instead of 512-byte ones), and the
// receiver sorts what arrived - n_seg segments, one per source rank - into the 2^sub_bits hash-prefix regions K3
// sweeps.  Same counting sort as K2a, the records are simply re-read instead of computed: 32 B of HBM traffic per
// record.  Also what keeps a region inside L2 when a table needs more regions than one K2a pass can address.
struct SplitArgs {
    const uint4 *in; const int64_t *seg_off, *seg_cnt; int n_seg; int64_t seg_cap;
    BucketOut out;
    int64_t *stats;        // optional table statistics: PG_STAT_LOST is raised when an input segment claims more than seg_cap
};
constexpr int KB_MAX_SEG = 64;
// 256-thread tiles (4096 records, 3 CTAs/SM) up to 256 regions; 512-thread tiles (8192 records, one CTA per SM) beyond:
// with 1024 regions a 4096-record tile leaves 4-record (64-byte) runs per bucket - measured 46 G records/s on BASELINE
// config 4 at 2 GPUs against 116 G records/s with 128 regions (profiles/r2d_*)
template <int KB_THREADS>
__global__ void __launch_bounds__(KB_THREADS, KB_THREADS == 256 ? 3 : 1)
k2b_split(SplitArgs a) {
    constexpr int KB_TILE = KB_THREADS * KP_G;
    extern __shared__ __align__(16) unsigned char smem[];
    TileSort<KB_TILE> ts;
    ts.carve(smem, a.out.n_parts);
    __shared__ uint32_t s_nrec;
    __shared__ uint32_t s_chunk[32];
    __shared__ long long s_tile0[KB_MAX_SEG + 1];       // first tile of every segment (exclusive prefix of ceil(cnt / tile))
    const uint64_t pol = pg_policy_evict_first();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int s = 0; s < a.n_seg; s++) {
            long long c = a.seg_cnt[s];
            if (c > a.seg_cap) {       // the sender dropped records (its wire bucket overflowed): never read past the segment
                if (blockIdx.x == 0 && a.stats) atomicExch(reinterpret_cast<unsigned long long *>(a.stats + PG_STAT_LOST), 1ull);
                c = a.seg_cap;
            }
            if (c < 0) c = 0;
            s_tile0[s] = t; t += (c + KB_TILE - 1) / KB_TILE;
        }
        s_tile0[a.n_seg] = t;
    }
    __syncthreads();
    const long long n_tiles = s_tile0[a.n_seg];
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        ts.reset(a.out.n_parts, KB_THREADS);
        int seg = 0;
        while (seg + 1 < a.n_seg && tile >= s_tile0[seg + 1]) seg++;
        long long cnt = a.seg_cnt[seg]; if (cnt > a.seg_cap) cnt = a.seg_cap;
        const long long i0 = (tile - s_tile0[seg]) * KB_TILE;
        const uint4 *src = a.in + a.seg_off[seg] + i0;
        const long long left = cnt - i0;                   // records of this tile: min(left, KB_TILE)
        __syncthreads();
        // two batches of eight 16-byte loads in flight per thread (sixteen would spill at 3 CTAs per SM)
#pragma unroll
        for (int h = 0; h < KP_G; h += 8) {
            uint4 r[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int slot = (h + q) * KB_THREADS + threadIdx.x;
                if (slot < left) r[q] = pg_ld_stream_l2first(src + slot, pol);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int slot = (h + q) * KB_THREADS + threadIdx.x;
                if (slot < left) {
                    const uint64_t key = (uint64_t)r[q].x | ((uint64_t)r[q].y << 32);
                    const uint32_t pid = a.out.sub_bits ? (uint32_t)(pg_mix64(key) >> (64 - a.out.sub_bits)) : 0u;
                    ts.emit(slot, pid, key, r[q].z, r[q].w);
                } else {
                    ts.s_pid[slot] = NOREC;
                }
            }
        }
        __syncthreads();
        ts.template sort_write<KB_THREADS>(a.out, s_chunk, &s_nrec, pol);
    }
}

// What K3 sweeps after K2a / K2b filled a local bucket set: per-bucket counts clamped to the capacities (the
// spill is region n_parts), and the sticky PG_STAT_LOST flag when records were dropped - a bucket overflowed
// without a spill, or the spill itself overflowed.  One tiny CTA; keeps the step free of host read-backs.
__global__ void k_buckets_plan(const unsigned long long *__restrict__ counts, int n_parts, int64_t part_cap, int64_t spill_cap,
                               int64_t *__restrict__ seg_cnt, int64_t *stats) {
    bool lost = false;
    for (int i = threadIdx.x; i <= n_parts; i += blockDim.x) {
        const unsigned long long c = counts[i];
        const int64_t cap = i < n_parts ? part_cap : spill_cap;
        if (i < n_parts) { if (spill_cap <= 0 && c > (unsigned long long)part_cap) lost = true; }
        else if (c > (unsigned long long)(spill_cap > 0 ? spill_cap : 0)) lost = true;
        seg_cnt[i] = c > (unsigned long long)cap ? cap : (int64_t)c;
    }
    if (lost && stats) atomicExch(reinterpret_cast<unsigned long long *>(stats + PG_STAT_LOST), 1ull);
}

// K3: insert update records region by region.  The records of table region b arrive as n_src
// segments (one per source rank / pipeline chunk): seg_off/seg_cnt are [n_regions][n_src] in
// region-major order and the threads of the whole grid stride over the CONCATENATION of a region's
// segments, so a region is swept once with every thread busy while its 8 MB of slots sit in L2.
// (Loading the next record before the current one is merged, or pipelining record / slot / CAS three deep,
// measured no faster: the kernel waits on random DRAM sectors, not on its own dependency chain.)
constexpr int K3_MAX_SRC = 64;
// record i of region b's concatenated segments (false past the end).  No shared memory and no
// barriers: every warp walks the regions at its own pace, the segment table is read through L1.
__device__ __forceinline__ bool k3_fetch(const uint4 *__restrict__ records, const int64_t *__restrict__ seg_off,
                                         const int64_t *__restrict__ seg_cnt, int n_src, int64_t seg_cap, int b, int64_t i, uint4 &r) {
    const int64_t *off = seg_off + (int64_t)b * n_src, *cnt = seg_cnt + (int64_t)b * n_src;
    for (int j = 0; j < n_src; j++) {
        int64_t c = __ldg(cnt + j);
        if (c > seg_cap) c = seg_cap;      // a count above the segment capacity means records were dropped (the host checks); never read past it
        if (i < c) { r = pg_ld_stream(records + __ldg(off + j) + i); return true; }
        i -= c;
    }
    return false;
}
template <int MINB, bool ROTATE>
__global__ void __launch_bounds__(256, MINB)
k3_insert_records(TableView t, const uint4 *__restrict__ records, const int64_t *__restrict__ seg_off,
                  const int64_t *__restrict__ seg_cnt, int n_regions, int n_src, int64_t seg_cap) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t rot = ROTATE ? (int64_t)((gridDim.x * 618u) / 1000u) * blockDim.x : 0;
    uint32_t n_claimed = 0;
    // A region holds a non-integral number of grid strides of records, so the threads with the lowest
    // start index do one more record than the rest.  The start index is rotated from region to region
    // (whole CTAs, golden-ratio steps) so that the extra record falls on different threads each time
    // and every thread ends up with the same total instead of the low ones carrying all the excess.
    int64_t start = i0;
    for (int b = 0; b < n_regions; b++) {
        uint4 r;
        for (int64_t i = start; k3_fetch(records, seg_off, seg_cnt, n_src, seg_cap, b, i, r); i += stride)
            table_upsert(t, (uint64_t)r.x | ((uint64_t)r.y << 32), r.z, r.w, n_claimed);
        start += rot; if (start >= stride) start -= stride;
    }
    publish_claims(t, n_claimed);
}

// K3 with the record stream one step ahead: the sweep above keeps ONE dependent chain per thread (record -> slot ->
// atomic), so a region pass costs at least that chain's latency (~3.5 us) however few records the region holds - the floor
// a multi-round build pays once per round and region.  Here the load of a thread's NEXT record (possibly of the next
// region) is in flight while the current one is inserted, which takes the HBM read of the stream off the chain.
// (More independent chains per thread do NOT help at full load: the random-slot micro-benchmark saturates at ~110 G
// sector requests/s whatever the threads x operations in flight - profiles/r2b_microbench.jsonl - and K3 already moves
// that many; a batched variant with 2..8 records in flight per thread measured 1.03-5.4 ms against 1.01 ms.)
struct K3Cursor {
    const uint4 *__restrict__ records; const int64_t *__restrict__ seg_off, *__restrict__ seg_cnt;
    int n_regions; int64_t seg_cap; uint32_t stride, rot, start, i; int b; uint32_t c; const uint4 *base;
    __device__ __forceinline__ void open_region() {
        int64_t c64 = __ldg(seg_cnt + b);
        if (c64 > seg_cap) c64 = seg_cap;
        c = c64 > 0x7FFFFFFF ? 0x7FFFFFFFu : (c64 < 0 ? 0u : (uint32_t)c64);
        base = records + __ldg(seg_off + b);
        i = start;
    }
    // load the next record of this thread's walk (false at the end of the sweep)
    __device__ __forceinline__ bool next(uint4 &r) {
        for (;;) {
            if (i < c) { r = pg_ld_stream(base + i); i += stride; return true; }
            if (++b >= n_regions) return false;
            start += rot; if (start >= stride) start -= stride;
            open_region();
        }
    }
};
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k3_insert_ahead(TableView t, const uint4 *__restrict__ records, const int64_t *__restrict__ seg_off,
                const int64_t *__restrict__ seg_cnt, int n_regions, int64_t seg_cap) {
    K3Cursor cur;
    cur.records = records; cur.seg_off = seg_off; cur.seg_cnt = seg_cnt; cur.n_regions = n_regions; cur.seg_cap = seg_cap;
    cur.stride = gridDim.x * blockDim.x; cur.rot = ((gridDim.x * 618u) / 1000u) * blockDim.x;
    cur.start = blockIdx.x * blockDim.x + threadIdx.x; cur.b = 0;
    uint32_t n_claimed = 0;
    if (n_regions > 0) {
        cur.open_region();
        uint4 r, rn;
        bool have = cur.next(r);
        while (have) {
            const bool have_next = cur.next(rn);            // in flight while r is inserted
            table_upsert(t, (uint64_t)r.x | ((uint64_t)r.y << 32), r.z, r.w, n_claimed);
            r = rn; have = have_next;
        }
    }
    publish_claims(t, n_claimed);
}

int part_smem_bytes(int mode, int n_parts, int threads) {
    int maxr = (mode == PG_MODE_LITERAL_RC ? 2 : 1) * threads * KP_G;
    return maxr * 12 + maxr * 2 * 2 + 4 * n_parts * 4 + 2 * n_parts * 8 + 16;       // == TileSort<maxr>::bytes(n_parts)
}

}  // namespace


// g_begin / g_end: absolute stream offsets, or - with d_counts - offsets relative to the first record (g_end < 0 = to the end)
static int partition_launch(const pg_table *t, const uint32_t *d_pk2, const uint32_t *d_amb, const int64_t *d_seq_off,
                            int64_t n_rec, int64_t g_begin, int64_t g_end, const pg_bucket_set *b, bool zero_counts,
                            uint64_t *d_sample_keys, int64_t sample_cap, int64_t *d_sample_count,
                            const int64_t *d_counts, int64_t cap_records, int64_t max_bases, pg_stream_t stream_) {
    if (!t || t->k < 1 || t->k > 27 || t->mode < 0 || t->mode > 2)
        return pg_fail(PG_ERR_INVALID, "pg_kmer_partition: bad table descriptor (only mode and k are used)");
    if (!d_pk2 || !d_amb || !d_seq_off || n_rec < 0 || g_begin < 0 || (!d_counts && g_end < g_begin))
        return pg_fail(PG_ERR_INVALID, "pg_kmer_partition: bad arguments");
    PartArgs a;
    int rc = make_bucket_out(b, "pg_kmer_partition", a.out); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    const int n_parts = a.out.n_parts;
    if (zero_counts) PG_CUDA(cudaMemsetAsync(b->d_part_counts, 0, (size_t)(n_parts + 1) * 8, st));
    int64_t span = g_end - g_begin;                                     // grid sizing only in device-argument mode
    if (d_counts) { n_rec = 1; span = (g_end < 0 || g_end - g_begin > max_bases) ? max_bases - (g_begin < max_bases ? g_begin : max_bases) : g_end - g_begin; }
    if (n_rec == 0 || span <= 0) return PG_OK;
    a.pk2 = reinterpret_cast<const uint64_t *>(d_pk2); a.amb = d_amb; a.n_words = ((g_end + 31) >> 5) + 4;
    a.seq_off = d_seq_off; a.n_rec = n_rec; a.g_begin = g_begin; a.g_end = g_end; a.k = t->k; a.pow5km1 = pg_pow5(t->k - 1);
    static int thr_env = -1;
    if (thr_env < 0) { const char *e = getenv("PG_K2A_THREADS"); thr_env = e ? atoi(e) : 0; }
    const bool peer = b->d_peer_bases != nullptr;
    const int threads = (thr_env == 128 || thr_env == 256 || thr_env == 512) ? thr_env : (peer ? (t->mode == PG_MODE_LITERAL_RC ? 256 : 512) : 256);
    const int tile = threads * KP_G;
    a.t_first = g_begin / tile; a.n_tiles = (g_begin + span + tile - 1) / tile - a.t_first;
    a.owner_bits = b->owner_bits;
    a.d_counts = d_counts; a.cap_records = cap_records;
    a.sample_keys = nullptr; a.sample_mask = 0; a.sample_count = nullptr;
    if (d_sample_keys) {
        if (!d_sample_count || sample_cap < 2 || (sample_cap & (sample_cap - 1)))
            return pg_fail(PG_ERR_INVALID, "pg_kmer_partition: sample set needs a power-of-two capacity and a counter");
        a.sample_keys = d_sample_keys; a.sample_mask = (uint64_t)sample_cap - 1;
        a.sample_count = reinterpret_cast<unsigned long long *>(d_sample_count);
    }
    int smem = part_smem_bytes(t->mode, n_parts, threads);
    int64_t maxg = (int64_t)pg_num_sms() * ctas_per_sm(smem, 2048 + 256);      // static: the 2 KB digit-group table + small arrays
    int grid = (int)(a.n_tiles < maxg ? a.n_tiles : maxg);
    if (grid < 1) grid = 1;
    const bool general = a.sample_keys != nullptr || b->owner_bits > 0;
#define K2A_LAUNCH(M, T)                                                                                                    \
    do {                                                                                                                    \
        if (general) {                                                                                                      \
            PG_CUDA(cudaFuncSetAttribute(k2a_partition<M, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
            k2a_partition<M, T, true><<<grid, T, smem, st>>>(a);                                                            \
        } else {                                                                                                            \
            PG_CUDA(cudaFuncSetAttribute(k2a_partition<M, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
            k2a_partition<M, T, false><<<grid, T, smem, st>>>(a);                                                           \
        }                                                                                                                   \
    } while (0)
    if (threads == 128) {
        if (t->mode == PG_MODE_LITERAL) K2A_LAUNCH(PG_MODE_LITERAL, 128);
        else if (t->mode == PG_MODE_LITERAL_RC) K2A_LAUNCH(PG_MODE_LITERAL_RC, 128);
        else K2A_LAUNCH(PG_MODE_CANONICAL, 128);
    } else if (threads == 256) {
        if (t->mode == PG_MODE_LITERAL) K2A_LAUNCH(PG_MODE_LITERAL, 256);
        else if (t->mode == PG_MODE_LITERAL_RC) K2A_LAUNCH(PG_MODE_LITERAL_RC, 256);
        else K2A_LAUNCH(PG_MODE_CANONICAL, 256);
    } else {
        if (t->mode == PG_MODE_LITERAL_RC) return pg_fail(PG_ERR_INVALID, "pg_kmer_partition: 512-thread tiles do not fit the literal-rc mode");
        if (t->mode == PG_MODE_LITERAL) K2A_LAUNCH(PG_MODE_LITERAL, 512);
        else K2A_LAUNCH(PG_MODE_CANONICAL, 512);
    }
#undef K2A_LAUNCH
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

static pg_bucket_set local_set(uint64_t *d_records, int64_t part_cap, int64_t *d_part_counts, int owner_bits, int sub_bits) {
    pg_bucket_set b;
    b.d_records = d_records; b.d_peer_bases = nullptr; b.d_part_counts = d_part_counts; b.part_cap = part_cap; b.spill_cap = 0;
    b.owner_bits = owner_bits; b.sub_bits = sub_bits; b.my_rank = 0; b.reserved = 0;
    return b;
}
// NB the classic entry points keep their contract: d_part_counts has n_parts counters (no spill counter), so the
// memset there covers n_parts words only
static int classic_zero(int64_t *d_part_counts, int owner_bits, int sub_bits, pg_stream_t stream_) {
    if (!d_part_counts || owner_bits < 0 || sub_bits < 0 || owner_bits + sub_bits > 10) return pg_fail(PG_ERR_INVALID, "pg_kmer_partition: bad arguments");
    PG_CUDA(cudaMemsetAsync(d_part_counts, 0, (size_t)(1 << (owner_bits + sub_bits)) * 8, (cudaStream_t)stream_));
    return PG_OK;
}

extern "C" int pg_kmer_partition(const pg_table *t, const uint32_t *d_pk2, const uint32_t *d_amb, const int64_t *d_seq_off,
                                 int64_t n_rec, int64_t g_begin, int64_t g_end, int owner_bits, int sub_bits,
                                 uint64_t *d_records, int64_t part_cap, int64_t *d_part_counts, uint64_t *d_sample_keys,
                                 int64_t sample_cap, int64_t *d_sample_count, pg_stream_t stream_) {
    int rc = classic_zero(d_part_counts, owner_bits, sub_bits, stream_); if (rc) return rc;
    pg_bucket_set b = local_set(d_records, part_cap, d_part_counts, owner_bits, sub_bits);
    return partition_launch(t, d_pk2, d_amb, d_seq_off, n_rec, g_begin, g_end, &b, false, d_sample_keys, sample_cap, d_sample_count,
                            nullptr, 0, 0, stream_);
}

// Same as pg_kmer_partition over ALL records of a packed stream, but n_rec and the stream range are read on
// the device from K1's outputs (d_counts[0], d_seq_off) - nothing K1 produced has to reach the host first.
extern "C" int pg_kmer_partition_dev(const pg_table *t, const uint32_t *d_pk2, const uint32_t *d_amb, const int64_t *d_seq_off,
                                     const int64_t *d_counts, int64_t cap_records, int64_t max_bases, int owner_bits, int sub_bits,
                                     uint64_t *d_records, int64_t part_cap, int64_t *d_part_counts, pg_stream_t stream_) {
    if (!d_counts || cap_records < 0 || max_bases < 0) return pg_fail(PG_ERR_INVALID, "pg_kmer_partition_dev: bad arguments");
    int rc = classic_zero(d_part_counts, owner_bits, sub_bits, stream_); if (rc) return rc;
    pg_bucket_set b = local_set(d_records, part_cap, d_part_counts, owner_bits, sub_bits);
    return partition_launch(t, d_pk2, d_amb, d_seq_off, 1, 0, -1, &b, false, nullptr, 0, nullptr, d_counts, cap_records, max_bases, stream_);
}

// ---- bucket-set API: rounds, spill, receiver-side split (the streaming builders) ---------------------------
extern "C" int pg_kmer_partition_to(const pg_table *t, const uint32_t *d_pk2, const uint32_t *d_amb, const int64_t *d_seq_off,
                                    int64_t n_rec, int64_t g_begin, int64_t g_end, const int64_t *d_counts, int64_t cap_records,
                                    int64_t max_bases, const pg_bucket_set *out, uint64_t *d_sample_keys, int64_t sample_cap,
                                    int64_t *d_sample_count, pg_stream_t stream_) {
    if (d_counts && (cap_records < 0 || max_bases < 0)) return pg_fail(PG_ERR_INVALID, "pg_kmer_partition_to: bad device-argument sizes");
    return partition_launch(t, d_pk2, d_amb, d_seq_off, n_rec, g_begin, g_end, out, true, d_sample_keys, sample_cap, d_sample_count,
                            d_counts, cap_records, max_bases, stream_);
}

extern "C" int pg_records_split(const uint64_t *d_records_in, const int64_t *d_seg_off, const int64_t *d_seg_cnt, int n_seg,
                                int64_t seg_cap, const pg_bucket_set *out, int64_t *d_table_stats, pg_stream_t stream_) {
    SplitArgs a;
    int rc = make_bucket_out(out, "pg_records_split", a.out); if (rc) return rc;
    if (out->d_peer_bases || out->owner_bits != 0) return pg_fail(PG_ERR_INVALID, "pg_records_split: the output must be a local bucket set with owner_bits 0");
    if (!d_records_in || !d_seg_off || !d_seg_cnt || n_seg < 1 || n_seg > KB_MAX_SEG || seg_cap < 1 || (reinterpret_cast<uintptr_t>(d_records_in) & 15))
        return pg_fail(PG_ERR_INVALID, "pg_records_split: bad arguments (1..%d segments, 16-byte aligned records)", KB_MAX_SEG);
    cudaStream_t st = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(out->d_part_counts, 0, (size_t)(a.out.n_parts + 1) * 8, st));
    a.in = reinterpret_cast<const uint4 *>(d_records_in); a.seg_off = d_seg_off; a.seg_cnt = d_seg_cnt; a.n_seg = n_seg; a.seg_cap = seg_cap;
    a.stats = d_table_stats;
    static int v1 = -1;
    if (v1 < 0) { const char *e = getenv("PG_SPLIT_V1"); v1 = e ? atoi(e) : 0; }
    if (!v1 && a.out.n_parts <= 256) {       // beyond 256 ways the 8192-record tile sort below is the faster one (profiles/r2k_*)
        PgMultiSplit m;
        m.in = a.in; m.seg_off = d_seg_off; m.seg_cnt = reinterpret_cast<const unsigned long long *>(d_seg_cnt); m.n_seg = n_seg; m.seg_cap = seg_cap;
        m.pass_seg = -1; m.pass_cap = 0; m.lost_on_clamp = 1;
        m.out = a.out.records; m.out_counts = a.out.part_counts; m.out_part_cap = a.out.part_cap; m.out_spill_cap = a.out.spill_cap;
        m.n_out = a.out.n_parts; m.skip_bits = 0; m.bits = out->sub_bits; m.sliced = 0; m.stats = d_table_stats;
        return pg_multisplit_launch(m, st);
    }
    static int thr_env = -1;
    if (thr_env < 0) { const char *e = getenv("PG_K2B_THREADS"); thr_env = e ? atoi(e) : 0; }
    const int threads = (thr_env == 256 || thr_env == 512) ? thr_env : (a.out.n_parts > 256 ? 512 : 256);
    const int tile = threads * KP_G;
    const int smem = threads == 256 ? TileSort<256 * KP_G>::bytes(a.out.n_parts) : TileSort<512 * KP_G>::bytes(a.out.n_parts);
    const int64_t max_tiles = (int64_t)n_seg * ((seg_cap + tile - 1) / tile);
    int64_t maxg = (int64_t)pg_num_sms() * ctas_per_sm(smem, 1024);
    int grid = (int)(max_tiles < maxg ? max_tiles : maxg);
    if (grid < 1) grid = 1;
    if (threads == 256) {
        PG_CUDA(cudaFuncSetAttribute(k2b_split<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k2b_split<256><<<grid, 256, smem, st>>>(a);
    } else {
        PG_CUDA(cudaFuncSetAttribute(k2b_split<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k2b_split<512><<<grid, 512, smem, st>>>(a);
    }
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_buckets_plan(const pg_bucket_set *b, int64_t *d_seg_cnt, int64_t *d_table_stats, pg_stream_t stream_) {
    BucketOut o;
    int rc = make_bucket_out(b, "pg_buckets_plan", o); if (rc) return rc;
    if (b->d_peer_bases || !d_seg_cnt) return pg_fail(PG_ERR_INVALID, "pg_buckets_plan: needs a local bucket set and an output array");
    k_buckets_plan<<<1, 256, 0, (cudaStream_t)stream_>>>(o.part_counts, o.n_parts, o.part_cap, o.spill_cap, d_seg_cnt, d_table_stats);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

// ---- peer memory for the fused exchange (CUDA IPC; one process per GPU) ------------------------
extern "C" int pg_peer_alloc(int64_t bytes, void **d_ptr, uint8_t *handle64) {
    if (bytes <= 0 || !d_ptr || !handle64) return pg_fail(PG_ERR_INVALID, "pg_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PG_CUDA(cudaMalloc(d_ptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    PG_CUDA(cudaIpcGetMemHandle(&h, *d_ptr));
    memcpy(handle64, &h, 64);
    return PG_OK;
}
extern "C" int pg_peer_open(const uint8_t *handle64, void **d_ptr) {
    if (!handle64 || !d_ptr) return pg_fail(PG_ERR_INVALID, "pg_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PG_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PG_OK;
}
extern "C" int pg_peer_close(void *d_ptr) { PG_CUDA(cudaIpcCloseMemHandle(d_ptr)); return PG_OK; }
extern "C" int pg_peer_free(void *d_ptr) { PG_CUDA(cudaFree(d_ptr)); return PG_OK; }

extern "C" int pg_insert_records(const pg_table *t, const uint64_t *d_records, const int64_t *d_seg_off,
                                 const int64_t *d_seg_cnt, int n_regions, int n_src, int64_t seg_cap, pg_stream_t stream_) {
    if (!t || !t->d_slots || !t->d_stats || t->capacity < 2 || (t->capacity & (t->capacity - 1)) || t->epoch < 1 || t->epoch > PG_EPOCH_MAX)
        return pg_fail(PG_ERR_INVALID, "pg_insert_records: bad table");
    if (n_regions < 0 || n_src < 1 || n_src > K3_MAX_SRC || (n_regions > 0 && (!d_records || !d_seg_off || !d_seg_cnt)))
        return pg_fail(PG_ERR_INVALID, "pg_insert_records: bad arguments (n_src must be 1..%d)", K3_MAX_SRC);
    if (n_regions == 0) return PG_OK;
    if (seg_cap <= 0) seg_cap = INT64_MAX;
    if (reinterpret_cast<uintptr_t>(d_records) & 15) return pg_fail(PG_ERR_INVALID, "pg_insert_records: records must be 16-byte aligned");
    TableView tv = make_view(t);
    static int gmul = -1, rotate = 1, ahead = 0;
    if (gmul < 0) {
        const char *e = getenv("PG_K3_GRID"); gmul = e ? atoi(e) : 5;
        e = getenv("PG_K3_ROTATE"); rotate = e ? atoi(e) : 1;
        e = getenv("PG_K3_AHEAD"); ahead = e ? atoi(e) : 0;
    }
    // the grid is exactly the resident CTAs (a region sweep must not leave a second wave behind): 8 per SM caps the
    // kernel at 32 registers, which costs it a few stack slots; 6 per SM runs it spill-free at 38
    int grid = pg_num_sms() * gmul;
    const uint4 *recs = reinterpret_cast<const uint4 *>(d_records);
    cudaStream_t st = (cudaStream_t)stream_;
#define K3_LAUNCH(M, R) k3_insert_records<M, R><<<grid, 256, 0, st>>>(tv, recs, d_seg_off, d_seg_cnt, n_regions, n_src, seg_cap)
#define K3A_LAUNCH(M) k3_insert_ahead<M><<<grid, 256, 0, st>>>(tv, recs, d_seg_off, d_seg_cnt, n_regions, seg_cap)
    if (ahead && n_src == 1) { if (gmul >= 6) K3A_LAUNCH(6); else if (gmul == 5) K3A_LAUNCH(5); else K3A_LAUNCH(4); }
    else if (gmul >= 7) { if (rotate) K3_LAUNCH(8, true); else K3_LAUNCH(8, false); }
    else { if (rotate) K3_LAUNCH(6, true); else K3_LAUNCH(6, false); }
#undef K3_LAUNCH
#undef K3A_LAUNCH
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
