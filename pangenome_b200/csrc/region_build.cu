// region_build.cu - K2c + K3s: the dBG table built region by region in SHARED MEMORY.
//
// Why: the atomic K3 (partition.cu) resolves every update record with a 16-byte L2 load plus one or two L2 atomics on
// a random slot; ncu and the slot micro-benchmark (profiles/r2b_microbench.jsonl) put it AT the chip's random-atomic
// ceiling (~50 G records/s) with 2.8 GB of DRAM traffic for 1.6 GB of algorithmic bytes.  The way past that ceiling
// is to make the piece of table under construction fit on chip:
//
//   K2a  (partition.cu)  records -> 2^cb coarse buckets by the TOP cb bits of mix64(key)          (as before)
//   K2c  (here)          every coarse bucket -> 2^fb fine buckets by the NEXT fb bits: one bucket per table REGION of
//                        2^RB slots.  Same shared-memory counting sort as K2a/K2b; a tile belongs to one coarse
//                        bucket, so a 4096-record tile leaves 4096 / 2^fb-record runs (2 KB at fb = 5)
//   K3s  (here)          one CTA per region: the region's slots live in shared memory (64 KB at RB = 12), the bucket
//                        is streamed once with coalesced 16-byte loads, keys are claimed with a 64-bit shared-memory
//                        CAS and merged with shared-memory atomicOr / atomicAdd, and the finished region is written
//                        to HBM once, sequentially, every slot exactly one 16-byte store
//   spill                records that overflowed a bucket (hash skew: poly-A, satellites) sit in the set's spill
//                        bucket and are upserted afterwards with the L2 atomics of the classic path
//
// The table this produces probes INSIDE a region (pg_table.region_bits = RB, table_dev.cuh tv_next), so a region holds
// every key that hashes into it and later upserts (spill, further rounds) find them.  HBM traffic per build: the
// record stream once more through K2c (32 B/record) + 16 B/record read and 16 B/slot written by K3s - all of it
// streaming.  A later round (FIRST = false) starts a region from what HBM holds instead of from empty.
// Replaces oakht.__setitem__ / resize (kmer_numba.py:423-474, 540-561) on the build path.
#include "tile_sort.cuh"

namespace {

struct RefineArgs {
    const uint4 *in; const unsigned long long *in_counts; int n_seg; int64_t seg_cap, in_spill_cap;     // coarse set: n_seg buckets + spill
    uint4 *out; unsigned long long *out_counts; int64_t out_part_cap, out_spill_cap; int coarse_bits, fine_bits;
    int64_t *stats;
};

// 16-byte shared-memory load that is re-issued every time (the slots change under the other threads' atomics)
__device__ __forceinline__ uint4 lds128(const void *p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return r;
}

__device__ __forceinline__ void raise_lost(int64_t *stats) {
    if (stats) atomicExch(reinterpret_cast<unsigned long long *>(stats + PG_STAT_LOST), 1ull);
}

template <int T>
__global__ void __launch_bounds__(T, T == 256 ? 3 : 1)
k2c_refine(RefineArgs a) {
    constexpr int TILE = T * KP_G;
    extern __shared__ __align__(16) unsigned char smem[];
    const int fan = 1 << a.fine_bits;
    TileSort<TILE> ts;
    ts.carve(smem, fan);
    long long *s_tile0 = reinterpret_cast<long long *>(smem + ((TileSort<TILE>::bytes(fan) + 15) & ~15));       // n_seg + 2 entries
    __shared__ uint32_t s_nrec;
    __shared__ uint32_t s_chunk[32];
    const uint64_t pol = pg_policy_evict_first();
    const int nseg1 = a.n_seg + 1;                  // the coarse spill is one more segment
    // ---- tiles per segment, then their exclusive prefix: s_tile0[s] = first tile of segment s
    for (int s = threadIdx.x; s < nseg1; s += T) {
        long long c = (long long)a.in_counts[s];
        const long long cap = s < a.n_seg ? a.seg_cap : a.in_spill_cap;
        if (c > cap) {        // a bucket's surplus went to the spill; a spill above ITS capacity dropped records
            if (s == a.n_seg && blockIdx.x == 0) raise_lost(a.stats);
            c = cap;
        }
        s_tile0[s + 1] = (c + TILE - 1) / TILE;
    }
    if (threadIdx.x == 0) s_tile0[0] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long carry = 0;
        for (int base = 0; base < nseg1; base += 32) {
            const int i = base + threadIdx.x;
            long long v = i < nseg1 ? s_tile0[i + 1] : 0, inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += y; }
            if (i < nseg1) s_tile0[i + 1] = carry + inc;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    const long long n_tiles = s_tile0[nseg1];
    const int64_t n_fine = (int64_t)a.n_seg << a.fine_bits;
    uint4 *const out_spill = a.out + n_fine * a.out_part_cap;
    unsigned long long *const out_spill_count = a.out_counts + n_fine;
    const int shift = 64 - a.coarse_bits - a.fine_bits;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int lo = 1, hi = nseg1;                      // first index in [1, nseg1] whose prefix exceeds the tile
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_tile0[mid] <= tile) lo = mid + 1; else hi = mid; }
        const int seg = lo - 1;
        long long cnt = (long long)a.in_counts[seg];
        const long long cap = seg < a.n_seg ? a.seg_cap : a.in_spill_cap;
        if (cnt > cap) cnt = cap;
        const long long i0 = (tile - s_tile0[seg]) * TILE;
        const uint4 *src = a.in + (int64_t)seg * a.seg_cap + i0;        // the spill follows the last bucket: same formula
        const long long left = cnt - i0;
        if (seg == a.n_seg) {
            // records the coarse pass could not bucket: pass them on to the fine spill (rare, scattered)
#pragma unroll 1
            for (int q = 0; q < KP_G; q++) {
                const int slot = q * T + threadIdx.x;
                if (slot < left) {
                    const uint4 r = pg_ld_stream_l2first(src + slot, pol);
                    const unsigned long long at = atomicAdd(out_spill_count, 1ull);
                    if ((int64_t)at < a.out_spill_cap) out_spill[at] = r;
                }
            }
            continue;
        }
        __syncthreads();
        ts.reset(fan, T);
        __syncthreads();
#pragma unroll
        for (int h = 0; h < KP_G; h += 8) {
            uint4 r[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int slot = (h + q) * T + threadIdx.x;
                if (slot < left) r[q] = pg_ld_stream_l2first(src + slot, pol);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int slot = (h + q) * T + threadIdx.x;
                if (slot < left) {
                    const uint64_t key = (uint64_t)r[q].x | ((uint64_t)r[q].y << 32);
                    const uint32_t pid = (uint32_t)(pg_mix64(key) >> shift) & (uint32_t)(fan - 1);
                    ts.emit(slot, pid, key, r[q].z, r[q].w);
                } else {
                    ts.s_pid[slot] = NOREC;
                }
            }
        }
        __syncthreads();
        BucketOut o;
        o.records = a.out + ((int64_t)seg << a.fine_bits) * a.out_part_cap; o.part_cap = a.out_part_cap; o.spill_cap = a.out_spill_cap;
        o.part_counts = a.out_counts + ((int64_t)seg << a.fine_bits); o.n_parts = fan; o.sub_bits = a.fine_bits;
        o.spill = out_spill; o.spill_count = out_spill_count; o.peers = nullptr; o.my_rank = 0;
        ts.template sort_write<T>(o, s_chunk, &s_nrec, pol);
    }
}

// ---- K3s ------------------------------------------------------------------------------------------------
struct RegionArgs {
    TableView t; const uint4 *records; const unsigned long long *counts; int64_t part_cap, spill_cap; int n_regions;
};

// THREADS x RPT records are in flight per CTA batch.  512 threads x 2 records: the shared-memory table (64 KB) limits an SM
// to 3 CTAs, so the warps that hide the LDS -> compare -> CAS dependency chains have to come from wider CTAs.
template <int RB, bool FIRST, int THREADS, int RPT>
__global__ void __launch_bounds__(THREADS, RB == 12 ? 3 : 4)
k3s_region_build(RegionArgs a) {
    constexpr int NS = 1 << RB;
    constexpr uint32_t CNT_MAX = (1u << 22) - 1u;
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *s_key = reinterpret_cast<uint64_t *>(smem);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_key + NS);
    uint32_t *s_cnt = s_mask + NS;
    const TableView &t = a.t;
    uint4 *const slots = reinterpret_cast<uint4 *>(t.slots);
    uint32_t n_claimed = 0;

    // what slot s of region r starts from: empty on the first round, else whatever earlier rounds left in HBM
    auto init_slot = [&](int64_t r, int s) {
        uint64_t key = PG_EMPTY; uint32_t m = 0, c = 0;
        if (!FIRST) {
            const uint4 g = slots[r * NS + s];
            const uint64_t hi = (uint64_t)g.z | ((uint64_t)g.w << 32);
            if ((hi & ~PG_VAL_MASK) == t.tag) { key = (uint64_t)g.x | ((uint64_t)g.y << 32); m = g.z; c = g.w & CNT_MAX; }
        }
        s_key[s] = key; s_mask[s] = m; s_cnt[s] = c;
    };

    int64_t r = blockIdx.x;
    if (r < a.n_regions)
        for (int s = threadIdx.x; s < NS; s += THREADS) init_slot(r, s);
    __syncthreads();
    for (; r < a.n_regions; r += gridDim.x) {
        // ---- the region's records: coalesced 16-byte loads, RPT in flight per thread
        unsigned long long c64 = a.counts[r];
        if (c64 > (unsigned long long)a.part_cap) {     // the surplus sits in the spill; without one it was dropped
            if (a.spill_cap <= 0 && threadIdx.x == 0) raise_lost(t.stats);
            c64 = (unsigned long long)a.part_cap;
        }
        const uint32_t c = (uint32_t)c64;
        const uint4 *src = a.records + r * a.part_cap;
        for (uint32_t base = 0; base < c; base += RPT * THREADS) {
            uint4 rec[RPT];
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                const uint32_t i = base + j * THREADS + threadIdx.x;
                if (i < c) rec[j] = pg_ld_stream(src + i);
            }
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                // One probe step looks at a GROUP of four consecutive slots (two 16-byte shared-memory loads): the probe
                // sequences of the 32 lanes of a warp differ in length and the warp pays for the longest, so what matters
                // is the tail - at load 0.5 the longest of 32 sequences is ~2 groups against ~6-9 single slots.  Match and
                // free-slot positions come from bit masks (no branch per slot), every lane that claims goes through ONE
                // CAS site, and the warp reconverges before the merge so the two atomics issue once per batch.
                const bool act = base + j * THREADS + threadIdx.x < c;
                const uint32_t klo = rec[j].x, khi = rec[j].y;
                const uint64_t key = (uint64_t)klo | ((uint64_t)khi << 32);
                uint32_t g = (uint32_t)(pg_mix64(key) >> t.shift) & (NS - 1) & ~(uint32_t)(PG_REGION_GROUP - 1);
                int s = -1;
                if (act) {
                    for (int probe = 0; probe < NS / PG_REGION_GROUP;) {
                        const uint4 a = lds128(s_key + g), b = lds128(s_key + g + 2);
                        const uint32_t mm = (uint32_t)(a.x == klo && a.y == khi) | ((uint32_t)(a.z == klo && a.w == khi) << 1) |
                                            ((uint32_t)(b.x == klo && b.y == khi) << 2) | ((uint32_t)(b.z == klo && b.w == khi) << 3);
                        if (mm) { s = (int)g + __ffs(mm) - 1; break; }
                        // free slots hold PG_EMPTY; a key's high word is never 0xFFFFFFFF (base-5 codes stay below 2^63)
                        const uint32_t em = (uint32_t)(a.y == 0xFFFFFFFFu) | ((uint32_t)(a.w == 0xFFFFFFFFu) << 1) |
                                            ((uint32_t)(b.y == 0xFFFFFFFFu) << 2) | ((uint32_t)(b.w == 0xFFFFFFFFu) << 3);
                        if (em) {
                            const int e = (int)g + __ffs(em) - 1;
                            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(s_key + e), (unsigned long long)PG_EMPTY,
                                                                     (unsigned long long)key);
                            if (old == PG_EMPTY) { n_claimed++; s = e; break; }
                            if (old == key) { s = e; break; }
                            continue;             // another key took that slot first: look at the group again
                        }
                        g = (g + PG_REGION_GROUP) & (NS - 1); probe++;
                    }
                    if (s < 0)                // the region is full: the table is too small for this input
                        atomicExch(reinterpret_cast<unsigned long long *>(t.stats + PG_STAT_OVERFLOW), 1ull);
                }
                __syncwarp();
                if (s >= 0) {
                    atomicOr(s_mask + s, rec[j].z);
                    atomicAdd(s_cnt + s, rec[j].w);
                }
            }
        }
        __syncthreads();
        // ---- write the region out (one 16-byte store per slot, consecutive threads consecutive slots) and start
        // the next one: a slot is re-initialised by the thread that just read it, no barrier in between
        const int64_t nxt = r + gridDim.x;
        for (int s = threadIdx.x; s < NS; s += THREADS) {
            const uint64_t key = s_key[s];
            uint4 g = make_uint4(0u, 0u, 0u, 0u);
            if (key != PG_EMPTY) {
                const uint32_t cnt = s_cnt[s];
                g = make_uint4((uint32_t)key, (uint32_t)(key >> 32), s_mask[s], (cnt < CNT_MAX ? cnt : CNT_MAX) | (uint32_t)(t.tag >> 32));
            }
            slots[r * NS + s] = g;
            if (nxt < a.n_regions) init_slot(nxt, s);
        }
        __syncthreads();
    }
    publish_claims(t, n_claimed);
}

// The spill bucket: upserts with L2 atomics, probing confined to the regions (TableView::rmask).
__global__ void __launch_bounds__(256)
k3s_spill_insert(TableView t, const uint4 *__restrict__ spill, const unsigned long long *__restrict__ count, int64_t cap) {
    unsigned long long c = *count;
    if (c > (unsigned long long)cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) raise_lost(t.stats);
        c = (unsigned long long)cap;
    }
    uint32_t n_claimed = 0;
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < c; i += gridDim.x * 256ull) {
        const uint4 r = pg_ld_stream(spill + i);
        table_upsert(t, (uint64_t)r.x | ((uint64_t)r.y << 32), r.z, r.w, n_claimed);
    }
    publish_claims(t, n_claimed);
}

template <int RB, int THREADS, int RPT>
int launch_regions_t(const RegionArgs &ra, bool first, cudaStream_t st) {
    constexpr int smem = (1 << RB) * 16;
    const int per_sm = RB == 12 ? 3 : (THREADS == 256 ? 6 : 4);
    int64_t maxg = (int64_t)pg_num_sms() * per_sm;
    const int grid = (int)(ra.n_regions < maxg ? ra.n_regions : maxg);
    if (first) {
        PG_CUDA(cudaFuncSetAttribute(k3s_region_build<RB, true, THREADS, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k3s_region_build<RB, true, THREADS, RPT><<<grid, THREADS, smem, st>>>(ra);
    } else {
        PG_CUDA(cudaFuncSetAttribute(k3s_region_build<RB, false, THREADS, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k3s_region_build<RB, false, THREADS, RPT><<<grid, THREADS, smem, st>>>(ra);
    }
    return PG_OK;
}
template <int RB>
int launch_regions(const RegionArgs &ra, bool first, cudaStream_t st) {
    static int thr = -1;
    if (thr < 0) { const char *e = getenv("PG_K3S_THREADS"); thr = e ? atoi(e) : 512; }
    if (thr == 256) return launch_regions_t<RB, 256, 4>(ra, first, st);
    return launch_regions_t<RB, 512, 2>(ra, first, st);
}

}  // namespace

extern "C" int pg_records_refine(const pg_bucket_set *coarse, int fine_bits, uint64_t *d_fine_records, int64_t *d_fine_counts,
                                 int64_t fine_part_cap, int64_t fine_spill_cap, int64_t *d_table_stats, pg_stream_t stream_) {
    BucketOut co;
    int rc = make_bucket_out(coarse, "pg_records_refine", co); if (rc) return rc;
    if (coarse->d_peer_bases || coarse->owner_bits != 0) return pg_fail(PG_ERR_INVALID, "pg_records_refine: the coarse set must be local with owner_bits 0");
    if (fine_bits < 1 || fine_bits > 8 || !d_fine_records || !d_fine_counts || fine_part_cap < 1 || fine_spill_cap < 0 ||
        (reinterpret_cast<uintptr_t>(d_fine_records) & 15))
        return pg_fail(PG_ERR_INVALID, "pg_records_refine: bad fine bucket set (fine_bits 1..8, 16-byte aligned records)");
    cudaStream_t st = (cudaStream_t)stream_;
    const int64_t n_fine = (int64_t)co.n_parts << fine_bits;
    PG_CUDA(cudaMemsetAsync(d_fine_counts, 0, (size_t)(n_fine + 1) * 8, st));
    static int v1 = -1;
    if (v1 < 0) { const char *e = getenv("PG_SPLIT_V1"); v1 = e ? atoi(e) : 0; }
    if (!v1) {
        PgMultiSplit m;
        m.in = co.records; m.seg_off = nullptr; m.seg_cnt = co.part_counts; m.n_seg = co.n_parts + 1; m.seg_cap = co.part_cap;
        m.pass_seg = co.n_parts; m.pass_cap = co.spill_cap; m.lost_on_clamp = 0;
        m.out = reinterpret_cast<uint4 *>(d_fine_records); m.out_counts = reinterpret_cast<unsigned long long *>(d_fine_counts);
        m.out_part_cap = fine_part_cap; m.out_spill_cap = fine_spill_cap; m.n_out = n_fine;
        m.skip_bits = coarse->sub_bits; m.bits = fine_bits; m.sliced = 1; m.stats = d_table_stats;
        return pg_multisplit_launch(m, st);
    }
    RefineArgs a;
    a.in = co.records; a.in_counts = co.part_counts; a.n_seg = co.n_parts; a.seg_cap = co.part_cap; a.in_spill_cap = co.spill_cap;
    a.out = reinterpret_cast<uint4 *>(d_fine_records); a.out_counts = reinterpret_cast<unsigned long long *>(d_fine_counts);
    a.out_part_cap = fine_part_cap; a.out_spill_cap = fine_spill_cap; a.coarse_bits = coarse->sub_bits; a.fine_bits = fine_bits; a.stats = d_table_stats;
    const int fan = 1 << fine_bits;
    const int smem = ((TileSort<256 * KP_G>::bytes(fan) + 15) & ~15) + (co.n_parts + 2) * 8;
    const int tile = 256 * KP_G;
    const int64_t max_tiles = (int64_t)co.n_parts * ((co.part_cap + tile - 1) / tile) + (co.spill_cap + tile - 1) / tile;
    int64_t maxg = (int64_t)pg_num_sms() * ctas_per_sm(smem, 1024);
    int grid = (int)(max_tiles < maxg ? max_tiles : maxg);
    if (grid < 1) grid = 1;
    PG_CUDA(cudaFuncSetAttribute(k2c_refine<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k2c_refine<256><<<grid, 256, smem, st>>>(a);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_region_build(const pg_table *t, const uint64_t *d_records, const int64_t *d_counts, int64_t part_cap, int64_t spill_cap,
                               int first_round, pg_stream_t stream_) {
    if (!t || !t->d_slots || !t->d_stats || t->capacity < 2 || (t->capacity & (t->capacity - 1)) || t->epoch < 1 || t->epoch > PG_EPOCH_MAX)
        return pg_fail(PG_ERR_INVALID, "pg_region_build: bad table");
    if (t->region_bits != 8 && t->region_bits != 12)
        return pg_fail(PG_ERR_INVALID, "pg_region_build: pg_table.region_bits must be 8 or 12 (slots per shared-memory region: 256 / 4096)");
    int bits = 0; while ((1ll << bits) < t->capacity) bits++;
    if (bits < t->region_bits) return pg_fail(PG_ERR_INVALID, "pg_region_build: the table is smaller than one region");
    if (bits - t->region_bits > 30) return pg_fail(PG_ERR_INVALID, "pg_region_build: too many regions");
    if (!d_records || !d_counts || part_cap < 1 || part_cap > 0x7FFFFFFF || spill_cap < 0 || (reinterpret_cast<uintptr_t>(d_records) & 15))
        return pg_fail(PG_ERR_INVALID, "pg_region_build: bad bucket set (16-byte aligned records, part_cap 1..2^31-1)");
    cudaStream_t st = (cudaStream_t)stream_;
    const int64_t n_regions = 1ll << (bits - t->region_bits);
    TableView tv = make_view(t);
    RegionArgs ra;
    ra.t = tv; ra.n_regions = (int)n_regions;
    ra.records = reinterpret_cast<const uint4 *>(d_records); ra.counts = reinterpret_cast<const unsigned long long *>(d_counts);
    ra.part_cap = part_cap; ra.spill_cap = spill_cap;
    int rc = t->region_bits == 12 ? launch_regions<12>(ra, first_round != 0, st) : launch_regions<8>(ra, first_round != 0, st);
    if (rc) return rc;
    PG_CUDA(cudaGetLastError());
    if (spill_cap > 0) {
        k3s_spill_insert<<<pg_num_sms(), 256, 0, st>>>(tv, ra.records + n_regions * part_cap, ra.counts + n_regions, spill_cap);
        PG_CUDA(cudaGetLastError());
    }
    return PG_OK;
}
