// multisplit.cu - K2b / K2c as a REGISTER-HELD multisplit.
//
// The counting sort K2a uses (tile_sort.cuh) stages every record in shared memory in natural order, ranks it with a second
// shared-memory atomic, builds a permutation and gathers through it: ~137 thread-instructions per record in ncu
// (profiles/r2f_*), 40 each for the permutation and the gather.  K2a needs that because its records come out of 16 unrolled
// positions of k-mer arithmetic; K2b and K2c only RE-READ records, so a thread can simply keep its 8 records in registers:
//
//   1. load 8 records (coalesced), bucket id from the hash, rank = atomicAdd(hist[bucket], 1)  - the histogram atomic
//      already hands out the rank inside the tile's run
//   2. scan the histogram, reserve room in every output bucket with one global atomicAdd per bucket per tile
//   3. store every record at ITS SORTED POSITION of a 16-byte shared-memory staging (offset[bucket] + rank), the bucket
//      id riding in the spare bits of the increment word
//   4. copy the staging out linearly: consecutive threads write consecutive records of a run
//
// ~45 thread-instructions per record, and the only shared memory is the staging itself (16 B per record of the tile).
// Tile = 8 records x 256 / 512 / 1024 threads, chosen so that a bucket's run in a tile is >= 8-16 records.
#include "tile_sort.cuh"

namespace {

constexpr int MS_RPT = 8;
constexpr int MS_MAX_PARTS = 1024;

template <int T>
__global__ void __launch_bounds__(T, T == 256 ? 3 : (T == 512 ? 2 : 1))
k2x_multisplit(PgMultiSplit a) {
    constexpr int TILE = T * MS_RPT;
    extern __shared__ __align__(16) unsigned char smem[];
    const int fan = 1 << a.bits;
    uint4 *s_sorted = reinterpret_cast<uint4 *>(smem);                                   // TILE records
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_sorted + TILE);   // fan: where the tile's run starts in its bucket
    uint4 **s_dst = reinterpret_cast<uint4 **>(s_base + fan);                             // fan: address of sorted position 0
    uint32_t *s_tile0 = reinterpret_cast<uint32_t *>(s_dst + fan);                        // n_seg + 1 prefix of tiles per segment
    uint32_t *s_hist = s_tile0 + a.n_seg + 1;                                             // fan
    uint32_t *s_off = s_hist + fan;                                                       // fan: exclusive offsets in the tile
    uint32_t *s_end = s_off + fan;                                                        // fan: end of the part of the run that fits the bucket
    __shared__ uint32_t s_chunk[32];
    __shared__ uint32_t s_nrec;
    const uint64_t pol = pg_policy_evict_first();
    const int lane = threadIdx.x & 31;

    auto seg_count = [&](int s) -> long long {
        long long c = (long long)a.seg_cnt[s];
        const long long cap = s == a.pass_seg ? a.pass_cap : a.seg_cap;
        if (c > cap) {
            // K2b: the sender's wire bucket overflowed.  K2c: a bucket's surplus went to the spill (fine), but a spill above
            // ITS capacity dropped records
            if ((a.lost_on_clamp || s == a.pass_seg) && blockIdx.x == 0 && threadIdx.x == 0 && a.stats)
                atomicExch(reinterpret_cast<unsigned long long *>(a.stats + PG_STAT_LOST), 1ull);
            c = cap;
        }
        return c < 0 ? 0 : c;
    };
    // ---- tiles per segment and their prefix
    for (int s = threadIdx.x; s < a.n_seg; s += T) s_tile0[s + 1] = (uint32_t)((seg_count(s) + TILE - 1) / TILE);
    if (threadIdx.x == 0) s_tile0[0] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t carry = 0;
        for (int base = 0; base < a.n_seg; base += 32) {
            const int i = base + threadIdx.x;
            uint32_t v = i < a.n_seg ? s_tile0[i + 1] : 0, inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
            if (i < a.n_seg) s_tile0[i + 1] = carry + inc;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    const uint32_t n_tiles = s_tile0[a.n_seg];
    uint4 *const out_spill = a.out + a.n_out * a.out_part_cap;
    unsigned long long *const out_spill_count = a.out_counts + a.n_out;
    const int shift = 64 - a.skip_bits - a.bits;

    int seg = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // a CTA's tiles only move forward through the segments: a short linear walk instead of a search
        while (s_tile0[seg + 1] <= tile) seg++;
        const long long i0 = (long long)(tile - s_tile0[seg]) * TILE;
        const long long left = seg_count(seg) - i0;
        const uint4 *src = a.in + (a.seg_off ? a.seg_off[seg] : (int64_t)seg * a.seg_cap) + i0;
        if (seg == a.pass_seg) {
            // records an earlier pass could not bucket: hand them on to the output spill (rare, scattered)
#pragma unroll 1
            for (int q = 0; q < MS_RPT; q++) {
                const int slot = q * T + threadIdx.x;
                if (slot < left) {
                    const uint4 r = pg_ld_stream_l2first(src + slot, pol);
                    const unsigned long long at = atomicAdd(out_spill_count, 1ull);
                    if ((int64_t)at < a.out_spill_cap) out_spill[at] = r;
                }
            }
            continue;
        }
        __syncthreads();                              // the previous tile's copy-out is done with the staging and the tables
        for (int i = threadIdx.x; i < fan; i += T) s_hist[i] = 0;
        __syncthreads();
        // ---- 1. load, bucket, rank
        uint4 rec[MS_RPT];
        uint32_t pr[MS_RPT];                          // bucket | rank << 10
#pragma unroll
        for (int q = 0; q < MS_RPT; q++) {
            const int slot = q * T + threadIdx.x;
            if (slot < left) rec[q] = pg_ld_stream_l2first(src + slot, pol);
        }
#pragma unroll
        for (int q = 0; q < MS_RPT; q++) {
            const int slot = q * T + threadIdx.x;
            pr[q] = 0xFFFFFFFFu;
            if (slot < left) {
                const uint64_t key = (uint64_t)rec[q].x | ((uint64_t)rec[q].y << 32);
                const uint32_t pid = (uint32_t)(pg_mix64(key) >> shift) & (uint32_t)(fan - 1);
                pr[q] = pid | (atomicAdd(&s_hist[pid], 1u) << 10);
            }
        }
        __syncthreads();
        // ---- 2. offsets in the tile, room in the output buckets
        unsigned long long *const counts = a.out_counts + (a.sliced ? ((int64_t)seg << a.bits) : 0);
        uint4 *const obase = a.out + (a.sliced ? ((int64_t)seg << a.bits) : 0) * a.out_part_cap;
        for (int base = 0; base < fan; base += T) {
            const int i = base + threadIdx.x;
            const uint32_t h = i < fan ? s_hist[i] : 0;
            if (i < fan) {
                const unsigned long long b = h ? atomicAdd(counts + i, (unsigned long long)h) : 0ull;
                s_base[i] = b;
            }
            uint32_t inc = h;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
            if (i < fan) s_off[i] = inc - h;
            if (lane == 31) s_chunk[i >> 5] = inc;            // chunks past `fan` hold 0
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int nchunk = (fan + 31) >> 5;
            uint32_t v = (int)threadIdx.x < nchunk ? s_chunk[threadIdx.x] : 0, inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
            s_chunk[threadIdx.x] = inc - v;
            if (threadIdx.x == 31) s_nrec = inc;
        }
        __syncthreads();
        // per bucket: first sorted position of its run (s_off), one past the last position that still fits the bucket
        // (s_end), and the address sorted position 0 would map to (s_dst): record at sorted position p -> s_dst[bucket][p]
        for (int i = threadIdx.x; i < fan; i += T) {
            const uint32_t off = s_off[i] + s_chunk[i >> 5], h = s_hist[i];
            const int64_t b = (int64_t)s_base[i];
            int64_t room = a.out_part_cap - b;
            if (room < 0) room = 0;
            s_off[i] = off;
            s_end[i] = off + (room < (int64_t)h ? (uint32_t)room : h);
            s_dst[i] = obase + (int64_t)i * a.out_part_cap + (b - (int64_t)off);
        }
        __syncthreads();
        // ---- 3. every record to its sorted position
#pragma unroll
        for (int q = 0; q < MS_RPT; q++) {
            if (pr[q] == 0xFFFFFFFFu) continue;
            const uint32_t pid = pr[q] & 1023u;
            const uint32_t p = s_off[pid] + (pr[q] >> 10);
            s_sorted[p] = make_uint4(rec[q].x, rec[q].y, rec[q].z, (rec[q].w & 0xFFu) | (pid << 8));
            if (p >= s_end[pid] && a.out_spill_cap > 0) {      // rare: the bucket is full - one scattered store into the spill
                const unsigned long long at = atomicAdd(out_spill_count, 1ull);
                if ((int64_t)at < a.out_spill_cap) out_spill[at] = rec[q];
            }
        }
        __syncthreads();
        // ---- 4. linear copy-out: consecutive threads, consecutive records of a run
        const uint32_t nrec = s_nrec;
        for (uint32_t p0 = threadIdx.x; p0 < nrec; p0 += 4 * T) {
            uint4 r[4];
#pragma unroll
            for (int j = 0; j < 4; j++) if (p0 + j * T < nrec) r[j] = s_sorted[p0 + j * T];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t p = p0 + j * T;
                if (p >= nrec) continue;
                const uint32_t pid = r[j].w >> 8;
                if (p < s_end[pid]) pg_st_stream_l2first(s_dst[pid] + p, make_uint4(r[j].x, r[j].y, r[j].z, r[j].w & 0xFFu), pol);
            }
        }
    }
}

template <int T>
int launch(const PgMultiSplit &a, cudaStream_t st) {
    const int fan = 1 << a.bits;
    const int tile = T * MS_RPT;
    const int smem = tile * 16 + fan * 16 + (a.n_seg + 1) * 4 + 3 * fan * 4 + 16;
    int64_t max_tiles = 0;
    max_tiles = (int64_t)(a.n_seg - (a.pass_seg >= 0 ? 1 : 0)) * ((a.seg_cap + tile - 1) / tile) + (a.pass_seg >= 0 ? (a.pass_cap + tile - 1) / tile : 0);
    int per_sm = 227 * 1024 / (smem + 1024 + 256);
    const int cap_sm = T == 256 ? 3 : (T == 512 ? 2 : 1);
    if (per_sm > cap_sm) per_sm = cap_sm;
    if (per_sm < 1) return pg_fail(PG_ERR_INVALID, "multisplit: %d bytes of shared memory do not fit an SM", smem);
    int64_t maxg = (int64_t)pg_num_sms() * per_sm;
    int grid = (int)(max_tiles < maxg ? max_tiles : maxg);
    if (grid < 1) grid = 1;
    PG_CUDA(cudaFuncSetAttribute(k2x_multisplit<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k2x_multisplit<T><<<grid, T, smem, st>>>(a);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

}  // namespace

int pg_multisplit_launch(const PgMultiSplit &a, cudaStream_t st) {
    if (a.bits < 0 || (1 << a.bits) > MS_MAX_PARTS || a.n_seg < 1 || a.skip_bits < 0 || a.skip_bits + a.bits > 40)
        return pg_fail(PG_ERR_INVALID, "multisplit: bad geometry");
    // tile size by fan-out: runs of >= 8-16 records per bucket and tile
    static int thr_env = -1;
    if (thr_env < 0) { const char *e = getenv("PG_SPLIT_THREADS"); thr_env = e ? atoi(e) : 0; }
    const int fan = 1 << a.bits;
    int threads = fan <= 64 ? 256 : (fan <= 256 ? 512 : 1024);
    if (thr_env == 256 || thr_env == 512 || thr_env == 1024) threads = thr_env;
    static int pol_env = -1;
    if (pol_env < 0) { const char *e = getenv("PG_SPLIT_POLICY"); pol_env = e ? atoi(e) : 0; }
    PgMultiSplit b = a;
    b.policy = pol_env;
    if (threads == 256) return launch<256>(b, st);
    if (threads == 512) return launch<512>(b, st);
    return launch<1024>(b, st);
}

extern "C" int pg_records_resplit(const uint64_t *d_in, const int64_t *d_in_counts, int in_bits, int64_t in_part_cap, int64_t in_spill_cap,
                                  int bits, uint64_t *d_out, int64_t *d_out_counts, int64_t out_part_cap, int64_t out_spill_cap,
                                  int64_t *d_table_stats, pg_stream_t stream_) {
    if (!d_in || !d_in_counts || !d_out || !d_out_counts || in_bits < 0 || in_bits > 13 || bits < 1 || bits > 8 || in_part_cap < 1 ||
        in_spill_cap < 0 || out_part_cap < 1 || out_spill_cap < 0 || ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15))
        return pg_fail(PG_ERR_INVALID, "pg_records_resplit: bad arguments (in_bits 0..13, bits 1..8, 16-byte aligned records)");
    cudaStream_t st = (cudaStream_t)stream_;
    const int64_t n_in = 1ll << in_bits, n_out = n_in << bits;
    PG_CUDA(cudaMemsetAsync(d_out_counts, 0, (size_t)(n_out + 1) * 8, st));
    PgMultiSplit m;
    m.in = reinterpret_cast<const uint4 *>(d_in); m.seg_off = nullptr; m.seg_cnt = reinterpret_cast<const unsigned long long *>(d_in_counts);
    m.n_seg = (int)n_in + 1; m.seg_cap = in_part_cap; m.pass_seg = (int)n_in; m.pass_cap = in_spill_cap; m.lost_on_clamp = 0;
    m.out = reinterpret_cast<uint4 *>(d_out); m.out_counts = reinterpret_cast<unsigned long long *>(d_out_counts);
    m.out_part_cap = out_part_cap; m.out_spill_cap = out_spill_cap; m.n_out = n_out;
    m.skip_bits = in_bits; m.bits = bits; m.sliced = 1; m.stats = d_table_stats; m.policy = 0;
    return pg_multisplit_launch(m, st);
}
