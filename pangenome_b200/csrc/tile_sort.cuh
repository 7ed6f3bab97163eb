// tile_sort.cuh - the shared-memory counting sort of one CTA tile of update records and the bucket sets it writes to:
// shared by K2a (partition.cu: records computed from the sequence), K2b (records re-read from the owners' wire segments)
// and K2c (region_build.cu: coarse hash-prefix buckets refined into one bucket per shared-memory table region).
#pragma once
#include <stdlib.h>
#include <string.h>
#include "table_dev.cuh"

namespace {

constexpr int KP_G = 16;                               // positions (K2a) / records (K2b) per thread
// CTA tile = THREADS x 16 positions: 256 threads (4096 positions, 75 KB of smem, 3 CTAs/SM) for local
// buckets; 512 threads (8192 positions, one CTA per SM) when buckets are peer memory - twice
// the run length per bucket on NVLink (2 GPUs: K2a + exchange 0.85 ms against 0.92 ms with 256 threads)
constexpr int KP_MAX_PARTS = 1024;
constexpr uint16_t NOREC = 0xFFFFu;

// Where a tile's sorted records go.  Bucket i holds part_cap records; local bucket sets carry one more
// bucket of spill_cap records at index n_parts (the SPILL): a record whose bucket is full goes there instead
// of being dropped, so hash skew (one k-mer making up a visible share of the input: poly-A, satellites) costs
// nothing but a few scattered stores.  part_counts[i] counts everything PRODUCED for bucket i (it may exceed
// part_cap: the surplus went to the spill), part_counts[n_parts] what was offered to the spill.
struct BucketOut {
    uint4 *records; int64_t part_cap, spill_cap; unsigned long long *part_counts; int n_parts, sub_bits;
    // the spill bucket and its counter: records + n_parts * part_cap / part_counts + n_parts for a plain local set; K2c points
    // every coarse bucket's slice of the fine set (its own records / part_counts / n_parts) at the ONE spill of the whole set
    uint4 *spill; unsigned long long *spill_count;
    // fused exchange: when `peers` is set, bucket (owner, sub) is written straight into rank `owner`'s
    // receive buffer over NVLink (peer-mapped pointer), at the slice reserved for source rank `my_rank`
    uint4 *const *peers; int my_rank;
};

// ---- the counting sort of one CTA tile, shared by K2a (records computed from the sequence) and K2b
// (records re-read from coarse buckets).  Staging: MAXR records of 12 bytes (the key, and the masks with the
// increment (1 or 2) folded into bit 31 - shared memory is what limits the CTAs per SM), their bucket ids,
// the permutation, and per bucket: histogram, tile offset, ticket, room, global base, destination pointer.
template <int MAXR>
struct TileSort {
    uint64_t *s_key; uint32_t *s_mi; uint16_t *s_pid, *s_perm;
    uint32_t *s_hist, *s_off, *s_tick, *s_room; unsigned long long *s_base; uint4 **s_dst;
    __device__ __forceinline__ void carve(unsigned char *smem, int n_parts) {
        s_key = reinterpret_cast<uint64_t *>(smem);
        s_mi = reinterpret_cast<uint32_t *>(s_key + MAXR);
        s_pid = reinterpret_cast<uint16_t *>(s_mi + MAXR);             // MAXR bucket ids (NOREC = slot unused)
        s_perm = s_pid + MAXR;                                         // sorted index -> natural index
        s_hist = reinterpret_cast<uint32_t *>(s_perm + MAXR);          // n_parts
        s_off = s_hist + n_parts;                                      // n_parts: exclusive offsets
        s_tick = s_off + n_parts;                                      // n_parts: tickets
        s_room = s_tick + n_parts;                                     // n_parts: records of this tile the bucket can still take
        s_base = reinterpret_cast<unsigned long long *>(s_room + n_parts);
        s_dst = reinterpret_cast<uint4 **>(s_base + n_parts);          // n_parts: address of sorted position 0
    }
    __host__ __device__ static int bytes(int n_parts) { return MAXR * 12 + MAXR * 2 * 2 + 4 * n_parts * 4 + 2 * n_parts * 8 + 16; }
    __device__ __forceinline__ void emit(int slot, uint32_t pid, uint64_t key, uint32_t masks, uint32_t inc) const {
        s_key[slot] = key;
        s_mi[slot] = masks | ((inc - 1u) << 31);
        s_pid[slot] = (uint16_t)pid;
        atomicAdd(&s_hist[pid], 1u);
    }
    __device__ __forceinline__ uint4 record(uint32_t i) const {
        const uint64_t key = s_key[i];
        const uint32_t mi = s_mi[i];
        return make_uint4((uint32_t)key, (uint32_t)(key >> 32), mi & 0x7FFFFFFFu, 1u + (mi >> 31));
    }
    __device__ __forceinline__ void reset(int n_parts, int threads) const {
        for (int i = threadIdx.x; i < n_parts; i += threads) { s_hist[i] = 0; s_tick[i] = 0; }
    }

    // Phases 2-4 (after a barrier that follows the emits): reserve space with one global atomicAdd per bucket
    // per tile, scan the histogram, build the permutation, write coalesced runs.  All threads must call.
    template <int THREADS>
    __device__ __forceinline__ void sort_write(const BucketOut &o, uint32_t *s_chunk, uint32_t *s_nrec, uint64_t pol) const {
        const int n_parts = o.n_parts;
        {
            const int lane = threadIdx.x & 31;
            for (int base = 0; base < n_parts; base += THREADS) {
                const int i = base + threadIdx.x;
                uint32_t h = i < n_parts ? s_hist[i] : 0;
                if (i < n_parts) s_base[i] = h ? atomicAdd(o.part_counts + i, (unsigned long long)h) : 0ull;
                uint32_t inc = h;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
                if (i < n_parts) s_off[i] = inc - h;                         // offset inside its 32-bucket chunk
                if (lane == 31) s_chunk[i >> 5] = inc;                       // chunk total (chunks past n_parts hold 0)
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {       // exclusive scan of the <= 32 chunk totals
            const int nchunk = (n_parts + 31) >> 5;
            uint32_t v = (int)threadIdx.x < nchunk ? s_chunk[threadIdx.x] : 0, inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += y; }
            s_chunk[threadIdx.x] = inc - v;
            if (threadIdx.x == 31) *s_nrec = inc;
        }
        __syncthreads();
        // final offsets, and per bucket the address its sorted position 0 would map to: record at sorted
        // position p goes to s_dst[pid][p] (its run in the bucket starts at s_base, its run in the tile at s_off)
        for (int i = threadIdx.x; i < n_parts; i += THREADS) {
            const uint32_t off = s_off[i] + s_chunk[i >> 5];
            s_off[i] = off;
            uint4 *bucket;
            if (o.peers) {     // [source rank][sub][part_cap] in the owner's memory: the layout an all-to-all would produce
                const uint32_t owner = (uint32_t)i >> o.sub_bits, sub = (uint32_t)i & ((1u << o.sub_bits) - 1u);
                bucket = o.peers[owner] + (((int64_t)o.my_rank << o.sub_bits) + sub) * o.part_cap;
            } else {
                bucket = o.records + (int64_t)i * o.part_cap;
            }
            const unsigned long long base = s_base[i];
            s_dst[i] = bucket + ((int64_t)base - (int64_t)off);
            // records of this tile that still fit the bucket (the rest go to the spill; without one they are dropped
            // and the host sees the overflow in part_counts)
            const int64_t room = o.part_cap - (int64_t)base;
            s_room[i] = room <= 0 ? 0u : (room > 0xFFFF ? 0xFFFFu : (uint32_t)room);
        }
        __syncthreads();
        const uint32_t nrec = *s_nrec;
        if (nrec == 0) return;
        // ---- 3. permutation: sorted position -> natural index ------------------------------------
        for (uint32_t i = threadIdx.x; i < (uint32_t)MAXR; i += THREADS) {
            const uint32_t pid = s_pid[i];
            if (pid == NOREC) continue;
            const uint32_t rank = atomicAdd(&s_tick[pid], 1u);
            const bool fits = rank < s_room[pid];
            s_perm[s_off[pid] + rank] = fits ? (uint16_t)i : NOREC;
            if (!fits && o.spill_cap > 0) {            // rare: the bucket is full - one scattered store into the spill
                const unsigned long long at = atomicAdd(o.spill_count, 1ull);
                if ((int64_t)at < o.spill_cap) o.spill[at] = record(i);
            }
        }
        __syncthreads();
        // ---- 4. coalesced write-out: consecutive threads write consecutive records of one bucket; four
        // independent chains (perm -> bucket id -> address, record) per thread hide the shared-memory latency
        const bool remote = o.peers != nullptr;
        for (uint32_t o0 = threadIdx.x; o0 < nrec; o0 += 4 * THREADS) {
            uint32_t idx[4]; uint4 *dst[4]; uint4 rec[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t p = o0 + j * THREADS;
                idx[j] = p < nrec ? (uint32_t)s_perm[p] : (uint32_t)NOREC;
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (idx[j] != NOREC) {
                    dst[j] = s_dst[s_pid[idx[j]]] + (o0 + j * THREADS);
                    rec[j] = record(idx[j]);
                }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (idx[j] != NOREC) {
                    if (remote) *dst[j] = rec[j];
                    else pg_st_stream_l2first(dst[j], rec[j], pol);
                }
        }
    }
};

// Validate a caller's bucket set and turn it into the kernels' view.
static int make_bucket_out(const pg_bucket_set *b, const char *who, BucketOut &o) {
    if (!b || !b->d_part_counts || b->part_cap < 1 || b->spill_cap < 0 || b->owner_bits < 0 || b->owner_bits > 6 || b->sub_bits < 0 ||
        b->sub_bits > 10 || (1 << (b->owner_bits + b->sub_bits)) > KP_MAX_PARTS)
        return pg_fail(PG_ERR_INVALID, "%s: bad bucket set (owner_bits 0..6, sub_bits 0..10, at most %d buckets)", who, KP_MAX_PARTS);
    if (!b->d_records == !b->d_peer_bases) return pg_fail(PG_ERR_INVALID, "%s: exactly one of d_records / d_peer_bases must be set", who);
    if (b->d_peer_bases && (b->spill_cap != 0 || b->my_rank < 0 || b->my_rank >= (1 << b->owner_bits)))
        return pg_fail(PG_ERR_INVALID, "%s: peer bucket sets have no spill and need 0 <= my_rank < 2^owner_bits", who);
    if (b->d_records && (reinterpret_cast<uintptr_t>(b->d_records) & 15)) return pg_fail(PG_ERR_INVALID, "%s: records must be 16-byte aligned", who);
    o.records = reinterpret_cast<uint4 *>(b->d_records); o.part_cap = b->part_cap; o.spill_cap = b->spill_cap;
    o.part_counts = reinterpret_cast<unsigned long long *>(b->d_part_counts);
    o.n_parts = 1 << (b->owner_bits + b->sub_bits); o.sub_bits = b->sub_bits;
    o.spill = o.records ? o.records + (int64_t)o.n_parts * o.part_cap : nullptr; o.spill_count = o.part_counts + o.n_parts;
    o.peers = reinterpret_cast<uint4 *const *>(b->d_peer_bases); o.my_rank = b->my_rank;
    return PG_OK;
}

static int ctas_per_sm(int smem_dynamic, int smem_static) {
    int per_sm = 227 * 1024 / (smem_dynamic + smem_static + 1024);      // 227 KB usable per SM, 1 KB reserved per CTA
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 12) per_sm = 12;
    return per_sm;
}

}  // namespace

// ---- the register-held multisplit (multisplit.cu): K2b and K2c without the permutation pass ----------------------
// Segments of records in, hash-prefix buckets out.  K2b: n_seg wire segments (seg_off given) -> one local set, bucket =
// top `bits` hash bits.  K2c: segment s = coarse bucket s (offset s * seg_cap, the set's spill as one more segment that is
// passed through to the output spill) -> buckets [s << bits, (s + 1) << bits) of the fine set, bucket = the next `bits`
// hash bits after the `skip_bits` the coarse pass used.
struct PgMultiSplit {
    const uint4 *in; const int64_t *seg_off; const unsigned long long *seg_cnt; int n_seg; int64_t seg_cap;
    int pass_seg; int64_t pass_cap;          // index of the pass-through (spill) segment or -1, and its capacity
    int lost_on_clamp;                       // a count above seg_cap means records were dropped upstream (K2b): raise PG_STAT_LOST
    uint4 *out; unsigned long long *out_counts; int64_t out_part_cap, out_spill_cap; int64_t n_out;   // n_out buckets, then the spill
    int skip_bits, bits, sliced;             // sliced: segment s writes buckets [s << bits, ...)
    int64_t *stats;
    int policy;                              // cache-hint experiment switch (PG_SPLIT_POLICY)
};
int pg_multisplit_launch(const PgMultiSplit &a, cudaStream_t st);

