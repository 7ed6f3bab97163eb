// fasta_pack.cu - K1: FASTA scan + 2-bit pack on the GPU.
//
// Replaces the reference's byte-at-a-time reader (readline_jit_ / seqio_jit_,
// kmer_numba.py:122-188) and the per-base table lookups (alpha / lastc, :736-768)
// with three launches over 16 KB tiles of the raw file:
//   A  k1_tile_summaries  every tile -> what it does to the line-state machine for each of the three
//                         possible entry states (Sum3) + its real '\n' count        [reads 1 B/byte]
//   B  k1_tile_scan       one CTA composes the tile summaries in file order -> per-tile entry state,
//                         base rank, record rank; zeroes the few words tiles share   [12 B/tile]
//   C  k1_tile_pack       every tile again: classify, rank, pack 2-bit digits + 1-bit ambiguity
//                         mask into shared memory, write whole words coalesced       [reads 1 B/byte,
//                         writes 0.375 B/base]
// HBM-bound streaming work: 128-bit coalesced loads, SWAR byte classification (no per-byte loops),
// shuffle scans, shared-memory staging so global stores are full words.
#include <stdlib.h>
#include "fasta_chunk.cuh"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_SUB = K1_THREADS * 16;   // bytes per sub-tile
constexpr int K1_NSUB = 4;
constexpr int K1_TILE = K1_SUB * K1_NSUB; // 16 KB

struct TileSum { uint32_t v[3]; uint32_t real_nl; };
struct TileEntry { uint64_t seq; uint64_t hdr; uint32_t state; uint32_t dead; };

__device__ __forceinline__ Sum3 shfl_up_sum3(const Sum3 &x, int o) {
    Sum3 y;
#pragma unroll
    for (int i = 0; i < 3; i++) y.v[i] = __shfl_up_sync(0xffffffffu, x.v[i], o);
    return y;
}

__device__ __forceinline__ ChunkCls load_classify(const uint8_t *fasta, int64_t nbytes, int64_t off) {
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t left = nbytes - off;
    if (left >= 16) {
        uint4 v = pg_ld_stream(reinterpret_cast<const uint4 *>(fasta + off));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else if (left > 0) {   // the single ragged chunk at the end of the file
        uint64_t lo = 0, hi = 0;
        for (int i = 0; i < (int)left; i++) {
            uint64_t b = fasta[off + i];
            if (i < 8) lo |= b << (8 * i); else hi |= b << (8 * (i - 8));
        }
        w[0] = (uint32_t)lo; w[1] = (uint32_t)(lo >> 32); w[2] = (uint32_t)hi; w[3] = (uint32_t)(hi >> 32);
    }
    int n_file = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
    return classify16(w, n_file, left <= 16);
}

// Ordered scan of the K1_NSUB x K1_THREADS chunk summaries of one tile (non-commutative compose).
// Chunk (c, t) covers bytes [(c * K1_THREADS + t) * 16, +16) of the tile, so every load instruction is
// perfectly coalesced and the four loads of a thread are in flight together.  One warp-shuffle scan per
// c, then warp 0 scans the 32 (c, warp) aggregates.  excl[c] = everything before chunk (c, t).
__device__ __forceinline__ void tile_scan4(const Sum3 mine[K1_NSUB], Sum3 *s_agg /* K1_NSUB*8 */, Sum3 excl[K1_NSUB], Sum3 &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Sum3 inc[K1_NSUB];
#pragma unroll
    for (int c = 0; c < K1_NSUB; c++) {
        inc[c] = mine[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Sum3 y = shfl_up_sum3(inc[c], o);
            if (lane >= o) inc[c] = sum3_compose(y, inc[c]);
        }
        if (lane == 31) s_agg[c * (K1_THREADS / 32) + warp] = inc[c];
    }
    __syncthreads();
    if (warp == 0) {   // 32 aggregates in (c, warp) order -> exclusive prefixes, total in slot 32
        Sum3 a = s_agg[lane], ainc = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Sum3 y = shfl_up_sum3(ainc, o);
            if (lane >= o) ainc = sum3_compose(y, ainc);
        }
        Sum3 prev = shfl_up_sum3(ainc, 1);
        __syncwarp();
        s_agg[lane] = lane == 0 ? sum3_identity() : prev;
        if (lane == 31) s_agg[32] = ainc;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < K1_NSUB; c++) {
        Sum3 wp = s_agg[c * (K1_THREADS / 32) + warp];
        Sum3 prev = shfl_up_sum3(inc[c], 1);
        excl[c] = (lane == 0) ? wp : sum3_compose(wp, prev);
    }
    total = s_agg[32];
    __syncthreads();
}

template <bool LINES_ONLY>
__device__ __forceinline__ void load_classify4(const uint8_t *fasta, int64_t nbytes, int64_t tile_off, ChunkCls c[K1_NSUB]) {
    uint4 v[K1_NSUB];
    const bool full = tile_off + K1_TILE <= nbytes;
    if (full) {
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++)
            v[i] = pg_ld_stream(reinterpret_cast<const uint4 *>(fasta + tile_off + ((int64_t)i * K1_THREADS + threadIdx.x) * 16));
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) {
            uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            int64_t left = nbytes - (tile_off + ((int64_t)i * K1_THREADS + threadIdx.x) * 16);
            c[i] = LINES_ONLY ? classify16_lines(w, 16, left <= 16) : classify16(w, 16, left <= 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) c[i] = load_classify(fasta, nbytes, tile_off + ((int64_t)i * K1_THREADS + threadIdx.x) * 16);
    }
}

__global__ void __launch_bounds__(K1_THREADS)
k1_tile_summaries(const uint8_t *__restrict__ fasta, int64_t nbytes, int64_t ntiles, TileSum *__restrict__ sums) {
    __shared__ Sum3 s_agg[K1_NSUB * (K1_THREADS / 32) + 1];
    __shared__ uint32_t s_nl;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_nl = 0;
        ChunkCls c[K1_NSUB];
        load_classify4<true>(fasta, nbytes, tile * K1_TILE, c);
        Sum3 mine[K1_NSUB], excl[K1_NSUB], total;
        uint32_t my_nl = 0;
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) { mine[i] = chunk_sum3(c[i]); my_nl += c[i].real_nl; }
        tile_scan4(mine, s_agg, excl, total);
        my_nl = __reduce_add_sync(0xffffffffu, my_nl);
        if ((threadIdx.x & 31) == 0 && my_nl) atomicAdd(&s_nl, my_nl);
        __syncthreads();
        if (threadIdx.x == 0) {
            TileSum t; t.v[0] = total.v[0]; t.v[1] = total.v[1]; t.v[2] = total.v[2]; t.real_nl = s_nl;
            sums[tile] = t;
        }
        __syncthreads();
    }
}

// 64-bit running state of the file-order scan
struct Run { uint64_t seq, hdr; uint32_t state; uint64_t nl; };
struct Agg3 { uint64_t seq[3], hdr[3]; uint32_t st[3]; uint64_t nl; };

__device__ __forceinline__ void agg3_push(Agg3 &a, const TileSum &t) {
#pragma unroll
    for (int s = 0; s < 3; s++) {
        uint32_t st_ = a.st[s]; uint32_t y = st_ == 0 ? t.v[0] : (st_ == 1 ? t.v[1] : t.v[2]);
        a.seq[s] += SV_SEQ(y); a.hdr[s] += SV_HDR(y); a.st[s] = SV_STATE(y);
    }
    a.nl += t.real_nl;
}

constexpr int K1B_THREADS = 256;

__global__ void __launch_bounds__(K1B_THREADS)
k1_tile_scan(const TileSum *__restrict__ sums, int64_t ntiles, TileEntry *__restrict__ entries,
             uint32_t *__restrict__ pk2, uint32_t *__restrict__ amb, int64_t *__restrict__ seq_off,
             int64_t cap_records, int64_t *__restrict__ counts) {
    __shared__ Agg3 s_agg[K1B_THREADS];
    __shared__ Run s_run[K1B_THREADS];
    const int t = threadIdx.x;
    int64_t per = (ntiles + K1B_THREADS - 1) / K1B_THREADS;
    int64_t lo = (int64_t)t * per, hi = lo + per < ntiles ? lo + per : ntiles;
    Agg3 a;
#pragma unroll
    for (int s = 0; s < 3; s++) { a.seq[s] = 0; a.hdr[s] = 0; a.st[s] = s; }
    a.nl = 0;
    for (int64_t i = lo; i < hi; i++) agg3_push(a, sums[i]);
    s_agg[t] = a;
    __syncthreads();
    if (t == 0) {   // 256 compositions, serial: a few microseconds
        Run r; r.seq = 0; r.hdr = 0; r.state = ST_LINE_START; r.nl = 0;
        for (int i = 0; i < K1B_THREADS; i++) {
            s_run[i] = r;
            const Agg3 &g = s_agg[i];
            uint32_t s = r.state;
            r.seq += g.seq[s]; r.hdr += g.hdr[s]; r.state = g.st[s]; r.nl += g.nl;
        }
        // A file without any real '\n' yields no line at all (readline_jit_ :129-132: `end > start > 0`)
        bool dead = (r.nl == 0);
        counts[0] = dead ? 0 : (int64_t)r.hdr;
        counts[1] = dead ? 0 : (int64_t)r.seq;
        counts[2] = (int64_t)r.nl;
        counts[3] = dead ? 1 : 0;
        if (!dead && (int64_t)r.hdr <= cap_records) seq_off[r.hdr] = (int64_t)r.seq;
        if (dead) seq_off[0] = 0;
        // zero the padding the k-mer kernels may read past the last base
        uint64_t tot = dead ? 0 : r.seq;
        for (int i = 0; i < 16; i++) { pk2[(tot >> 4) + i] = 0; amb[(tot >> 5) + i] = 0; }
    }
    __syncthreads();
    bool dead = counts[3] != 0;
    Run r = s_run[t];
    for (int64_t i = lo; i < hi; i++) {
        TileEntry e; e.seq = r.seq; e.hdr = r.hdr; e.state = r.state; e.dead = dead ? 1u : 0u;
        entries[i] = e;
        if (!dead) { pk2[r.seq >> 4] = 0; amb[r.seq >> 5] = 0; }   // words shared between neighbouring tiles
        TileSum ts = sums[i];
        uint32_t y = r.state == 0 ? ts.v[0] : (r.state == 1 ? ts.v[1] : ts.v[2]);
        r.seq += SV_SEQ(y); r.hdr += SV_HDR(y); r.state = SV_STATE(y);
    }
}

constexpr int K1_PKW = (K1_TILE + 32) / 16 + 2;   // staged pk2 words per tile
constexpr int K1_AMW = (K1_TILE + 32) / 32 + 2;

__global__ void __launch_bounds__(K1_THREADS)
k1_tile_pack(const uint8_t *__restrict__ fasta, int64_t nbytes, int64_t ntiles,
             const TileEntry *__restrict__ entries, uint32_t *__restrict__ pk2, uint32_t *__restrict__ amb,
             int64_t *__restrict__ hdr_off, int64_t *__restrict__ seq_off, int64_t cap_records) {
    __shared__ Sum3 s_agg[K1_NSUB * (K1_THREADS / 32) + 1];
    __shared__ uint32_t s_pk[K1_PKW];
    __shared__ uint32_t s_am[K1_AMW];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TileEntry e = entries[tile];
        if (e.dead) return;
        ChunkCls c[K1_NSUB];
        load_classify4<false>(fasta, nbytes, tile * K1_TILE, c);   // loads in flight while the staging is cleared
        for (int i = threadIdx.x; i < K1_PKW; i += K1_THREADS) s_pk[i] = 0;
        for (int i = threadIdx.x; i < K1_AMW; i += K1_THREADS) s_am[i] = 0;
        Sum3 mine[K1_NSUB], excl[K1_NSUB], total;
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) mine[i] = chunk_sum3(c[i]);
        tile_scan4(mine, s_agg, excl, total);                      // contains the barrier that publishes the zeroed staging
        const uint32_t a0 = (uint32_t)(e.seq & 31);
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) {
            const int64_t off = tile * K1_TILE + ((int64_t)i * K1_THREADS + threadIdx.x) * 16;
            uint32_t x = sum3_sel(excl[i], e.state);
            ChunkRun r = chunk_run(c[i], SV_STATE(x));
            uint32_t rank = SV_SEQ(x);                 // bases of this tile before my chunk
            uint32_t cnt = pg_popc(r.seqmask);
            if (cnt) {
                uint32_t d = pext16_2bit(c[i].dig, r.seqmask);
                uint32_t m = pext16_1bit(c[i].amb, r.seqmask);
                if (cnt < 16) d &= (1u << (2 * cnt)) - 1u;
                uint32_t q = a0 + rank;
                uint32_t sh = 2 * (q & 15);
                atomicOr(&s_pk[q >> 4], d << sh);
                if (sh && (d >> (32 - sh))) atomicOr(&s_pk[(q >> 4) + 1], d >> (32 - sh));
                if (m) {
                    uint32_t sh1 = q & 31;
                    atomicOr(&s_am[q >> 5], m << sh1);
                    if (sh1 > 16 && (m >> (32 - sh1))) atomicOr(&s_am[(q >> 5) + 1], m >> (32 - sh1));
                }
            }
            if (r.hs) {   // rare: this chunk starts record(s)
                uint32_t hs = r.hs; uint64_t idx = e.hdr + SV_HDR(x);
                while (hs) {
                    int j = pg_ctz(hs); hs &= hs - 1;
                    if ((int64_t)idx < cap_records) {
                        hdr_off[idx] = off + j;
                        seq_off[idx] = (int64_t)(e.seq + rank + pg_popc(r.seqmask & ((1u << j) - 1u)));
                    }
                    idx++;
                }
            }
        }
        const uint32_t cseq = SV_SEQ(sum3_sel(total, e.state));
        __syncthreads();
        // write-out: words fully owned by this tile are stored, shared boundary words are OR-ed
        const uint32_t end = a0 + cseq;
        uint32_t *gpk = pk2 + ((e.seq >> 5) << 1);
        uint32_t *gam = amb + (e.seq >> 5);
        const uint32_t npk = (end + 15) >> 4, nam = (end + 31) >> 5;
        for (uint32_t w = threadIdx.x; w < npk; w += K1_THREADS) {
            uint32_t v = s_pk[w];
            if (16 * w >= a0 && 16 * (w + 1) <= end) gpk[w] = v;
            else if (v) atomicOr(&gpk[w], v);
        }
        for (uint32_t w = threadIdx.x; w < nam; w += K1_THREADS) {
            uint32_t v = s_am[w];
            if (32 * w >= a0 && 32 * (w + 1) <= end) gam[w] = v;
            else if (v) atomicOr(&gam[w], v);
        }
        __syncthreads();
    }
}

// =================================================================================================
// Single-pass K1: one launch reads the file ONCE.
// Tiles are handed out by an atomic ticket (so every earlier tile is already running or done), and
// the prefix a tile needs - entry line state, base rank, record rank - comes from a decoupled
// look-back over the tiles before it: each tile first publishes what it does for each of the three
// entry states (tile_sum3: built from quantities that do not depend on the entry state), later its
// inclusive prefix.  Inside a tile the line state of every 16-byte chunk follows from a max-scan of
// "position of the last newline" (one word per chunk) and the base ranks from a plain sum scan - no
// 3-variant composition in the per-chunk scans any more.
struct K1Pref { unsigned long long seq, hdr, nl; uint32_t state, pad; };
struct K1Entry { unsigned long long seq, hdr, nl; uint32_t state; };

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// executed by warp 0: the inclusive prefix of tiles [0, tile), or ok = false if a predecessor never showed up
__device__ __forceinline__ K1Entry k1_look_back(int64_t tile, const uint4 *agg, const uint32_t *tile_nl, const K1Pref *pref,
                                                const uint32_t *pflag, bool &ok) {
    const int lane = threadIdx.x & 31;
    // composite of the tiles between the window and `tile`: state map + counts per entry state
    uint32_t gst[3] = {0, 1, 2}; unsigned long long gseq[3] = {0, 0, 0}, ghdr[3] = {0, 0, 0}, gnl = 0;
    K1Entry e; e.seq = e.hdr = e.nl = 0; e.state = ST_LINE_START;
    ok = true;
    int64_t j = tile - 1;
    while (j >= 0) {
        const int64_t idx = j - lane;
        uint32_t pf = 0; uint4 a = make_uint4(0, 0, 0, 0); uint32_t nl = 0;
        int spins = 0;
        unsigned need;
        for (;;) {
            if (idx >= 0) {
                pf = ld_acquire_u32(pflag + idx);
                if (pf != 2) a = ld_volatile_u4(agg + idx);
            }
            const unsigned has_pref = __ballot_sync(0xffffffffu, idx >= 0 && pf == 2);
            const unsigned has_agg = __ballot_sync(0xffffffffu, idx < 0 || pf == 2 || a.w == 1);
            const int first_pref = has_pref ? __ffs(has_pref) - 1 : 32;
            need = first_pref >= 32 ? 0xffffffffu : ((1u << first_pref) - 1u);       // lanes nearer than the first prefix must show an aggregate
            if ((has_agg & need) == need) break;
            if (++spins > (1 << 20)) { ok = false; return e; }
            __nanosleep(20);
        }
        if (idx >= 0 && pf != 2) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(nl) : "l"(tile_nl + idx));   // never through a stale L1 line
        const unsigned has_pref = __ballot_sync(0xffffffffu, idx >= 0 && pf == 2);
        const int first_pref = has_pref ? __ffs(has_pref) - 1 : 32;
        const int n_agg = first_pref < 32 ? first_pref : (int)((j + 1 < 32) ? j + 1 : 32);
        for (int l = 0; l < n_agg; l++) {          // nearest tile first: new composite = (tile j-l) then (old composite)
            const uint32_t v0 = __shfl_sync(0xffffffffu, a.x, l), v1 = __shfl_sync(0xffffffffu, a.y, l), v2 = __shfl_sync(0xffffffffu, a.z, l);
            const uint32_t tnl = __shfl_sync(0xffffffffu, nl, l);
            uint32_t nst[3]; unsigned long long nseq[3], nhdr[3];
#pragma unroll
            for (int s = 0; s < 3; s++) {
                const uint32_t x = s == 0 ? v0 : (s == 1 ? v1 : v2);
                const uint32_t mid = SV_STATE(x);
                nst[s] = mid == 0 ? gst[0] : (mid == 1 ? gst[1] : gst[2]);
                nseq[s] = SV_SEQ(x) + (mid == 0 ? gseq[0] : (mid == 1 ? gseq[1] : gseq[2]));
                nhdr[s] = SV_HDR(x) + (mid == 0 ? ghdr[0] : (mid == 1 ? ghdr[1] : ghdr[2]));
            }
#pragma unroll
            for (int s = 0; s < 3; s++) { gst[s] = nst[s]; gseq[s] = nseq[s]; ghdr[s] = nhdr[s]; }
            gnl += tnl;
        }
        if (first_pref < 32) {
            K1Pref p;
            const K1Pref *pp = pref + (j - first_pref);
            asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(p.seq), "=l"(p.hdr) : "l"(pp));
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(p.nl) : "l"(&pp->nl));
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(p.state) : "l"(&pp->state));
            e.state = p.state == 0 ? gst[0] : (p.state == 1 ? gst[1] : gst[2]);
            e.seq = p.seq + (p.state == 0 ? gseq[0] : (p.state == 1 ? gseq[1] : gseq[2]));
            e.hdr = p.hdr + (p.state == 0 ? ghdr[0] : (p.state == 1 ? ghdr[1] : ghdr[2]));
            e.nl = p.nl + gnl;
            return e;
        }
        j -= 32;
    }
    e.state = gst[ST_LINE_START]; e.seq = gseq[ST_LINE_START]; e.hdr = ghdr[ST_LINE_START]; e.nl = gnl;   // start of file
    return e;
}

__global__ void __launch_bounds__(K1_THREADS)
k1_fused_pack(const uint8_t *__restrict__ fasta, int64_t nbytes, int64_t ntiles, unsigned int *ticket, uint4 *agg, uint32_t *tile_nl,
              K1Pref *pref, uint32_t *pflag, uint32_t *__restrict__ pk2, uint32_t *__restrict__ amb,
              int64_t *__restrict__ hdr_off, int64_t *__restrict__ seq_off, int64_t cap_records, int64_t *__restrict__ counts) {
    constexpr int NW = K1_THREADS / 32, NAGG = K1_NSUB * NW;
    __shared__ uint16_t s_gt[K1_TILE / 16];
    __shared__ int s_imax[NAGG + 1];
    __shared__ uint32_t s_add[NAGG + 1];
    __shared__ uint32_t s_pk[K1_PKW], s_am[K1_AMW];
    __shared__ int s_first_nl; __shared__ uint32_t s_post_seq, s_post_hdr, s_real_nl;
    __shared__ long long s_tile; __shared__ K1Entry s_entry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_tile = (long long)atomicAdd(ticket, 1u); s_first_nl = 0x7FFFFFFF; s_post_seq = s_post_hdr = s_real_nl = 0; }
        for (int i = threadIdx.x; i < K1_PKW; i += K1_THREADS) s_pk[i] = 0;
        for (int i = threadIdx.x; i < K1_AMW; i += K1_THREADS) s_am[i] = 0;
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= ntiles) break;
        const int64_t tile_off = tile * K1_TILE;
        // ---- S1: load + classify (four 128-bit loads in flight per thread)
        ChunkCls c[K1_NSUB];
        load_classify4<false>(fasta, nbytes, tile_off, c);
        int off[K1_NSUB], prev[K1_NSUB];
        // ---- S2: max-scan of the last newline position, tile-wide first newline, real newline count
        int my_first = 0x7FFFFFFF; uint32_t my_nl = 0;
        int inc[K1_NSUB];
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) {
            off[i] = (i * K1_THREADS + threadIdx.x) * 16;
            s_gt[off[i] >> 4] = (uint16_t)c[i].gt;
            inc[i] = chunk_last_nl(c[i].nl, off[i]);
            int f = chunk_first_nl(c[i].nl, off[i]); my_first = f < my_first ? f : my_first;
            my_nl += c[i].real_nl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc[i], o); if (lane >= o && y > inc[i]) inc[i] = y; }
            if (lane == 31) s_imax[i * NW + warp] = inc[i];
        }
        my_first = __reduce_min_sync(0xffffffffu, my_first);
        my_nl = __reduce_add_sync(0xffffffffu, my_nl);
        if (lane == 0) { atomicMin(&s_first_nl, my_first); if (my_nl) atomicAdd(&s_real_nl, my_nl); }
        __syncthreads();
        if (warp == 0) {       // exclusive max-scan of the 32 (sub-tile, warp) aggregates; total in slot NAGG
            int v = s_imax[lane], x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o && y > x) x = y; }
            int ex = __shfl_up_sync(0xffffffffu, x, 1);
            __syncwarp();
            s_imax[lane] = lane == 0 ? -1 : ex;
            if (lane == 31) s_imax[NAGG] = x;
        }
        __syncthreads();
        const int last_nl = s_imax[NAGG], first_nl = s_first_nl;
        // ---- S3/S4: chunk states where they do not depend on the tile's entry state; entry-independent sums
        ChunkRun run[K1_NSUB];
        uint32_t psum_seq = 0, psum_hdr = 0;
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) {
            int ex = __shfl_up_sync(0xffffffffu, inc[i], 1);
            int wp = s_imax[i * NW + warp];
            prev[i] = lane == 0 ? wp : (ex > wp ? ex : wp);
            run[i] = chunk_run(c[i], chunk_entry_state(prev[i], off[i], s_gt, ST_HEADER));
            psum_seq += pg_popc(run[i].seqmask); psum_hdr += pg_popc(run[i].hs);
        }
        psum_seq = __reduce_add_sync(0xffffffffu, psum_seq); psum_hdr = __reduce_add_sync(0xffffffffu, psum_hdr);
        if (lane == 0) { if (psum_seq) atomicAdd(&s_post_seq, psum_seq); if (psum_hdr) atomicAdd(&s_post_hdr, psum_hdr); }
        __syncthreads();
        // ---- S5: publish the tile summary, look back, publish the inclusive prefix (warp 0)
        if (warp == 0) {
            TileLocal tl; tl.post_seq = s_post_seq; tl.post_hdr = s_post_hdr; tl.first_nl = first_nl; tl.last_nl = last_nl; tl.tile_len = K1_TILE;
            tl.first_gt = s_gt[0] & 1u;
            tl.gt_after_last = (last_nl >= 0 && last_nl + 1 < K1_TILE) ? ((s_gt[(last_nl + 1) >> 4] >> ((last_nl + 1) & 15)) & 1u) : 0u;
            const Sum3 sm = tile_sum3(tl);
            if (lane == 0) {
                tile_nl[tile] = s_real_nl;
                __threadfence();
                asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(agg + tile), "r"(sm.v[0]), "r"(sm.v[1]), "r"(sm.v[2]), "r"(1u) : "memory");
            }
            bool ok;
            K1Entry e = k1_look_back(tile, agg, tile_nl, pref, pflag, ok);
            if (lane == 0) {
                const uint32_t y = sum3_sel(sm, e.state);
                K1Pref p; p.seq = e.seq + SV_SEQ(y); p.hdr = e.hdr + SV_HDR(y); p.nl = e.nl + s_real_nl; p.state = SV_STATE(y); p.pad = 0;
                pref[tile] = p;
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(pflag + tile), "r"(2u) : "memory");
                s_entry = e;
                if (tile == ntiles - 1) {       // whole-file totals
                    const bool dead = (p.nl == 0) || !ok;     // a file without any real newline yields no line (readline_jit_ :129-132)
                    counts[0] = dead ? 0 : (int64_t)p.hdr; counts[1] = dead ? 0 : (int64_t)p.seq; counts[2] = (int64_t)p.nl;
                    counts[3] = !ok ? 2 : (dead ? 1 : 0);
                    if (dead) seq_off[0] = 0;
                    else if ((int64_t)p.hdr <= cap_records) seq_off[p.hdr] = (int64_t)p.seq;
                }
            }
        }
        __syncthreads();
        const K1Entry e = s_entry;
        // ---- S6: final chunk states, ranks by a plain sum scan (bases | headers << 16)
        const uint32_t fallback = e.state == ST_LINE_START ? ((s_gt[0] & 1u) ? (uint32_t)ST_HEADER : (uint32_t)ST_SEQ) : e.state;
        uint32_t cnt[K1_NSUB], ainc[K1_NSUB];
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) {
            if (prev[i] < 0) {
                const uint32_t st = (off[i] == 0 && e.state == ST_LINE_START) ? (uint32_t)ST_LINE_START : fallback;
                run[i] = chunk_run(c[i], st);
            }
            cnt[i] = pg_popc(run[i].seqmask) | (pg_popc(run[i].hs) << 16);
            ainc[i] = cnt[i];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, ainc[i], o); if (lane >= o) ainc[i] += y; }
            if (lane == 31) s_add[i * NW + warp] = ainc[i];
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t v = s_add[lane], x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            s_add[lane] = x - v;
            if (lane == 31) s_add[NAGG] = x;
        }
        __syncthreads();
        // ---- S7: pack into shared memory, record index, coalesced write-out
        const uint32_t a0 = (uint32_t)(e.seq & 31);
#pragma unroll
        for (int i = 0; i < K1_NSUB; i++) {
            const uint32_t excl = s_add[i * NW + warp] + ainc[i] - cnt[i];
            const uint32_t rank = excl & 0xFFFFu, hrank = excl >> 16;
            const uint32_t n = cnt[i] & 0xFFFFu;
            if (n) {
                uint32_t d = pext16_2bit(c[i].dig, run[i].seqmask);
                uint32_t m = pext16_1bit(c[i].amb, run[i].seqmask);
                if (n < 16) d &= (1u << (2 * n)) - 1u;
                const uint32_t q = a0 + rank, sh = 2 * (q & 15);
                atomicOr(&s_pk[q >> 4], d << sh);
                if (sh && (d >> (32 - sh))) atomicOr(&s_pk[(q >> 4) + 1], d >> (32 - sh));
                if (m) {
                    const uint32_t sh1 = q & 31;
                    atomicOr(&s_am[q >> 5], m << sh1);
                    if (sh1 > 16 && (m >> (32 - sh1))) atomicOr(&s_am[(q >> 5) + 1], m >> (32 - sh1));
                }
            }
            if (run[i].hs) {
                uint32_t hs = run[i].hs; unsigned long long idx = e.hdr + hrank;
                while (hs) {
                    const int j = pg_ctz(hs); hs &= hs - 1;
                    if ((int64_t)idx < cap_records) {
                        hdr_off[idx] = tile_off + off[i] + j;
                        seq_off[idx] = (int64_t)(e.seq + rank + pg_popc(run[i].seqmask & ((1u << j) - 1u)));
                    }
                    idx++;
                }
            }
        }
        __syncthreads();
        const uint32_t tot = s_add[NAGG] & 0xFFFFu;            // bases of this tile (<= 16384 fits 16 bits? 16384 = 0x4000: yes)
        const uint32_t end = a0 + tot;
        uint32_t *gpk = pk2 + ((e.seq >> 5) << 1);
        uint32_t *gam = amb + (e.seq >> 5);
        const uint32_t npk = (end + 15) >> 4, nam = (end + 31) >> 5;
        for (uint32_t w = threadIdx.x; w < npk; w += K1_THREADS) {
            const uint32_t v = s_pk[w];
            if (16 * w >= a0 && 16 * (w + 1) <= end) gpk[w] = v;
            else if (v) atomicOr(&gpk[w], v);                      // word shared with a neighbouring tile (buffers are pre-zeroed)
        }
        for (uint32_t w = threadIdx.x; w < nam; w += K1_THREADS) {
            const uint32_t v = s_am[w];
            if (32 * w >= a0 && 32 * (w + 1) <= end) gam[w] = v;
            else if (v) atomicOr(&gam[w], v);
        }
    }
}


// =================================================================================================
// K1 in 32-byte chunks with LOCAL header detection (fasta_chunk.cuh, "K1 in 32-byte chunks"): the default.
//   A  k1x_tile_aggs   16 KB tiles, one 32-byte chunk per thread: newline / '>' masks only; the in-header bit carried
//                      across chunks by two ballots per warp and once more across the 16 warps; per tile ONE uint4:
//                      bases if the tile is entered outside a header, how many of them precede its first line end (they
//                      are header bytes if it is entered inside one), headers started, the state it hands on, real '\n's
//   B  k1x_scan        one CTA composes the tiles as functions of their entry state (TileFn) with a shuffle scan
//   C  k1x_pack        every tile again with the full classification: ranks from a plain sum scan, 2-bit / 1-bit pack
//                      staged in shared memory, whole-word stores (the write-out of k1_tile_pack)
// Against the three launches above: no 3-variant summary per chunk, no ordered scan over them (the per-chunk work of
// passes A and C drops from ~70 to ~15 thread-instructions per input byte), half as many scan participants.
constexpr int X_THREADS = 512;
constexpr int X_TILE = X_THREADS * 32;
constexpr int X_NW = X_THREADS / 32;
static_assert(X_TILE == K1_TILE, "the staging arrays and the workspace are sized for 16 KB tiles");

__device__ __forceinline__ void x_load32(const uint8_t *fasta, int64_t nbytes, int64_t off, uint4 &a, uint4 &b) {
    const int64_t left = nbytes - off;
    if (left >= 32) {
        a = __ldg(reinterpret_cast<const uint4 *>(fasta + off)); b = __ldg(reinterpret_cast<const uint4 *>(fasta + off) + 1);
    } else {
        uint64_t q[4] = {0, 0, 0, 0};
        for (int i = 0; i < (int)left; i++) {           // the ragged chunk(s) at the end of the file
            const uint64_t v = (uint64_t)fasta[off + i] << (8 * (i & 7));
            if (i < 8) q[0] |= v; else if (i < 16) q[1] |= v; else if (i < 24) q[2] |= v; else q[3] |= v;
        }
        a = make_uint4((uint32_t)q[0], (uint32_t)(q[0] >> 32), (uint32_t)q[1], (uint32_t)(q[1] >> 32));
        b = make_uint4((uint32_t)q[2], (uint32_t)(q[2] >> 32), (uint32_t)q[3], (uint32_t)(q[3] >> 32));
    }
}
template <bool LINES_ONLY>
__device__ __forceinline__ Cls32 x_classify(const uint4 &a, const uint4 &b, int64_t left) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    return classify32<LINES_ONLY>(w, left);
}
// the byte before the chunk ends a line (or the chunk starts the file): lane 0 looks, the others take the neighbour's bit 31
__device__ __forceinline__ uint32_t x_prev_nl(const uint8_t *fasta, int64_t nbytes, int64_t off, uint32_t nl) {
    uint32_t p = __shfl_up_sync(0xffffffffu, nl >> 31, 1);
    if ((threadIdx.x & 31) == 0) p = off == 0 ? 1u : (off - 1 < nbytes ? (uint32_t)(fasta[off - 1] == '\n') : 1u);
    return p;
}
// entry state of this thread's chunk: two ballots inside the warp, the warps' kinds through shared memory.
// Contains one __syncthreads.  tile_kind: what the whole tile hands on.
__device__ __forceinline__ uint32_t x_chunk_entry(uint32_t kind, uint32_t tile_entry, uint32_t *s_wkind, uint32_t &tile_kind) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t fixed = __ballot_sync(0xffffffffu, kind != HK_PASS), set = __ballot_sync(0xffffffffu, kind == HK_SET);
    if (lane == 0) s_wkind[warp] = fixed ? (set >> pg_msb(fixed)) & 1u : (uint32_t)HK_PASS;
    __syncthreads();
    const uint32_t k2 = lane < X_NW ? s_wkind[lane] : (uint32_t)HK_PASS;
    const uint32_t tf = __ballot_sync(0xffffffffu, k2 != HK_PASS), ts = __ballot_sync(0xffffffffu, k2 == HK_SET);
    tile_kind = tf ? (ts >> pg_msb(tf)) & 1u : (uint32_t)HK_PASS;
    const uint32_t wentry = hdr_entry_from_masks(tf, ts, warp, tile_entry);
    return hdr_entry_from_masks(fixed, set, lane, wentry);
}

__global__ void __launch_bounds__(X_THREADS)
k1x_tile_aggs(const uint8_t *__restrict__ fasta, int64_t nbytes, int64_t ntiles, uint4 *__restrict__ aggs) {
    __shared__ uint32_t s_wkind[X_NW];
    __shared__ uint32_t s_seq, s_hdr, s_nl, s_first;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) { s_seq = 0; s_hdr = 0; s_nl = 0; s_first = X_TILE; }
        const int64_t off = tile * X_TILE + (int64_t)threadIdx.x * 32;
        uint4 wa, wb;
        x_load32(fasta, nbytes, off, wa, wb);
        const Cls32 c = x_classify<true>(wa, wb, nbytes - off);
        const uint32_t hs = hdr_starts32(c.nl, c.gt, x_prev_nl(fasta, nbytes, off, c.nl));
        uint32_t tile_kind;
        const uint32_t entry = x_chunk_entry(hdr_kind32(c.nl, hs), 0u, s_wkind, tile_kind);      // barrier inside: the zeroed sums are visible
        const uint32_t h = hdr_fill32(c.nl, hs, entry);
        uint32_t seq = pg_popc(~c.nl & ~h), nh = pg_popc(hs), rnl = c.real_nl;
        uint32_t first = c.nl ? (uint32_t)threadIdx.x * 32u + (uint32_t)pg_ctz(c.nl) : (uint32_t)X_TILE;
        seq = __reduce_add_sync(0xffffffffu, seq); nh = __reduce_add_sync(0xffffffffu, nh); rnl = __reduce_add_sync(0xffffffffu, rnl);
        first = __reduce_min_sync(0xffffffffu, first);
        if ((threadIdx.x & 31) == 0) {
            if (seq) atomicAdd(&s_seq, seq);
            if (nh) atomicAdd(&s_hdr, nh);
            if (rnl) atomicAdd(&s_nl, rnl);
            if (first < (uint32_t)X_TILE) atomicMin(&s_first, first);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t pre = (hs & 1u) ? 0u : s_first;       // a tile entered inside a header loses the bases before its first line end
            aggs[tile] = make_uint4(s_seq, pre, s_hdr | (tile_kind << 30), s_nl);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ TileFn x_fn_of(const uint4 a) { return tilefn_make(a.x, a.y, a.z & 0x3FFFFFFFu, a.z >> 30, a.w); }
__device__ __forceinline__ TileFn x_fn_shfl_up(const TileFn &f, int d) {
    TileFn g;
    g.seq0 = __shfl_up_sync(0xffffffffu, f.seq0, d); g.seq1 = __shfl_up_sync(0xffffffffu, f.seq1, d);
    g.hdr = __shfl_up_sync(0xffffffffu, f.hdr, d); g.nl = __shfl_up_sync(0xffffffffu, f.nl, d);
    const uint32_t e = __shfl_up_sync(0xffffffffu, f.exit0 | (f.exit1 << 1), d);
    g.exit0 = e & 1u; g.exit1 = e >> 1;
    return g;
}
constexpr int XB_THREADS = 1024;

__global__ void __launch_bounds__(XB_THREADS)
k1x_scan(const uint4 *__restrict__ aggs, int64_t ntiles, TileEntry *__restrict__ entries, uint32_t *__restrict__ pk2,
         uint32_t *__restrict__ amb, int64_t *__restrict__ seq_off, int64_t cap_records, int64_t *__restrict__ counts) {
    __shared__ TileFn s_w[XB_THREADS / 32];
    __shared__ TileFn s_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t per = (ntiles + XB_THREADS - 1) / XB_THREADS;
    const int64_t lo = (int64_t)t * per < ntiles ? (int64_t)t * per : ntiles, hi = lo + per < ntiles ? lo + per : ntiles;
    TileFn mine = tilefn_identity();
    for (int64_t i = lo; i < hi; i++) mine = tilefn_compose(mine, x_fn_of(aggs[i]));
    // inclusive scan of the (non-commutative) composition over the block: inside the warps, then over the warp totals
    TileFn inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const TileFn y = x_fn_shfl_up(inc, d); if (lane >= d) inc = tilefn_compose(y, inc); }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        TileFn a = s_w[lane], ainc = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const TileFn y = x_fn_shfl_up(ainc, d); if (lane >= d) ainc = tilefn_compose(y, ainc); }
        const TileFn ex = x_fn_shfl_up(ainc, 1);
        __syncwarp();
        s_w[lane] = lane == 0 ? tilefn_identity() : ex;           // exclusive over the warps
        if (lane == 31) s_total = ainc;
    }
    __syncthreads();
    TileFn before = x_fn_shfl_up(inc, 1);
    if (lane == 0) before = tilefn_identity();
    before = tilefn_compose(s_w[warp], before);                   // everything before this thread's first tile
    const TileFn total = s_total;
    // a file without any real '\n' yields no line at all (readline_jit_ :129-132: `end > start > 0`)
    const bool dead = total.nl == 0;
    if (t == 0) {
        counts[0] = dead ? 0 : (int64_t)total.hdr;
        counts[1] = dead ? 0 : (int64_t)total.seq0;
        counts[2] = (int64_t)total.nl;
        counts[3] = dead ? 1 : 0;
        if (!dead && (int64_t)total.hdr <= cap_records) seq_off[total.hdr] = (int64_t)total.seq0;
        if (dead) seq_off[0] = 0;
        const uint64_t tot = dead ? 0 : total.seq0;               // zero the padding the k-mer kernels may read past the last base
        for (int i = 0; i < 16; i++) { pk2[(tot >> 4) + i] = 0; amb[(tot >> 5) + i] = 0; }
    }
    // the file starts outside a header: only the entry-0 branch of `before` is ever taken
    uint64_t seq = before.seq0, hdr = before.hdr; uint32_t st = before.exit0;
    for (int64_t i = lo; i < hi; i++) {
        TileEntry e; e.seq = seq; e.hdr = hdr; e.state = st; e.dead = dead ? 1u : 0u;
        entries[i] = e;
        if (!dead) { pk2[seq >> 4] = 0; amb[seq >> 5] = 0; }      // words shared between neighbouring tiles (the tiles OR into them)
        const TileFn f = x_fn_of(aggs[i]);
        seq += st ? f.seq1 : f.seq0; hdr += f.hdr; st = st ? f.exit1 : f.exit0;
    }
}

__global__ void __launch_bounds__(X_THREADS)
k1x_pack(const uint8_t *__restrict__ fasta, int64_t nbytes, int64_t ntiles, const TileEntry *__restrict__ entries,
         uint32_t *__restrict__ pk2, uint32_t *__restrict__ amb, int64_t *__restrict__ hdr_off, int64_t *__restrict__ seq_off,
         int64_t cap_records) {
    __shared__ uint32_t s_wkind[X_NW];
    __shared__ uint32_t s_wtot[X_NW + 1];
    __shared__ uint32_t s_pk[K1_PKW];
    __shared__ uint32_t s_am[K1_AMW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileEntry e = entries[tile];
        if (e.dead) return;
        const int64_t off = tile * X_TILE + (int64_t)threadIdx.x * 32;
        uint4 wa, wb;
        x_load32(fasta, nbytes, off, wa, wb);                     // loads in flight while the staging is cleared
        for (int i = threadIdx.x; i < K1_PKW; i += X_THREADS) s_pk[i] = 0;
        for (int i = threadIdx.x; i < K1_AMW; i += X_THREADS) s_am[i] = 0;
        const Cls32 c = x_classify<false>(wa, wb, nbytes - off);
        const uint32_t hs = hdr_starts32(c.nl, c.gt, x_prev_nl(fasta, nbytes, off, c.nl));
        uint32_t tile_kind;
        const uint32_t entry = x_chunk_entry(hdr_kind32(c.nl, hs), e.state, s_wkind, tile_kind);      // barrier inside: staging zeroed
        const uint32_t seqmask = ~c.nl & ~hdr_fill32(c.nl, hs, entry);
        // ranks: bases | headers << 16 before this chunk, plain sum scan
        const uint32_t cnt = pg_popc(seqmask) | (pg_popc(hs) << 16);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t wt = lane < X_NW ? s_wtot[lane] : 0u, winc = wt;
#pragma unroll
        for (int d = 1; d < X_NW; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += y; }
        const uint32_t wbase = __shfl_sync(0xffffffffu, winc - wt, warp);
        const uint32_t total = __shfl_sync(0xffffffffu, winc, X_NW - 1);
        const uint32_t excl = wbase + inc - cnt;
        const uint32_t rank = excl & 0xFFFFu, hrank = excl >> 16;
        const uint32_t a0 = (uint32_t)(e.seq & 31);
        // ---- pack the two 16-byte halves into the shared-memory staging
        uint32_t r2 = rank;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t sm = (seqmask >> (16 * half)) & 0xFFFFu, n = pg_popc(sm);
            if (n) {
                uint32_t d = pext16_2bit(half ? c.dig_hi : c.dig_lo, sm);
                const uint32_t m = pext16_1bit((c.amb >> (16 * half)) & 0xFFFFu, sm);
                if (n < 16) d &= (1u << (2 * n)) - 1u;
                const uint32_t q = a0 + r2, sh = 2 * (q & 15);
                atomicOr(&s_pk[q >> 4], d << sh);
                if (sh && (d >> (32 - sh))) atomicOr(&s_pk[(q >> 4) + 1], d >> (32 - sh));
                if (m) {
                    const uint32_t sh1 = q & 31;
                    atomicOr(&s_am[q >> 5], m << sh1);
                    if (sh1 > 16 && (m >> (32 - sh1))) atomicOr(&s_am[(q >> 5) + 1], m >> (32 - sh1));
                }
            }
            r2 += n;
        }
        if (hs) {   // rare: this chunk starts record(s)
            uint32_t x = hs; uint64_t idx = e.hdr + hrank;
            while (x) {
                const int j = pg_ctz(x); x &= x - 1;
                if ((int64_t)idx < cap_records) {
                    hdr_off[idx] = off + j;
                    seq_off[idx] = (int64_t)(e.seq + rank + pg_popc(seqmask & ((1u << j) - 1u)));
                }
                idx++;
            }
        }
        __syncthreads();
        // ---- write-out: words fully owned by this tile are stored, shared boundary words are OR-ed
        const uint32_t end = a0 + (total & 0xFFFFu);
        uint32_t *gpk = pk2 + ((e.seq >> 5) << 1);
        uint32_t *gam = amb + (e.seq >> 5);
        const uint32_t npk = (end + 15) >> 4, nam = (end + 31) >> 5;
        for (uint32_t x = threadIdx.x; x < npk; x += X_THREADS) {
            const uint32_t v = s_pk[x];
            if (16 * x >= a0 && 16 * (x + 1) <= end) gpk[x] = v;
            else if (v) atomicOr(&gpk[x], v);
        }
        for (uint32_t x = threadIdx.x; x < nam; x += X_THREADS) {
            const uint32_t v = s_am[x];
            if (32 * x >= a0 && 32 * (x + 1) <= end) gam[x] = v;
            else if (v) atomicOr(&gam[x], v);
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int64_t pg_pack_words(int64_t cap_bases) { return (cap_bases < 0 ? 0 : cap_bases) / 16 + 32; }

extern "C" int64_t pg_fasta_workspace_bytes(int64_t nbytes) {
    int64_t ntiles = (nbytes + K1_TILE - 1) / K1_TILE;
    if (ntiles < 1) ntiles = 1;
    int64_t legacy = ntiles * (int64_t)(sizeof(TileSum) + sizeof(TileEntry)) + 256;
    int64_t fused = 256 + ntiles * (int64_t)(sizeof(uint4) + sizeof(K1Pref) + 2 * sizeof(uint32_t)) + 64;
    return legacy > fused ? legacy : fused;
}

extern "C" int pg_fasta_scan_pack(const uint8_t *d_fasta, int64_t nbytes, uint32_t *d_pk2, uint32_t *d_amb,
                                  int64_t cap_bases, int64_t *d_hdr_off, int64_t *d_seq_off, int64_t cap_records,
                                  int64_t *d_counts, void *d_ws, int64_t ws_bytes, pg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nbytes < 0 || !d_pk2 || !d_amb || !d_hdr_off || !d_seq_off || !d_counts || !d_ws || cap_records < 0)
        return pg_fail(PG_ERR_INVALID, "pg_fasta_scan_pack: null buffer or negative size");
    if (nbytes > 0 && !d_fasta) return pg_fail(PG_ERR_INVALID, "pg_fasta_scan_pack: d_fasta is null");
    if (cap_bases < nbytes) return pg_fail(PG_ERR_CAPACITY, "pg_fasta_scan_pack: cap_bases %lld < nbytes %lld",
                                           (long long)cap_bases, (long long)nbytes);
    if (ws_bytes < pg_fasta_workspace_bytes(nbytes))
        return pg_fail(PG_ERR_WORKSPACE, "pg_fasta_scan_pack: workspace %lld < %lld", (long long)ws_bytes,
                       (long long)pg_fasta_workspace_bytes(nbytes));
    if ((reinterpret_cast<uintptr_t>(d_fasta) & 15) != 0)
        return pg_fail(PG_ERR_INVALID, "pg_fasta_scan_pack: d_fasta must be 16-byte aligned");
    int64_t ntiles = (nbytes + K1_TILE - 1) / K1_TILE;
    // PG_K1_SINGLE_PASS=1 selects the one-launch kernel below.  It reads the file once (no second pass, 67 M
    // instead of 112 M warp instructions) and is bit-exact, but measured 267 us against 189 us for the three
    // launches on the 50 MB config-2 file: at 100 registers only two CTAs fit an SM and seven of a CTA's eight
    // warps sit at the barrier while warp 0 looks back (ncu: barrier stall 8.5).  It becomes the default once the
    // look-back runs on a helper warp under the digit classification (round-2 work).
    static int single_pass = -1;
    if (single_pass < 0) { const char *e = getenv("PG_K1_SINGLE_PASS"); single_pass = e ? atoi(e) : 0; }
    if (single_pass) {
        // single pass: zero the output planes (tiles OR into the words they share) and the look-back state
        PG_CUDA(cudaMemsetAsync(d_pk2, 0, (size_t)pg_pack_words(cap_bases) * 4, stream));
        PG_CUDA(cudaMemsetAsync(d_amb, 0, (size_t)pg_pack_words(cap_bases) * 4, stream));
        if (ntiles == 0) {
            PG_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int64_t), stream));
            PG_CUDA(cudaMemsetAsync(d_seq_off, 0, sizeof(int64_t), stream));
            return PG_OK;
        }
        char *w = reinterpret_cast<char *>(d_ws);
        size_t used = 256 + (size_t)ntiles * (sizeof(uint4) + sizeof(K1Pref) + 2 * sizeof(uint32_t));
        PG_CUDA(cudaMemsetAsync(d_ws, 0, used, stream));
        unsigned int *ticket = reinterpret_cast<unsigned int *>(w);
        uint4 *agg = reinterpret_cast<uint4 *>(w + 256);
        K1Pref *pref = reinterpret_cast<K1Pref *>(w + 256 + (size_t)ntiles * sizeof(uint4));
        uint32_t *tile_nl = reinterpret_cast<uint32_t *>(w + 256 + (size_t)ntiles * (sizeof(uint4) + sizeof(K1Pref)));
        uint32_t *pflag = tile_nl + ntiles;
        int64_t maxg = (int64_t)pg_num_sms() * 5;
        int grid1 = (int)(ntiles < maxg ? ntiles : maxg);
        k1_fused_pack<<<grid1, K1_THREADS, 0, stream>>>(d_fasta, nbytes, ntiles, ticket, agg, tile_nl, pref, pflag, d_pk2, d_amb,
                                                        d_hdr_off, d_seq_off, cap_records, d_counts);
        PG_CUDA(cudaGetLastError());
        return PG_OK;
    }
    static int legacy = -1;
    if (legacy < 0) { const char *e = getenv("PG_K1_LEGACY"); legacy = e ? atoi(e) : 0; }
    if (!legacy) {
        // 32-byte chunks, local header detection: aggregates (uint4 per tile), one-CTA scan, pack
        uint4 *aggs = reinterpret_cast<uint4 *>(d_ws);
        TileEntry *ent = reinterpret_cast<TileEntry *>(reinterpret_cast<char *>(d_ws) + (ntiles < 1 ? 1 : ntiles) * sizeof(uint4));
        // one CTA per tile up to 2^20 CTAs (then a grid-stride loop): with a resident-size grid of 4 CTAs per SM the 3052
        // tiles of the 50 MB config-2 file are 5.16 per CTA - a sixth, mostly empty wave; PG_K1_GRID=<CTAs per SM> restores that
        static int k1_grid = -1;
        if (k1_grid < 0) { const char *e = getenv("PG_K1_GRID"); k1_grid = e ? atoi(e) : 0; }
        const int64_t cap_x = k1_grid > 0 ? (int64_t)pg_num_sms() * k1_grid : (1ll << 20);
        const int grid_x = (int)(ntiles < cap_x ? (ntiles < 1 ? 1 : ntiles) : cap_x);
        if (ntiles > 0) k1x_tile_aggs<<<grid_x, X_THREADS, 0, stream>>>(d_fasta, nbytes, ntiles, aggs);
        k1x_scan<<<1, XB_THREADS, 0, stream>>>(aggs, ntiles, ent, d_pk2, d_amb, d_seq_off, cap_records, d_counts);
        if (ntiles > 0) k1x_pack<<<grid_x, X_THREADS, 0, stream>>>(d_fasta, nbytes, ntiles, ent, d_pk2, d_amb, d_hdr_off, d_seq_off, cap_records);
        PG_CUDA(cudaGetLastError());
        return PG_OK;
    }
    TileSum *sums = reinterpret_cast<TileSum *>(d_ws);
    TileEntry *entries = reinterpret_cast<TileEntry *>(reinterpret_cast<char *>(d_ws) + ((ntiles < 1 ? 1 : ntiles) * sizeof(TileSum) + 15) / 16 * 16);
    int sms = pg_num_sms();
    int grid = (int)(ntiles < (int64_t)sms * 8 ? (ntiles < 1 ? 1 : ntiles) : (int64_t)sms * 8);
    if (ntiles > 0) k1_tile_summaries<<<grid, K1_THREADS, 0, stream>>>(d_fasta, nbytes, ntiles, sums);
    k1_tile_scan<<<1, K1B_THREADS, 0, stream>>>(sums, ntiles, entries, d_pk2, d_amb, d_seq_off, cap_records, d_counts);
    if (ntiles > 0)
        k1_tile_pack<<<grid, K1_THREADS, 0, stream>>>(d_fasta, nbytes, ntiles, entries, d_pk2, d_amb, d_hdr_off,
                                                     d_seq_off, cap_records);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
