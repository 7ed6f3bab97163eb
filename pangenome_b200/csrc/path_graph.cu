// path_graph.cu - K5..K8: compressed-path hits, rdBG edges + weights, connected components,
// per-record breakpoint (region) labelling.
//
// Replaces the reference's sequential walks rdbg_edge_weight (kmer_numba.py:1446-1518),
// the label-dict build around the external `mcl` call (:1893-1944) and seq2path_jit_ (:1523-1573).
//
//   K5  pg_path_hits      one rdBG-table lookup per position -> hit bitmap (1 bit / base) + per-tile counts,
//                         scan, then an ordered emit of the hits (position, node key, record, v6)
//   K6  pg_graph_add_hits node table / edge table / per-record visit set - each a single-word
//                         atomicCAS open-addressing table - give edge weights = #records containing the edge
//   K7  pg_graph_components  atomic-hooking union-find over node slots + full path compression
//   K8  pg_label_regions  label lookup per hit, greedy non-overlap chain by pointer doubling, run ends -> rows
//
// Node identity follows the reference exactly: (code, v5) with v5 = (lastc[prev] << 5) | lastc[next]
// as a NUMBER (offbit 5, quirk Q7 / F7: bit 5 is shared by prev=A and next=$).  A node key packs
// (rdBG slot, orientation, v5) in 64 bits; an edge key packs two 32-bit node slots; a visit key packs
// (edge slot, record-strand id).
#include "table_dev.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int NODE_V_BITS = 11;           // v5 <= 544, v6 <= 1056
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// generic exclusive scan of uint32 counts into int64 offsets (three small launches)
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const uint32_t *__restrict__ in, int64_t n, int64_t *__restrict__ bsum) {
    __shared__ unsigned long long s[32];
    int64_t i = blockIdx.x * (int64_t)SCAN_THREADS + threadIdx.x;
    unsigned long long v = i < n ? in[i] : 0;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = s[threadIdx.x];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) bsum[blockIdx.x] = (int64_t)v;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_top(int64_t *bsum, int64_t nb, int64_t *total) {
    // single CTA: exclusive scan of the block sums, chunk by chunk
    __shared__ long long s[SCAN_THREADS];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
        int64_t i = base + threadIdx.x;
        long long v = i < nb ? bsum[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_THREADS; o <<= 1) {
            long long y = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += y;
            __syncthreads();
        }
        long long incl = s[threadIdx.x];
        if (i < nb) bsum[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const uint32_t *__restrict__ in, int64_t n, const int64_t *__restrict__ bsum, int64_t *__restrict__ out) {
    __shared__ unsigned long long s[32];
    int64_t i = blockIdx.x * (int64_t)SCAN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long v = i < n ? in[i] : 0, inc = v;
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = s[lane], winc = w;
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
        s[lane] = winc - w;
    }
    __syncthreads();
    if (i < n) out[i] = bsum[blockIdx.x] + (int64_t)(s[warp] + inc - v);
}
// out[i] = exclusive prefix of in[0..i), *total = sum.  bsum: ceil(n/1024) int64 of workspace.
int exclusive_scan(const uint32_t *in, int64_t n, int64_t *out, int64_t *bsum, int64_t *total, cudaStream_t st) {
    int64_t nb = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    if (nb > 0) scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, bsum);
    scan_top<<<1, SCAN_THREADS, 0, st>>>(bsum, nb, total);
    if (nb > 0) scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, bsum, out);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// rdBG lookup of the literal code the walked strand shows at a position.
// Returns true on a hit; slot_o = (slot << 1) | orientation; the phantom code 0 (Q6: has_key(0) is
// always true) hits even when absent from the table, with slot = capacity.
__device__ __forceinline__ bool rdbg_hit(const TableView &rd, int mode, uint64_t lit, uint64_t other, uint64_t &slot_o) {
    uint64_t key = lit; uint32_t o = 0;
    if (mode == PG_MODE_CANONICAL && other < lit) { key = other; o = 1; }
    uint64_t s = tv_home(rd, key);
    for (uint32_t probe = 0; probe < PG_MAX_PROBE; probe++) {
        uint64_t ck, cv;
        pg_ld_slot(rd.slots + 2 * s, rd.tag, ck, cv);
        if (ck == key) {
            uint32_t f = (uint32_t)(cv >> 32);
            slot_o = (s << 1) | o;
            return ((f >> o) & 1u) || lit == 0;
        }
        if (ck == PG_EMPTY) break;
        s = (s + 1) & rd.capmask;
    }
    if (lit == 0) { slot_o = ((rd.capmask + 1) << 1); return true; }
    return false;
}

struct SeqArgs {
    const uint64_t *pk2; const uint32_t *amb; int64_t n_words;
    const int64_t *seq_off; int64_t n_rec, g_begin, g_end; int k; uint64_t pow5km1; int64_t w_first, n_tiles;
};

// K5 pass 1: one lookup per position; thread t owns the 32-position word w0+t -> one bitmap word.
__global__ void __launch_bounds__(K2_THREADS)
k5_mark(TableView rd, int mode, SeqArgs a, int strand, uint32_t *__restrict__ hitbits, uint32_t *__restrict__ tile_counts) {
    __shared__ __align__(16) uint64_t s_pk[K2_TILE_WORDS + 4];
    __shared__ __align__(16) uint32_t s_am[K2_TILE_WORDS + 8];
    __shared__ uint32_t s_cnt;
    __shared__ uint16_t s_lut5[PG_LUT5_SIZE];
    for (int i = threadIdx.x; i < PG_LUT5_SIZE; i += K2_THREADS) s_lut5[i] = (uint16_t)pg_lut5_entry(i);      // visible after the first barrier
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int64_t w0 = a.w_first + tile * K2_TILE_WORDS;
        __syncthreads();
        if (threadIdx.x == 0) s_cnt = 0;
        stage_tile(a.pk2, a.amb, w0, a.n_words, s_pk, s_am);
        __syncthreads();
        const int64_t g0 = (w0 + threadIdx.x) * 32;
        uint32_t bits = 0;
        if (g0 < a.g_end && g0 + 32 > a.g_begin) {
            PgWindow w;
            w.prv = s_pk[threadIdx.x + 1]; w.cur = s_pk[threadIdx.x + 2]; w.nxt = s_pk[threadIdx.x + 3];
            w.aprv = s_am[threadIdx.x + 3]; w.acur = s_am[threadIdx.x + 4]; w.anxt = s_am[threadIdx.x + 5];
            int64_t r = find_record(a.seq_off, a.n_rec, g0);
            int64_t rs = r >= 0 ? __ldg(a.seq_off + r) : 0, re = __ldg(a.seq_off + r + 1);
            if (pg_is_interior(w, g0, 32, a.k, rs, re, r >= 0, a.g_begin, a.g_end)) {
                pg_interior_visit<32>(w, 0, a.k, a.pow5km1, nullptr, s_lut5, [&](int q, uint64_t F, uint64_t R, uint32_t) {
                    uint64_t so;
                    if (rdbg_hit(rd, mode, strand ? R : F, strand ? F : R, so)) bits |= 1u << q;
                });
            } else {
                uint64_t F, R;
                pg_codes_init(w, 0, a.k, F, R);
#pragma unroll 1
                for (int j = 0; j < 32; j++) {
                    const int64_t g = g0 + j;
                    if (g >= a.g_end) break;
                    while (r + 1 < a.n_rec && g >= re) { r++; re = __ldg(a.seq_off + r + 1); }
                    if (g >= a.g_begin && r >= 0 && g + a.k <= re) {
                        uint64_t so;
                        if (rdbg_hit(rd, mode, strand ? R : F, strand ? F : R, so)) bits |= 1u << j;
                    }
                    pg_codes_roll(w, j, a.k, a.pow5km1, F, R);
                }
            }
        }
        hitbits[w0 - a.w_first + threadIdx.x] = bits;
        uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(bits));
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (threadIdx.x == 0) tile_counts[tile] = s_cnt;
    }
}

// K5 pass 2: ordered emit.  Only hit positions recompute their codes (a few % of all positions).
__global__ void __launch_bounds__(K2_THREADS)
k5_emit(TableView rd, int mode, SeqArgs a, int strand, const uint32_t *__restrict__ hitbits,
        const int64_t *__restrict__ tile_off, int64_t *__restrict__ hit_g, uint64_t *__restrict__ hit_node,
        int32_t *__restrict__ hit_rec, uint16_t *__restrict__ hit_v6, int64_t cap_hits) {
    __shared__ __align__(16) uint64_t s_pk[K2_TILE_WORDS + 4];
    __shared__ __align__(16) uint32_t s_am[K2_TILE_WORDS + 8];
    __shared__ uint32_t s_warp[K2_THREADS / 32];
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int64_t w0 = a.w_first + tile * K2_TILE_WORDS;
        __syncthreads();
        stage_tile(a.pk2, a.amb, w0, a.n_words, s_pk, s_am);
        uint32_t bits = hitbits[w0 - a.w_first + threadIdx.x];
        // block-exclusive rank of my first hit
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t c = __popc(bits), inc = c;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wp = 0;
        for (int q = 0; q < warp; q++) wp += s_warp[q];
        int64_t at = tile_off[tile] + wp + inc - c;
        if (!bits) continue;
        const int64_t g0 = (w0 + threadIdx.x) * 32;
        PgWindow w;
        w.prv = s_pk[threadIdx.x + 1]; w.cur = s_pk[threadIdx.x + 2]; w.nxt = s_pk[threadIdx.x + 3];
        w.aprv = s_am[threadIdx.x + 3]; w.acur = s_am[threadIdx.x + 4]; w.anxt = s_am[threadIdx.x + 5];
        int64_t r = find_record(a.seq_off, a.n_rec, g0);
        int64_t rs = r >= 0 ? __ldg(a.seq_off + r) : 0, re = __ldg(a.seq_off + r + 1);
        while (bits) {
            int j = __ffs((int)bits) - 1; bits &= bits - 1;
            const int64_t g = g0 + j;
            while (r + 1 < a.n_rec && g >= re) { r++; rs = re; re = __ldg(a.seq_off + r + 1); }
            uint64_t F, R;
            pg_codes_init(w, j, a.k, F, R);
            uint64_t so = 0;
            rdbg_hit(rd, mode, strand ? R : F, strand ? F : R, so);
            uint32_t vf, vr;
            pg_occ_vals(w, j, g - rs, re - rs, a.k, vf, vr);
            uint32_t v12 = strand ? vr : vf;                       // (lp << 6) | ln of the walked strand
            uint32_t lp = v12 >> 6, ln = v12 & 63u;
            uint32_t v5 = (lp << 5) | ln;                          // edge stage: offbit 5 (F7)
            if (at < cap_hits) {
                hit_g[at] = g; hit_node[at] = (so << NODE_V_BITS) | v5; hit_rec[at] = (int32_t)r; hit_v6[at] = (uint16_t)v12;
            }
            at++;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-word CAS tables
__device__ __forceinline__ int64_t set_insert(uint64_t *keys, uint64_t capmask, uint64_t key, bool *fresh) {
    uint64_t s = pg_mix64(key) & capmask;
    for (uint64_t probe = 0; probe <= capmask; probe++) {
        uint64_t ck = keys[s];
        if (ck == PG_EMPTY) {
            uint64_t old = atomicCAS(reinterpret_cast<unsigned long long *>(keys + s), (unsigned long long)PG_EMPTY, (unsigned long long)key);
            if (old == PG_EMPTY) { if (fresh) *fresh = true; return (int64_t)s; }
            ck = old;
        }
        if (ck == key) { if (fresh) *fresh = false; return (int64_t)s; }
        s = (s + 1) & capmask;
    }
    return -1;
}
__device__ __forceinline__ int64_t set_find(const uint64_t *keys, uint64_t capmask, uint64_t key) {
    uint64_t s = pg_mix64(key) & capmask;
    for (uint64_t probe = 0; probe <= capmask; probe++) {
        uint64_t ck = keys[s];
        if (ck == key) return (int64_t)s;
        if (ck == PG_EMPTY) return -1;
        s = (s + 1) & capmask;
    }
    return -1;
}

__global__ void k_graph_clear(pg_graph g) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < g.node_cap; i += stride) { g.d_node_keys[i] = PG_EMPTY; g.d_node_parent[i] = (uint32_t)i; g.d_node_label[i] = -1; }
    for (int64_t i = i0; i < g.edge_cap; i += stride) { g.d_edge_keys[i] = PG_EMPTY; g.d_edge_w[i] = 0; g.d_edge_first[i] = PG_EMPTY; }
    for (int64_t i = i0; i < g.visit_cap; i += stride) g.d_visit_keys[i] = PG_EMPTY;
    if (i0 < 8) g.d_stats[i0] = 0;
}

// K6a: every hit -> node slot
__global__ void k6_nodes(pg_graph g, const uint64_t *__restrict__ hit_node, int64_t n, uint32_t *__restrict__ hit_nslot) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool fresh;
    int64_t s = set_insert(g.d_node_keys, (uint64_t)g.node_cap - 1, hit_node[i], &fresh);
    if (s < 0) { atomicExch(reinterpret_cast<unsigned long long *>(g.d_stats), 1ull); s = 0; }
    else if (fresh) atomicAdd(reinterpret_cast<unsigned long long *>(g.d_stats + 1), 1ull);
    hit_nslot[i] = (uint32_t)s;
}
// K6b: consecutive hits of one record-strand -> edge; first sighting per record-strand adds 1 to its weight
// (the `visit` dict, :1479-1484); `first` = smallest walk ordinal (file order of the .xyz)
__global__ void k6_edges(pg_graph g, const uint32_t *__restrict__ hit_nslot, const int32_t *__restrict__ hit_rec, int64_t n,
                         int strand, int n_strands) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    int32_t rec = hit_rec[i];
    if (hit_rec[i + 1] != rec) return;
    // the forward walk goes i -> i+1; the rc-strand walk visits hits in descending position: i+1 -> i
    uint32_t a = strand ? hit_nslot[i + 1] : hit_nslot[i], b = strand ? hit_nslot[i] : hit_nslot[i + 1];
    bool fresh;
    int64_t e = set_insert(g.d_edge_keys, (uint64_t)g.edge_cap - 1, ((uint64_t)a << 32) | b, &fresh);
    if (e < 0) { atomicExch(reinterpret_cast<unsigned long long *>(g.d_stats), 1ull); return; }
    if (fresh) atomicAdd(reinterpret_cast<unsigned long long *>(g.d_stats + 2), 1ull);
    uint64_t rs_id = (uint64_t)rec * n_strands + strand;
    int64_t v = set_insert(g.d_visit_keys, (uint64_t)g.visit_cap - 1, ((uint64_t)e << 32) | rs_id, &fresh);
    if (v < 0) { atomicExch(reinterpret_cast<unsigned long long *>(g.d_stats), 1ull); return; }
    if (fresh) atomicAdd(g.d_edge_w + e, 1u);
    // only nodes that appear in an edge exist upstream (label_dct is built from the .xyz, :1918-1944):
    // -2 = "in the graph, label pending"; hits that never form an edge keep -1 and can never match in K8
    g.d_node_label[a] = -2; g.d_node_label[b] = -2;
    // walk ordinal: records in order, forward strand before rc strand (rdbg_edge_weight_jit_ :1814-1817), then walk order
    uint64_t ord = (rs_id << 36) | (uint64_t)(strand ? (n - 2 - i) : i);
    atomicMin(reinterpret_cast<unsigned long long *>(g.d_edge_first + e), (unsigned long long)ord);
}

// K7: union-find with atomic hooking (smaller slot index becomes the root)
__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x) {
    uint32_t p = parent[x];
    while (p != x) { uint32_t gp = parent[p]; if (gp != p) parent[x] = gp; x = p; p = gp; }   // path halving (benign race)
    return x;
}
__global__ void k7_union(pg_graph g, uint32_t min_weight) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= g.edge_cap) return;
    uint64_t ek = g.d_edge_keys[i];
    if (ek == PG_EMPTY || g.d_edge_w[i] < min_weight) return;
    uint32_t a = (uint32_t)(ek >> 32), b = (uint32_t)ek;
    for (;;) {
        a = uf_find(g.d_node_parent, a); b = uf_find(g.d_node_parent, b);
        if (a == b) break;
        if (a < b) { uint32_t t = a; a = b; b = t; }                  // hook the larger root under the smaller
        if (atomicCAS(g.d_node_parent + a, a, b) == a) break;
    }
}
__global__ void k7_flatten(pg_graph g) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= g.node_cap || g.d_node_keys[i] == PG_EMPTY) return;
    uint32_t r = (uint32_t)i;
    while (g.d_node_parent[r] != r) r = g.d_node_parent[r];
    g.d_node_parent[i] = r;      // racing writers only ever store a valid ancestor chain end
}

// decode a node key into the reference's (code, v5)
__device__ __forceinline__ void node_decode(uint64_t nk, const uint64_t *rd_slots, int64_t rd_cap, int mode, int k, uint64_t &code, uint32_t &v5) {
    v5 = (uint32_t)(nk & ((1u << NODE_V_BITS) - 1u));
    uint64_t so = nk >> NODE_V_BITS;
    uint64_t slot = so >> 1; uint32_t o = (uint32_t)(so & 1u);
    if ((int64_t)slot >= rd_cap) { code = 0; return; }            // phantom key 0 that is in no table
    uint64_t key = rd_slots[2 * slot];
    code = (o && mode == PG_MODE_CANONICAL) ? pg_rc_code(key, k) : key;
}

__global__ void k_export_edges(pg_graph g, const uint64_t *__restrict__ rd_slots, int64_t rd_cap, int mode, int k,
                               uint64_t *c0, uint32_t *v0, uint64_t *c1, uint32_t *v1, uint32_t *w, uint64_t *first,
                               int64_t cap, unsigned long long *n_out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= g.edge_cap) return;
    uint64_t ek = g.d_edge_keys[i];
    if (ek == PG_EMPTY) return;
    unsigned long long at = atomicAdd(n_out, 1ull);
    if ((int64_t)at >= cap) return;
    uint64_t ca, cb; uint32_t va, vb;
    node_decode(g.d_node_keys[(uint32_t)(ek >> 32)], rd_slots, rd_cap, mode, k, ca, va);
    node_decode(g.d_node_keys[(uint32_t)ek], rd_slots, rd_cap, mode, k, cb, vb);
    c0[at] = ca; v0[at] = va; c1[at] = cb; v1[at] = vb; w[at] = g.d_edge_w[i]; first[at] = g.d_edge_first[i];
}
__global__ void k_export_nodes(pg_graph g, const uint64_t *__restrict__ rd_slots, int64_t rd_cap, int mode, int k,
                               uint32_t *nslot, uint64_t *code, uint32_t *v5, uint32_t *root, int64_t cap, unsigned long long *n_out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= g.node_cap) return;
    uint64_t nk = g.d_node_keys[i];
    if (nk == PG_EMPTY || g.d_node_label[i] == -1) return;     // slot empty / hit that is in no edge
    unsigned long long at = atomicAdd(n_out, 1ull);
    if ((int64_t)at >= cap) return;
    uint64_t c; uint32_t v;
    node_decode(nk, rd_slots, rd_cap, mode, k, c, v);
    nslot[at] = (uint32_t)i; code[at] = c; v5[at] = v; root[at] = g.d_node_parent[i];
}

// ---------------------------------------------------------------------------------------------
// K8.  label_dct lookup uses offbit 6 (seq2path_jit_ :1541-1549) against keys built with offbit 5:
// a hit matches iff the NUMBER (lp << 6) | ln equals some node's v5 for the same code (Q7).
__global__ void k8_match(pg_graph g, const uint64_t *__restrict__ hit_node, const uint16_t *__restrict__ hit_v6, int64_t n,
                         int32_t *__restrict__ hit_label, uint32_t *__restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t nk = (hit_node[i] & ~(uint64_t)((1u << NODE_V_BITS) - 1u)) | hit_v6[i];
    int32_t lab = -1;
    if (hit_v6[i] < (1u << NODE_V_BITS)) {
        int64_t s = set_find(g.d_node_keys, (uint64_t)g.node_cap - 1, nk);
        if (s >= 0) lab = g.d_node_label[s];
    }
    hit_label[i] = lab;
    flag[i] = lab >= 0;
}
// Compact the matched hits in WALK order: ascending position on the forward strand; on the rc strand
// the walk runs through descending forward positions, so the order is reversed and positions become
// q = n - k - p (seqs2path_jit_ walks reverse_jit_(seq), :1840-1845).  m_g is a coordinate that grows
// along the walk with the same spacing as the positions; m_p the position inside the walked strand.
__global__ void k8_compact(const uint32_t *__restrict__ flag, const int64_t *__restrict__ off, int64_t n,
                           const int64_t *__restrict__ hit_g, const int32_t *__restrict__ hit_rec, const int32_t *__restrict__ hit_label,
                           const int64_t *__restrict__ seq_off, int k, int strand, const int64_t *__restrict__ m_total,
                           int64_t *__restrict__ m_g, int64_t *__restrict__ m_p, int32_t *__restrict__ m_rec, int32_t *__restrict__ m_label) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    int64_t o = off[i], g = hit_g[i];
    int32_t rec = hit_rec[i];
    if (strand) {
        o = *m_total - 1 - o;
        m_g[o] = ((int64_t)1 << 62) - g;
        m_p[o] = seq_off[rec + 1] - k - g;
    } else {
        m_g[o] = g;
        m_p[o] = g - seq_off[rec];
    }
    m_rec[o] = rec; m_label[o] = hit_label[i];
}
// next[i] = first matched hit j of the same record-strand with pos_j > pos_i + k ("starts[-1] < idx", :1552);
// reach[i] = 1 for the first matched hit of each record-strand with pos > 0 (starts = [0] initially)
__global__ void k8_next(const int64_t *__restrict__ m_g, const int64_t *__restrict__ m_p, const int32_t *__restrict__ m_rec, int64_t m, int k,
                        uint32_t *__restrict__ next, uint32_t *__restrict__ reach) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t g = m_g[i]; int32_t rec = m_rec[i];
    int64_t lo = i + 1, hi = m;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (m_g[mid] > g + k) hi = mid; else lo = mid + 1; }
    next[i] = (lo < m && m_rec[lo] == rec) ? (uint32_t)lo : NONE32;
    bool first_of_rec = (i == 0) || (m_rec[i - 1] != rec);
    uint32_t r = 0;
    if (first_of_rec) r = m_p[i] > 0;
    else if (m_rec[i - 1] == rec && (i == 1 || m_rec[i - 2] != rec) && m_p[i - 1] == 0) r = 1;   // the record's first hit sat at position 0
    reach[i] = r;
}
__global__ void k8_jump(const uint32_t *__restrict__ jin, uint32_t *__restrict__ jout, uint32_t *__restrict__ reach, int64_t m) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t j = jin[i];
    if (j != NONE32) { if (reach[i]) reach[j] = 1; jout[i] = jin[j]; } else jout[i] = NONE32;
}
__global__ void k8_run_ends(const uint32_t *__restrict__ next, const uint32_t *__restrict__ reach, const int32_t *__restrict__ m_label,
                            int64_t m, uint32_t *__restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t j = next[i];
    flag[i] = reach[i] && (j == NONE32 || m_label[j] != m_label[i]);
}
__global__ void k8_rows(const uint32_t *__restrict__ flag, const int64_t *__restrict__ off, int64_t m, const int64_t *__restrict__ m_p,
                        const int32_t *__restrict__ m_rec, const int32_t *__restrict__ m_label, int k,
                        int32_t *__restrict__ row_rec, int64_t *__restrict__ row_end, int32_t *__restrict__ row_label, int64_t cap_rows) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m || !flag[i]) return;
    int64_t o = off[i];
    if (o >= cap_rows) return;
    int32_t rec = m_rec[i];
    row_rec[o] = rec; row_end[o] = m_p[i] + k; row_label[o] = m_label[i];
}

// ---- raw slot transfer + rank-independent hits (multi-GPU path stages) --------------------------
__global__ void k_export_raw(const uint64_t *__restrict__ slots, int64_t cap, uint64_t tag, uint64_t *keys, uint64_t *vals, int64_t out_cap,
                             unsigned long long *n_out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= cap) return;
    uint64_t k, v;
    pg_ld_slot(slots + 2 * i, tag, k, v);              // v comes back without the generation tag
    if (k == PG_EMPTY) return;
    unsigned long long at = atomicAdd(n_out, 1ull);
    if ((int64_t)at < out_cap) { keys[at] = k; vals[at] = v; }
}
__global__ void k_insert_raw(TableView t, const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    table_put_or(t, keys[i], vals[i]);
}
// node key (local rdBG slot, orientation, v5) -> the reference's literal code: what another rank can re-key
__global__ void k_hits_decode(const uint64_t *__restrict__ hit_node, int64_t n, const uint64_t *__restrict__ rd_slots, int64_t rd_cap,
                              int mode, int k, uint64_t *__restrict__ hit_code) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t c; uint32_t v;
    node_decode(hit_node[i], rd_slots, rd_cap, mode, k, c, v);
    hit_code[i] = c;
}
// literal code + v5 -> node key of THIS rank's rdBG table (the code is a member or the phantom key 0 by construction)
__global__ void k_hits_rekey(TableView rd, int mode, int k, const uint64_t *__restrict__ hit_code, const uint32_t *__restrict__ hit_v5,
                             int64_t n, uint64_t *__restrict__ hit_node) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t lit = hit_code[i], other = mode == PG_MODE_CANONICAL ? pg_rc_code(lit, k) : lit, so = 0;
    if (!rdbg_hit(rd, mode, lit, other, so)) so = ((rd.capmask + 1) << 1);      // cannot happen for gathered hits; keep it total
    hit_node[i] = (so << NODE_V_BITS) | hit_v5[i];
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads > 0 ? (n + threads - 1) / threads : 1); }

int check_graph(const pg_graph *g, const char *who) {
    if (!g || !g->d_node_keys || !g->d_node_parent || !g->d_node_label || !g->d_edge_keys || !g->d_edge_w || !g->d_edge_first ||
        !g->d_visit_keys || !g->d_stats) return pg_fail(PG_ERR_INVALID, "%s: null graph buffer", who);
    int64_t caps[3] = {g->node_cap, g->edge_cap, g->visit_cap};
    for (int i = 0; i < 3; i++)
        if (caps[i] < 2 || (caps[i] & (caps[i] - 1)) || caps[i] > (1ll << 32))
            return pg_fail(PG_ERR_INVALID, "%s: capacities must be powers of two in [2, 2^32]", who);
    return PG_OK;
}

int check_tab(const pg_table *t, const char *who) {
    if (!t || !t->d_slots || !t->d_stats || t->capacity < 2 || (t->capacity & (t->capacity - 1)) || t->k < 1 || t->k > 27 ||
        t->epoch < 1 || t->epoch > PG_EPOCH_MAX)
        return pg_fail(PG_ERR_INVALID, "%s: bad table", who);
    return PG_OK;
}

}  // namespace

extern "C" int64_t pg_path_workspace_bytes(int64_t n_bases) {
    int64_t words = n_bases / 32 + K2_TILE_WORDS + 8;
    int64_t tiles = words / K2_TILE_WORDS + 2;
    // hit bitmap (one word per 32 bases, whole tiles) + per-tile counts + per-tile offsets + scan block sums
    return tiles * K2_TILE_WORDS * 4 + tiles * 4 + 16 + tiles * 8 + (tiles / SCAN_THREADS + 2) * 8 + 256;
}

extern "C" int pg_path_hits(const pg_table *rdbg, const uint32_t *d_pk2, const uint32_t *d_amb, const int64_t *d_seq_off,
                            int64_t n_rec, int64_t g_begin, int64_t g_end, int strand, int64_t *d_hit_g, uint64_t *d_hit_node,
                            int32_t *d_hit_rec, uint16_t *d_hit_v6, int64_t cap_hits, int64_t *d_n_hits, void *d_ws,
                            int64_t ws_bytes, pg_stream_t stream_) {
    int rc = check_tab(rdbg, "pg_path_hits"); if (rc) return rc;
    if (!d_pk2 || !d_amb || !d_seq_off || !d_hit_g || !d_hit_node || !d_hit_rec || !d_hit_v6 || !d_n_hits || !d_ws || n_rec < 0 ||
        g_begin < 0 || g_end < g_begin || cap_hits < 0 || strand < 0 || strand > 1)
        return pg_fail(PG_ERR_INVALID, "pg_path_hits: bad arguments");
    if (rdbg->capacity > (1ll << 40)) return pg_fail(PG_ERR_INVALID, "pg_path_hits: rdBG capacity too large for node keys");
    if (ws_bytes < pg_path_workspace_bytes(g_end)) return pg_fail(PG_ERR_WORKSPACE, "pg_path_hits: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_rec == 0 || g_end == g_begin) { PG_CUDA(cudaMemsetAsync(d_n_hits, 0, 8, st)); return PG_OK; }
    SeqArgs a;
    a.pk2 = reinterpret_cast<const uint64_t *>(d_pk2); a.amb = d_amb; a.seq_off = d_seq_off; a.n_rec = n_rec;
    a.g_begin = g_begin; a.g_end = g_end; a.k = rdbg->k; a.pow5km1 = pg_pow5(rdbg->k - 1);
    a.w_first = (g_begin >> 5) & ~(int64_t)3;
    int64_t w_last = (g_end + 31) >> 5;
    a.n_tiles = (w_last - a.w_first + K2_TILE_WORDS - 1) / K2_TILE_WORDS;
    a.n_words = ((g_end + 31) >> 5) + 4;
    char *ws = reinterpret_cast<char *>(d_ws);
    uint32_t *hitbits = reinterpret_cast<uint32_t *>(ws);
    int64_t nbits_words = a.n_tiles * K2_TILE_WORDS;
    uint32_t *tile_counts = hitbits + nbits_words;
    int64_t *tile_off = reinterpret_cast<int64_t *>(ws + ((nbits_words + a.n_tiles) * 4 + 15) / 16 * 16);
    int64_t *bsum = tile_off + a.n_tiles;
    TableView rd = make_view(rdbg);
    int grid = (int)(a.n_tiles < (int64_t)pg_num_sms() * 8 ? a.n_tiles : (int64_t)pg_num_sms() * 8);
    k5_mark<<<grid, K2_THREADS, 0, st>>>(rd, rdbg->mode, a, strand, hitbits, tile_counts);
    exclusive_scan(tile_counts, a.n_tiles, tile_off, bsum, d_n_hits, st);
    k5_emit<<<grid, K2_THREADS, 0, st>>>(rd, rdbg->mode, a, strand, hitbits, tile_off, d_hit_g, d_hit_node, d_hit_rec, d_hit_v6, cap_hits);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_graph_clear(const pg_graph *g, pg_stream_t stream_) {
    int rc = check_graph(g, "pg_graph_clear"); if (rc) return rc;
    int64_t n = g->node_cap > g->edge_cap ? g->node_cap : g->edge_cap;
    if (g->visit_cap > n) n = g->visit_cap;
    int64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    k_graph_clear<<<(unsigned)b, 256, 0, (cudaStream_t)stream_>>>(*g);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_graph_add_hits(const pg_graph *g, const uint64_t *d_hit_node, const int32_t *d_hit_rec, int64_t n_hits,
                                 uint32_t *d_hit_nslot, int strand, int n_strands, pg_stream_t stream_) {
    int rc = check_graph(g, "pg_graph_add_hits"); if (rc) return rc;
    if (n_hits < 0 || (n_hits > 0 && (!d_hit_node || !d_hit_rec || !d_hit_nslot))) return pg_fail(PG_ERR_INVALID, "pg_graph_add_hits: bad arguments");
    if (n_hits == 0) return PG_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    k6_nodes<<<blocks_for(n_hits, 256), 256, 0, st>>>(*g, d_hit_node, n_hits, d_hit_nslot);
    k6_edges<<<blocks_for(n_hits, 256), 256, 0, st>>>(*g, d_hit_nslot, d_hit_rec, n_hits, strand, n_strands);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_graph_components(const pg_graph *g, uint32_t min_weight, pg_stream_t stream_) {
    int rc = check_graph(g, "pg_graph_components"); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    k7_union<<<blocks_for(g->edge_cap, 256), 256, 0, st>>>(*g, min_weight);
    k7_flatten<<<blocks_for(g->node_cap, 256), 256, 0, st>>>(*g);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_graph_export_edges(const pg_graph *g, const pg_table *rdbg, uint64_t *d_c0, uint32_t *d_v0, uint64_t *d_c1,
                                     uint32_t *d_v1, uint32_t *d_w, uint64_t *d_first, int64_t cap, int64_t *d_n, pg_stream_t stream_) {
    int rc = check_graph(g, "pg_graph_export_edges"); if (rc) return rc;
    rc = check_tab(rdbg, "pg_graph_export_edges"); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_n, 0, 8, st));
    k_export_edges<<<blocks_for(g->edge_cap, 256), 256, 0, st>>>(*g, rdbg->d_slots, rdbg->capacity, rdbg->mode, rdbg->k, d_c0, d_v0, d_c1, d_v1,
                                                               d_w, d_first, cap, reinterpret_cast<unsigned long long *>(d_n));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_graph_export_nodes(const pg_graph *g, const pg_table *rdbg, uint32_t *d_nslot, uint64_t *d_code, uint32_t *d_v5,
                                     uint32_t *d_root, int64_t cap, int64_t *d_n, pg_stream_t stream_) {
    int rc = check_graph(g, "pg_graph_export_nodes"); if (rc) return rc;
    rc = check_tab(rdbg, "pg_graph_export_nodes"); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_n, 0, 8, st));
    k_export_nodes<<<blocks_for(g->node_cap, 256), 256, 0, st>>>(*g, rdbg->d_slots, rdbg->capacity, rdbg->mode, rdbg->k, d_nslot, d_code, d_v5,
                                                               d_root, cap, reinterpret_cast<unsigned long long *>(d_n));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int64_t pg_label_workspace_bytes(int64_t n_hits) {
    int64_t n = n_hits + 8;
    // hit_label i32, flag u32, off i64, m_g i64, m_rec i32, m_label i32, next u32 x2, reach u32, bsum
    return n * (4 + 4 + 8 + 8 + 8 + 4 + 4 + 4 + 4 + 4) + (n / SCAN_THREADS + 2) * 8 + 1024;
}

extern "C" int pg_label_regions(const pg_graph *g, const int64_t *d_hit_g, const uint64_t *d_hit_node, const int32_t *d_hit_rec,
                                const uint16_t *d_hit_v6, int64_t n_hits, const int64_t *d_seq_off, int k, int strand,
                                int32_t *d_row_rec, int64_t *d_row_end, int32_t *d_row_label, int64_t cap_rows, int64_t *d_n_rows,
                                void *d_ws, int64_t ws_bytes, pg_stream_t stream_) {
    int rc = check_graph(g, "pg_label_regions"); if (rc) return rc;
    if (n_hits < 0 || !d_n_rows || !d_ws || !d_seq_off || k < 1 || k > 27) return pg_fail(PG_ERR_INVALID, "pg_label_regions: bad arguments");
    if (ws_bytes < pg_label_workspace_bytes(n_hits)) return pg_fail(PG_ERR_WORKSPACE, "pg_label_regions: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_n_rows, 0, 16, st));
    if (n_hits == 0) return PG_OK;
    if (n_hits >= (1ll << 32) - 1) return pg_fail(PG_ERR_CAPACITY, "pg_label_regions: more than 2^32 hits per call");
    int64_t n = n_hits + 8;
    char *p = reinterpret_cast<char *>(d_ws);
    auto take = [&](int64_t bytes) { char *q = p; p += (bytes + 15) / 16 * 16; return q; };
    int64_t *off = reinterpret_cast<int64_t *>(take(n * 8));
    int64_t *m_g = reinterpret_cast<int64_t *>(take(n * 8));
    int64_t *m_p = reinterpret_cast<int64_t *>(take(n * 8));
    int64_t *bsum = reinterpret_cast<int64_t *>(take((n / SCAN_THREADS + 2) * 8));
    int32_t *hit_label = reinterpret_cast<int32_t *>(take(n * 4));
    uint32_t *flag = reinterpret_cast<uint32_t *>(take(n * 4));
    int32_t *m_rec = reinterpret_cast<int32_t *>(take(n * 4));
    int32_t *m_label = reinterpret_cast<int32_t *>(take(n * 4));
    uint32_t *nxt = reinterpret_cast<uint32_t *>(take(n * 4));
    uint32_t *ja = reinterpret_cast<uint32_t *>(take(n * 4));
    uint32_t *reach = reinterpret_cast<uint32_t *>(take(n * 4));
    // d_n_rows[1] receives the number of matched hits (device-side), d_n_rows[0] the rows
    unsigned nb = blocks_for(n_hits, 256);
    k8_match<<<nb, 256, 0, st>>>(*g, d_hit_node, d_hit_v6, n_hits, hit_label, flag);
    exclusive_scan(flag, n_hits, off, bsum, d_n_rows + 1, st);
    k8_compact<<<nb, 256, 0, st>>>(flag, off, n_hits, d_hit_g, d_hit_rec, hit_label, d_seq_off, k, strand, d_n_rows + 1, m_g, m_p, m_rec, m_label);
    // the matched count sizes the remaining launches: one small D2H + stream synchronisation
    PG_CUDA(cudaGetLastError());
    int64_t m_host = 0;
    PG_CUDA(cudaMemcpyAsync(&m_host, d_n_rows + 1, 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    if (m_host == 0) return PG_OK;
    unsigned mb = blocks_for(m_host, 256);
    k8_next<<<mb, 256, 0, st>>>(m_g, m_p, m_rec, m_host, k, nxt, reach);
    // pointer doubling: after r rounds every node within 2^r chain steps of a start is marked
    PG_CUDA(cudaMemcpyAsync(ja, nxt, (size_t)m_host * 4, cudaMemcpyDeviceToDevice, st));
    uint32_t *jin = ja, *jout = reinterpret_cast<uint32_t *>(hit_label);   // hit_label is free after k8_compact
    int rounds = 1; while ((1ll << rounds) < m_host) rounds++;
    for (int r = 0; r <= rounds; r++) {
        k8_jump<<<mb, 256, 0, st>>>(jin, jout, reach, m_host);
        uint32_t *t = jin; jin = jout; jout = t;
    }
    k8_run_ends<<<mb, 256, 0, st>>>(nxt, reach, m_label, m_host, flag);
    exclusive_scan(flag, m_host, off, bsum, d_n_rows, st);
    k8_rows<<<mb, 256, 0, st>>>(flag, off, m_host, m_p, m_rec, m_label, k, d_row_rec, d_row_end, d_row_label, cap_rows);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_table_export_raw(const pg_table *t, uint64_t *d_keys, uint64_t *d_vals, int64_t cap, int64_t *d_n, pg_stream_t stream_) {
    int rc = check_tab(t, "pg_table_export_raw"); if (rc) return rc;
    if (!d_keys || !d_vals || !d_n || cap < 0) return pg_fail(PG_ERR_INVALID, "pg_table_export_raw: bad arguments");
    cudaStream_t st = (cudaStream_t)stream_;
    PG_CUDA(cudaMemsetAsync(d_n, 0, 8, st));
    k_export_raw<<<blocks_for(t->capacity, 256), 256, 0, st>>>(t->d_slots, t->capacity, pg_tag(t), d_keys, d_vals, cap, reinterpret_cast<unsigned long long *>(d_n));
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_table_insert_raw(const pg_table *t, const uint64_t *d_keys, const uint64_t *d_vals, int64_t n, pg_stream_t stream_) {
    int rc = check_tab(t, "pg_table_insert_raw"); if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_keys || !d_vals))) return pg_fail(PG_ERR_INVALID, "pg_table_insert_raw: bad arguments");
    if (n == 0) return PG_OK;
    k_insert_raw<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(make_view(t), d_keys, d_vals, n);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_hits_decode(const pg_table *rdbg, const uint64_t *d_hit_node, int64_t n, uint64_t *d_hit_code, pg_stream_t stream_) {
    int rc = check_tab(rdbg, "pg_hits_decode"); if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_hit_node || !d_hit_code))) return pg_fail(PG_ERR_INVALID, "pg_hits_decode: bad arguments");
    if (n == 0) return PG_OK;
    k_hits_decode<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(d_hit_node, n, rdbg->d_slots, rdbg->capacity, rdbg->mode, rdbg->k, d_hit_code);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_hits_rekey(const pg_table *rdbg, const uint64_t *d_hit_code, const uint32_t *d_hit_v5, int64_t n, uint64_t *d_hit_node,
                             pg_stream_t stream_) {
    int rc = check_tab(rdbg, "pg_hits_rekey"); if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_hit_code || !d_hit_v5 || !d_hit_node))) return pg_fail(PG_ERR_INVALID, "pg_hits_rekey: bad arguments");
    if (n == 0) return PG_OK;
    k_hits_rekey<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(make_view(rdbg), rdbg->mode, rdbg->k, d_hit_code, d_hit_v5, n, d_hit_node);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
