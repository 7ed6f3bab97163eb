// host_emul.cu - CPU emulation of the K1/K2/K4 kernels' per-thread logic, built from the SAME
// host+device functions the kernels use (fasta_chunk.cuh, kmer_core.cuh).  Test infrastructure for
// the `-m "not gpu"` suite: it lets the quirk-heavy integer logic be checked against the oracle in a
// container without a GPU.  It does not emulate warps/CTAs/atomics - the `-m gpu` tests cover those.
#include <unordered_map>
#include <vector>
#include <cstring>
#include "../pangenome_b200/csrc/fasta_chunk.cuh"
#include "../pangenome_b200/csrc/kmer_core.cuh"

thread_local char pg_err_buf[512] = "";
int pg_fail(int code, const char *, ...) { return code; }
int pg_num_sms() { return 148; }
void pg_tune_once() {}

static const int TILE = 16384;

static ChunkCls load_cls(const uint8_t *f, int64_t n, int64_t off) {
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t left = n - off;
    for (int i = 0; i < 16 && i < left; i++) w[i >> 2] |= (uint32_t)f[off + i] << (8 * (i & 3));
    int n_file = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
    return classify16(w, n_file, left <= 16);
}

// mirrors k1_tile_summaries + k1_tile_scan + k1_tile_pack
extern "C" int64_t emul_pack(const uint8_t *fasta, int64_t nbytes, uint32_t *pk2, uint32_t *amb, int64_t n_words,
                             int64_t *hdr_off, int64_t *seq_off, int64_t cap_rec, int64_t *counts) {
    memset(pk2, 0, (size_t)n_words * 4); memset(amb, 0, (size_t)n_words * 4);
    int64_t ntiles = (nbytes + TILE - 1) / TILE;
    std::vector<Sum3> sums((size_t)ntiles);
    uint64_t real_nl = 0;
    for (int64_t t = 0; t < ntiles; t++) {
        Sum3 carry = sum3_identity();
        for (int c = 0; c < TILE / 16; c++) {
            ChunkCls cls = load_cls(fasta, nbytes, t * TILE + c * 16);
            real_nl += cls.real_nl;
            carry = sum3_compose(carry, chunk_sum3(cls));
        }
        sums[(size_t)t] = carry;
    }
    bool dead = real_nl == 0;
    uint64_t seq = 0, hdr = 0; uint32_t state = ST_LINE_START;
    for (int64_t t = 0; t < ntiles && !dead; t++) {
        uint64_t tseq = seq, thdr = hdr; uint32_t cstate = state, cseq = 0, chdr = 0;
        Sum3 excl = sum3_identity();
        for (int c = 0; c < TILE / 16; c++) {
            int64_t off = t * TILE + c * 16;
            ChunkCls cls = load_cls(fasta, nbytes, off);
            uint32_t x = sum3_sel(excl, cstate);
            ChunkRun r = chunk_run(cls, SV_STATE(x));
            uint32_t rank = cseq + SV_SEQ(x), cnt = pg_popc(r.seqmask);
            if (cnt) {
                uint32_t d = pext16_2bit(cls.dig, r.seqmask), m = pext16_1bit(cls.amb, r.seqmask);
                if (cnt < 16) d &= (1u << (2 * cnt)) - 1u;
                uint64_t g = tseq + rank;
                uint32_t sh = 2 * (g & 15);
                pk2[g >> 4] |= d << sh;
                if (sh && (d >> (32 - sh))) pk2[(g >> 4) + 1] |= d >> (32 - sh);
                uint32_t sh1 = g & 31;
                amb[g >> 5] |= m << sh1;
                if (sh1 > 16 && (m >> (32 - sh1))) amb[(g >> 5) + 1] |= m >> (32 - sh1);
            }
            uint32_t hs = r.hs; uint64_t idx = thdr + chdr + SV_HDR(x);
            while (hs) {
                int j = pg_ctz(hs); hs &= hs - 1;
                if ((int64_t)idx < cap_rec) { hdr_off[idx] = off + j; seq_off[idx] = (int64_t)(tseq + rank + pg_popc(r.seqmask & ((1u << j) - 1u))); }
                idx++;
            }
            excl = sum3_compose(excl, chunk_sum3(cls));
        }
        uint32_t y = sum3_sel(sums[(size_t)t], state);
        seq += SV_SEQ(y); hdr += SV_HDR(y); state = SV_STATE(y);
        (void)cseq; (void)chdr;
    }
    counts[0] = dead ? 0 : (int64_t)hdr; counts[1] = dead ? 0 : (int64_t)seq; counts[2] = (int64_t)real_nl;
    if (!dead && (int64_t)hdr <= cap_rec) seq_off[hdr] = (int64_t)seq;
    if (dead) seq_off[0] = 0;
    return counts[1];
}

struct Slot { uint32_t masks; uint64_t cnt; };

// 1: interior positions go through the compact 8-byte record (pg_interior_visit_c -> pg_crec_pack -> what K3s-c does with
// it: value word from the context bits, base-5 key from the 2-bit code) instead of the base-5 fast path
static int g_compact = 0;
static int64_t g_compact_bad = 0;
extern "C" void emul_set_compact(int on) { g_compact = on; g_compact_bad = 0; }
extern "C" int64_t emul_compact_bad() { return g_compact_bad; }

// mirrors k2_kmer_insert (one "thread" per 32-base word) + pg_table_export / pg_rdbg_export
extern "C" int64_t emul_dbg(const uint32_t *pk2_32, const uint32_t *amb, int64_t n_words32, const int64_t *seq_off, int64_t n_rec,
                            int k, int mode, uint64_t *keys, uint16_t *vals, uint8_t *cnts, int64_t cap,
                            uint64_t *rkeys, uint16_t *rvals, int64_t rcap, int64_t *n_rdbg) {
    const uint64_t *pk2 = reinterpret_cast<const uint64_t *>(pk2_32);
    int64_t n_words = n_words32 / 2;
    std::unordered_map<uint64_t, Slot> tab;
    auto upsert = [&](uint64_t key, uint32_t masks, uint32_t inc) { Slot &s = tab[key]; s.masks |= masks; s.cnt += inc; };
    int64_t g_begin = 0, g_end = seq_off[n_rec];
    int strands = mode == PG_MODE_LITERAL ? 1 : 2;
    int64_t n_short = 0;
    for (int64_t i = 0; i < n_rec; i++) if (seq_off[i + 1] - seq_off[i] < k) n_short += strands;
    uint64_t p5 = pg_pow5(k - 1);
    for (int64_t wi = 0; wi * 32 < g_end; wi++) {
        PgWindow w;
        auto W = [&](int64_t x) -> uint64_t { return (x >= 0 && x < n_words) ? pk2[x] : 0; };
        auto A = [&](int64_t x) -> uint32_t { return (x >= 0 && x < n_words32) ? amb[x] : 0; };
        w.prv = W(wi - 1); w.cur = W(wi); w.nxt = W(wi + 1);
        w.aprv = A(wi - 1); w.acur = A(wi); w.anxt = A(wi + 1);
        int64_t g0 = wi * 32;
        int64_t lo = 0, hi = n_rec;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (seq_off[mid] <= g0) lo = mid + 1; else hi = mid; }
        int64_t r = lo - 1;
        int64_t rs = r >= 0 ? seq_off[r] : 0, re = seq_off[r + 1];
        if (pg_is_interior(w, g0, 32, k, rs, re, r >= 0, g_begin, g_end)) {     // the kernels' fast path
            uint32_t vlut[16];                       // the kernels keep these tables in shared memory
            for (uint32_t i = 0; i < 16; i++) vlut[i] = pg_vlut_entry(i);
            static uint16_t lut5[PG_LUT5_SIZE];
            static bool lut_ready = false;
            if (!lut_ready) { for (uint32_t i = 0; i < PG_LUT5_SIZE; i++) lut5[i] = (uint16_t)pg_lut5_entry(i); lut_ready = true; }
            if (g_compact && mode == PG_MODE_CANONICAL) {
                auto take = [&](int, uint64_t F2, uint64_t R2, uint32_t ctx4) {
                    const uint64_t rec = pg_crec_pack(F2, R2, ctx4);
                    const uint64_t key2 = rec & PG_C_KEYMASK;
                    const uint32_t ctx = (uint32_t)(rec >> PG_C_KEYBITS);
                    uint32_t masks, inc;
                    pg_crec_vals(ctx, vlut[ctx & 15u], masks, inc);
                    const uint64_t key5 = pg_code5_of2(key2, lut5);
                    uint64_t back;
                    if (key5 != pg_code5_of2_loop(key2, k) || !pg_code2_of5(key5, k, back) || back != key2 ||
                        pg_hash_kind1(key5, k) != pg_mix64(key2) || (rec >> 60) != 0)
                        g_compact_bad++;
                    upsert(key5, masks, inc);
                };
                // the compile-time-k form K2a-c runs for the common k (two 16-position halves), else the rolling form
                if (g_compact == 2 && k == 27) { pg_interior_visit_ck<27>(w, 0, take); pg_interior_visit_ck<27>(w, 16, take); }
                else if (g_compact == 2 && k == 21) { pg_interior_visit_ck<21>(w, 0, take); pg_interior_visit_ck<21>(w, 16, take); }
                else if (g_compact == 2 && k == 17) { pg_interior_visit_ck<17>(w, 0, take); pg_interior_visit_ck<17>(w, 16, take); }
                else if (g_compact == 2 && k == 26) { pg_interior_visit_ck<26>(w, 0, take); pg_interior_visit_ck<26>(w, 16, take); }
                else pg_interior_visit_c<32>(w, 0, k, take);
                continue;
            }
            // alternate between the table-driven and the loop forms of both helpers
            pg_interior_visit<32>(w, 0, k, p5, (wi & 1) ? vlut : nullptr, (wi & 2) ? nullptr : lut5, [&](int, uint64_t F, uint64_t R, uint32_t vw) {
                if (mode == PG_MODE_CANONICAL) { PgUpdate u = pg_canonical_update_w(F, R, vw); upsert(u.key, u.masks, u.inc); }
                else { upsert(F, vw & 0xFFFFu, 1); if (mode == PG_MODE_LITERAL_RC) upsert(R, vw >> 16, 1); }
            });
            continue;
        }
        uint64_t F, R;
        pg_codes_init(w, 0, k, F, R);
        for (int j = 0; j < 32; j++) {
            int64_t g = g0 + j;
            if (g >= g_end) break;
            while (r + 1 < n_rec && g >= re) { r++; rs = re; re = seq_off[r + 1]; }
            if (g >= g_begin && r >= 0 && g + k <= re) {
                uint32_t vf, vr;
                pg_occ_vals(w, j, g - rs, re - rs, k, vf, vr);
                if (mode == PG_MODE_CANONICAL) { PgUpdate u = pg_canonical_update(F, R, vf, vr); upsert(u.key, u.masks, u.inc); }
                else { upsert(F, vf, 1); if (mode == PG_MODE_LITERAL_RC) upsert(R, vr, 1); }
            }
            pg_codes_roll(w, j, k, p5, F, R);
        }
    }
    int64_t n = 0, nr = 0;
    for (auto &kv : tab) {
        uint64_t c = kv.second.cnt > 0xFFFFFFFFull ? 0xFFFFFFFFull : kv.second.cnt;
        uint64_t v = (uint64_t)kv.second.masks | (c << 32);
        PgEntry e[2];
        int m = pg_slot_entries(kv.first, v, mode, k, e);
        for (int q = 0; q < m; q++) { if (n < cap) { keys[n] = e[q].key; vals[n] = (uint16_t)e[q].val; cnts[n] = (uint8_t)e[q].cnt; } n++; }
        uint32_t f = pg_rdbg_flags(kv.first, v, mode, k);
        if (f & 1u) { if (nr < rcap) { rkeys[nr] = e[0].key; rvals[nr] = (uint16_t)e[0].val; } nr++; }
        if (f & 2u) { if (nr < rcap) { rkeys[nr] = e[1].key; rvals[nr] = (uint16_t)e[1].val; } nr++; }
    }
    if (n_short > 0) {
        if (n < cap) { keys[n] = PG_EMPTY; vals[n] = 32; cnts[n] = (uint8_t)(n_short < 255 ? n_short : 255); } n++;
        if (nr < rcap) { rkeys[nr] = PG_EMPTY; rvals[nr] = 32; } nr++;
    }
    *n_rdbg = nr;
    return n;
}


// mirrors k1_fused_pack: per tile, states from a max-scan of the last newline, tile summary from
// entry-independent quantities (tile_sum3), then ranks from a plain prefix sum
extern "C" int64_t emul_pack2(const uint8_t *fasta, int64_t nbytes, uint32_t *pk2, uint32_t *amb, int64_t n_words,
                              int64_t *hdr_off, int64_t *seq_off, int64_t cap_rec, int64_t *counts) {
    memset(pk2, 0, (size_t)n_words * 4); memset(amb, 0, (size_t)n_words * 4);
    const int NCH = TILE / 16;
    int64_t ntiles = (nbytes + TILE - 1) / TILE;
    uint64_t seq = 0, hdr = 0, real_nl = 0; uint32_t state = ST_LINE_START;
    for (int64_t t = 0; t < ntiles; t++) {
        std::vector<ChunkCls> cls(NCH); std::vector<uint16_t> gt(NCH); std::vector<int> prev(NCH);
        int run = -1, first_nl = 0x7FFFFFFF, last_nl = -1;
        for (int c = 0; c < NCH; c++) {
            cls[c] = load_cls(fasta, nbytes, t * TILE + c * 16);
            gt[c] = (uint16_t)cls[c].gt; real_nl += cls[c].real_nl;
            prev[c] = run;
            int l = chunk_last_nl(cls[c].nl, c * 16); if (l > run) run = l;
            int f = chunk_first_nl(cls[c].nl, c * 16); if (f < first_nl) first_nl = f;
        }
        last_nl = run;
        TileLocal tl; tl.post_seq = tl.post_hdr = 0; tl.first_nl = first_nl; tl.last_nl = last_nl; tl.tile_len = TILE;
        tl.first_gt = gt[0] & 1u;
        tl.gt_after_last = (last_nl >= 0 && last_nl + 1 < TILE) ? ((gt[(last_nl + 1) >> 4] >> ((last_nl + 1) & 15)) & 1u) : 0u;
        for (int c = 0; c < NCH; c++) {
            ChunkRun r = chunk_run(cls[c], chunk_entry_state(prev[c], c * 16, gt.data(), ST_HEADER));
            tl.post_seq += pg_popc(r.seqmask); tl.post_hdr += pg_popc(r.hs);
        }
        Sum3 agg = tile_sum3(tl);
        // ---- entry known (sequential emulation of the look-back): final states, ranks, pack
        const uint32_t fallback = state == ST_LINE_START ? (tl.first_gt ? (uint32_t)ST_HEADER : (uint32_t)ST_SEQ) : state;
        uint32_t rank = 0, hrank = 0;
        for (int c = 0; c < NCH; c++) {
            int64_t off = t * TILE + c * 16;
            uint32_t st = (prev[c] < 0 && c == 0 && state == ST_LINE_START) ? (uint32_t)ST_LINE_START
                                                                             : chunk_entry_state(prev[c], c * 16, gt.data(), fallback);
            ChunkRun r = chunk_run(cls[c], st);
            uint32_t cnt = pg_popc(r.seqmask);
            if (cnt) {
                uint32_t d = pext16_2bit(cls[c].dig, r.seqmask), m = pext16_1bit(cls[c].amb, r.seqmask);
                if (cnt < 16) d &= (1u << (2 * cnt)) - 1u;
                uint64_t g = seq + rank;
                uint32_t sh = 2 * (g & 15);
                pk2[g >> 4] |= d << sh;
                if (sh && (d >> (32 - sh))) pk2[(g >> 4) + 1] |= d >> (32 - sh);
                uint32_t sh1 = g & 31;
                amb[g >> 5] |= m << sh1;
                if (sh1 > 16 && (m >> (32 - sh1))) amb[(g >> 5) + 1] |= m >> (32 - sh1);
            }
            uint32_t hs = r.hs; uint64_t idx = hdr + hrank;
            while (hs) {
                int j = pg_ctz(hs); hs &= hs - 1;
                if ((int64_t)idx < cap_rec) { hdr_off[idx] = off + j; seq_off[idx] = (int64_t)(seq + rank + pg_popc(r.seqmask & ((1u << j) - 1u))); }
                idx++;
            }
            rank += cnt; hrank += pg_popc(r.hs);
        }
        uint32_t y = sum3_sel(agg, state);
        if (SV_SEQ(y) != rank || SV_HDR(y) != hrank) return -1000 - t;       // the summary must agree with the final pass
        seq += rank; hdr += hrank; state = SV_STATE(y);
    }
    bool dead = real_nl == 0;
    counts[0] = dead ? 0 : (int64_t)hdr; counts[1] = dead ? 0 : (int64_t)seq; counts[2] = (int64_t)real_nl;
    if (!dead && (int64_t)hdr <= cap_rec) seq_off[hdr] = (int64_t)seq;
    if (dead) seq_off[0] = 0;
    return counts[1];
}


// mirrors k1x_tile_aggs + k1x_scan + k1x_pack: 32-byte chunks, local header detection, the in-header carry resolved per
// warp from two masks (hdr_entry_from_masks) and per tile from the warps' kinds, tiles composed as functions of the
// state they are handed (TileFn)
namespace {
struct X1Chunk { Cls32 c; uint32_t hs, kind; };
template <bool LINES_ONLY>
static void x1_load(const uint8_t *f, int64_t n, int64_t off, uint32_t prev_nl, X1Chunk &o) {
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t left = n - off;
    for (int i = 0; i < 32 && i < left; i++) w[i >> 2] |= (uint32_t)f[off + i] << (8 * (i & 3));
    o.c = classify32<LINES_ONLY>(w, left);
    o.hs = hdr_starts32(o.c.nl, o.c.gt, prev_nl);
    o.kind = hdr_kind32(o.c.nl, o.hs);
}
// entry state of every chunk of a tile (NCH chunks, warps of 32) given the tile's entry state
static void x1_entries(const std::vector<X1Chunk> &ch, uint32_t tile_entry, std::vector<uint32_t> &entry, uint32_t &tile_kind) {
    const int nch = (int)ch.size(), nw = nch / 32;
    std::vector<uint32_t> wfixed(nw), wset(nw), wkind(nw);
    for (int wv = 0; wv < nw; wv++) {
        uint32_t fixed = 0, set = 0;
        for (int l = 0; l < 32; l++) { uint32_t k = ch[wv * 32 + l].kind; if (k != HK_PASS) fixed |= 1u << l; if (k == HK_SET) set |= 1u << l; }
        wfixed[wv] = fixed; wset[wv] = set;
        wkind[wv] = fixed ? (set >> pg_msb(fixed)) & 1u : (uint32_t)HK_PASS;
    }
    uint32_t tf = 0, tsx = 0;
    for (int wv = 0; wv < nw; wv++) { if (wkind[wv] != HK_PASS) tf |= 1u << wv; if (wkind[wv] == HK_SET) tsx |= 1u << wv; }
    tile_kind = tf ? (tsx >> pg_msb(tf)) & 1u : (uint32_t)HK_PASS;
    entry.resize(nch);
    for (int wv = 0; wv < nw; wv++) {
        const uint32_t wentry = hdr_entry_from_masks(tf, tsx, wv, tile_entry);
        for (int l = 0; l < 32; l++) entry[wv * 32 + l] = hdr_entry_from_masks(wfixed[wv], wset[wv], l, wentry);
    }
}
}  // namespace

extern "C" int64_t emul_pack3(const uint8_t *fasta, int64_t nbytes, uint32_t *pk2, uint32_t *amb, int64_t n_words,
                              int64_t *hdr_off, int64_t *seq_off, int64_t cap_rec, int64_t *counts) {
    memset(pk2, 0, (size_t)n_words * 4); memset(amb, 0, (size_t)n_words * 4);
    const int NCH = TILE / 32;
    const int64_t ntiles = (nbytes + TILE - 1) / TILE;
    std::vector<TileFn> fn((size_t)ntiles);
    auto prev_nl_of = [&](int64_t off) -> uint32_t { return off == 0 ? 1u : (uint32_t)(fasta[off - 1] == '\n'); };
    // ---- pass A
    for (int64_t t = 0; t < ntiles; t++) {
        std::vector<X1Chunk> ch(NCH);
        for (int c = 0; c < NCH; c++) {
            const int64_t off = t * TILE + c * 32;
            x1_load<true>(fasta, nbytes, off, c == 0 ? (off < nbytes ? prev_nl_of(off) : 1u) : (ch[c - 1].c.nl >> 31), ch[c]);
        }
        std::vector<uint32_t> entry; uint32_t kind;
        x1_entries(ch, 0, entry, kind);
        uint32_t seq0 = 0, nh = 0, rnl = 0; int first_nl = TILE;
        for (int c = 0; c < NCH; c++) {
            const uint32_t h = hdr_fill32(ch[c].c.nl, ch[c].hs, entry[c]);
            seq0 += pg_popc(~ch[c].c.nl & ~h); nh += pg_popc(ch[c].hs); rnl += ch[c].c.real_nl;
            if (ch[c].c.nl && first_nl == TILE) first_nl = c * 32 + pg_ctz(ch[c].c.nl);
        }
        const uint32_t pre = (ch[0].hs & 1u) ? 0u : (uint32_t)first_nl;
        fn[(size_t)t] = tilefn_make(seq0, pre, nh, kind, rnl);
    }
    // ---- pass B
    TileFn run = tilefn_identity();
    std::vector<TileFn> before((size_t)ntiles);
    for (int64_t t = 0; t < ntiles; t++) { before[(size_t)t] = run; run = tilefn_compose(run, fn[(size_t)t]); }
    const bool dead = run.nl == 0;
    counts[0] = dead ? 0 : (int64_t)run.hdr; counts[1] = dead ? 0 : (int64_t)run.seq0; counts[2] = (int64_t)run.nl;
    if (!dead && (int64_t)run.hdr <= cap_rec) seq_off[run.hdr] = (int64_t)run.seq0;
    if (dead) { seq_off[0] = 0; return 0; }
    // ---- pass C
    for (int64_t t = 0; t < ntiles; t++) {
        const uint64_t tseq = before[(size_t)t].seq0, thdr = before[(size_t)t].hdr; const uint32_t tentry = before[(size_t)t].exit0;
        std::vector<X1Chunk> ch(NCH);
        for (int c = 0; c < NCH; c++) {
            const int64_t off = t * TILE + c * 32;
            x1_load<false>(fasta, nbytes, off, c == 0 ? (off < nbytes ? prev_nl_of(off) : 1u) : (ch[c - 1].c.nl >> 31), ch[c]);
        }
        std::vector<uint32_t> entry; uint32_t kind;
        x1_entries(ch, tentry, entry, kind);
        uint32_t rank = 0, hrank = 0;
        for (int c = 0; c < NCH; c++) {
            const int64_t off = t * TILE + c * 32;
            const uint32_t h = hdr_fill32(ch[c].c.nl, ch[c].hs, entry[c]);
            const uint32_t seqmask = ~ch[c].c.nl & ~h;
            uint32_t r2 = rank;
            for (int half = 0; half < 2; half++) {
                const uint32_t sm = (seqmask >> (16 * half)) & 0xFFFFu, cnt = pg_popc(sm);
                if (cnt) {
                    uint32_t d = pext16_2bit(half ? ch[c].c.dig_hi : ch[c].c.dig_lo, sm), m = pext16_1bit((ch[c].c.amb >> (16 * half)) & 0xFFFFu, sm);
                    if (cnt < 16) d &= (1u << (2 * cnt)) - 1u;
                    const uint64_t g = tseq + r2;
                    const uint32_t sh = 2 * (g & 15);
                    pk2[g >> 4] |= d << sh;
                    if (sh && (d >> (32 - sh))) pk2[(g >> 4) + 1] |= d >> (32 - sh);
                    const uint32_t sh1 = g & 31;
                    amb[g >> 5] |= m << sh1;
                    if (sh1 > 16 && (m >> (32 - sh1))) amb[(g >> 5) + 1] |= m >> (32 - sh1);
                }
                r2 += cnt;
            }
            uint32_t hs = ch[c].hs; uint64_t idx = thdr + hrank;
            while (hs) {
                const int j = pg_ctz(hs); hs &= hs - 1;
                if ((int64_t)idx < cap_rec) { hdr_off[idx] = off + j; seq_off[idx] = (int64_t)(tseq + rank + pg_popc(seqmask & ((1u << j) - 1u))); }
                idx++;
            }
            rank = r2; hrank += pg_popc(ch[c].hs);
        }
        const TileFn &f = fn[(size_t)t];
        if ((tentry ? f.seq1 : f.seq0) != rank || f.hdr != hrank) return -1000 - t;      // pass A's summary must agree with pass C
    }
    return counts[1];
}
