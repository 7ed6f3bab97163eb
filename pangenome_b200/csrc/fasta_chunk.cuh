// fasta_chunk.cuh - the per-16-byte FASTA line-state machine of K1, written as
// host+device functions so tests/host_emul.cu can check it on the CPU.
//
// Restates readline_jit_ / seqio_jit_ (kmer_numba.py:122-168):
//   * lines end at '\n'; the last byte of EVERY line is dropped, also when the
//     final line has no '\n' (Q8) - modelled as a virtual newline on the last
//     byte of the file;
//   * a line whose first byte is '>' is a header and starts a record;
//   * every other byte of a non-header line is a base ('\r' included, Q9).
#pragma once
#include "common.cuh"

enum { ST_LINE_START = 0, ST_HEADER = 1, ST_SEQ = 2 };

// 3-variant summary word: what a span of bytes does for one given entry state.
//   bits 0..14 bases, bits 15..28 headers started, bits 30..31 exit state.
#define SV_CNT_MASK 0x3FFFFFFFu
#define SV_SEQ(x) ((x) & 0x7FFFu)
#define SV_HDR(x) (((x) >> 15) & 0x3FFFu)
#define SV_STATE(x) ((x) >> 30)
#define SV_MAKE(st, hdr, seq) (((uint32_t)(st) << 30) | ((uint32_t)(hdr) << 15) | (uint32_t)(seq))

struct Sum3 { uint32_t v[3]; };
// register-friendly v[s] (a dynamic index would push the array to local memory)
PG_HD uint32_t sum3_sel(const Sum3 &a, uint32_t s) { return s == 0 ? a.v[0] : (s == 1 ? a.v[1] : a.v[2]); }

// "a then b"
PG_HD Sum3 sum3_compose(const Sum3 &a, const Sum3 &b) {
    Sum3 c;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        uint32_t x = a.v[s];
        uint32_t y = sum3_sel(b, SV_STATE(x));
        c.v[s] = ((x & SV_CNT_MASK) + (y & SV_CNT_MASK)) | (y & ~SV_CNT_MASK);
    }
    return c;
}
PG_HD Sum3 sum3_identity() { Sum3 c; c.v[0] = SV_MAKE(0, 0, 0); c.v[1] = SV_MAKE(1, 0, 0); c.v[2] = SV_MAKE(2, 0, 0); return c; }

struct ChunkCls {
    uint32_t nl;    // bit i: byte i ends a line (real '\n', the last byte of the file, or past the end)
    uint32_t gt;    // bit i: byte i == '>'
    uint32_t amb;   // bit i: byte i is not one of ACGTacgt
    uint32_t dig;   // 2 bits per byte: base-5 digit for ACGT (A0 G1 C2 T3); 0 for N/n; 1 for other bytes
    uint32_t real_nl;  // number of real '\n' bytes
};

PG_HD uint32_t pg_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
PG_HD int pg_ctz(uint32_t v) {   // v != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
PG_HD int pg_msb(uint32_t v) {   // v != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// SWAR helpers on 4 packed bytes
PG_HD uint32_t swar_nonzero80(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }   // 0x80 where byte != 0
PG_HD uint32_t swar_flags4(uint32_t b80) { return (((b80 >> 7) * 0x01020408u) >> 24) & 0xFu; }                // byte i -> bit i

// Classify 16 bytes held in 4 little-endian words.  `n_file` = how many of the 16 bytes lie inside
// the file (0..16); `has_last` = the chunk reaches the end of the file.  The last byte of the file
// and everything after it are treated as line terminators (Q8 virtual newline).
PG_HD ChunkCls classify16(const uint32_t w[4], int n_file, bool has_last) {
    ChunkCls c;
    c.nl = c.gt = c.amb = c.dig = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t x = w[j];
        uint32_t nl80 = ~swar_nonzero80(x ^ 0x0A0A0A0Au) & 0x80808080u;
        uint32_t gt80 = ~swar_nonzero80(x ^ 0x3E3E3E3Eu) & 0x80808080u;
        uint32_t u = x & 0xDFDFDFDFu;
        uint32_t n = (x >> 1) & 0x03030303u;
        uint32_t hi = (n >> 1) & 0x01010101u, lo = n & 0x01010101u;
        uint32_t expect = 0x41414141u + 2u * lo + 0x13u * hi - 0x0Fu * (hi & lo);   // A C T G by (c>>1)&3
        uint32_t amb80 = swar_nonzero80(u ^ expect);
        uint32_t d = ((hi ^ lo) << 1) | hi;                                         // A0 C2 T3 G1
        c.nl |= swar_flags4(nl80) << (4 * j);
        c.gt |= swar_flags4(gt80) << (4 * j);
        c.amb |= swar_flags4(amb80) << (4 * j);
        c.dig |= (((d * 0x01041040u) >> 24) & 0xFFu) << (8 * j);
    }
    uint32_t infile = n_file >= 16 ? 0xFFFFu : ((1u << (n_file < 0 ? 0 : n_file)) - 1u);
    c.real_nl = pg_popc(c.nl & infile);
    if (has_last) {
        int first_forced = n_file > 0 ? n_file - 1 : 0;
        uint32_t forced = (0xFFFFu << first_forced) & 0xFFFFu;
        c.nl |= forced; c.gt &= ~forced;
    }
    if (c.amb) {   // rare: N/n -> digit 0, any other byte -> digit 1
        uint32_t m = c.amb;
        while (m) {
            int i = pg_ctz(m); m &= m - 1;
            uint32_t wi = (i < 4) ? w[0] : (i < 8) ? w[1] : (i < 12) ? w[2] : w[3];
            uint32_t byte = (wi >> (8 * (i & 3))) & 0xFFu;
            uint32_t d = ((byte & 0xDFu) == 'N') ? 0u : 1u;
            c.dig = (c.dig & ~(3u << (2 * i))) | (d << (2 * i));
        }
    }
    return c;
}

// Pass A only needs the line structure: newline and '>' masks (no digits, no ambiguity plane).
PG_HD ChunkCls classify16_lines(const uint32_t w[4], int n_file, bool has_last) {
    ChunkCls c;
    c.nl = c.gt = c.amb = c.dig = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t x = w[j];
        uint32_t nl80 = ~swar_nonzero80(x ^ 0x0A0A0A0Au) & 0x80808080u;
        uint32_t gt80 = ~swar_nonzero80(x ^ 0x3E3E3E3Eu) & 0x80808080u;
        c.nl |= swar_flags4(nl80) << (4 * j);
        c.gt |= swar_flags4(gt80) << (4 * j);
    }
    uint32_t infile = n_file >= 16 ? 0xFFFFu : ((1u << (n_file < 0 ? 0 : n_file)) - 1u);
    c.real_nl = pg_popc(c.nl & infile);
    if (has_last) {
        int first_forced = n_file > 0 ? n_file - 1 : 0;
        uint32_t forced = (0xFFFFu << first_forced) & 0xFFFFu;
        c.nl |= forced; c.gt &= ~forced;
    }
    return c;
}

struct ChunkRun {
    uint32_t seqmask;   // bytes that are bases
    uint32_t hs;        // bytes that start a header line
    uint32_t exit_state;
};

// Run the line-state machine over one classified chunk from a known entry state.
PG_HD ChunkRun chunk_run(const ChunkCls &c, uint32_t entry) {
    ChunkRun r;
    uint32_t nl = c.nl & 0xFFFFu;
    if (c.gt == 0) {   // fast path (almost every chunk): no '>' byte, so no header can start here
        r.hs = 0;
        if (nl == 0) {
            r.seqmask = entry == ST_HEADER ? 0u : 0xFFFFu;
            r.exit_state = entry == ST_HEADER ? ST_HEADER : ST_SEQ;
        } else {
            uint32_t after_first = ~((2u << pg_ctz(nl)) - 1u);
            r.seqmask = ~nl & 0xFFFFu & (entry == ST_HEADER ? after_first : 0xFFFFFFFFu);
            r.exit_state = (nl & 0x8000u) ? ST_LINE_START : ST_SEQ;
        }
        return r;
    }
    uint32_t ls = ((nl << 1) | (entry == ST_LINE_START ? 1u : 0u)) & 0xFFFFu;   // line starts
    r.hs = ls & c.gt & ~nl;
    uint32_t starts = r.hs | (entry == ST_HEADER ? 1u : 0u);
    uint32_t hdr = 0;
    while (starts) {
        int s = pg_ctz(starts);
        uint32_t above = nl & ~((1u << s) - 1u);
        if (above) {
            int t = pg_ctz(above);
            uint32_t span = ((2u << t) - 1u) & ~((1u << s) - 1u);
            hdr |= span; starts &= ~span;
        } else {
            hdr |= 0xFFFFu & ~((1u << s) - 1u); starts = 0;   // header line continues past the chunk
        }
    }
    r.seqmask = ~nl & ~hdr & 0xFFFFu;
    if (nl == 0) r.exit_state = (entry == ST_LINE_START) ? ((c.gt & 1u) ? ST_HEADER : ST_SEQ) : entry;
    else {
        int ln = pg_msb(nl);
        r.exit_state = (ln == 15) ? ST_LINE_START : (((c.gt >> (ln + 1)) & 1u) ? ST_HEADER : ST_SEQ);
    }
    return r;
}

PG_HD Sum3 chunk_sum3(const ChunkCls &c) {
    Sum3 s;
    if (c.gt == 0) {   // fast path: the entry state only decides whether the bytes before the first '\n' are bases
        uint32_t nl = c.nl & 0xFFFFu;
        if (nl == 0) {
            s.v[ST_LINE_START] = SV_MAKE(ST_SEQ, 0, 16); s.v[ST_HEADER] = SV_MAKE(ST_HEADER, 0, 0); s.v[ST_SEQ] = SV_MAKE(ST_SEQ, 0, 16);
        } else {
            uint32_t ex = (nl & 0x8000u) ? ST_LINE_START : ST_SEQ;
            uint32_t all = 16u - pg_popc(nl);
            uint32_t tail = pg_popc(~nl & 0xFFFFu & ~((2u << pg_ctz(nl)) - 1u));
            s.v[ST_LINE_START] = SV_MAKE(ex, 0, all); s.v[ST_HEADER] = SV_MAKE(ex, 0, tail); s.v[ST_SEQ] = SV_MAKE(ex, 0, all);
        }
        return s;
    }
#pragma unroll
    for (uint32_t e = 0; e < 3; e++) {
        ChunkRun r = chunk_run(c, e);
        s.v[e] = SV_MAKE(r.exit_state, pg_popc(r.hs), pg_popc(r.seqmask));
    }
    return s;
}

// software PEXT over 16 lanes: keep the 2-bit fields / bits selected by `mask`
PG_HD uint32_t pext16_2bit(uint32_t dig, uint32_t mask) {
    if (mask == 0xFFFFu) return dig;
    uint32_t out = 0; int o = 0;
    while (mask) {
        int i = pg_ctz(mask);
        // take the run of consecutive set bits starting at i
        uint32_t run = mask >> i;
        int len = pg_ctz(~run);
        uint32_t field = (len >= 16) ? (dig >> (2 * i)) : ((dig >> (2 * i)) & ((1u << (2 * len)) - 1u));
        out |= field << (2 * o);
        o += len;
        mask = (len + i >= 32) ? 0 : (mask & ~(((1u << len) - 1u) << i));
    }
    return out;
}
PG_HD uint32_t pext16_1bit(uint32_t bits, uint32_t mask) {
    if (mask == 0xFFFFu) return bits & 0xFFFFu;
    uint32_t out = 0; int o = 0;
    while (mask) {
        int i = pg_ctz(mask);
        uint32_t run = mask >> i;
        int len = pg_ctz(~run);
        out |= ((bits >> i) & ((1u << len) - 1u)) << o;
        o += len;
        mask &= ~(((1u << len) - 1u) << i);
    }
    return out;
}

// ---- single-pass K1 helpers ---------------------------------------------------------------------
// The line state of a chunk follows from the LAST newline before it: the line starts right after it,
// and the line is a header iff its first byte is '>'.  gt_masks[c] = '>' mask of chunk c of the tile.
PG_HD uint32_t chunk_entry_state(int prev_nl, int chunk_off, const uint16_t *gt_masks, uint32_t fallback) {
    if (prev_nl < 0) return fallback;                    // no newline in the tile before this chunk
    const int ls = prev_nl + 1;
    if (ls == chunk_off) return ST_LINE_START;
    return ((gt_masks[ls >> 4] >> (ls & 15)) & 1u) ? ST_HEADER : ST_SEQ;
}
PG_HD int chunk_last_nl(uint32_t nl, int chunk_off) { return nl ? chunk_off + pg_msb(nl & 0xFFFFu) : -1; }
PG_HD int chunk_first_nl(uint32_t nl, int chunk_off) { return nl ? chunk_off + pg_ctz(nl & 0xFFFFu) : 0x7FFFFFFF; }

// What a whole tile does for each entry state, from quantities that do not depend on the entry state:
// bases / headers after the tile's first newline, the position of the first and last newline, and
// whether the tile's first byte is '>'.
struct TileLocal { uint32_t post_seq, post_hdr; int first_nl, last_nl; uint32_t first_gt, gt_after_last; int tile_len; };
PG_HD Sum3 tile_sum3(const TileLocal &t) {
    const bool has_nl = t.last_nl >= 0;
    const uint32_t pre = has_nl ? (uint32_t)t.first_nl : (uint32_t)t.tile_len;      // bytes before the first newline (none of them is one)
    const uint32_t exit_local = !has_nl ? 0u : (t.last_nl == t.tile_len - 1 ? (uint32_t)ST_LINE_START : (t.gt_after_last ? (uint32_t)ST_HEADER : (uint32_t)ST_SEQ));
    Sum3 s;
#pragma unroll
    for (uint32_t e = 0; e < 3; e++) {
        const bool pre_hdr = (e == ST_HEADER) || (e == ST_LINE_START && t.first_gt && pre > 0);
        const uint32_t seq = t.post_seq + (pre_hdr ? 0u : pre);
        const uint32_t hdr = t.post_hdr + ((e == ST_LINE_START && t.first_gt && pre > 0) ? 1u : 0u);
        const uint32_t ex = has_nl ? exit_local : (pre == 0 ? e : (pre_hdr ? (uint32_t)ST_HEADER : (uint32_t)ST_SEQ));
        s.v[e] = SV_MAKE(ex, hdr, seq);
    }
    return s;
}

// ---- K1 in 32-byte chunks with LOCAL header detection (k1x_* kernels) ------------------------------------------------
// A header starts where a '>' sits at a line start, and a line starts after a line-ending byte (or at byte 0 of the
// file): that is a property of two neighbouring bytes, no state machine needed.  What does carry over from chunk to
// chunk is one bit - "inside a header line" - and it behaves like the carry of an adder: a header start GENERATES it, a
// line end KILLS it, every other byte PROPAGATES it.  So the in-header mask of a chunk is one 64-bit addition
// (hdr_fill32), the state a chunk hands on is decided by its last event alone (or, without events, by its entry
// state), and a warp resolves its 32 chunks with two ballots instead of a scan over 3-variant summaries.
struct Cls32 { uint32_t nl, gt, amb; uint32_t dig_lo, dig_hi; uint32_t real_nl; };
// classify 32 bytes = two classify16 halves; left = bytes of the file from the chunk's first byte on (may be <= 0)
template <bool LINES_ONLY>
PG_HD Cls32 classify32(const uint32_t w[8], int64_t left) {
    const int n_lo = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
    const int64_t left_hi = left - 16;
    const int n_hi = left_hi >= 16 ? 16 : (left_hi > 0 ? (int)left_hi : 0);
    const uint32_t wl[4] = {w[0], w[1], w[2], w[3]}, wh[4] = {w[4], w[5], w[6], w[7]};
    const ChunkCls a = LINES_ONLY ? classify16_lines(wl, n_lo, left <= 16) : classify16(wl, n_lo, left <= 16);
    const ChunkCls b = LINES_ONLY ? classify16_lines(wh, n_hi, left_hi <= 16) : classify16(wh, n_hi, left_hi <= 16);
    Cls32 c;
    c.nl = (a.nl & 0xFFFFu) | (b.nl << 16); c.gt = (a.gt & 0xFFFFu) | (b.gt << 16); c.amb = (a.amb & 0xFFFFu) | (b.amb << 16);
    c.dig_lo = a.dig; c.dig_hi = b.dig; c.real_nl = a.real_nl + b.real_nl;
    return c;
}
// header starts of a chunk; prev_nl = 1 when the byte before the chunk ends a line or the chunk starts the file
PG_HD uint32_t hdr_starts32(uint32_t nl, uint32_t gt, uint32_t prev_nl) { return gt & ~nl & ((nl << 1) | (prev_nl & 1u)); }
// in-header mask: h_i = hs_i | (~nl_i & ~hs_i & h_{i-1}), h_{-1} = cin  - the carries of (g | p) + g + cin
PG_HD uint32_t hdr_fill32(uint32_t nl, uint32_t hs, uint32_t cin) {
    const uint64_t x = (uint64_t)(uint32_t)(~nl | hs), y = hs;       // g | p = hs | (~nl & ~hs) = hs | ~nl
    const uint64_t c = (x + y + (cin & 1u)) ^ x ^ y;                 // carry INTO every bit
    return (uint32_t)(c >> 1);                                       // carry OUT of bit i = in-header at byte i
}
// what a span hands on: 0 = not in a header, 1 = in a header, 2 = whatever it was handed (no event inside)
enum { HK_RESET = 0, HK_SET = 1, HK_PASS = 2 };
PG_HD uint32_t hdr_kind32(uint32_t nl, uint32_t hs) {
    const uint32_t ev = nl | hs;
    if (!ev) return HK_PASS;
    return (hs >> pg_msb(ev)) & 1u;
}
// entry state of element `idx` (lane of a warp / warp of a tile) from two masks over the elements: fixed = has an event,
// set = its last event is a header start; `before` = the state handed to element 0
PG_HD uint32_t hdr_entry_from_masks(uint32_t fixed, uint32_t set, int idx, uint32_t before) {
    const uint32_t m = idx >= 32 ? fixed : (fixed & ((1u << idx) - 1u));
    return m ? (set >> pg_msb(m)) & 1u : before;
}
// A tile as a function of the state it is handed: bases for entry 0 / entry 1, exit state for both, headers, real '\n's.
struct TileFn { unsigned long long seq0, seq1, hdr, nl; uint32_t exit0, exit1; };
PG_HD TileFn tilefn_identity() { TileFn f; f.seq0 = f.seq1 = f.hdr = f.nl = 0; f.exit0 = 0; f.exit1 = 1; return f; }
PG_HD TileFn tilefn_make(uint32_t seq0, uint32_t pre_seq, uint32_t hdr, uint32_t kind, uint32_t real_nl) {
    TileFn f; f.seq0 = seq0; f.seq1 = seq0 - pre_seq; f.hdr = hdr; f.nl = real_nl;
    f.exit0 = kind == HK_PASS ? 0u : kind; f.exit1 = kind == HK_PASS ? 1u : kind;
    return f;
}
PG_HD TileFn tilefn_compose(const TileFn &a, const TileFn &b) {      // a then b
    TileFn c;
    c.seq0 = a.seq0 + (a.exit0 ? b.seq1 : b.seq0); c.exit0 = a.exit0 ? b.exit1 : b.exit0;
    c.seq1 = a.seq1 + (a.exit1 ? b.seq1 : b.seq0); c.exit1 = a.exit1 ? b.exit1 : b.exit0;
    c.hdr = a.hdr + b.hdr; c.nl = a.nl + b.nl;
    return c;
}
