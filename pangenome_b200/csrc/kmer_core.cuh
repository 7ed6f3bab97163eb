// kmer_core.cuh - per-position k-mer arithmetic shared by the insert (K2/K3) and path (K5) kernels.
// Host+device so tests/host_emul.cu can run the same code on the CPU against the oracle.
//
// Restates build_dbg / seq2ns_jit_ (kmer_numba.py:1052-1090, 991-1033) in position-parallel form:
// for a forward position p of a record of length n (k-mer = s[p..p+k)), the forward strand holds the
// occurrence (code, prev, next) and the reverse-complement strand holds, at q = n-k-p, the occurrence
// (rc(code), prev', next') with prev' = comp(s[p+k]), next' = comp(s[p-1]).  Quirks kept:
//   Q1  the LAST occurrence of each strand takes its predecessor one base too far left (the reference
//       reuses a loop variable, :1080/:1019)  -> fwd p = n-k uses s[p-2]; rc q = n-k (p = 0) uses s[k+1];
//   Q3  '#' (no in-bit) at a strand start, '$' (out-bit 32) at a strand end;
//   Q4  a non-ACGTN byte has lastc 0 on the forward strand but is 'N' (16) on the rc strand;
//   Q2  n == k+1 is undefined upstream; here the true predecessor is used.
#pragma once
#include "common.cuh"

// A 96-base window of the packed stream around a 32-base word: bases [-32, 64) relative to its first.
struct PgWindow {
    uint64_t prv, cur, nxt;     // 2-bit digits, 32 bases each
    uint32_t aprv, acur, anxt;  // ambiguity bits
    PG_HD uint32_t dig2(int j) const {   // j in [-32, 64)
        uint64_t w = j < 0 ? prv : (j < 32 ? cur : nxt);
        return (uint32_t)(w >> (2 * (j & 31))) & 3u;
    }
    PG_HD uint32_t ambbit(int j) const {
        uint32_t w = j < 0 ? aprv : (j < 32 ? acur : anxt);
        return (w >> (j & 31)) & 1u;
    }
    // base-5 digit: A0 G1 C2 T3, anything else 4 (alpha, kmer_numba.py:763-768)
    PG_HD uint32_t dig5(int j) const { return ambbit(j) ? 4u : dig2(j); }
    // symbol: 0..3 ACGT digit, 4 = N/n, 5 = other byte
    PG_HD uint32_t sym(int j) const { return ambbit(j) ? 4u + dig2(j) : dig2(j); }
    PG_HD bool any_amb() const { return (aprv | acur | anxt) != 0; }
};

// 12-bit values of the forward-strand occurrence (vf) and of the paired rc-strand occurrence (vr).
// `j` = window index of the k-mer's first base, p = its position in the record, n = record length.
PG_HD void pg_occ_vals(const PgWindow &w, int j, int64_t p, int64_t n, int k, uint32_t &vf, uint32_t &vr) {
    const bool first = (p == 0), last = (p == n - k), q1 = (n >= (int64_t)k + 2);
    uint32_t lpf = first ? 0u : pg_lastc_f(w.sym(j + ((last && q1) ? -2 : -1)));
    uint32_t lnf = last ? 32u : pg_lastc_f(w.sym(j + k));
    uint32_t lpr = last ? 0u : pg_lastc_r(w.sym(j + ((first && q1) ? k + 1 : k)));
    uint32_t lnr = first ? 32u : pg_lastc_r(w.sym(j - 1));
    vf = (lpf << 6) | lnf;
    vr = (lpr << 6) | lnr;
}

// forward / rc base-5 codes of the window [j, j+k) from scratch
PG_HD void pg_codes_init(const PgWindow &w, int j, int k, uint64_t &F, uint64_t &R) {
    F = 0; R = 0;
    uint64_t p5 = 1;
    for (int i = 0; i < k; i++) {
        uint32_t d = w.dig5(j + i);
        F += (uint64_t)d * p5;            // sum d[i] * 5^i
        R = R * 5 + pg_cdig(d);           // sum cdig(d[k-1-i]) * 5^i
        p5 *= 5;
    }
}
// slide both codes one base to the right: window [j, j+k) -> [j+1, j+k+1)
PG_HD void pg_codes_roll(const PgWindow &w, int j, int k, uint64_t pow5km1, uint64_t &F, uint64_t &R) {
    uint32_t dout = w.dig5(j), din = w.dig5(j + k);
    F = (F - dout) * PG_INV5 + (uint64_t)din * pow5km1;       // Nu // 5 + c * 5^(k-1)   (:1070)
    R = (R - (uint64_t)pg_cdig(dout) * pow5km1) * 5 + pg_cdig(din);
}

// ---- interior fast path -------------------------------------------------------------------------
// 64 bits of the digit stream starting at base j of the window (j in [-1, 63]); only valid where the
// window holds real bases.
PG_HD uint64_t pg_win64(const PgWindow &w, int j) {
    if (j < 0) return (w.cur << 2) | (w.prv >> 62);
    if (j == 0) return w.cur;
    if (j < 32) return (w.cur >> (2 * j)) | (w.nxt << (64 - 2 * j));
    return w.nxt >> (2 * (j - 32));
}
// 4-entry lastc tables for ACGT digits (A0 G1 C2 T3): forward A1 G4 C8 T2, complemented T2 C8 G4 A1
PG_HD uint32_t pg_lastc4_f(uint32_t d) { return (0x02080401u >> (8 * d)) & 0xffu; }
PG_HD uint32_t pg_lastc4_r(uint32_t d) { return (0x01040802u >> (8 * d)) & 0xffu; }

// Can G positions starting at stream offset g0 take the fast path?  All of them (and their prev / next
// bases) must be ACGT and lie strictly inside one record [rs, re): no '#', '$', Q1 or ambiguity.
PG_HD bool pg_is_interior(const PgWindow &w, int64_t g0, int G, int k, int64_t rs, int64_t re, bool have_rec,
                          int64_t g_begin, int64_t g_end) {
    return !w.any_amb() && have_rec && g0 - 2 >= rs && g0 + G + k + 1 <= re && g0 >= g_begin && g0 + G <= g_end;
}
template <bool NARROW> struct PgView { typedef uint64_t type; };
template <> struct PgView<true> { typedef uint32_t type; };
// The two 12-bit values of an interior position depend on its previous and next base only: 16 words
// vf | vr << 16 indexed by dprev * 4 + dnext.  Kernels keep them in shared memory (one LDS per position
// instead of four shift-table lookups); pg_fill_vlut is called by the first 16 threads of a CTA.
PG_HD uint32_t pg_vlut_entry(uint32_t i) {
    const uint32_t dp = i >> 2, din = i & 3u;
    const uint32_t vf = (pg_lastc4_f(dp) << 6) | pg_lastc4_f(din);
    const uint32_t vr = (pg_lastc4_r(din) << 6) | pg_lastc4_r(dp);
    return vf | (vr << 16);
}
// ---- table-driven initial codes ---------------------------------------------------------------
// base-5 value of five 2-bit digits: lut5[x] = sum_{i<5} ((x >> 2i) & 3) * 5^i, x < 1024 (2 KB as uint16).
// Kernels keep it in shared memory (pg_fill_lut5 by the whole CTA); with it the first window's two codes
// cost 12 look-ups and 10 multiply-adds instead of a 27-step digit loop per code.
#define PG_LUT5_SIZE 1024
PG_HD uint32_t pg_lut5_entry(uint32_t x) {
    return (x & 3u) + 5u * ((x >> 2) & 3u) + 25u * ((x >> 4) & 3u) + 125u * ((x >> 6) & 3u) + 625u * ((x >> 8) & 3u);
}
// reverse the order of the 32 two-bit groups of a word
PG_HD uint64_t pg_rev2(uint64_t x) {
#ifdef __CUDA_ARCH__
    x = __brevll(x);
#else
    x = ((x >> 32) | (x << 32));
    x = ((x & 0xFFFF0000FFFF0000ull) >> 16) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = ((x & 0xFF00FF00FF00FF00ull) >> 8) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x & 0xF0F0F0F0F0F0F0F0ull) >> 4) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x & 0xCCCCCCCCCCCCCCCCull) >> 2) | ((x & 0x3333333333333333ull) << 2);
    x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
#endif
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);     // bit-reversed pairs back in order
}
// base-5 code of the k (<= 27) two-bit digits in the low 2k bits of x (higher bits must be zero)
PG_HD uint64_t pg_code5_of2(uint64_t x, const uint16_t *lut5) {
    uint64_t c = lut5[(x >> 50) & 1023u];
#pragma unroll
    for (int g = 4; g >= 0; g--) c = c * 3125u + lut5[(x >> (10 * g)) & 1023u];
    return c;
}
// forward and reverse-complement codes of the ACGT-only window whose digits start at bit 0 of `view`
PG_HD void pg_codes_init_lut(uint64_t view, int k, const uint16_t *lut5, uint64_t &F, uint64_t &R) {
    const uint64_t x = k < 32 ? (view & ((1ull << (2 * k)) - 1ull)) : view;
    F = pg_code5_of2(x, lut5);
    R = pg_code5_of2(pg_rev2(~x) >> (64 - 2 * k), lut5);      // digit i of the rc strand = 3 - d[k-1-i]
}

// Visit G consecutive interior positions (window index j0 .. j0+G-1, j0 + G <= 32): f(q, F, R, vw) with
// vw = vf | vr << 16.  Every digit is a constant-shift field of three 64-bit views; codes roll with one
// multiply each.  ``vlut``: the 16-word table above (shared memory), or nullptr to compute the values;
// ``lut5``: the 1024-entry digit-group table (shared memory), or nullptr to run the digit loop.
template <int G, class Fn>
PG_HD void pg_interior_visit(const PgWindow &w, int j0, int k, uint64_t pow5km1, const uint32_t *vlut, const uint16_t *lut5, Fn &&f) {
    const uint64_t dout64 = pg_win64(w, j0);
    uint64_t F = 0, R = 0;
    if (lut5) {
        pg_codes_init_lut(dout64, k, lut5, F, R);
    } else {
        uint64_t p5 = 1;
        for (int i = 0; i < k; i++) {
            uint32_t d = (uint32_t)(dout64 >> (2 * i)) & 3u;      // k <= 27 digits: all inside the 64-bit view
            F += (uint64_t)d * p5; R = R * 5 + (3u - d); p5 *= 5;
        }
    }
    // the G per-position digit fields: 32-bit views are enough (and half the shift work) for G <= 16
    using view_t = typename PgView<(G <= 16)>::type;
    const view_t dprev_w = (view_t)pg_win64(w, j0 - 1), dout_w = (view_t)dout64, din_w = (view_t)pg_win64(w, j0 + k);
#pragma unroll
    for (int q = 0; q < G; q++) {
        const uint32_t dp = (uint32_t)(dprev_w >> (2 * q)) & 3u, dout = (uint32_t)(dout_w >> (2 * q)) & 3u,
                       din = (uint32_t)(din_w >> (2 * q)) & 3u;
        const uint32_t vw = vlut ? vlut[dp * 4 + din] : pg_vlut_entry(dp * 4 + din);
        f(q, F, R, vw);
        F = (F - dout) * PG_INV5 + (uint64_t)din * pow5km1;
        R = (R - (uint64_t)(3u - dout) * pow5km1) * 5 + (3u - din);
    }
}

// ---- compact (8-byte) update records: the 2-bit form of an ACGT-only k-mer -----------------------------------------
// An interior position (pg_is_interior: the k-mer, its previous and its next base are ACGT inside one record) needs
// 2k <= 54 bits of key and 4 bits of context, so the streaming build (compact_build.cu) moves 8-byte records
//     [ key2 : 54 | ctx : 6 | 0 : 4 ]     key2 = min(F2, R2), the 2-bit codes of the window and of its reverse complement
//     ctx bits 0..3 = dprev * 4 + dnext (index of the 16-word value table pg_vlut_entry), bit 4 = the rc strand's code is
//     the key (swap the two 12-bit values), bit 5 = palindrome (F2 == R2, even k only: fold both values, count twice)
// instead of the 16-byte {base-5 key, masks, inc}.  2-bit and base-5 codes are positional over the same digits
// (A0 G1 C2 T3, first base least significant), so F2 < R2 <=> F < R and the canonical choice is the same as
// pg_canonical_update_w's.  Every other position (record edges with '#' / '$' / Q1, ambiguity codes) travels as a 16-byte
// "wide" record.  Tables built from compact records are placed by the hash of the 2-bit code (pg_table.hash_kind 1).
#define PG_C_KEYBITS 54
#define PG_C_KEYMASK ((1ull << PG_C_KEYBITS) - 1ull)
#define PG_C_SWAP 16u
#define PG_C_PAL 32u
#define PG_WIDE_FLAG 0x8000000000000000ull      // shared-memory region tables: bit 63 marks a base-5 key that has no 2-bit form

PG_HD uint64_t pg_crec_pack(uint64_t F2, uint64_t R2, uint32_t ctx4) {
    const bool lt = F2 < R2, eq = F2 == R2;
    const uint32_t c = ctx4 | ((lt || eq) ? 0u : PG_C_SWAP) | (eq ? PG_C_PAL : 0u);
    return (lt ? F2 : R2) | ((uint64_t)c << PG_C_KEYBITS);
}
// masks / increment of a compact record; vw = value-table word of its ctx & 15
PG_HD void pg_crec_vals(uint32_t ctx, uint32_t vw, uint32_t &masks, uint32_t &inc) {
    const uint32_t sw = (vw >> 16) | (vw << 16);
    masks = (ctx & PG_C_PAL) ? ((vw | sw) & 0xFFFFu) : ((ctx & PG_C_SWAP) ? sw : vw);
    inc = (ctx & PG_C_PAL) ? 2u : 1u;
}
// base-5 code of the k two-bit digits of x, digit loop (rare paths; the kernels' common path uses pg_code5_of2 + lut5)
PG_HD uint64_t pg_code5_of2_loop(uint64_t x, int k) {
    uint64_t c = 0;
    for (int i = k - 1; i >= 0; i--) c = c * 5 + ((x >> (2 * i)) & 3u);
    return c;
}
// 2-bit code of a base-5 code; false when a digit is the ambiguity digit 4 (no 2-bit form).  9-digit limbs as pg_rc_code.
PG_HD bool pg_code2_of5(uint64_t code, int k, uint64_t &x2) {
    const uint32_t P9 = 1953125u;
    uint32_t limb[3];
    limb[0] = (uint32_t)(code % P9); code /= P9;
    limb[1] = (uint32_t)(code % P9); code /= P9;
    limb[2] = (uint32_t)code;
    uint64_t x = 0; bool ok = true;
    int left = k;
#pragma unroll
    for (int l = 0; l < 3; l++) {
        uint32_t v = limb[l];
        const int n = left < 9 ? left : 9;
        for (int i = 0; i < n; i++) {
            const uint32_t d = v % 5u; v /= 5u;
            ok = ok && d != 4u;
            x |= (uint64_t)(d & 3u) << (2 * (9 * l + i));
        }
        left -= n;
    }
    x2 = x;
    return ok;
}
// slot placement hash of a base-5 key in a hash_kind-1 table: the hash of its 2-bit form when it has one
PG_HD uint64_t pg_hash_kind1(uint64_t code5, int k) {
    uint64_t x2;
    return pg_code2_of5(code5, k, x2) ? pg_mix64(x2) : pg_mix64(~code5);
}
// Visit G consecutive interior positions like pg_interior_visit, in the 2-bit domain: f(q, F2, R2, ctx4).  No multiplies:
// both codes roll with shifts.
template <int G, class Fn>
PG_HD void pg_interior_visit_c(const PgWindow &w, int j0, int k, Fn &&f) {
    const uint64_t x0 = pg_win64(w, j0);
    const uint64_t M = (1ull << (2 * k)) - 1ull;              // k <= 27
    uint64_t F2 = x0 & M;
    uint64_t R2 = pg_rev2(~x0) >> (64 - 2 * k);               // digit i of the rc strand = 3 - d[k-1-i]
    using view_t = typename PgView<(G <= 16)>::type;
    const view_t dprev_w = (view_t)pg_win64(w, j0 - 1), din_w = (view_t)pg_win64(w, j0 + k);
    const int top = 2 * (k - 1);
#pragma unroll
    for (int q = 0; q < G; q++) {
        const uint32_t dp = (uint32_t)(dprev_w >> (2 * q)) & 3u, din = (uint32_t)(din_w >> (2 * q)) & 3u;
        f(q, F2, R2, dp * 4u + din);
        F2 = (F2 >> 2) | ((uint64_t)din << top);
        R2 = ((R2 << 2) & M) | (uint64_t)(3u - din);
    }
}

// The same visit for a COMPILE-TIME k and 16 positions starting at a multiple of 16 (K2a-c's thread): no rolling state at
// all.  The 16 + K + 1 digits a thread looks at lie in three 32-bit words W0..W2 of the digit stream; the window of
// position q is a funnel shift of two of them by the constant 2q, the reverse-complement window the same on the
// digit-reversed complement Y0..Y2 (Y digit m = 3 - d[47 - m], so the rc code of position q starts at Y digit 48 - K - q).
PG_HD uint32_t pg_rev2c32(uint32_t x) {          // complement and reverse the 16 two-bit digits of a word
    x = ~x;
#ifdef __CUDA_ARCH__
    x = __brev(x);
#else
    x = (x >> 16) | (x << 16);
    x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
    x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
#endif
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
PG_HD uint32_t pg_funnel_r(uint32_t lo, uint32_t hi, int sh) {      // bits [sh, sh + 32) of hi:lo, 0 <= sh < 32
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
// K digits of the 48-digit stream (w0, w1, w2) starting at digit s (0 <= s, s + K <= 48, 16 < K <= 32)
template <int K>
PG_HD uint64_t pg_take_digits(uint32_t w0, uint32_t w1, uint32_t w2, int s) {
    const uint32_t hmask = K >= 32 ? 0xFFFFFFFFu : ((1u << (2 * (K - 16))) - 1u);
    uint32_t lo, hi;
    if (s < 16) { lo = pg_funnel_r(w0, w1, 2 * s); hi = pg_funnel_r(w1, w2, 2 * s); }
    else { lo = pg_funnel_r(w1, w2, 2 * (s - 16)); hi = w2 >> (2 * (s - 16)); }
    return (uint64_t)lo | ((uint64_t)(hi & hmask) << 32);
}
template <int K, class Fn>
PG_HD void pg_interior_visit_ck(const PgWindow &w, int j0, Fn &&f) {
    static_assert(K > 16 && K <= 27, "compile-time k of the compact extraction: 17..27");
    const bool up = j0 != 0;                                   // j0 is 0 or 16
    const uint32_t W0 = up ? (uint32_t)(w.cur >> 32) : (uint32_t)w.cur, W1 = up ? (uint32_t)w.nxt : (uint32_t)(w.cur >> 32),
                   W2 = up ? (uint32_t)(w.nxt >> 32) : (uint32_t)w.nxt;
    const uint32_t Y0 = pg_rev2c32(W2), Y1 = pg_rev2c32(W1), Y2 = pg_rev2c32(W0);
    const uint32_t dm1 = up ? ((uint32_t)w.cur >> 30) & 3u : (uint32_t)(w.prv >> 62) & 3u;      // digit j0 - 1
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const uint64_t F2 = pg_take_digits<K>(W0, W1, W2, q);
        const uint64_t R2 = pg_take_digits<K>(Y0, Y1, Y2, 48 - K - q);
        const uint32_t dp = q ? (W0 >> (2 * (q - 1))) & 3u : dm1;
        const int n = q + K;                                   // digit after the window: 17 <= n <= 42
        const uint32_t din = n < 32 ? (W1 >> (2 * (n - 16))) & 3u : (W2 >> (2 * (n - 32))) & 3u;
        f(q, F2, R2, dp * 4u + din);
    }
}

// What one position contributes to a table in each mode.
struct PgUpdate { uint64_t key; uint32_t masks; uint32_t inc; };
// canonical pairing: slot key = min(F, R); masks = m(orientation 0) | m(orientation 1) << 16,
// orientation 0 being the one whose literal code equals the slot key.  Palindromes (F == R, only
// possible with ambiguity digits or even k) fold both strands into orientation 0 and count twice.
// same from the packed pair vw = vf | vr << 16: orientation 0 first means "swap the halves unless F < R"
PG_HD PgUpdate pg_canonical_update_w(uint64_t F, uint64_t R, uint32_t vw) {
    PgUpdate u;
    const bool lt = F < R, eq = F == R;
    const uint32_t sw = (vw >> 16) | (vw << 16);
    u.key = lt ? F : R;
    u.masks = eq ? ((vw | sw) & 0xFFFFu) : (lt ? vw : sw);
    u.inc = eq ? 2u : 1u;
    return u;
}
PG_HD PgUpdate pg_canonical_update(uint64_t F, uint64_t R, uint32_t vf, uint32_t vr) {
    // selects, not branches: the three cases are spread at random over the lanes of a warp
    PgUpdate u;
    const bool lt = F < R, eq = F == R;
    const uint32_t v0 = lt ? vf : vr, v1 = lt ? vr : vf;       // orientation 0 = the strand whose code is the key
    u.key = lt ? F : R;
    u.masks = eq ? (vf | vr) : (v0 | (v1 << 16));
    u.inc = eq ? 2u : 1u;
    return u;
}

// ---- read-out: one occupied slot -> its entries in the reference's convention --------------------
struct PgEntry { uint64_t key; uint32_t val, cnt; };

// slot value word: low 32 = masks (orientation 0 | orientation 1 << 16), high 32 = count.
// Returns the number of entries (0..2): both orientations are separate keys upstream (F3).
PG_HD int pg_slot_entries(uint64_t key, uint64_t v, int mode, int k, PgEntry e[2]) {
    if (key == PG_EMPTY) return 0;
    uint32_t masks = (uint32_t)v, cnt = (uint32_t)(v >> 32);
    uint32_t c255 = cnt < 255u ? cnt : 255u;          // uint8 saturation (kmer_numba.py:551)
    e[0].key = key; e[0].val = masks & 0xFFFu; e[0].cnt = c255;
    if (mode != PG_MODE_CANONICAL) return 1;
    uint64_t rk = pg_rc_code(key, k);
    if (rk == key) return 1;      // palindrome: both strands were folded into orientation 0 (count += 2 each)
    e[1].key = rk; e[1].val = (masks >> 16) & 0xFFFu; e[1].cnt = c255;
    return 2;
}
// Same, without materialising the rc key (callers that only need the values): n = 1 or 2 entries,
// vals[o] = 12-bit value of orientation o.  pow5_mid = 5^(k/2) lets odd-k keys skip the rc computation.
PG_HD int pg_slot_vals(uint64_t key, uint64_t v, int mode, int k, uint64_t pow5_mid, uint32_t vals[2]) {
    if (key == PG_EMPTY) return 0;
    uint32_t masks = (uint32_t)v;
    vals[0] = masks & 0xFFFu; vals[1] = (masks >> 16) & 0xFFFu;
    if (mode != PG_MODE_CANONICAL) return 1;
    // a palindrome folds both strands into orientation 0 (pg_canonical_update_w), so masks in the upper half prove a
    // proper pair - no 64-bit division by 5^(k/2) on the common path (it was most of K4's 150 instructions per slot)
    if (masks >> 16) return 2;
    if (pg_maybe_palindrome(key, k, pow5_mid) && pg_rc_code(key, k) == key) return 1;
    return 2;
}
PG_HD uint32_t pg_rdbg_flags_fast(uint64_t key, uint64_t v, int mode, int k, uint64_t pow5_mid) {
    uint32_t vals[2];
    int n = pg_slot_vals(key, v, mode, k, pow5_mid, vals);
    if (n == 0) return 0;
    uint32_t f = pg_is_rdbg(vals[0]) ? 1u : 0u;
    if (n == 2 && pg_is_rdbg(vals[1])) f |= 2u;
    if (key == 0) f |= 4u;
    return f;
}

// flags of an rdBG slot: bit0/bit1 = orientation 0/1 is a member, bit2 = phantom key 0 (Q6)
PG_HD uint32_t pg_rdbg_flags(uint64_t key, uint64_t v, int mode, int k) {
    PgEntry e[2];
    int n = pg_slot_entries(key, v, mode, k, e);
    if (n == 0) return 0;
    uint32_t f = pg_is_rdbg(e[0].val) ? 1u : 0u;
    if (n == 2 && pg_is_rdbg(e[1].val)) f |= 2u;
    if (key == 0) f |= 4u;    // has_key(0) is true for every table (equality before occupancy, :532/:599)
    return f;
}
