// common.cuh - shared device/host helpers for libpgdbg (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/pgdbg.h"

#define PG_EMPTY 0xFFFFFFFFFFFFFFFFull
// Slot value word: masks [0,32) | count or rdBG flags [32,54) | generation tag [54,64).  A slot is live only
// while its tag equals the table's epoch, so "clearing" a table is epoch += 1 (pg_table_reset): no HBM traffic.
// pg_table_clear writes tag 0 everywhere; epochs 1..PG_EPOCH_MAX are the live ones.  The count field saturates
// at 255 by "skip the add once >= 255 was seen"; 22 bits leave room for every add that can be in flight before
// a thread sees the saturated value (<= 2 per resident thread).
#define PG_TAG_SHIFT 54
#define PG_VAL_MASK ((1ull << PG_TAG_SHIFT) - 1ull)
#define PG_EPOCH_MAX 1023
#define PG_HD __host__ __device__ __forceinline__

extern thread_local char pg_err_buf[512];
int pg_fail(int code, const char *fmt, ...);

#define PG_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return pg_fail(PG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                           cudaGetErrorString(e_));                                        \
    } while (0)

int pg_num_sms();
void pg_tune_once();

// ---- hashing ---------------------------------------------------------------
PG_HD uint64_t pg_mix64(uint64_t x) {   // murmur3 fmix64: a bijection on u64
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// the same without the last xor-shift, which only changes the low 31 bits: for callers that use nothing below bit 31
// (bucket / region / slot indices of tables up to 2^33 slots are the TOP bits)
PG_HD uint64_t pg_mix64_top(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    return x;
}

// ---- base-5 k-mer codes (kmer_numba.py:975-985: code = sum alpha[s[i]] * 5^i) ----
#define PG_INV5 0xCCCCCCCCCCCCCCCDull   // 5^-1 mod 2^64: exact division of multiples of 5
PG_HD uint64_t pg_pow5(int e) {
    uint64_t p = 1;
    for (int i = 0; i < e; i++) p *= 5;
    return p;
}
// complement digit: A0<->T3, G1<->C2, other 4 -> 4 (reverse_jit_ maps every non-ACGT byte to N)
PG_HD uint32_t pg_cdig(uint32_t d) { return d == 4 ? 4u : 3u - d; }

// reverse-complement of a base-5 code (the key the rc strand inserts for the same window).
// The code is split into 9-digit limbs (5^9 < 2^21) so the digit loop runs in 32-bit arithmetic.
PG_HD uint64_t pg_rc_code(uint64_t code, int k) {
    const uint32_t P9 = 1953125u;                       // 5^9
    uint32_t limb[3];
    limb[0] = (uint32_t)(code % P9); code /= P9;
    limb[1] = (uint32_t)(code % P9); code /= P9;
    limb[2] = (uint32_t)code;                            // < 5^9 for k <= 27
    uint64_t r = 0;
    int left = k;
#pragma unroll
    for (int l = 0; l < 3; l++) {
        uint32_t x = limb[l];
        int n = left < 9 ? left : 9;
        for (int i = 0; i < n; i++) { uint32_t d = x % 5u; x /= 5u; r = r * 5 + pg_cdig(d); }
        left -= n;
    }
    return r;
}
// can the code be its own reverse complement?  Only if the middle digit of an odd-k code is the
// ambiguity digit 4 (cdig(d) == d only for d == 4); for even k any code might be.
PG_HD bool pg_maybe_palindrome(uint64_t code, int k, uint64_t pow5_mid) {
    if ((k & 1) == 0) return true;
    return (code / pow5_mid) % 5 == 4;
}

// ---- symbols: 0..3 = A G C T (base-5 digit), 4 = N/n, 5 = any other byte ------------
// lastc on the forward strand (kmer_numba.py:736-743): A1 T2 G4 C8 N16 other 0
PG_HD uint32_t pg_lastc_f(uint32_t sym) { return (uint32_t)((0x001002080401ull >> (8 * sym)) & 0xff); }
// lastc of the complemented byte (rc strand; every non-ACGT byte became 'N')
PG_HD uint32_t pg_lastc_r(uint32_t sym) { return (uint32_t)((0x101001040802ull >> (8 * sym)) & 0xff); }

PG_HD int pg_popc12(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// rdBG membership of one orientation's 12-bit value (build_rdbg_jit_ :1299-1303)
PG_HD bool pg_is_rdbg(uint32_t val12) { return !(pg_popc12(val12 >> 6) == 1 && pg_popc12(val12 & 63) == 1); }

// ---- order-independent table checksum (mirrors oracle.table_checksum) ----------------
PG_HD uint64_t pg_entry_mix(uint64_t key, uint32_t val12, uint32_t cnt) {
    uint64_t w = ((uint64_t)val12 << 8) | cnt;
    return pg_mix64(key ^ (w * 0x9E3779B97F4A7C15ull));
}

// ---- ASCII classification for K1 ---------------------------------------------------
// returns the 3-bit symbol of a byte
PG_HD uint32_t pg_sym_of_byte(uint32_t c) {
    uint32_t u = c & 0xDF;
    if (u == 'A') return 0;
    if (u == 'G') return 1;
    if (u == 'C') return 2;
    if (u == 'T') return 3;
    if (u == 'N') return 4;
    return 5;
}

#ifdef __CUDACC__
// 128-bit streaming load (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 pg_ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 16-byte table-slot load at L2 (slots are mutated by atomics, never trust L1).  One .b128 access: the
// key and the value word (which carries the generation tag) are observed together, like atom.cas.b128 writes them.
__device__ __forceinline__ void pg_ld_slot_raw(const uint64_t *p, uint64_t &lo, uint64_t &hi) {
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p));
}
// The slot as the current generation sees it: anything written under another tag reads as EMPTY.
__device__ __forceinline__ void pg_ld_slot(const uint64_t *p, uint64_t tag, uint64_t &key, uint64_t &val) {
    uint64_t lo, hi;
    pg_ld_slot_raw(p, lo, hi);
    const bool live = (hi & ~PG_VAL_MASK) == tag;
    key = live ? lo : PG_EMPTY;
    val = live ? (hi & PG_VAL_MASK) : 0ull;
}
// read-once streams (update records): keep them from displacing the table region in L2
__device__ __forceinline__ uint64_t pg_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 pg_ld_stream_l2first(const uint4 *p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void pg_st_stream_l2first(uint4 *p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
// 128-bit compare-and-swap of a whole slot {key, value}: claims an empty slot and deposits the first
// occurrence's masks + count in ONE L2 transaction (instead of CAS + red.or + red.add)
__device__ __forceinline__ void pg_cas128(uint64_t *p, uint64_t cmp_lo, uint64_t cmp_hi, uint64_t new_lo, uint64_t new_hi,
                                          uint64_t &old_lo, uint64_t &old_hi) {
    asm volatile("{\n\t.reg .b128 c, s, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\t"
                 "atom.global.cas.b128 d, [%6], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(old_lo), "=l"(old_hi) : "l"(cmp_lo), "l"(cmp_hi), "l"(new_lo), "l"(new_hi), "l"(p) : "memory");
}
__device__ __forceinline__ void pg_red_or32(uint32_t *p, uint32_t v) {
    asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void pg_red_add32(uint32_t *p, uint32_t v) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif
