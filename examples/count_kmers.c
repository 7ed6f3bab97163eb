/* count_kmers.c - the C-ABI of libpgdbg.so used from plain C, no Python and no torch:
 * FASTA file -> K1 pack -> fused k-mer insert (both strands, canonical pairs) -> table statistics,
 * then the same table again through the two-phase build after an O(1) reset.
 *
 *   gcc -O2 -Iinclude examples/count_kmers.c -o count_kmers \
 *       -Lpangenome_b200 -lpgdbg -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/pangenome_b200
 *   ./count_kmers input.fasta 27
 *
 * Prints: records, bases, distinct canonical k-mers, entries in the reference's convention (both
 * orientations are keys, kmer_numba.py:1036-1093), and the order-independent table checksum that
 * tests/ compare with the oracle.  Device memory comes straight from the CUDA runtime.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "pgdbg.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)
#define PG(call) do { int r_ = (call); if (r_ != PG_OK) { \
    fprintf(stderr, "%s: %s\n", #call, pg_last_error()); return 3; } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s input.fasta [k]\n", argv[0]); return 1; }
    int k = argc > 2 ? atoi(argv[2]) : 27;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    int64_t nbytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *host = (uint8_t *)malloc(nbytes > 0 ? nbytes : 1);
    if (fread(host, 1, nbytes, f) != (size_t)nbytes) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);

    /* K1: raw bytes -> 2-bit digit plane + 1-bit ambiguity plane + record index */
    const int64_t cap_records = 1 << 16, words = pg_pack_words(nbytes), ws_bytes = pg_fasta_workspace_bytes(nbytes);
    uint8_t *d_fasta; uint32_t *d_pk2, *d_amb; int64_t *d_hdr, *d_off, *d_cnt; void *d_ws;
    CK(cudaMalloc((void **)&d_fasta, nbytes + 16));
    CK(cudaMalloc((void **)&d_pk2, words * 4)); CK(cudaMalloc((void **)&d_amb, words * 4));
    CK(cudaMalloc((void **)&d_hdr, (cap_records + 1) * 8)); CK(cudaMalloc((void **)&d_off, (cap_records + 2) * 8));
    CK(cudaMalloc((void **)&d_cnt, 4 * 8)); CK(cudaMalloc(&d_ws, ws_bytes > 0 ? ws_bytes : 16));
    CK(cudaMemcpy(d_fasta, host, nbytes, cudaMemcpyHostToDevice));
    PG(pg_fasta_scan_pack(d_fasta, nbytes, d_pk2, d_amb, nbytes, d_hdr, d_off, cap_records, d_cnt, d_ws, ws_bytes, NULL));
    int64_t cnt[4];
    CK(cudaMemcpy(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost));
    const int64_t n_rec = cnt[0], n_bases = cnt[1];
    if (n_rec > cap_records) { fprintf(stderr, "more than %lld records: enlarge cap_records\n", (long long)cap_records); return 1; }
    int64_t *off = (int64_t *)malloc((n_rec + 1) * 8);
    CK(cudaMemcpy(off, d_off, (n_rec + 1) * 8, cudaMemcpyDeviceToHost));

    /* the table: a power of two >= 2 slots per base keeps the load under 0.5 without rehashing */
    int64_t cap = 1024;
    while (cap < 2 * n_bases) cap <<= 1;
    pg_table t = {0};
    CK(cudaMalloc((void **)&t.d_slots, pg_table_bytes(cap))); CK(cudaMalloc((void **)&t.d_stats, PG_STAT_WORDS * 8));
    t.capacity = cap; t.mode = PG_MODE_CANONICAL; t.k = k; t.epoch = 1;
    PG(pg_table_clear(&t, NULL));                        /* fresh memory: write every slot once */
    if (n_rec > 0) PG(pg_kmer_insert(&t, d_pk2, d_amb, d_off, n_rec, off[0], off[n_rec], NULL));
    PG(pg_table_count(&t, NULL));
    int64_t st[PG_STAT_WORDS];
    uint64_t sum1[3], *d_sum;
    CK(cudaMalloc((void **)&d_sum, 3 * 8));
    PG(pg_table_checksum(&t, d_sum, NULL));
    CK(cudaMemcpy(st, t.d_stats, sizeof st, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sum1, d_sum, sizeof sum1, cudaMemcpyDeviceToHost));
    if (st[PG_STAT_OVERFLOW]) { fprintf(stderr, "table overflow\n"); return 4; }
    printf("records %lld  bases %lld  distinct canonical %d-mers %lld  entries %lld  checksum %llu %llu %llu\n",
           (long long)n_rec, (long long)n_bases, k, (long long)st[PG_STAT_USED], (long long)st[PG_STAT_ENTRIES],
           (unsigned long long)sum1[0], (unsigned long long)sum1[1], (unsigned long long)sum1[2]);

    /* the same table once more: O(1) reset (epoch bump), then the two-phase build -
     * K2a buckets one 16-byte update record per position by table region, K3 inserts region by region */
    PG(pg_table_reset(&t, NULL));
    const int sub_bits = cap >= (1 << 19) ? 8 : 0;       /* 256 regions once the table is worth sweeping in pieces */
    const int64_t n_parts = 1ll << sub_bits, part_cap = (n_bases / n_parts) * 5 / 4 + 4096;
    uint64_t *d_rec; int64_t *d_pcnt, *d_seg;
    CK(cudaMalloc((void **)&d_rec, n_parts * part_cap * 16)); CK(cudaMalloc((void **)&d_pcnt, n_parts * 8));
    CK(cudaMalloc((void **)&d_seg, n_parts * 8));
    CK(cudaMemset(d_pcnt, 0, n_parts * 8));
    int64_t *seg = (int64_t *)malloc(n_parts * 8);
    for (int64_t i = 0; i < n_parts; i++) seg[i] = i * part_cap;
    CK(cudaMemcpy(d_seg, seg, n_parts * 8, cudaMemcpyHostToDevice));
    if (n_rec > 0) {
        PG(pg_kmer_partition(&t, d_pk2, d_amb, d_off, n_rec, off[0], off[n_rec], 0, sub_bits, d_rec, part_cap, d_pcnt, NULL, 0, NULL, NULL));
        PG(pg_count_short(&t, d_off, n_rec, off[0], off[n_rec], NULL));
        PG(pg_insert_records(&t, d_rec, d_seg, d_pcnt, (int)n_parts, 1, part_cap, NULL));
    }
    uint64_t sum2[3];
    PG(pg_table_checksum(&t, d_sum, NULL));
    CK(cudaMemcpy(sum2, d_sum, sizeof sum2, cudaMemcpyDeviceToHost));
    if (sum2[0] != sum1[0] || sum2[1] != sum1[1] || sum2[2] != sum1[2]) { fprintf(stderr, "two-phase build disagrees\n"); return 5; }
    printf("two-phase build after pg_table_reset: same checksum\n");
    return 0;
}
