// host_io.cu - host-side text writers for the side files the reference leaves next to its input
// (seq2graph, kmer_numba.py:1893-1904 writes <qry>_rdbg_weight.xyz line by line from Python; `mcl`
// writes <xyz>.mcl).  Plain C++ on the host: formatting millions of short lines is the slowest part
// of the drop-in CLI when done in Python.
#include <stdio.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace {
inline char *put_u64(char *p, uint64_t v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
}  // namespace

// "%d_%d\t%d_%d\t%d\n" per edge, in the order given (the caller sorts by first-insertion ordinal)
extern "C" int pg_host_write_xyz(const char *path, const uint64_t *c0, const uint32_t *v0, const uint64_t *c1,
                                 const uint32_t *v1, const uint32_t *w, int64_t n) {
    FILE *f = fopen(path, "wb");
    if (!f) return pg_fail(PG_ERR_INVALID, "pg_host_write_xyz: cannot open %s", path);
    std::vector<char> buf(1 << 22);
    char *p = buf.data(), *end = buf.data() + buf.size() - 128;
    for (int64_t i = 0; i < n; i++) {
        p = put_u64(p, c0[i]); *p++ = '_'; p = put_u64(p, v0[i]); *p++ = '\t';
        p = put_u64(p, c1[i]); *p++ = '_'; p = put_u64(p, v1[i]); *p++ = '\t';
        p = put_u64(p, w[i]); *p++ = '\n';
        if (p > end) { fwrite(buf.data(), 1, (size_t)(p - buf.data()), f); p = buf.data(); }
    }
    fwrite(buf.data(), 1, (size_t)(p - buf.data()), f);
    fclose(f);
    return PG_OK;
}

// one tab-separated line of node names "code_v5" per label; nodes must arrive sorted by (label, code, v5)
extern "C" int pg_host_write_mcl(const char *path, const uint64_t *code, const uint32_t *v5, const int64_t *label, int64_t n) {
    FILE *f = fopen(path, "wb");
    if (!f) return pg_fail(PG_ERR_INVALID, "pg_host_write_mcl: cannot open %s", path);
    std::vector<char> buf(1 << 22);
    char *p = buf.data(), *end = buf.data() + buf.size() - 128;
    for (int64_t i = 0; i < n; i++) {
        if (i) *p++ = (label[i] != label[i - 1]) ? '\n' : '\t';
        p = put_u64(p, code[i]); *p++ = '_'; p = put_u64(p, v5[i]);
        if (p > end) { fwrite(buf.data(), 1, (size_t)(p - buf.data()), f); p = buf.data(); }
    }
    if (n) *p++ = '\n';
    fwrite(buf.data(), 1, (size_t)(p - buf.data()), f);
    fclose(f);
    return PG_OK;
}
