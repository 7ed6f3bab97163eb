// microbench.cu - measurement support: the memory-access pattern of K3 with the key arithmetic
// taken out.  Every operation touches one pseudo-random 16-byte slot of a table region; regions are swept in order
// by the whole grid exactly like k3_insert_records does, so the numbers are the ceiling the memory system (L2
// atomic units, L2 <-> HBM sector traffic, the fabric between the two dies) sets for "one slot load + one atomic per
// update record" - what bench.py reports K3 against beside the HBM byte roofline.
//   mode 0  slot load only (ld.global.cg 16 B)
//   mode 1  load + red.add.u32                      (a repeat occurrence: count += 1)
//   mode 2  load + atom.cas.b128                    (a first occurrence: claim the slot)
//   mode 3  load + (1/3 cas.b128, 2/3 red.add)      (the mix of BASELINE config 2: 3 occurrences per distinct key)
//   mode 4  red.add only (fire and forget, no load)
//   mode 5  load + red.or + red.add
//   +8      additionally stream one 16-byte record per operation from d_records (K3's input stream)
#include "common.cuh"

namespace {
template <int ILP>
__global__ void __launch_bounds__(256)
k_mb_slots(uint64_t *slots, int cap_bits, int region_bits, int64_t ops_per_region, int mode, const uint4 *__restrict__ records,
           uint64_t seed, unsigned long long *sink) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int n_regions = 1 << (cap_bits - region_bits);
    const uint64_t rmask = (1ull << region_bits) - 1ull;
    const bool stream_in = (mode & 8) != 0;
    const int m = mode & 7;
    uint64_t acc = 0;
    for (int b = 0; b < n_regions; b++) {
        for (int64_t j0 = t0; j0 < ops_per_region; j0 += stride * ILP) {
            uint64_t r[ILP], lo[ILP], hi[ILP]; uint64_t *p[ILP];
#pragma unroll
            for (int q = 0; q < ILP; q++) {
                const int64_t j = j0 + q * stride;
                if (j >= ops_per_region) continue;
                const uint64_t op = (uint64_t)b * (uint64_t)ops_per_region + (uint64_t)j;
                r[q] = pg_mix64(op * 0x9E3779B97F4A7C15ull + seed);
                if (stream_in) { uint4 rec = pg_ld_stream(records + op); r[q] ^= rec.x & 1u; }
            }
#pragma unroll
            for (int q = 0; q < ILP; q++) {
                if (j0 + q * stride >= ops_per_region) continue;
                p[q] = slots + 2 * (((uint64_t)b << region_bits) | (r[q] & rmask));
                lo[q] = hi[q] = 0;
                if (m != 4) pg_ld_slot_raw(p[q], lo[q], hi[q]);
            }
#pragma unroll
            for (int q = 0; q < ILP; q++) {
                if (j0 + q * stride >= ops_per_region) continue;
                acc ^= lo[q] + hi[q];
                const bool claim = (m == 2) || (m == 3 && (r[q] >> 40) % 3 == 0);
                if (claim) {
                    uint64_t olo, ohi;
                    pg_cas128(p[q], lo[q], hi[q], r[q], hi[q] + 1, olo, ohi);
                    acc ^= olo;
                } else if (m == 1 || m == 3 || m == 4) {
                    pg_red_add32(reinterpret_cast<uint32_t *>(p[q] + 1) + 1, 1u);
                } else if (m == 5) {
                    pg_red_or32(reinterpret_cast<uint32_t *>(p[q] + 1), (uint32_t)r[q] & 0xFFFu);
                    pg_red_add32(reinterpret_cast<uint32_t *>(p[q] + 1) + 1, 1u);
                }
            }
        }
    }
    if (acc == 0x1234567ull) atomicAdd(sink, 1ull);      // keep the loads alive
}
}  // namespace

// capacity, region_slots: powers of two; n_ops operations in total, spread evenly over the regions; ilp (1, 2, 4, 8)
// independent operations in flight per thread.  d_records (n_ops 16-byte records, any contents) is only read in the
// +8 modes.  grid = ctas_per_sm x SMs.
extern "C" int pg_microbench_slots(uint64_t *d_slots, int64_t capacity, int64_t region_slots, int64_t n_ops, int mode,
                                   int ctas_per_sm, int ilp, const uint64_t *d_records, uint64_t *d_sink, pg_stream_t stream_) {
    if (!d_slots || !d_sink || capacity < 2 || (capacity & (capacity - 1)) || region_slots < 1 || (region_slots & (region_slots - 1)) ||
        region_slots > capacity || n_ops < 1 || ctas_per_sm < 1 || ctas_per_sm > 8 || mode < 0 || mode > 15 || ((mode & 8) && !d_records) ||
        (ilp != 1 && ilp != 2 && ilp != 4 && ilp != 8))
        return pg_fail(PG_ERR_INVALID, "pg_microbench_slots: bad arguments");
    int cap_bits = 0; while ((1ll << cap_bits) < capacity) cap_bits++;
    int region_bits = 0; while ((1ll << region_bits) < region_slots) region_bits++;
    const int64_t ops_per_region = n_ops >> (cap_bits - region_bits);
    if (ops_per_region < 1) return pg_fail(PG_ERR_INVALID, "pg_microbench_slots: fewer operations than regions");
    const int grid = pg_num_sms() * ctas_per_sm;
    cudaStream_t st = (cudaStream_t)stream_;
    const uint4 *recs = reinterpret_cast<const uint4 *>(d_records);
    unsigned long long *sink = reinterpret_cast<unsigned long long *>(d_sink);
#define MB_LAUNCH(I) k_mb_slots<I><<<grid, 256, 0, st>>>(d_slots, cap_bits, region_bits, ops_per_region, mode, recs, 0x51ed270b1ull, sink)
    if (ilp == 1) MB_LAUNCH(1); else if (ilp == 2) MB_LAUNCH(2); else if (ilp == 4) MB_LAUNCH(4); else MB_LAUNCH(8);
#undef MB_LAUNCH
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
