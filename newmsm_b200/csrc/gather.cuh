// Warp-level row gather through bulk asynchronous copies (sm_100a): the device side of gather.cu (barycentric maps and CSR rows).
// See gather.cu for the data path and the measurements behind it.
#pragma once
#include "common.cuh"

namespace msm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// one row: global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_row_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// per-warp slice of shared memory: NST barriers, the (weight, column) records of the staged entries, then the row slots
template <int G, int NST>
struct GatherSmem {
    static __host__ __device__ constexpr size_t header_bytes() { return ((size_t)NST * 8 + (size_t)NST * G * (8 + 4) + 127) / 128 * 128; }
    // slot pitch in 16-byte units = the row length: padding the slots to 128-byte alignment was measured and is SLOWER (0.78 vs 0.62 ms for
    // the barycentric maps, profiles/r2h_tune_gather.txt: the larger ring costs more than the alignment gains)
    static __host__ __device__ int slot_f4(int D4) { return D4; }
    static __host__ __device__ size_t warp_bytes(int D4) { return header_bytes() + (size_t)NST * G * slot_f4(D4) * 16; }
};

// The ring of one warp. n_issued / n_done count the groups issued / consumed since the barriers were initialised
// (stage = n % NST, phase parity = (n / NST) & 1): the ring can be reused across tiles without re-initialisation.
template <int G, int NST>
struct GatherRing {
    unsigned long long* bars;
    double* s_w;
    int* s_c;
    float4* slots;
    uint32_t n_issued, n_done;
    __device__ __forceinline__ void init(unsigned char* base, int lane) {
        bars = reinterpret_cast<unsigned long long*>(base);
        s_w = reinterpret_cast<double*>(base + NST * 8);
        s_c = reinterpret_cast<int*>(base + NST * 8 + NST * G * 8);
        slots = reinterpret_cast<float4*>(base + GatherSmem<G, NST>::header_bytes());
        n_issued = n_done = 0;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NST; ++i) mbar_init(smem_u32(bars + i), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // initialised barriers visible to the async proxy
        }
        __syncwarp();
    }
};

// Entry source of the stand-alone kernels: (col, val) arrays in global memory. The entries are read 32 at a time (one coalesced load
// per lane), ONE WINDOW AHEAD of the group being issued, and handed to the issuing lanes by shuffles: the address of a row copy
// never waits for a column load (the first version loaded col[e] right before issuing the copy of entry e and spent most of its
// time in that dependent DRAM round trip; profiles/r2c_tune_gather_first_version.txt).
struct EntryWindow {
    const int* __restrict__ col;
    const double* __restrict__ val;
    int e_begin, e_end;      // absolute entry range of the tile
    int k;                   // window index of (c0, w0); (c1, w1) is window k + 1
    int c0, c1;
    double w0, w1;
    __device__ __forceinline__ void load(int kk, int lane, int& c, double& w) const {
        const int e = e_begin + kk * 32 + lane;
        const bool ok = e < e_end;
        c = ok ? __ldg(col + e) : -1;
        w = ok ? __ldg(val + e) : 0.0;
    }
    __device__ __forceinline__ void start(const int* col_, const double* val_, int e_begin_, int e_end_, int lane) {
        col = col_; val = val_; e_begin = e_begin_; e_end = e_end_; k = 0;
        load(0, lane, c0, w0);
        load(1, lane, c1, w1);
    }
    // entries [first, first + G) relative to e_begin are about to be requested (G <= 32: they lie in windows k, k + 1 afterwards)
    __device__ __forceinline__ void advance(int first, int lane) {
        while ((first >> 5) > k) {
            c0 = c1; w0 = w1; ++k;
            load(k + 1, lane, c1, w1);
        }
    }
    // every lane of the warp calls this (shuffles); e_rel may differ per lane
    __device__ __forceinline__ void get(int e_rel, int& c, double& w) const {
        const int src = e_rel & 31;
        const int a0 = __shfl_sync(0xffffffffu, c0, src), a1 = __shfl_sync(0xffffffffu, c1, src);
        const double b0 = __shfl_sync(0xffffffffu, w0, src), b1 = __shfl_sync(0xffffffffu, w1, src);
        const bool first = (e_rel >> 5) == k;
        c = first ? a0 : a1;
        w = first ? b0 : b1;
    }
};

// Gathers one tile: `n_entries` entries (numbered from 0) feeding `rows` output rows starting at out4 (row stride D4 float4).
//   BARY: row r owns the entries [3r, 3r + 3), absent entries have col < 0 (G is a multiple of 3).
//   CSR : row r owns [rp_r - rp_0, rp_{r+1} - rp_0) where lane l holds rp = rowptr[r0 + min(l, rows)] and rp_end = rowptr[r0 + rows];
//         rows <= 32.
// FP64 accumulation in entry order, one rounding to FP32 per output value. All 32 lanes call this together.
// The record of a staged entry is its column with two flags: bit 30 = last entry of its row (the consumer stores the row after it),
// negative = absent (-2: absent AND last, barycentric maps only). The flags are computed by the issuing lanes (CSR: a five-step binary
// search over the row ends held by the lanes), so the consumer's per-entry work is straight-line code.
constexpr int kLastFlag = 1 << 30;

template <int G, int NST, int CH, bool BARY, typename Src>
__device__ __forceinline__ void gather_tile(GatherRing<G, NST>& ring, Src& src, int n_entries, int rows, int rp, int rp_end,
                                            const float4* __restrict__ in4, float4* __restrict__ out4, int D4, int lane) {
    static_assert(!BARY || G % 3 == 0, "barycentric maps: three entry slots per row");
    static_assert(G <= 32, "one lane issues one row copy");
    const uint32_t row_bytes = (uint32_t)D4 * 16u;
    const int SP = GatherSmem<G, NST>::slot_f4(D4);      // slot pitch in 16-byte units
    const int n_groups = (n_entries + G - 1) / G;

    // CSR: end of this lane's row relative to the tile's first entry (lanes >= rows: the tile end), rows without entries
    int my_end = 0;
    unsigned todo = 0;          // CSR: rows that still have to be stored, in order (bit r = row r has entries)
    if (!BARY) {
        const int rp0 = __shfl_sync(0xffffffffu, rp, 0);
        const int nxt = __shfl_down_sync(0xffffffffu, rp, 1);
        my_end = (lane + 1 < rows ? nxt : rp_end) - rp0;
        const int my_begin = rp - rp0;
        todo = __ballot_sync(0xffffffffu, lane < rows && my_end > my_begin);
        unsigned empty = __ballot_sync(0xffffffffu, lane < rows && my_end == my_begin);
        while (empty) {          // an empty map resamples to zeros
            const int r = __ffs((int)empty) - 1;
            empty &= empty - 1;
            for (int c = lane; c < D4; c += 32) __stcs(out4 + (size_t)r * D4 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }

    auto issue = [&](int g) {
        const uint32_t st = ring.n_issued % NST;
        const uint32_t bar = smem_u32(ring.bars + st);
        src.advance(g * G, lane);
        const int e_rel = g * G + min(lane, G - 1);
        int c;
        double w;
        src.get(e_rel, c, w);
        const bool present = lane < G && e_rel < n_entries && c >= 0;
        bool last;
        if (BARY) {
            last = (e_rel % 3) == 2;
        } else {   // is e_rel + 1 the end of a row? lower bound over the (ascending) row ends held by the lanes
            int pos = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, my_end, pos + step - 1);
                if (v < e_rel + 1) pos += step;
            }
            last = __shfl_sync(0xffffffffu, my_end, pos) == e_rel + 1;
        }
        const uint32_t n_present = (uint32_t)__popc(__ballot_sync(0xffffffffu, present));
        if (lane == 0) mbar_arrive_expect_tx(bar, n_present * row_bytes);
        if (lane < G) {
            ring.s_w[st * G + lane] = w;
            ring.s_c[st * G + lane] = present ? (c | (last ? kLastFlag : 0)) : (last && e_rel < n_entries ? -2 : -1);
        }
        if (present) bulk_row_g2s(smem_u32(ring.slots + (size_t)(st * G + lane) * SP), in4 + (size_t)c * D4, row_bytes, bar);
        ++ring.n_issued;
    };

    const int ahead = min(NST - 1, n_groups);
    for (int g = 0; g < ahead; ++g) issue(g);
    __syncwarp();

    double acc[CH][4];
#pragma unroll
    for (int h = 0; h < CH; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.0;
    int bary_row = 0;

    for (int g = 0; g < n_groups; ++g) {
        if (g + NST - 1 < n_groups) issue(g + NST - 1);          // re-arms the stage consumed in the previous iteration
        const uint32_t st = ring.n_done % NST;
        mbar_wait(smem_u32(ring.bars + st), (ring.n_done / NST) & 1u);
        const float4* __restrict__ sl = ring.slots + (size_t)st * G * SP;
        // all shared-memory reads of the group are issued before the first conversion (the per-entry form, LDS -> F2F -> DMUL -> DADD
        // in a chain, was 27 % slower at 16 resident warps: profiles/r2_gather_summary.md)
        float4 vv[G][CH];
        double ww[G];
        int cc[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            cc[j] = ring.s_c[st * G + j];
            ww[j] = ring.s_w[st * G + j];
#pragma unroll
            for (int h = 0; h < CH; ++h) {
                const int c = lane + 32 * h;
                vv[j][h] = c < D4 ? sl[j * SP + c] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int cj = cc[j];
            const double wj = ww[j];
            if (cj >= 0) {      // a present entry multiplies even when its weight is 0 (NaN * 0 = NaN, resampler.cpp:46-48)
#pragma unroll
                for (int h = 0; h < CH; ++h) {
                    const float4 v = vv[j][h];
                    acc[h][0] += (double)v.x * wj; acc[h][1] += (double)v.y * wj;
                    acc[h][2] += (double)v.z * wj; acc[h][3] += (double)v.w * wj;
                }
            }
            if (cj == -2 || (cj >= 0 && (cj & kLastFlag))) {     // the row is complete: one rounding to FP32, 16-byte chunks per lane
                int row;
                if (BARY) { row = bary_row++; }
                else { row = __ffs((int)todo) - 1; todo &= todo - 1; }
#pragma unroll
                for (int h = 0; h < CH; ++h) {
                    const int c = lane + 32 * h;
                    if (c < D4) __stcs(out4 + (size_t)row * D4 + c, make_float4((float)acc[h][0], (float)acc[h][1], (float)acc[h][2], (float)acc[h][3]));
                    acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.0;
                }
            }
        }
        ++ring.n_done;
        __syncwarp();       // every lane has read the stage (and its entry records) before it is re-armed
    }
}

}  // namespace msm
