// Row gather with bulk asynchronous copies: out[r][:] = sum over the entries (col, val) of row r, in list order, of in[col][:] * val.
//
// This is the interpolation loop of Resampler::barycentric_data_interpolation (msm-newresampler/src/resampler.cpp:40-52) for BOTH
// resampling methods of the batch path: the three-entry weight maps of get_barycentric_weights (resampler.cpp:142-167; "BARY": three
// slots per row, an absent map entry has col < 0) and the CSR rows of get_adaptive_barycentric_weights (resampler.cpp:72-140).
// FP64 accumulation in ascending-column order (the std::map iteration order of the reference), one rounding to FP32 on output.
//
// sm_100a data path. A source row is D*4 contiguous bytes at a data-dependent address. Every warp owns a private ring of NST stages in
// shared memory; a stage holds G rows. The warp issues one `cp.async.bulk.shared::cluster.global` per row
// (UBLKCP in SASS; the instruction takes uniform operands, so the compiler serialises the lanes of a group with ELECT / R2UR) — the
// copy engine moves the whole row, no register or L1 line is held while the bytes are in flight — and the copies of a stage complete
// on the stage's mbarrier (`mbarrier.arrive.expect_tx` by lane 0 with the byte count of the stage = SYNCS.ARRIVE.TRANS64, complete_tx
// by the copies). The warp then waits on the barrier's phase (`mbarrier.try_wait.parity` = SYNCS.PHASECHK.TRANS64.TRYWAIT), reads the
// rows back as 128-bit words (lane c owns the 16-byte chunk c of every row: conflict-free; the accumulators of a row live in
// registers across stages) and re-arms the stage NST-1 groups ahead. The (column, weight) entries are read 32 at a time one window
// AHEAD of the group being issued and handed round by shuffles, so a copy's address never waits for a column load.
// The grid is persistent (CTAs = SMs * resident CTAs per SM); warp w takes the row tiles w, w + W, w + 2W, ...
//
// Measured on B200 (profiles/r2_gather_summary.md; 64 subjects, 32 492 targets, D = 100, tools/tune_gather.py):
//   barycentric maps : 0.615 ms = 5.5 TB/s of algorithmic bytes = 0.845 of the measured HBM peak (ncu: DRAM bytes 1.10x algorithmic);
//                      the register-path phase B it replaces could not be timed alone, the fused kernel runs at 0.56.
//   CSR rows         : 3.3 - 3.5 ms against 1.72 ms for k_csr_apply_f32x4 — NEGATIVE result, so weights.cu keeps the register path
//                      (knob "gather_csr"). The adaptive matrix re-reads every source row ~3x (from L2) and has ~15 entries per
//                      output row: the kernel is bound by instruction issue, not by bytes in flight. ncu: 70 warp instructions per
//                      entry of which 12 are the FP32->FP64 conversions (XU pipe, 16 lanes/clk) and FP64 mul/add that both paths
//                      need; per-copy issue costs ~10 instructions (ELECT + 4 R2UR + UBLKCP + branch), and only 25 of 32 lanes hold a
//                      chunk of a 400-byte row. The copy engine itself is not the limit (tools/bulk_copy_probe.cu: 33 G copies/s from
//                      L2, 13.8 G copies/s = 5.5 TB/s from HBM for random 400-byte rows).
#include "gather.cuh"

#include <cstdlib>

namespace msm {

// CH = ceil(D4 / 32): 16-byte chunks of a row per lane. BARY: rows have three entry slots at [3r, 3r+3), absent entries col < 0
// (G must then be a multiple of 3); otherwise CSR with absolute offsets rowptr[r] .. rowptr[r+1] into col / val.
template <int G, int NST, int CH, bool BARY, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (CH == 1 && G <= 8) ? (WARPS == 8 ? 2 : 4) : 1) k_gather_rows_bulk(const GatherJob* __restrict__ jobs, int n_jobs, int n_rows, int D4, int R_csr) {
    // R = rows per warp tile (CSR: <= 32, one row start per lane). The tiles in flight at any moment are consecutive: #warps * R rows.
    // For CSR rows that window decides the L2 reuse of source rows shared by neighbouring targets (profiles/r2f)
    const int R = BARY ? 96 : R_csr;     // barycentric maps: 288 entries per tile, a compile-time constant (the unrolled consumer depends on it)
    extern __shared__ __align__(128) unsigned char g_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GatherRing<G, NST> ring;
    ring.init(g_smem + (size_t)warp * GatherSmem<G, NST>::warp_bytes(D4), lane);

    const int tiles_per_job = (n_rows + R - 1) / R;
    const long long total = (long long)tiles_per_job * n_jobs;
    const long long gwarp = (long long)blockIdx.x * WARPS + warp, nwarps = (long long)gridDim.x * WARPS;
    for (long long tile = gwarp; tile < total; tile += nwarps) {
        const GatherJob job = jobs[tile / tiles_per_job];
        const int r0 = (int)(tile % tiles_per_job) * R;
        const int rows = min(R, n_rows - r0);
        int rp = 0, rp_end = 0, e_begin, e_end;
        if (BARY) {
            e_begin = 3 * r0;
            e_end = 3 * (r0 + rows);
        } else {
            rp = __ldg(job.rowptr + r0 + min(lane, rows));       // lane l: start of row r0 + l
            rp_end = __ldg(job.rowptr + r0 + rows);
            e_begin = __shfl_sync(0xffffffffu, rp, 0);
            e_end = rp_end;
        }
        EntryWindow src;
        src.start(job.col, job.val, e_begin, e_end, lane);
        gather_tile<G, NST, CH, BARY>(ring, src, e_end - e_begin, rows, rp, rp_end, reinterpret_cast<const float4*>(job.in),
                                      reinterpret_cast<float4*>(job.out) + (size_t)r0 * D4, D4, lane);
    }
}

struct GatherConfig { int G, NST, warps; };

template <int G, int NST, int CH, bool BARY, int WARPS>
static msmgpu_status launch_one(const GatherJob* d_jobs, int n_jobs, int n_rows, int D4, int device, cudaStream_t s, int cap) {
    auto* kern = k_gather_rows_bulk<G, NST, CH, BARY, WARPS>;
    const size_t smem = (size_t)WARPS * GatherSmem<G, NST>::warp_bytes(D4);
    if (smem > 227 * 1024) return MSMGPU_ERR_CAPACITY;
    MSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 0;
    MSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    MSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) return MSMGPU_ERR_CAPACITY;
    if (cap > 0) per_sm = std::min(per_sm, cap);
    const int R = BARY ? 96 : std::max(1, std::min(32, tuning_get("gather_rows", "MSMGPU_GATHER_ROWS", 8)));
    const long long tiles = (long long)((n_rows + R - 1) / R) * n_jobs;
    const long long want = (tiles + WARPS - 1) / WARPS;
    const unsigned grid = (unsigned)std::min<long long>((long long)sms * per_sm, std::max<long long>(want, 1));
    kern<<<grid, WARPS * 32, smem, s>>>(d_jobs, n_jobs, n_rows, D4, R);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

template <int CH, bool BARY>
static msmgpu_status launch_ch(int variant, const GatherJob* d_jobs, int n_jobs, int n_rows, int D4, int device, cudaStream_t s, int cap) {
    // variants = (rows per stage, stages, warps per CTA); tuning knobs MSMGPU_GATHER_VARIANT / MSMGPU_GATHER_VARIANT_BARY (profiles/).
    // Rows longer than 512 bytes (CH > 1) only have the small-stage form: the staged rows of a group live in registers while they are summed.
    if constexpr (CH > 1) {
        if constexpr (BARY) return launch_one<3, 4, CH, true, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
        else return launch_one<4, 4, CH, false, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
    } else if constexpr (BARY) {
        switch (variant) {
            case 1: return launch_one<12, 4, 1, true, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
            case 3: return launch_one<3, 4, 1, true, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
            case 5: return launch_one<6, 3, 1, true, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
            default: return launch_one<6, 4, 1, true, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
        }
    } else {
        switch (variant) {
            case 2: return launch_one<8, 3, 1, false, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
            case 4: return launch_one<4, 4, 1, false, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
            default: return launch_one<8, 4, 1, false, 8>(d_jobs, n_jobs, n_rows, D4, device, s, cap);
        }
    }
}

// D must be a multiple of 4 and every in / out pointer 16-byte aligned (bulk copies move 16-byte units); rows of at least 128 bytes
// (below that a bulk copy per row is all overhead) and at most 2 KB. The callers keep their register-path kernels for other shapes.
bool gather_bulk_supported(int D) { return (D & 3) == 0 && D >= 32 && D <= 512; }

// tuning / A-B knob: MSMGPU_GATHER=0 keeps the register-path kernels (k_csr_apply_f32x4, fused phase B)
bool gather_bulk_enabled() {
    return tuning_get("gather", "MSMGPU_GATHER", 1) != 0;
}

msmgpu_status launch_gather_rows_bulk(const GatherJob* d_jobs, int n_jobs, int n_rows, int D, bool bary, int device, cudaStream_t s, int max_ctas_per_sm) {
    if (n_jobs <= 0 || n_rows <= 0) return MSMGPU_OK;
    const int variant_csr = tuning_get("gather_variant", "MSMGPU_GATHER_VARIANT", 0);
    const int variant_bary = tuning_get("gather_variant_bary", "MSMGPU_GATHER_VARIANT_BARY", 0);
    const int D4 = D >> 2, CH = (D4 + 31) / 32;
    const int v = bary ? variant_bary : variant_csr;
    const int cap = max_ctas_per_sm;
#define MSM_GATHER_CH(CH_)                                                                             \
    case CH_:                                                                                          \
        return bary ? launch_ch<CH_, true>(v, d_jobs, n_jobs, n_rows, D4, device, s, cap)                   \
                    : launch_ch<CH_, false>(v, d_jobs, n_jobs, n_rows, D4, device, s, cap);
    switch (CH) {
        MSM_GATHER_CH(1)
        MSM_GATHER_CH(2)
        MSM_GATHER_CH(3)
        MSM_GATHER_CH(4)
    }
#undef MSM_GATHER_CH
    return fail(MSMGPU_ERR_INVALID, "gather_rows_bulk: unsupported row length");
}

}  // namespace msm
