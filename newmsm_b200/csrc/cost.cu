// Discrete-optimisation cost evaluation: patch membership and unary cost tables for the three
// non-HO cost classes of msm-newmeshreg/src/DiscreteCostFunction.cpp:
//   Univariate   (cpp:326-383), Multivariate (385-458), PatchwiseMultivariate (620-692),
// with the similarities of similarities.cpp:129-188 (weighted Pearson, weighted SSD) selected by
// get_sim_for_min (similarities.h:48-58).
//
// Work decomposition: one CTA per (control point, label). Phase 1 — the lanes, in groups of G,
// rotate the patch's source points by R(cp,label), find their nearest target triangle and the
// UNPROJECTED barycentric weights (triangle.cpp:145-157) and park (ids, weights) in shared
// memory. Phase 2 — the similarity. Every floating-point sum of the reference is a sequential
// FP64 sum over the patch (P ~ 65) or over the channels (D ~ 40); a lane evaluates one such sum in
// the reference's order, so costs are reproduced to the last bit and the label choices of the host
// solver cannot flip. The sums are short and independent across the 48 678 (cp,label) CTAs, which
// is where the parallelism comes from.
#include "cost.cuh"

#include <cmath>
#include <limits>

namespace msm {

bool host_rotation_matrix(const double* ci, const double* index, double* R);

} // namespace msm

namespace msm {

constexpr unsigned kFullMask = kFull;

// ------------------------------------------------------------------------------------------
// patch membership (DiscreteCostFunction.cpp:102-107, 334-351)
//   member <=> 2 R asin(|cp - src| / 2R) < range * MAXSEP(cp)
// The left side is a non-decreasing function of the chord |cp - src|, so per control point there
// is one threshold chord x* with  member <=> chord < x*.  x* is found on the host by bisection
// over the doubles with the host libm (patch_chord_threshold), which keeps asin -- whose CUDA
// implementation is not bit-identical to glibc's -- out of the N_cp x N_src device loop while
// reproducing the reference's decision for every representable chord.
// ------------------------------------------------------------------------------------------
static double geodesic_of_chord(double chord) { return 2 * kRad * std::asin(chord / (2 * kRad)); }

double patch_chord_threshold(double limit) {
    // smallest double x in [0, 2R] with geodesic(x) >= limit, or nextafter(2R) if there is none
    const double hi0 = 2 * kRad;
    if (!(geodesic_of_chord(0.0) < limit)) return 0.0;
    if (geodesic_of_chord(hi0) < limit) return std::nextafter(hi0, std::numeric_limits<double>::infinity());
    uint64_t lo, hi;   // invariant: g(lo) < limit, g(hi) >= limit (positive doubles order like their bit patterns)
    double dlo = 0.0, dhi = hi0;
    memcpy(&lo, &dlo, 8);
    memcpy(&hi, &dhi, 8);
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        double dm;
        memcpy(&dm, &mid, 8);
        if (geodesic_of_chord(dm) < limit) lo = mid; else hi = mid;
    }
    memcpy(&dhi, &hi, 8);
    return dhi;
}

// The same decision on the SQUARED chord: sqrt is correctly rounded, hence monotone over the doubles, so there is a largest s with
// sqrt(s) < thr and `sqrt(s) < thr  <=>  s < chord_sq_threshold(thr)` for every representable s. Found by stepping ulps around thr^2.
double chord_sq_threshold(double thr) {
    if (!(thr > 0.0)) return thr == thr ? 0.0 : thr;            // nobody is a member (NaN stays NaN: every compare false)
    const double inf = std::numeric_limits<double>::infinity();
    if (thr == inf) return inf;
    double s = thr * thr;
    if (s == inf) s = std::numeric_limits<double>::max();
    while (s > 0.0 && std::sqrt(s) >= thr) s = std::nextafter(s, 0.0);       // now sqrt(s) < thr (or s = 0)
    if (std::sqrt(s) >= thr) return 0.0;
    while (std::sqrt(std::nextafter(s, inf)) < thr) s = std::nextafter(s, inf);
    return std::nextafter(s, inf);
}

__device__ __forceinline__ bool in_patch(const V3& c, const double* __restrict__ src, int i, double thr) {
    const V3 s{__ldg(src + 3 * (size_t)i), __ldg(src + 3 * (size_t)i + 1), __ldg(src + 3 * (size_t)i + 2)};
    return vnorm(vsub(c, s)) < thr;
}

// one CTA per control point; pass 0 counts, pass 1 writes the members in ascending source id
template <int PASS>
__global__ void __launch_bounds__(256) k_patch_members(int nsrc, const double* __restrict__ cp, const double* __restrict__ src,
                                                       const double* __restrict__ thr, int* __restrict__ count,
                                                       const int* __restrict__ rowptr, int* __restrict__ members) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int k = blockIdx.x;
    const V3 c{cp[3 * (size_t)k], cp[3 * (size_t)k + 1], cp[3 * (size_t)k + 2]};
    const double t = thr[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = PASS ? rowptr[k] : 0;
    __syncthreads();
    for (int base = 0; base < nsrc; base += 256) {
        const int i = base + threadIdx.x;
        const bool m = i < nsrc && in_patch(c, src, i, t);
        const unsigned bal = __ballot_sync(kFullMask, m);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (PASS && m) members[off + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (!PASS && threadIdx.x == 0) count[k] = s_base;
}

__global__ void k_max_i32(int n, const int* __restrict__ v, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(out, v[i]);
}

// ------------------------------------------------------------------------------------------
// unary cost table
// ------------------------------------------------------------------------------------------
struct UnaryArgs {
    TreeView tree;
    int kind, simmeasure, ncp, L, nsrc, D, nvt, cfw_rows, n_patch;
    double percentile;
    const double* R;          // [L][ncp][9]
    const double* src_xyz;    // [nsrc][3]
    const int* prow;          // [ncp+1]
    const int* pmem;          // [n_patch]
    const double* src_feat;   // [nsrc][D]
    const double* ref_feat;   // [nvt][D]
    const double* cfw;        // [nsrc][cfw_rows]
    const double* absw;       // [ncp]
    double* out;              // [L][ncp]
    int* tri_out;             // [L][n_patch] or NULL
    int* err;                 // set to 1 when a query finds no triangle
    // multivariate kind: per-thread staging tiles [D][kCostThreads] doubles in shared memory at byte offset stage_off
    // (0 = none, 1 = blended target values, 2 = target + source values); see the MULTIVARIATE branch of the kernel
    int stage;
    unsigned stage_off;
    int lb;                   // labels per CTA (<= kMaxLabelBlock)
};

constexpr int kCostThreads = 64;
constexpr int kMaxLabelBlock = 8;   // labels per CTA (UnaryArgs::lb)

// One CTA (64 threads) per (control point k, block of `lb` labels). A patch of this size (~67 points at range 1) fills a 64-thread
// round only once and a bit, so with one label per CTA the second round of BOTH phases ran nearly empty; the (label, point) items of a
// label block are therefore flattened: item q -> label l0 + q / P, patch point q % P. The sequential sums that follow (one lane per
// independent sum, the reference's order) then also run for the block's labels side by side in neighbouring lanes.
template <int G>
__global__ void __launch_bounds__(kCostThreads) k_unary_table(UnaryArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x, l0 = blockIdx.y * a.lb, nl = min(a.lb, a.L - l0);
    const int p0 = a.prow[k], P = a.prow[k + 1] - p0;
    const int NP = nl * P;
    const int D = a.D;
    // shared: w[NP][3] doubles | sims[max(NP, nl * D)] doubles | idx[NP][3] ints | staging tiles at stage_off (launch_unary)
    double* s_w = reinterpret_cast<double*>(smem_raw);
    const int n_sims = NP > nl * D ? NP : nl * D;
    double* s_sim = s_w + 3 * (size_t)NP;
    int* s_idx = reinterpret_cast<int*>(s_sim + n_sims);
    __shared__ int s_bad[kMaxLabelBlock];
    if (threadIdx.x < kMaxLabelBlock) s_bad[threadIdx.x] = 0;
    __syncthreads();

    // phase 1 (get_target_data, cpp:353-376 / 410-442 / 652-678): rotate, locate, weights
    const int gl = threadIdx.x % G;
    const int rounds = (NP + kCostThreads / G - 1) / (kCostThreads / G);
    for (int r = 0; r < rounds; ++r) {
        const int q = r * (kCostThreads / G) + threadIdx.x / G;
        const bool active = q < NP;
        const int lq = active ? q / P : 0, i = active ? q - lq * P : 0;
        V3 tmp{0, 0, 0};
        if (active) {
            const double* Rm = a.R + ((size_t)(l0 + lq) * a.ncp + k) * 9;
            double R[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) R[j] = __ldg(Rm + j);
            const int sv = __ldg(a.pmem + p0 + i);
            tmp = mat_apply(R, V3{__ldg(a.src_xyz + 3 * (size_t)sv), __ldg(a.src_xyz + 3 * (size_t)sv + 1), __ldg(a.src_xyz + 3 * (size_t)sv + 2)});
        }
        int st;
        const int t = nearest_triangle<G>(a.tree, tmp, active, gl, st);
        if (active && gl == 0) {
            if (a.tri_out) a.tri_out[(size_t)(l0 + lq) * a.n_patch + p0 + i] = t;
            if (t < 0) {
                s_bad[lq] = 1;
            } else {
                const double* v = a.tree.rec[t].v;
                double w[3];
                bary_weights_raw(tmp, V3{v[0], v[1], v[2]}, V3{v[3], v[4], v[5]}, V3{v[6], v[7], v[8]}, w);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    s_w[3 * q + j] = w[j];
                    s_idx[3 * q + j] = __ldg(a.tree.tri + 3 * (size_t)t + j);
                }
            }
        }
    }
    __syncthreads();
    bool any_bad = false;
    for (int j = 0; j < nl; ++j) any_bad = any_bad || s_bad[j];
    if (any_bad && threadIdx.x == 0) *a.err = 1;   // the reference throws here (octree.cpp:211); the entries of that label are NaN

    const double* __restrict__ rf = a.ref_feat;
    const double* __restrict__ sf = a.src_feat;
    // target value of item q in channel d (triangle.cpp:156: Aa*va1 + Ab*va2 + Ac*va3); items of a label with a failed query are skipped
    auto tgt = [&](int q, int d) -> double {
        return s_w[3 * q] * __ldg(rf + (size_t)s_idx[3 * q] * D + d) + s_w[3 * q + 1] * __ldg(rf + (size_t)s_idx[3 * q + 1] * D + d) +
               s_w[3 * q + 2] * __ldg(rf + (size_t)s_idx[3 * q + 2] * D + d);
    };
    auto srcv = [&](int i) -> int { return __ldg(a.pmem + p0 + i); };
    const int cr = a.cfw_rows;
    double cost = 0.0;
    const int lt = threadIdx.x;   // the label of this block whose sequential sums this lane owns (lt < nl)
    if (a.kind == MSMGPU_COST_UNIVARIATE) {   // cpp:378-383
        for (int q = threadIdx.x; q < NP; q += kCostThreads)
            if (!s_bad[q / P]) s_sim[q] = tgt(q, 0);   // parallel gather, sequential sums below
        __syncthreads();
        if (lt < nl && !s_bad[lt]) {
            const double* sim = s_sim + (size_t)lt * P;
            cost = sim_for_min(a.simmeasure, P,
                               [&](int i) { return __ldg(sf + (size_t)srcv(i) * D); },
                               [&](int i) { return sim[i]; },
                               [&](int i) { return cr >= 1 ? __ldg(a.cfw + (size_t)srcv(i) * cr) : 1.0; }, a.percentile);
        }
    } else if (a.kind == MSMGPU_COST_MULTIVARIATE) {   // cpp:444-458: per-vertex similarity across channels, mean over the patch
        // A thread owns an item and walks its D channels sequentially (the reference's summation order). sim_corr reads every
        // value twice (means, then covariances) and neighbouring lanes read different rows, so straight from global memory the kernel
        // is bound by L1 wavefronts (ncu: 96 % l1tex, profiles/s3_cost_kernels_ncu.md). The rows are therefore read ONCE, with 16-byte
        // loads, into a thread-private column of a shared tile (no barrier: a thread only reads back what it wrote); the arithmetic on
        // the staged values is the same expression in the same order.
        double* s_b = a.stage >= 1 ? reinterpret_cast<double*>(smem_raw + a.stage_off) + threadIdx.x : nullptr;
        double* s_a = a.stage >= 2 ? s_b + (size_t)D * kCostThreads : nullptr;
        for (int q = threadIdx.x; q < NP; q += kCostThreads) {
            if (s_bad[q / P]) continue;
            const int sv = srcv(q % P);
            if (s_b) {
                const double w0 = s_w[3 * q], w1 = s_w[3 * q + 1], w2 = s_w[3 * q + 2];
                const double *r0 = rf + (size_t)s_idx[3 * q] * D, *r1 = rf + (size_t)s_idx[3 * q + 1] * D, *r2 = rf + (size_t)s_idx[3 * q + 2] * D;
                const double* ra = sf + (size_t)sv * D;
                if ((D & 1) == 0) {   // rows start on 16-byte boundaries
                    for (int d = 0; d < D; d += 2) {
                        const double2 x0 = __ldg(reinterpret_cast<const double2*>(r0 + d)), x1 = __ldg(reinterpret_cast<const double2*>(r1 + d)),
                                      x2 = __ldg(reinterpret_cast<const double2*>(r2 + d));
                        s_b[(size_t)d * kCostThreads] = w0 * x0.x + w1 * x1.x + w2 * x2.x;
                        s_b[(size_t)(d + 1) * kCostThreads] = w0 * x0.y + w1 * x1.y + w2 * x2.y;
                        if (s_a) {
                            const double2 xa = __ldg(reinterpret_cast<const double2*>(ra + d));
                            s_a[(size_t)d * kCostThreads] = xa.x;
                            s_a[(size_t)(d + 1) * kCostThreads] = xa.y;
                        }
                    }
                } else {
                    for (int d = 0; d < D; ++d) {
                        s_b[(size_t)d * kCostThreads] = w0 * __ldg(r0 + d) + w1 * __ldg(r1 + d) + w2 * __ldg(r2 + d);
                        if (s_a) s_a[(size_t)d * kCostThreads] = __ldg(ra + d);
                    }
                }
            }
            s_sim[q] = sim_for_min(a.simmeasure, D,
                                   [&](int d) { return s_a ? s_a[(size_t)d * kCostThreads] : __ldg(sf + (size_t)sv * D + d); },
                                   [&](int d) { return s_b ? s_b[(size_t)d * kCostThreads] : tgt(q, d); },
                                   [&](int d) { return cr >= d + 1 ? __ldg(a.cfw + (size_t)sv * cr + d) : 1.0; }, a.percentile);
        }
        __syncthreads();
        if (lt < nl && !s_bad[lt]) {
            for (int i = 0; i < P; ++i) cost += s_sim[(size_t)lt * P + i];
            if (P > 0) cost /= P;
        }
    } else {   // cpp:681-692: per-channel similarity across the patch, mean over channels
        for (int q = threadIdx.x; q < nl * D; q += kCostThreads) {
            const int lq = q / D, d = q - lq * D;
            if (s_bad[lq]) continue;
            s_sim[q] = sim_for_min(a.simmeasure, P,
                                   [&](int i) { return __ldg(sf + (size_t)srcv(i) * D + d); },
                                   [&](int i) { return tgt(lq * P + i, d); },
                                   [&](int i) { return cr >= 1 ? __ldg(a.cfw + (size_t)srcv(i) * cr) : 1.0; }, a.percentile);
        }
        __syncthreads();
        if (lt < nl && !s_bad[lt]) {
            for (int d = 0; d < D; ++d) cost += s_sim[(size_t)lt * D + d];
            cost /= D;
        }
    }
    if (lt < nl) a.out[(size_t)(l0 + lt) * a.ncp + k] = s_bad[lt] ? nan("") : __ldg(a.absw + k) * cost;
}

// CSR lists {i : |cp_k - src_i| < thr_k}, ascending i, for n_cp centres (count pass, scan, write pass); synchronises
msmgpu_status build_patch_lists(int n_cp, const double* d_cp, int n_src, const double* d_src, const double* d_thr, DevBuf<int>& prow,
                                DevBuf<int>& pmem, int& total, int& max_len, cudaStream_t s) {
    DevBuf<int> count, d_tot;
    MSM_CUDA(count.alloc(n_cp, s));
    MSM_CUDA(d_tot.alloc(2, s));
    MSM_CUDA(prow.alloc((size_t)n_cp + 1, s));
    k_patch_members<0><<<n_cp, 256, 0, s>>>(n_src, d_cp, d_src, d_thr, count.p, nullptr, nullptr);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(count.p, prow.p, n_cp, d_tot.p, s));
    MSM_CUDA(cudaMemcpyAsync(prow.p + n_cp, d_tot.p, sizeof(int), cudaMemcpyDeviceToDevice, s));
    MSM_CUDA(cudaMemsetAsync(d_tot.p + 1, 0, sizeof(int), s));
    k_max_i32<<<(n_cp + 255) / 256, 256, 0, s>>>(n_cp, count.p, d_tot.p + 1);
    MSM_LAUNCH_CHECK();
    int h[2];
    MSM_CUDA(cudaMemcpyAsync(h, d_tot.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    total = h[0];
    max_len = h[1];
    MSM_CUDA(pmem.alloc((size_t)total, s));
    if (total > 0) {
        k_patch_members<1><<<n_cp, 256, 0, s>>>(n_src, d_cp, d_src, d_thr, nullptr, prow.p, pmem.p);
        MSM_LAUNCH_CHECK();
    }
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

static size_t unary_smem_bytes(int max_patch, int D) {
    const size_t n_sims = (size_t)(max_patch > D ? max_patch : D);
    return 3 * (size_t)max_patch * sizeof(double) + n_sims * sizeof(double) + 3 * (size_t)max_patch * sizeof(int) + 16;
}

template <int G>
static msmgpu_status launch_unary_g(const UnaryArgs& a, size_t smem, cudaStream_t s) {
    if (smem > 48 * 1024) MSM_CUDA(cudaFuncSetAttribute(k_unary_table<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_unary_table<G><<<dim3((unsigned)a.ncp, (unsigned)((a.L + a.lb - 1) / a.lb)), kCostThreads, smem, s>>>(a);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

static msmgpu_status launch_unary(UnaryArgs& a, int max_patch, cudaStream_t s) {
    // labels per CTA: up to `unary_labels` (default 4) while the per-item arrays of the largest patch stay within 24 KB
    a.lb = tuning_get("unary_labels", "MSMGPU_UNARY_LABELS", 4);
    if (a.lb > kMaxLabelBlock) a.lb = kMaxLabelBlock;
    if (a.lb > a.L) a.lb = a.L;
    if (a.lb < 1) a.lb = 1;
    while (a.lb > 1 && unary_smem_bytes(max_patch * a.lb, a.D * a.lb) > 24 * 1024) --a.lb;
    a.stage = 0;
    const size_t kStageBudget = 28 * 1024;   // eight CTAs per SM
    if (a.kind == MSMGPU_COST_MULTIVARIATE) {
        // Staging tiles while eight CTAs still fit an SM, and the label block shrinks before the tile is given up (measured, ico4 grid on
        // ico6 data, 19 labels, D = 40, whole call: no tile 4.37 ms; target tile 3.18 ms; target + source tiles 3.56 ms; tile with 2 labels
        // per CTA 3.14 ms; 4 labels per CTA without room for the tile 4.46 ms. D = 100, where a tile leaves four CTAs per SM: 7.66 ms
        // without, 8.35 ms with). Knob unary_stage = 0 / 1 / 2 tiles.
        const size_t tile = (size_t)a.D * kCostThreads * sizeof(double);
        const int want = tuning_get("unary_stage", "MSMGPU_UNARY_STAGE", 1);
        if (want >= 1 && unary_smem_bytes(max_patch, a.D) + 16 + tile <= kStageBudget) {
            while (a.lb > 1 && ((unary_smem_bytes(max_patch * a.lb, a.D * a.lb) + 15) & ~(size_t)15) + tile > kStageBudget) --a.lb;
            const size_t base = (unary_smem_bytes(max_patch * a.lb, a.D * a.lb) + 15) & ~(size_t)15;
            a.stage = (want >= 2 && base + 2 * tile <= kStageBudget) ? 2 : 1;
        }
    }
    size_t smem = unary_smem_bytes(max_patch * a.lb, a.D * a.lb);
    if (smem > 200 * 1024) return fail(MSMGPU_ERR_CAPACITY, "unary_table: patch too large for shared memory");
    a.stage_off = (unsigned)((smem + 15) & ~(size_t)15);
    if (a.stage) smem = a.stage_off + a.stage * (size_t)a.D * kCostThreads * sizeof(double);
    switch (query_group_width()) {
        case 1: return launch_unary_g<1>(a, smem, s);
        case 2: return launch_unary_g<2>(a, smem, s);
        case 4: return launch_unary_g<4>(a, smem, s);
        case 16: return launch_unary_g<16>(a, smem, s);
        case 32: return launch_unary_g<32>(a, smem, s);
        default: return launch_unary_g<8>(a, smem, s);
    }
}

template <typename T>
static msmgpu_status upload_vec(DevBuf<T>& b, const T* host, size_t n, cudaStream_t s) {
    MSM_CUDA(b.alloc(n, s));
    if (n) MSM_CUDA(cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return MSMGPU_OK;
}

// host channel-major [D][n] doubles -> device rows [n][D]
static msmgpu_status upload_rows(DevBuf<double>& rows, const double* host_cm, int D, int n, cudaStream_t s) {
    DevBuf<double> cm;
    MSM_TRY(upload_vec(cm, host_cm, (size_t)D * n, s));
    MSM_CUDA(rows.alloc((size_t)D * n, s));
    MSM_TRY(launch_transpose_f64(D, n, cm.p, rows.p, s));
    MSM_CUDA(cudaStreamSynchronize(s));   // `cm` is released after the transpose has consumed it
    return MSMGPU_OK;
}

} // namespace msm

using namespace msm;

extern "C" {

msmgpu_status msmgpu_costfn_create(msmgpu_octree* target_tree, msmgpu_cost_kind kind, int simmeasure, int nsrc, const double* source_xyz,
                                   int D, const double* src_feat, const double* ref_feat, msmgpu_costfn** out) {
    if (!target_tree || !out || nsrc <= 0 || D <= 0 || !source_xyz || !src_feat || !ref_feat || kind < 0 || kind > MSMGPU_COST_HO_MULTIVARIATE)
        return fail(MSMGPU_ERR_INVALID, "costfn_create: bad arguments");
    if (simmeasure != 1 && simmeasure != 2 && simmeasure != 4 && simmeasure != 5)
        return fail(MSMGPU_ERR_INVALID, "Unknown similarity metric");   // similarities.h:57 (1 SSD, 2 correlation, 4 DICE, 5 genDICE)
    *out = nullptr;
    msmgpu_ctx* ctx = target_tree->mesh->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    auto c = std::unique_ptr<msmgpu_costfn>(new msmgpu_costfn());
    c->ctx = ctx; c->tree = target_tree; c->kind = kind; c->simmeasure = simmeasure;
    c->nsrc = nsrc; c->D = D; c->nvt = target_tree->mesh->nv;
    MSM_TRY(upload_vec(c->src_xyz, source_xyz, 3 * (size_t)nsrc, s));
    MSM_TRY(upload_rows(c->src_feat, src_feat, D, nsrc, s));
    MSM_TRY(upload_rows(c->ref_feat, ref_feat, D, c->nvt, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    *out = c.release();
    return MSMGPU_OK;
}

void msmgpu_costfn_destroy(msmgpu_costfn* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    delete c;
}

msmgpu_status msmgpu_costfn_reset_source(msmgpu_costfn* c, const double* source_xyz) {
    if (!c || !source_xyz) return fail(MSMGPU_ERR_INVALID, "costfn_reset_source: bad arguments");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    MSM_CUDA(cudaMemcpyAsync(c->src_xyz.p, source_xyz, 3 * (size_t)c->nsrc * sizeof(double), cudaMemcpyHostToDevice, c->ctx->stream));
    MSM_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return MSMGPU_OK;
}

msmgpu_status msmgpu_costfn_set_cpgrid(msmgpu_costfn* c, int ncp, const double* cp_xyz, const double* maxsep, double range,
                                       int cfw_rows, const double* cfw, const double* absw) {
    if (!c || ncp <= 0 || !cp_xyz || !maxsep || !absw || cfw_rows < 0 || (cfw_rows > 0 && !cfw))
        return fail(MSMGPU_ERR_INVALID, "costfn_set_cpgrid: bad arguments");
    if (c->kind >= MSMGPU_COST_HO_UNIVARIATE) return fail(MSMGPU_ERR_INVALID, "costfn_set_cpgrid: HO cost functions take msmgpu_costfn_set_cpgrid_ho");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    c->ncp = ncp; c->range = range; c->cfw_rows = cfw_rows;
    c->h_cp.assign(cp_xyz, cp_xyz + 3 * (size_t)ncp);
    std::vector<double> thr(ncp);
#pragma omp parallel for
    for (int k = 0; k < ncp; ++k) thr[k] = patch_chord_threshold(range * maxsep[k]);   // cpp:104: < _controlptrange * MAXSEP(k+1)
    MSM_TRY(upload_vec(c->cp_xyz, cp_xyz, 3 * (size_t)ncp, s));
    MSM_TRY(upload_vec(c->absw, absw, (size_t)ncp, s));
    MSM_TRY(upload_vec(c->chord_thr, thr.data(), (size_t)ncp, s));
    if (cfw_rows > 0) MSM_TRY(upload_rows(c->cfw, cfw, cfw_rows, c->nsrc, s));
    MSM_TRY(build_patch_lists(ncp, c->cp_xyz.p, c->nsrc, c->src_xyz.p, c->chord_thr.p, c->prow, c->pmem, c->n_patch, c->max_patch, s));
    c->n_patch_rows = ncp;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_costfn_set_percentile(msmgpu_costfn* c, double percentile) {
    if (!c || !(percentile >= 0.0 && percentile <= 1.0)) return fail(MSMGPU_ERR_INVALID, "costfn_set_percentile: percentile must be in [0, 1]");
    c->percentile = percentile;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_costfn_patches(msmgpu_costfn* c, int32_t* rowptr, int32_t* members) {
    if (!c || !rowptr || c->n_patch_rows <= 0) return fail(MSMGPU_ERR_INVALID, "costfn_patches: set_cpgrid first");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    MSM_CUDA(cudaMemcpyAsync(rowptr, c->prow.p, ((size_t)c->n_patch_rows + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (members && c->n_patch) MSM_CUDA(cudaMemcpyAsync(members, c->pmem.p, (size_t)c->n_patch * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

// d_out [L][ncp]; d_tri_out [L][n_patch] or NULL; labels / rotations are HOST arrays (see host_rotation_matrix)
msmgpu_status msmgpu_costfn_unary_table_dev(msmgpu_costfn* c, int L, const double* labels, const double* rotations, double* d_out, int32_t* d_tri_out) {
    if (!c || L <= 0 || !labels || !rotations || !d_out) return fail(MSMGPU_ERR_INVALID, "costfn_unary_table: bad arguments");
    if (c->ncp <= 0) return fail(MSMGPU_ERR_INVALID, "costfn_unary_table: set_cpgrid first");
    if (c->kind >= MSMGPU_COST_HO_UNIVARIATE) return fail(MSMGPU_ERR_INVALID, "costfn_unary_table: the HO classes have no unary term (DiscreteCostFunction.h:46 returns 0)");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    const int ncp = c->ncp;
    // R(l,k) = estimate_rotation_matrix(CP_k, ROT_k * label_l)  (cpp:379-380, 445-446, 682-683)
    std::vector<double> R((size_t)L * ncp * 9);
    bool ok = true;
#pragma omp parallel for collapse(2) reduction(&& : ok)
    for (int l = 0; l < L; ++l)
        for (int k = 0; k < ncp; ++k) {
            const double* M = rotations + 9 * (size_t)k;
            const double* lb = labels + 3 * (size_t)l;
            const V3 dest = mat_apply(M, V3{lb[0], lb[1], lb[2]});
            const double de[3] = {dest.x, dest.y, dest.z};
            ok = host_rotation_matrix(c->h_cp.data() + 3 * (size_t)k, de, R.data() + ((size_t)l * ncp + k) * 9) && ok;
        }
    if (!ok) return fail(MSMGPU_ERR_INVALID, "rotation angle is greater than 90 degrees");
    DevBuf<double> d_R;
    MSM_TRY(upload_vec(d_R, R.data(), R.size(), s));
    UnaryArgs a;
    a.tree = c->tree->view();
    a.kind = c->kind; a.simmeasure = c->simmeasure; a.percentile = c->percentile; a.ncp = ncp; a.L = L; a.nsrc = c->nsrc; a.D = c->D; a.nvt = c->nvt;
    a.cfw_rows = c->cfw_rows; a.n_patch = c->n_patch;
    a.R = d_R.p; a.src_xyz = c->src_xyz.p; a.prow = c->prow.p; a.pmem = c->pmem.p;
    a.src_feat = c->src_feat.p; a.ref_feat = c->ref_feat.p; a.cfw = c->cfw.p; a.absw = c->absw.p;
    a.out = d_out; a.tri_out = d_tri_out;
    DevBuf<int> d_err;
    MSM_CUDA(d_err.alloc(1, s));
    MSM_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));
    a.err = d_err.p;
    MSM_TRY(launch_unary(a, c->max_patch, s));
    int h_err = 0;
    MSM_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));   // R (pageable host vector) and d_R are released on return
    if (h_err) return status_to_error(MSMGPU_ERR_NO_TRIANGLE);   // the reference throws (octree.cpp:211); failed entries are NaN
    return MSMGPU_OK;
}

msmgpu_status msmgpu_costfn_unary_table(msmgpu_costfn* c, int L, const double* labels, const double* rotations, double* out, int32_t* tri_out) {
    if (!c || L <= 0 || !out) return fail(MSMGPU_ERR_INVALID, "costfn_unary_table: bad arguments");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    DevBuf<double> d_out;
    DevBuf<int> d_tri;
    MSM_CUDA(d_out.alloc((size_t)L * c->ncp, s));
    if (tri_out) MSM_CUDA(d_tri.alloc((size_t)L * c->n_patch, s));
    const msmgpu_status st = msmgpu_costfn_unary_table_dev(c, L, labels, rotations, d_out.p, tri_out ? d_tri.p : nullptr);
    if (st != MSMGPU_OK && st != MSMGPU_ERR_NO_TRIANGLE) return st;
    MSM_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)L * c->ncp * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (tri_out) MSM_CUDA(cudaMemcpyAsync(tri_out, d_tri.p, (size_t)L * c->n_patch * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return st;
}

} // extern "C"
