// Groupwise registration (gMSM) hot paths.
//
// Reference behaviour restated (msm-newmeshreg/src):
//   DiscreteGroupModel::get_patch_data             DiscreteGroupModel.cpp:88-121
//   DiscreteGroupCostFunction::computePairwiseCost DiscreteGroupCostFunction.cpp:54-97
// called by Fusion::optimize with 4 label combinations per pair and candidate label (Fusion.h:164-174).
//
// The reference materialises, for every (subject, control point, label), a std::map<template vertex,
// feature vector> ("patch") and intersects two maps per pair cost: tens of GB for 64 subjects. Here the
// (subject, label) resampled FIELDS live once on the device ([S][L][N_t][D] rows) and patches are implicit:
//   patch(s,v,l) = { p : 2R asin(|ROT(s,v) label_l - tpl_p| / 2R) < range * spacing(s,v) }
// is a chord test against a per-node threshold (same host bisection as the unary patches, cost.cu), and the
// map intersection of a pair is "p passes both tests", walked in ascending p over a per-node SUPERSET list
// (all labels of that node) so a pair cost touches ~150 candidates instead of the whole template.
// Sums are the reference's sequential FP64 sums (one lane per channel).
//
// Multi-GPU: the fields of a shard of subjects are produced by msmgpu_group_fields on each
// rank and all-gathered by the host side (NCCL); pair blocks are independent, any rank can evaluate any
// pair once it holds all fields.
#include "cost.cuh"

#include <algorithm>
#include <cmath>

struct msmgpu_group {
    msmgpu_ctx* ctx = nullptr;
    int simmeasure = 2, S = 0, ncp = 0, L = 0, D = 0, n_tpl = 0;
    const double* d_fields = nullptr;   // [S][L][n_tpl][D], owned by the caller
    msm::DevBuf<double> tpl_xyz;        // [n_tpl][3]
    msm::DevBuf<double> rcp;            // [S*ncp][L][3] rotated control points
    msm::DevBuf<double> thr;            // [S*ncp] chord thresholds
    msm::DevBuf<double> thr2;           // [S*ncp] the same decision on the squared chord (chord_sq_threshold)
    msm::DevBuf<double> mask;           // [n_tpl] |mask value| per template vertex, empty: unit weights (msmgpu_group_set_mask)
    msm::DevBuf<int> sup_ptr, sup_mem;  // superset candidate lists per node
    int n_sup = 0, max_sup = 0;
    msm::DevBuf<int> pairs;             // optional device-resident pair list [P][2] (msmgpu_group_set_pairs)
    msm::DevBuf<int> labeling;          // scratch for the device-output batches
    int P = 0;
};

namespace msm {

// rcp[node][l] = ROT[node] * label_l ; anchor = rcp[node][0] ; maxdisp[node] = max_l |rcp_l - anchor|
__global__ void k_rotated_cps(int n_nodes, int L, const double* __restrict__ rot, const double* __restrict__ labels, double* __restrict__ rcp,
                              double* __restrict__ anchor, const double* __restrict__ thr, double* __restrict__ thr_super) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) return;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = rot[9 * (size_t)k + i];
    V3 a{0, 0, 0};
    double md = 0.0;
    for (int l = 0; l < L; ++l) {
        const V3 p = mat_apply(R, V3{labels[3 * l], labels[3 * l + 1], labels[3 * l + 2]});
        rcp[((size_t)k * L + l) * 3] = p.x; rcp[((size_t)k * L + l) * 3 + 1] = p.y; rcp[((size_t)k * L + l) * 3 + 2] = p.z;
        if (l == 0) a = p;
        md = fmax(md, vnorm(vsub(p, a)));
    }
    anchor[3 * (size_t)k] = a.x; anchor[3 * (size_t)k + 1] = a.y; anchor[3 * (size_t)k + 2] = a.z;
    // |rcp_l - t| < thr  =>  |anchor - t| < thr + |rcp_l - anchor|  (triangle inequality, padded for rounding)
    thr_super[k] = (thr[k] + md) * (1.0 + 1e-12) + 1e-9;
}

struct PairArgs {
    int simmeasure, ncp, L, D, n_tpl, n;
    const double* fields; const double* tpl; const double* rcp;
    const double* thr2;      // [S*ncp] squared-chord thresholds: chord < thr  <=>  chord^2 < thr2 (chord_sq_threshold, cost.cu)
    const double* mask;      // [n_tpl] |_MASK.get_pvalue(p)| (DiscreteGroupCostFunction.cpp:77), or NULL: unit weights
    const int* sup_ptr; const int* sup_mem;
    const int* pairs;        // [P][2]
    const int* req_pair; const int* req_la; const int* req_lb;   // list mode, or NULL:
    const int* labeling; int label;                              // Fusion mode: request r = 4 * pair + combo
    double* out;
};

constexpr int kPairWarps = 4;

__device__ __forceinline__ double chord2(const V3& c, const V3& t) {   // the argument of the square root in Point::norm (vnorm)
    const V3 d = vsub(c, t);
    return d.x * d.x + d.y * d.y + d.z * d.z;
}

__device__ __forceinline__ void pair_request(const PairArgs& a, int r, int& pair, int& la, int& lb) {
    if (a.req_pair) {
        pair = a.req_pair[r]; la = a.req_la[r]; lb = a.req_lb[r];
    } else {   // Fusion.h:170-173: (cur,cur), (cur,label), (label,cur), (label,label)
        pair = r >> 2;
        const int combo = r & 3;
        la = (combo & 2) ? a.label : a.labeling[a.pairs[2 * (size_t)pair]];
        lb = (combo & 1) ? a.label : a.labeling[a.pairs[2 * (size_t)pair + 1]];
    }
}

// ---- D > 8: one warp per request. Lane d owns channel d (d + 32, ...): the intersection is found 32 candidates at a time
// (one chord test pair per lane, ballot), then every member is visited in ascending template id and each lane
// adds its channel's value -> the reference's sequential sums, coalesced over the channels.
__global__ void __launch_bounds__(kPairWarps * 32) k_group_pair_costs(PairArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kPairWarps + warp;
    if (r >= a.n) return;
    int pair, la, lb;
    pair_request(a, r, pair, la, lb);
    const int nA = a.pairs[2 * (size_t)pair], nB = a.pairs[2 * (size_t)pair + 1];
    const V3 cA = load_pt(a.rcp, nA * a.L + la), cB = load_pt(a.rcp, nB * a.L + lb);
    const double tA = a.thr2[nA], tB = a.thr2[nB];
    const int sA = nA / a.ncp, sB = nB / a.ncp;
    const double* __restrict__ fA = a.fields + ((size_t)sA * a.L + la) * a.n_tpl * a.D;
    const double* __restrict__ fB = a.fields + ((size_t)sB * a.L + lb) * a.n_tpl * a.D;
    const double* __restrict__ mask = a.mask;
    const int b = a.sup_ptr[nA], e = a.sup_ptr[nA + 1];

    // every member of the intersection, in ascending template id, to f(template vertex)
    auto members = [&](auto f) {
        for (int i0 = b; i0 < e; i0 += 32) {
            const int i = i0 + lane;
            int p = -1;
            bool in = false;
            if (i < e) {
                p = __ldg(a.sup_mem + i);
                const V3 t = load_pt(a.tpl, p);
                in = chord2(cA, t) < tA && chord2(cB, t) < tB;
            }
            unsigned m = __ballot_sync(kFull, in);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                f(__shfl_sync(kFull, p, src));
            }
        }
    };

    double cost = 0.0;   // lane 0 accumulates the per-channel similarities in channel order
    int n_common = 0;
    for (int d0 = 0; d0 < a.D; d0 += 32) {
        const int d = d0 + lane;
        const bool has = d < a.D;
        // pass 1: count, weight sum, weighted sums of A and B (similarities.cpp:132-137)
        double sum = 0.0, meanA = 0.0, meanB = 0.0;
        int cnt = 0;
        members([&](int pp) {
            ++cnt;
            const double w = mask ? __ldg(mask + pp) : 1.0;
            sum += w;
            if (has) { meanA += w * fA[(size_t)pp * a.D + d]; meanB += w * fB[(size_t)pp * a.D + d]; }
        });
        n_common = cnt;
        double sim = 0.0;
        if (cnt > 0) {
            if (a.simmeasure == 2) {   // similarities.cpp:139-158
                if (sum > 0.0) { meanA /= sum; meanB /= sum; }
                double prod = 0.0, varA = 0.0, varB = 0.0;
                members([&](int pp) {
                    if (has) {
                        const double w = mask ? __ldg(mask + pp) : 1.0;
                        const double x = fA[(size_t)pp * a.D + d] - meanA, y = fB[(size_t)pp * a.D + d] - meanB;
                        prod += w * x * y; varA += w * x * x; varB += w * y * y;
                    }
                });
                if (sum > 0.0) { prod /= sum; varA /= sum; varB /= sum; }
                const double corr = (varA == 0.0 || varB == 0.0) ? 0.0 : prod / (sqrt(varA) * sqrt(varB));
                sim = 1 - (1 + corr) * 0.5;
            } else {   // SSD, similarities.cpp:179-188
                double prod = 0.0;
                members([&](int pp) {
                    if (has) {
                        const double w = mask ? __ldg(mask + pp) : 1.0;
                        const double dd = fA[(size_t)pp * a.D + d] - fB[(size_t)pp * a.D + d];
                        prod += w * dd * dd;
                    }
                });
                sim = sqrt(prod) / cnt;
            }
        }
        // pair_cost += sim, channel by channel (DiscreteGroupCostFunction.cpp:86-91)
        const int nd = min(32, a.D - d0);
        for (int k = 0; k < nd; ++k) cost += __shfl_sync(kFull, sim, k);
    }
    if (lane == 0) a.out[r] = n_common > 0 ? cost / a.D : nan("");   // empty intersection: the reference reads patch_data_A[0] of an empty vector
}

// ---- D <= 8: one THREAD per request. With a handful of channels a warp per request leaves 31 lanes idle in every sequential sum
// (ncu of the warp kernel at D = 1: profiles/s2_pair_costs_ncu.md). A thread tests its node's superset once (membership bits kept in
// a small private bit array), then walks the members twice -- weighted sums, centred products -- DC channels at a time, with the
// reference's sequential FP64 order per channel. The 4 requests of a pair (Fusion's combinations) sit in neighbouring lanes and read
// the same superset list and template points.
constexpr int kFlagWords = 32;   // membership bits for supersets of up to 1024 candidates; longer lists are re-tested in every walk

template <int DC>
__global__ void __launch_bounds__(128) k_group_pair_costs_thread(PairArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    int pair, la, lb;
    pair_request(a, r, pair, la, lb);
    const int nA = a.pairs[2 * (size_t)pair], nB = a.pairs[2 * (size_t)pair + 1];
    const V3 cA = load_pt(a.rcp, nA * a.L + la), cB = load_pt(a.rcp, nB * a.L + lb);
    const double tA = a.thr2[nA], tB = a.thr2[nB];
    const int sA = nA / a.ncp, sB = nB / a.ncp;
    const int D = a.D;
    const double* __restrict__ fA = a.fields + ((size_t)sA * a.L + la) * a.n_tpl * D;
    const double* __restrict__ fB = a.fields + ((size_t)sB * a.L + lb) * a.n_tpl * D;
    const double* __restrict__ mask = a.mask;
    const int* __restrict__ sup = a.sup_mem + a.sup_ptr[nA];
    const int len = a.sup_ptr[nA + 1] - a.sup_ptr[nA];
    const bool cached = len <= 32 * kFlagWords;
    unsigned flags[kFlagWords];

    auto test = [&](int p) {
        const V3 t = load_pt(a.tpl, p);
        return chord2(cA, t) < tA && chord2(cB, t) < tB;
    };
    int cnt = 0;
    for (int w0 = 0; w0 < len; w0 += 32) {
        unsigned m = 0;
        const int nb = min(32, len - w0);
        for (int k = 0; k < nb; ++k) m |= (unsigned)test(__ldg(sup + w0 + k)) << k;
        if (cached) flags[w0 >> 5] = m;
        cnt += __popc(m);
    }
    if (cnt == 0) { a.out[r] = nan(""); return; }   // empty intersection: the reference reads patch_data_A[0] of an empty vector
    auto members = [&](auto f) {   // ascending template id
        for (int w0 = 0; w0 < len; w0 += 32) {
            unsigned m;
            if (cached) {
                m = flags[w0 >> 5];
            } else {
                m = 0;
                const int nb = min(32, len - w0);
                for (int k = 0; k < nb; ++k) m |= (unsigned)test(__ldg(sup + w0 + k)) << k;
            }
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                f(__ldg(sup + w0 + k));
            }
        }
    };

    double cost = 0.0;
    for (int d0 = 0; d0 < D; d0 += DC) {
        double sum = 0.0, meanA[DC], meanB[DC];
#pragma unroll
        for (int c = 0; c < DC; ++c) meanA[c] = meanB[c] = 0.0;
        members([&](int p) {   // similarities.cpp:132-137
            const double w = mask ? __ldg(mask + p) : 1.0;
            sum += w;
#pragma unroll
            for (int c = 0; c < DC; ++c)
                if (d0 + c < D) { meanA[c] += w * fA[(size_t)p * D + d0 + c]; meanB[c] += w * fB[(size_t)p * D + d0 + c]; }
        });
        double acc[3 * DC];
#pragma unroll
        for (int c = 0; c < 3 * DC; ++c) acc[c] = 0.0;
        if (a.simmeasure == 2) {   // similarities.cpp:139-158
            if (sum > 0.0) {
#pragma unroll
                for (int c = 0; c < DC; ++c) { meanA[c] /= sum; meanB[c] /= sum; }
            }
            members([&](int p) {
                const double w = mask ? __ldg(mask + p) : 1.0;
#pragma unroll
                for (int c = 0; c < DC; ++c)
                    if (d0 + c < D) {
                        const double x = fA[(size_t)p * D + d0 + c] - meanA[c], y = fB[(size_t)p * D + d0 + c] - meanB[c];
                        acc[3 * c] += w * x * y; acc[3 * c + 1] += w * x * x; acc[3 * c + 2] += w * y * y;
                    }
            });
#pragma unroll
            for (int c = 0; c < DC; ++c)
                if (d0 + c < D) {
                    double prod = acc[3 * c], varA = acc[3 * c + 1], varB = acc[3 * c + 2];
                    if (sum > 0.0) { prod /= sum; varA /= sum; varB /= sum; }
                    const double corr = (varA == 0.0 || varB == 0.0) ? 0.0 : prod / (sqrt(varA) * sqrt(varB));
                    cost += 1 - (1 + corr) * 0.5;   // pair_cost += sim, channel by channel (DiscreteGroupCostFunction.cpp:86-91)
                }
        } else {   // SSD, similarities.cpp:179-188
            members([&](int p) {
                const double w = mask ? __ldg(mask + p) : 1.0;
#pragma unroll
                for (int c = 0; c < DC; ++c)
                    if (d0 + c < D) {
                        const double dd = fA[(size_t)p * D + d0 + c] - fB[(size_t)p * D + d0 + c];
                        acc[c] += w * dd * dd;
                    }
            });
#pragma unroll
            for (int c = 0; c < DC; ++c)
                if (d0 + c < D) cost += sqrt(acc[c]) / cnt;
        }
    }
    a.out[r] = cost / D;
}

static msmgpu_status launch_pair_costs(const PairArgs& a, cudaStream_t s) {
    const char* force = std::getenv("MSMGPU_PAIR_KERNEL");   // "warp" / "thread": A/B runs (tools, profiles)
    const bool thread = force ? force[0] == 't' : a.D <= 8;
    if (thread) {
        const unsigned grid = (unsigned)((a.n + 127) / 128);
        if (a.D == 1) k_group_pair_costs_thread<1><<<grid, 128, 0, s>>>(a);
        else if (a.D == 2) k_group_pair_costs_thread<2><<<grid, 128, 0, s>>>(a);
        else k_group_pair_costs_thread<4><<<grid, 128, 0, s>>>(a);
    } else {
        k_group_pair_costs<<<(unsigned)((a.n + kPairWarps - 1) / kPairWarps), kPairWarps * 32, 0, s>>>(a);
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

template <typename T>
static msmgpu_status upg(DevBuf<T>& b, const T* host, size_t n, cudaStream_t s) {
    MSM_CUDA(b.alloc(n, s));
    if (n) MSM_CUDA(cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return MSMGPU_OK;
}

// rotated data-mesh coordinates for every label: out[l][p] = R_p * label_l (l > 0), original for l = 0 (DiscreteGroupModel.cpp:98-103)
__global__ void k_rotate_data(int nv, int L, const double* __restrict__ xyz, const double* __restrict__ Rp, const double* __restrict__ labels,
                              double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nv) return;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rp[9 * (size_t)p + i];
    for (int l = 0; l < L; ++l) {
        V3 q;
        if (l == 0) q = load_pt(xyz, p);
        else q = mat_apply(R, V3{labels[3 * l], labels[3 * l + 1], labels[3 * l + 2]});
        double* o = out + ((size_t)l * nv + p) * 3;
        o[0] = q.x; o[1] = q.y; o[2] = q.z;
    }
}

static msmgpu_status pair_run(msmgpu_group* g, int P, const int32_t* pairs, int n, const int32_t* req_pair, const int32_t* req_la,
                              const int32_t* req_lb, const int32_t* labeling, int label, double* out) {
    if (!g || P <= 0 || !pairs || n <= 0 || !out) return fail(MSMGPU_ERR_INVALID, "group_pair: bad arguments");
    MSM_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = g->ctx->stream;
    DevBuf<int> d_pairs, d_rp, d_la, d_lb, d_labeling;
    DevBuf<double> d_out;
    MSM_TRY(upg(d_pairs, pairs, 2 * (size_t)P, s));
    if (req_pair) {
        MSM_TRY(upg(d_rp, req_pair, (size_t)n, s));
        MSM_TRY(upg(d_la, req_la, (size_t)n, s));
        MSM_TRY(upg(d_lb, req_lb, (size_t)n, s));
    } else {
        MSM_TRY(upg(d_labeling, labeling, (size_t)g->S * g->ncp, s));
    }
    MSM_CUDA(d_out.alloc((size_t)n, s));
    PairArgs a;
    a.simmeasure = g->simmeasure; a.ncp = g->ncp; a.L = g->L; a.D = g->D; a.n_tpl = g->n_tpl; a.n = n;
    a.fields = g->d_fields; a.tpl = g->tpl_xyz.p; a.rcp = g->rcp.p; a.thr2 = g->thr2.p; a.mask = g->mask.p; a.sup_ptr = g->sup_ptr.p; a.sup_mem = g->sup_mem.p;
    a.pairs = d_pairs.p; a.req_pair = req_pair ? d_rp.p : nullptr; a.req_la = d_la.p; a.req_lb = d_lb.p;
    a.labeling = d_labeling.p; a.label = label; a.out = d_out.p;
    MSM_TRY(launch_pair_costs(a, s));
    MSM_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

} // namespace msm

using namespace msm;

extern "C" {

msmgpu_status msmgpu_weights_apply_batch_f64_dev(msmgpu_ctx* ctx, int n, msmgpu_weights* const* ws, int D, const double* const* d_in, double* const* d_out);

msmgpu_status msmgpu_group_fields(msmgpu_ctx* ctx, int n_subjects, int nv, const double* data_xyz, int nt, const int32_t* tri, int D,
                                  const double* feat_cm, int L, const double* labels, const double* centre, msmgpu_mesh* tpl,
                                  msmgpu_octree* tpl_tree, double* d_fields) {
    if (!ctx || n_subjects <= 0 || nv <= 0 || !data_xyz || nt <= 0 || !tri || D <= 0 || !feat_cm || L <= 0 || !labels || !centre || !tpl || !d_fields)
        return fail(MSMGPU_ERR_INVALID, "group_fields: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int n_tpl = tpl->nv;
    std::unique_ptr<msmgpu_octree, void (*)(msmgpu_octree*)> own_tree(nullptr, msmgpu_octree_destroy);
    if (!tpl_tree) {
        msmgpu_octree* t = nullptr;
        MSM_TRY(msmgpu_octree_build(tpl, &t));
        own_tree.reset(t);
        tpl_tree = t;
    }
    DevBuf<int> d_tri;
    DevBuf<double> d_labels;
    MSM_TRY(upg(d_tri, tri, 3 * (size_t)nt, s));
    MSM_TRY(upg(d_labels, labels, 3 * (size_t)L, s));
    std::vector<double> Rp(9 * (size_t)nv);
    for (int sub = 0; sub < n_subjects; ++sub) {
        const double* xyz = data_xyz + 3 * (size_t)sub * nv;
        // estimate_rotation_matrix(centre, vertex) per data vertex, host libm
        bool ok = true;
#pragma omp parallel for reduction(&& : ok)
        for (int p = 0; p < nv; ++p) ok = host_rotation_matrix(centre, xyz + 3 * (size_t)p, Rp.data() + 9 * (size_t)p) && ok;
        if (!ok) return fail(MSMGPU_ERR_INVALID, "rotation angle is greater than 90 degrees");
        DevBuf<double> d_xyz, d_Rp, d_rot, d_feat_cm, d_feat_rows;
        MSM_TRY(upg(d_xyz, xyz, 3 * (size_t)nv, s));
        MSM_TRY(upg(d_Rp, Rp.data(), Rp.size(), s));
        MSM_CUDA(d_rot.alloc(3 * (size_t)L * nv, s));
        k_rotate_data<<<(nv + 255) / 256, 256, 0, s>>>(nv, L, d_xyz.p, d_Rp.p, d_labels.p, d_rot.p);
        MSM_LAUNCH_CHECK();
        MSM_TRY(upg(d_feat_cm, feat_cm + (size_t)sub * D * nv, (size_t)D * nv, s));
        MSM_CUDA(d_feat_rows.alloc((size_t)D * nv, s));
        MSM_TRY(launch_transpose_f64(D, nv, d_feat_cm.p, d_feat_rows.p, s));
        // L rotated meshes -> one forest, one adaptive-weights batch, one apply
        std::vector<msmgpu_mesh*> meshes(L, nullptr);
        std::vector<msmgpu_weights*> ws(L, nullptr);
        msmgpu_status st = MSMGPU_OK;
        {   // the L rotated copies are views of d_rot (no second copy), their tables share one allocation
            std::vector<const double*> rot_ptrs(L);
            for (int l = 0; l < L; ++l) rot_ptrs[l] = d_rot.p + 3 * (size_t)l * nv;
            st = msmgpu_mesh_create_view_batch(ctx, L, nv, rot_ptrs.data(), nt, d_tri.p, meshes.data());
        }
        // `rotated_mesh` is a COPY of the data mesh that is then moved with set_coord (DiscreteGroupModel.cpp:94-103): its cached
        // triangle areas, hence the source vertex areas metric_resample uses, are those of the un-moved mesh = meshes[0]
        for (int l = 1; l < L && st == MSMGPU_OK; ++l) st = msmgpu_mesh_set_area_source(meshes[l], meshes[0]);
        if (st == MSMGPU_OK) st = msmgpu_adaptive_weights_batch(ctx, L, meshes.data(), nullptr, tpl, tpl_tree, ws.data());
        if (st == MSMGPU_OK) {
            std::vector<const double*> in(L, d_feat_rows.p);
            std::vector<double*> outp(L);
            for (int l = 0; l < L; ++l) outp[l] = d_fields + (((size_t)sub * L + l) * n_tpl) * D;
            st = msmgpu_weights_apply_batch_f64_dev(ctx, L, ws.data(), D, in.data(), outp.data());
        }
        cudaStreamSynchronize(s);
        for (auto* w : ws) msmgpu_weights_destroy(w);
        for (auto* m : meshes) msmgpu_mesh_destroy(m);
        if (st != MSMGPU_OK) return st;
    }
    return MSMGPU_OK;
}

msmgpu_status msmgpu_group_create(msmgpu_ctx* ctx, int simmeasure, int S, int ncp, int L, int D, msmgpu_mesh* tpl, const double* d_fields,
                                  const double* rotations, const double* labels, const double* spacings, double range, msmgpu_group** out) {
    if (!ctx || !out || S <= 0 || ncp <= 0 || L <= 0 || D <= 0 || !tpl || !d_fields || !rotations || !labels || !spacings)
        return fail(MSMGPU_ERR_INVALID, "group_create: bad arguments");
    if (simmeasure != 1 && simmeasure != 2) return fail(MSMGPU_ERR_INVALID, "group_create: simmeasure must be 1 (SSD) or 2 (correlation)");
    *out = nullptr;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    auto g = std::unique_ptr<msmgpu_group>(new msmgpu_group());
    g->ctx = ctx; g->simmeasure = simmeasure; g->S = S; g->ncp = ncp; g->L = L; g->D = D; g->n_tpl = tpl->nv; g->d_fields = d_fields;
    const int n_nodes = S * ncp;
    std::vector<double> thr(n_nodes), thr2(n_nodes);
#pragma omp parallel for
    for (int k = 0; k < n_nodes; ++k) {
        thr[k] = patch_chord_threshold(range * spacings[k]);   // DiscreteGroupModel.cpp:111
        thr2[k] = chord_sq_threshold(thr[k]);
    }
    DevBuf<double> d_rot, d_labels, anchor, thr_super;
    MSM_TRY(upg(d_rot, rotations, 9 * (size_t)n_nodes, s));
    MSM_TRY(upg(d_labels, labels, 3 * (size_t)L, s));
    MSM_TRY(upg(g->thr, thr.data(), (size_t)n_nodes, s));
    MSM_TRY(upg(g->thr2, thr2.data(), (size_t)n_nodes, s));
    MSM_CUDA(g->rcp.alloc(3 * (size_t)n_nodes * L, s));
    MSM_CUDA(anchor.alloc(3 * (size_t)n_nodes, s));
    MSM_CUDA(thr_super.alloc((size_t)n_nodes, s));
    MSM_CUDA(g->tpl_xyz.alloc(3 * (size_t)tpl->nv, s));
    MSM_CUDA(cudaMemcpyAsync(g->tpl_xyz.p, tpl->xyz.p, 3 * (size_t)tpl->nv * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_rotated_cps<<<(n_nodes + 255) / 256, 256, 0, s>>>(n_nodes, L, d_rot.p, d_labels.p, g->rcp.p, anchor.p, g->thr.p, thr_super.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(build_patch_lists(n_nodes, anchor.p, tpl->nv, g->tpl_xyz.p, thr_super.p, g->sup_ptr, g->sup_mem, g->n_sup, g->max_sup, s));
    *out = g.release();
    return MSMGPU_OK;
}

void msmgpu_group_destroy(msmgpu_group* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    delete g;
}

msmgpu_status msmgpu_group_set_mask(msmgpu_group* g, const double* mask) {
    if (!g) return fail(MSMGPU_ERR_INVALID, "group_set_mask: bad arguments");
    MSM_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = g->ctx->stream;
    if (!mask) { g->mask.release(); return MSMGPU_OK; }
    std::vector<double> w((size_t)g->n_tpl);
    for (int p = 0; p < g->n_tpl; ++p) w[p] = std::fabs(mask[p]);   // std::abs(_MASK.get_pvalue(e.first)), DiscreteGroupCostFunction.cpp:77
    MSM_TRY(upg(g->mask, w.data(), w.size(), s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

msmgpu_status msmgpu_group_pair_costs(msmgpu_group* g, int P, const int32_t* pairs, int n, const int32_t* req_pair, const int32_t* req_la,
                                      const int32_t* req_lb, double* out) {
    if (!req_pair || !req_la || !req_lb) return fail(MSMGPU_ERR_INVALID, "group_pair_costs: bad arguments");
    return pair_run(g, P, pairs, n, req_pair, req_la, req_lb, nullptr, 0, out);
}

msmgpu_status msmgpu_group_set_pairs(msmgpu_group* g, int P, const int32_t* pairs) {
    if (!g || P <= 0 || !pairs) return fail(MSMGPU_ERR_INVALID, "group_set_pairs: bad arguments");
    MSM_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = g->ctx->stream;
    MSM_TRY(upg(g->pairs, pairs, 2 * (size_t)P, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    g->P = P;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_group_pair_batch_dev(msmgpu_group* g, int first_pair, int n_pairs, const int32_t* labeling, int label, double* d_out) {
    if (!g || !labeling || !d_out || label < 0 || label >= g->L || first_pair < 0 || n_pairs <= 0 || first_pair + (long long)n_pairs > g->P)
        return fail(MSMGPU_ERR_INVALID, "group_pair_batch_dev: bad arguments (msmgpu_group_set_pairs first)");
    MSM_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = g->ctx->stream;
    // pageable source: staged before the call returns, so one scratch buffer per group is enough for back-to-back batches
    MSM_TRY(upg(g->labeling, labeling, (size_t)g->S * g->ncp, s));
    PairArgs a;
    a.simmeasure = g->simmeasure; a.ncp = g->ncp; a.L = g->L; a.D = g->D; a.n_tpl = g->n_tpl; a.n = 4 * n_pairs;
    a.fields = g->d_fields; a.tpl = g->tpl_xyz.p; a.rcp = g->rcp.p; a.thr2 = g->thr2.p; a.mask = g->mask.p; a.sup_ptr = g->sup_ptr.p; a.sup_mem = g->sup_mem.p;
    a.pairs = g->pairs.p + 2 * (size_t)first_pair; a.req_pair = nullptr; a.req_la = nullptr; a.req_lb = nullptr;
    a.labeling = g->labeling.p; a.label = label; a.out = d_out;
    return launch_pair_costs(a, s);
}

msmgpu_status msmgpu_group_pair_batch(msmgpu_group* g, int P, const int32_t* pairs, const int32_t* labeling, int label, double* out) {
    if (!g || !labeling || label < 0 || label >= g->L) return fail(MSMGPU_ERR_INVALID, "group_pair_batch: bad arguments");
    return pair_run(g, P, pairs, 4 * P, nullptr, nullptr, nullptr, labeling, label, out);
}

} // extern "C"
