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
    const double* fields; const double* tpl; const double* rcp; const double* thr;
    const int* sup_ptr; const int* sup_mem;
    const int* pairs;        // [P][2]
    const int* req_pair; const int* req_la; const int* req_lb;   // list mode, or NULL:
    const int* labeling; int label;                              // Fusion mode: request r = 4 * pair + combo
    double* out;
};

constexpr int kPairWarps = 4;

// one warp per request. Lane d owns channel d (d + 32, ...): the intersection is found 32 candidates at a time
// (one chord test pair per lane, ballot), then every member is visited in ascending template id and each lane
// adds its channel's value -> the reference's sequential sums, coalesced over the channels.
__global__ void __launch_bounds__(kPairWarps * 32) k_group_pair_costs(PairArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kPairWarps + warp;
    if (r >= a.n) return;
    int pair, la, lb;
    if (a.req_pair) {
        pair = a.req_pair[r]; la = a.req_la[r]; lb = a.req_lb[r];
    } else {   // Fusion.h:170-173: (cur,cur), (cur,label), (label,cur), (label,label)
        pair = r >> 2;
        const int combo = r & 3;
        la = (combo & 2) ? a.label : a.labeling[a.pairs[2 * (size_t)pair]];
        lb = (combo & 1) ? a.label : a.labeling[a.pairs[2 * (size_t)pair + 1]];
    }
    const int nA = a.pairs[2 * (size_t)pair], nB = a.pairs[2 * (size_t)pair + 1];
    const V3 cA = load_pt(a.rcp, nA * a.L + la), cB = load_pt(a.rcp, nB * a.L + lb);
    const double tA = a.thr[nA], tB = a.thr[nB];
    const int sA = nA / a.ncp, sB = nB / a.ncp;
    const double* __restrict__ fA = a.fields + ((size_t)sA * a.L + la) * a.n_tpl * a.D;
    const double* __restrict__ fB = a.fields + ((size_t)sB * a.L + lb) * a.n_tpl * a.D;
    const int b = a.sup_ptr[nA], e = a.sup_ptr[nA + 1];

    double cost = 0.0;   // lane 0 accumulates the per-channel similarities in channel order
    int n_common = 0;
    for (int d0 = 0; d0 < a.D; d0 += 32) {
        const int d = d0 + lane;
        const bool has = d < a.D;
        // pass 1: count, sum A, sum B
        double sumA = 0.0, sumB = 0.0;
        int cnt = 0;
        for (int i0 = b; i0 < e; i0 += 32) {
            const int i = i0 + lane;
            int p = -1;
            bool in = false;
            if (i < e) {
                p = __ldg(a.sup_mem + i);
                const V3 t = load_pt(a.tpl, p);
                in = vnorm(vsub(cA, t)) < tA && vnorm(vsub(cB, t)) < tB;
            }
            unsigned m = __ballot_sync(kFull, in);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int pp = __shfl_sync(kFull, p, src);
                ++cnt;
                if (has) { sumA += fA[(size_t)pp * a.D + d]; sumB += fB[(size_t)pp * a.D + d]; }
            }
        }
        n_common = cnt;
        double sim = 0.0;
        if (cnt > 0) {
            if (a.simmeasure == 2) {   // similarities.cpp:129-158 with unit weights
                const double sum = (double)cnt;   // sum of cnt ones, exact
                const double meanA = sumA / sum, meanB = sumB / sum;
                double prod = 0.0, varA = 0.0, varB = 0.0;
                for (int i0 = b; i0 < e; i0 += 32) {
                    const int i = i0 + lane;
                    int p = -1;
                    bool in = false;
                    if (i < e) {
                        p = __ldg(a.sup_mem + i);
                        const V3 t = load_pt(a.tpl, p);
                        in = vnorm(vsub(cA, t)) < tA && vnorm(vsub(cB, t)) < tB;
                    }
                    unsigned m = __ballot_sync(kFull, in);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int pp = __shfl_sync(kFull, p, src);
                        if (has) {
                            const double x = fA[(size_t)pp * a.D + d] - meanA, y = fB[(size_t)pp * a.D + d] - meanB;
                            prod += 1.0 * x * y; varA += 1.0 * x * x; varB += 1.0 * y * y;
                        }
                    }
                }
                prod /= sum; varA /= sum; varB /= sum;
                const double corr = (varA == 0.0 || varB == 0.0) ? 0.0 : prod / (sqrt(varA) * sqrt(varB));
                sim = 1 - (1 + corr) * 0.5;
            } else {   // SSD, similarities.cpp:179-188
                double prod = 0.0;
                for (int i0 = b; i0 < e; i0 += 32) {
                    const int i = i0 + lane;
                    int p = -1;
                    bool in = false;
                    if (i < e) {
                        p = __ldg(a.sup_mem + i);
                        const V3 t = load_pt(a.tpl, p);
                        in = vnorm(vsub(cA, t)) < tA && vnorm(vsub(cB, t)) < tB;
                    }
                    unsigned m = __ballot_sync(kFull, in);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int pp = __shfl_sync(kFull, p, src);
                        if (has) {
                            const double dd = fA[(size_t)pp * a.D + d] - fB[(size_t)pp * a.D + d];
                            prod += 1.0 * dd * dd;
                        }
                    }
                }
                sim = sqrt(prod) / cnt;
            }
        }
        // pair_cost += sim, channel by channel (DiscreteGroupCostFunction.cpp:86-91)
        const int nd = min(32, a.D - d0);
        for (int k = 0; k < nd; ++k) cost += __shfl_sync(kFull, sim, k);
    }
    if (lane == 0) a.out[r] = n_common > 0 ? cost / a.D : nan("");   // empty intersection: the reference reads patch_data_A[0] of an empty vector
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
    a.fields = g->d_fields; a.tpl = g->tpl_xyz.p; a.rcp = g->rcp.p; a.thr = g->thr.p; a.sup_ptr = g->sup_ptr.p; a.sup_mem = g->sup_mem.p;
    a.pairs = d_pairs.p; a.req_pair = req_pair ? d_rp.p : nullptr; a.req_la = d_la.p; a.req_lb = d_lb.p;
    a.labeling = d_labeling.p; a.label = label; a.out = d_out.p;
    k_group_pair_costs<<<(unsigned)((n + kPairWarps - 1) / kPairWarps), kPairWarps * 32, 0, s>>>(a);
    MSM_LAUNCH_CHECK();
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
    std::vector<double> thr(n_nodes);
#pragma omp parallel for
    for (int k = 0; k < n_nodes; ++k) thr[k] = patch_chord_threshold(range * spacings[k]);   // DiscreteGroupModel.cpp:111
    DevBuf<double> d_rot, d_labels, anchor, thr_super;
    MSM_TRY(upg(d_rot, rotations, 9 * (size_t)n_nodes, s));
    MSM_TRY(upg(d_labels, labels, 3 * (size_t)L, s));
    MSM_TRY(upg(g->thr, thr.data(), (size_t)n_nodes, s));
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
    a.fields = g->d_fields; a.tpl = g->tpl_xyz.p; a.rcp = g->rcp.p; a.thr = g->thr.p; a.sup_ptr = g->sup_ptr.p; a.sup_mem = g->sup_mem.p;
    a.pairs = g->pairs.p + 2 * (size_t)first_pair; a.req_pair = nullptr; a.req_la = nullptr; a.req_lb = nullptr;
    a.labeling = g->labeling.p; a.label = label; a.out = d_out;
    k_group_pair_costs<<<(unsigned)((a.n + kPairWarps - 1) / kPairWarps), kPairWarps * 32, 0, s>>>(a);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status msmgpu_group_pair_batch(msmgpu_group* g, int P, const int32_t* pairs, const int32_t* labeling, int label, double* out) {
    if (!g || !labeling || label < 0 || label >= g->L) return fail(MSMGPU_ERR_INVALID, "group_pair_batch: bad arguments");
    return pair_run(g, P, pairs, 4 * P, nullptr, nullptr, nullptr, labeling, label, out);
}

} // extern "C"
