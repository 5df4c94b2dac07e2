// Shared by cost.cu (unary tables) and triplet.cu (HO patches, triplet costs): the device state of one
// DiscreteCostFunction and the sequential FP64 similarities.
#pragma once
#include <memory>
#include "query.cuh"

struct msmgpu_costfn {
    msmgpu_ctx* ctx = nullptr;
    msmgpu_octree* tree = nullptr;
    int kind = 0, simmeasure = 2;
    double percentile = 0.75;       // sparsesimkernel::percentile (similarities.h:68), DICE only
    int nsrc = 0, D = 0, nvt = 0;
    msm::DevBuf<double> src_xyz;    // [nsrc][3]
    msm::DevBuf<double> src_feat;   // [nsrc][D] rows
    msm::DevBuf<double> ref_feat;   // [nvt][D] rows
    int ncp = 0, cfw_rows = 0;
    double range = 0;
    std::vector<double> h_cp;       // host copy: the rotation matrices are built on the host (api.cu)
    msm::DevBuf<double> cp_xyz, cfw /* [nsrc][cfw_rows] rows */, absw, chord_thr;
    msm::DevBuf<int> prow, pmem;    // patches: CSR over control points (or CP-grid triangles for the HO kinds), ascending source id
    int n_patch = 0, max_patch = 0, n_patch_rows = 0;
    // HO (triclique) state: the CP-grid triangles the patches hang on (DiscreteCostFunction.cpp:468-485)
    int n_cp_tri = 0;
    // anatomical strain (regoption 4/5, DiscreteCostFunction.cpp:169-181, 245-301): msmgpu_costfn_set_anatomical
    struct Anat {
        int ntrip = 0, n_av = 0, max_u = 0, max_f = 0;
        msmgpu_mesh* thi = nullptr;          // _TARGEThi
        msmgpu_octree* tree = nullptr;       // anattree
        msm::DevBuf<double> asource_xyz, atarget_xyz, bary_w;
        msm::DevBuf<int> asource_tri, face_ptr, face_ids, face_local /* [faces][3]: corner -> index in the triplet's vertex list */,
            uv_ptr, uv_ids /* per triplet: the distinct _aSOURCE vertices of its faces, first-seen order */, bary_ptr, bary_key;
        std::vector<int> h_face_ptr;         // host copy (the host finish without device pow needs the face counts)
        ~Anat();
    };
    std::unique_ptr<Anat> anat;
};

namespace msm {

double patch_chord_threshold(double limit);   // cost.cu: host bisection, member <=> chord < threshold
double chord_sq_threshold(double thr);         // cost.cu: chord < thr <=> chord^2 < chord_sq_threshold(thr), exactly
bool host_rotation_matrix(const double* ci, const double* index, double* R);   // api.cu
msmgpu_status build_patch_lists(int n_cp, const double* d_cp, int n_src, const double* d_src, const double* d_thr, DevBuf<int>& prow,
                                DevBuf<int>& pmem, int& total, int& max_len, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// similarities (sequential FP64, similarities.cpp:129-188), element access through functors
// ------------------------------------------------------------------------------------------
template <class FA, class FB, class FW>
__device__ __forceinline__ double sim_corr(int n, FA A, FB B, FW W) {
    double prod = 0.0, varA = 0.0, varB = 0.0, meanA = 0.0, meanB = 0.0, sum = 0.0;
    for (int i = 0; i < n; ++i) sum += W(i);
    for (int i = 0; i < n; ++i) {
        const double w = W(i);
        meanA += w * A(i);
        meanB += w * B(i);
    }
    if (sum > 0.0) { meanA /= sum; meanB /= sum; }
    for (int i = 0; i < n; ++i) {
        const double w = W(i), a = A(i) - meanA, b = B(i) - meanB;
        prod += w * a * b;
        varA += w * a * a;
        varB += w * b * b;
    }
    if (sum > 0.0) { prod /= sum; varA /= sum; varB /= sum; }
    if (varA == 0.0 || varB == 0.0) return 0.0;
    return prod / (sqrt(varA) * sqrt(varB));
}
template <class FA, class FB, class FW>
__device__ __forceinline__ double sim_ssd(int n, FA A, FB B, FW W) {
    double prod = 0.0;
    for (int i = 0; i < n; ++i) {
        const double d = A(i) - B(i);
        prod += W(i) * d * d;
    }
    return sqrt(prod) / n;
}
// DICE / genDICE (similarities.cpp:201-253): the reference sorts copies of A and B and thresholds both at the order statistic
// idx = floor(percentile * n). Only that one order statistic is needed: v = X(i) is it iff #(X < v) <= idx < #(X <= v). O(n^2)
// compares in one lane (n = patch size or channel count; an "experimental" measure in the reference, mesh_registration.cpp:471).
template <class F>
__device__ __forceinline__ double order_statistic(int n, int idx, F X) {
    for (int i = 0; i < n; ++i) {
        const double v = X(i);
        int lt = 0, le = 0;
        for (int j = 0; j < n; ++j) {
            const double x = X(j);
            lt += x < v;
            le += x <= v;
        }
        if (lt <= idx && idx < le) return v;
    }
    return nan("");   // idx >= n (percentile = 1): the reference reads past the end of its sorted copy
}
template <class FA, class FB>
__device__ __forceinline__ double sim_dice(int n, FA A, FB B, double percentile, bool general) {
    const int idx = (int)floor(percentile * n);
    const double ta = order_statistic(n, idx, A), tb = order_statistic(n, idx, B);
    int size_A = n, size_B = n, common = 0;
    for (int i = 0; i < n; ++i) {
        const bool la = A(i) < ta, lb = B(i) < tb;
        size_A -= la;
        size_B -= lb;
        common += !(la || lb);
    }
    if (!general) return 1.0 - ((2.0 * common) / (size_A + size_B));
    const double b2 = (double)size_B * (double)size_B;   // pow(size_B, 2): exact for these integers
    return 1.0 - (2.0 * (((common / b2)) / ((size_A + size_B) / b2)));
}
template <class FA, class FB, class FW>
__device__ __forceinline__ double sim_for_min(int simmeasure, int n, FA A, FB B, FW W, double percentile = 0.75) {
    if (simmeasure == 1) return sim_ssd(n, A, B, W);
    if (simmeasure == 2) return 1 - (1 + sim_corr(n, A, B, W)) * 0.5;
    if (simmeasure == 4) return sim_dice(n, A, B, percentile, false);
    if (simmeasure == 5) return sim_dice(n, A, B, percentile, true);
    return nan("");
}

} // namespace msm
