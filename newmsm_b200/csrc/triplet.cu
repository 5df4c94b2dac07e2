// Triplet (higher-order clique) costs: the strain regulariser and the HO likelihoods.
//
// Reference behaviour restated (msm-newmeshreg/src):
//   NonLinearSRegDiscreteCostFunction::computeTripletCost       DiscreteCostFunction.cpp:135-188 (regoption 2/3: spherical strain)
//   HOUnivariate / HOMultivariate get_source_data                cpp:468-485, 541-563  (patch = sources whose nearest CP-grid triangle it is)
//   HOUnivariate / HOMultivariate get_target_data + likelihood   cpp:487-531, 565-618
//   calculate_triangular_strain / triangle_strain / calculate_tri reg_tools.cpp:698-743, 551-646, 267-313
// called by Fusion::optimize with 8 label combinations per triplet and candidate label (Fusion.h:181-196).
//
// The 2x2 / 3x3 NEWMAT algebra of the regulariser is written out (2x2 inverse by adjugate over determinant,
// 3x3 determinant by first-row cofactors, products summed left to right). FSL's NEWMAT is not part of the
// reference tree, so parity for this part is against the CPU restatement of the test suite ("parity unpinned").
//
// Work decomposition: one warp per request (triplet, la, lb, lc). HO kinds: the lanes, in groups of G, move
// the triangle's source points with the barycentric blend of the three displaced control points, locate them
// on the target (query.cuh) and park (ids, weights) in the warp's shared-memory slice; similarity sums are the
// reference's sequential FP64 sums (cost.cuh). Lane 0 evaluates the strain energy.
#include "cost.cuh"
#include "hostpow.cuh"

#include <algorithm>

namespace msm {

// ------------------------------------------------------------------------------------------
// strain energy
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ V3 tri_normal(const V3& v0, const V3& v1, const V3& v2) {   // triangle.cpp:42-47
    return vnormalized(vcross(vsub(v2, v0), vsub(v1, v0)));
}

__host__ __device__ __forceinline__ void tangents(const V3& a, V3& e1, V3& e2) {   // reg_tools.cpp:267-313
    V3 b{1.0, 0.0, 0.0};
    V3 c = vcross(a, b);
    double len = c.x * c.x + c.y * c.y + c.z * c.z;
    if (len == 0.0) {
        b = V3{0.0, 1.0, 0.0};
        c = vcross(a, b);
        len = c.x * c.x + c.y * c.y + c.z * c.z;
    }
    len = sqrt(len);
    if (len == 0.0) len = 1;
    e1 = V3{c.x / len, c.y / len, c.z / len};
    b = vcross(a, c);
    len = sqrt(b.x * b.x + b.y * b.y + b.z * b.z);
    if (len == 0) len = 1;
    e2 = V3{b.x / len, b.y / len, b.z / len};
}

__host__ __device__ __forceinline__ double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// The three pow() calls of the path (reg_tools.cpp:596, DiscreteCostFunction.cpp:187): glibc's pow(x, 2.0) is not always the correctly
// rounded x*x and CUDA's pow is not glibc's, while the costs feed a discrete optimiser. Default: the kernel evaluates them itself with
// the host library's own algorithm and tables (hostpow.cuh: host_pow, enabled after a self-test against std::pow). Otherwise
// (MSMGPU_DEVICE_POW=0, or a host library the self-test does not recognise) the kernel stops at the pow arguments (R = major/minor
// stretch ratio, J = area ratio) and finish_on_host() completes W and the cost with the host libm in the reference's expression order.

__device__ void triangle_strain_dev(const double* AA, const double* BB, double& R_out, double& J_out) {   // reg_tools.cpp:551-594
    const double c0 = AA[3] - AA[0], c1 = AA[4] - AA[1], c4 = AA[6] - AA[0], c5 = AA[7] - AA[1];
    const double c0c = BB[3] - BB[0], c1c = BB[4] - BB[1], c4c = BB[6] - BB[0], c5c = BB[7] - BB[1];
    const double det = c0 * c5 - c4 * c1;
    const double i11 = c5 / det, i12 = -c4 / det, i21 = -c1 / det, i22 = c0 / det;
    const double F11 = c0c * i11 + c4c * i21, F12 = c0c * i12 + c4c * i22;
    const double F21 = c1c * i11 + c5c * i21, F22 = c1c * i12 + c5c * i22;
    const double G11 = F11 * F11 + F21 * F21 + 0.0 * 0.0, G12 = F11 * F12 + F21 * F22 + 0.0 * 0.0;
    const double G21 = F12 * F11 + F22 * F21 + 0.0 * 0.0, G22 = F12 * F12 + F22 * F22 + 0.0 * 0.0;
    const double G33 = 0.0 * 0.0 + 0.0 * 0.0 + 1.0 * 1.0;
    const double I1 = G11 + G22 + G33;
    const double g[9] = {G11, G12, 0.0, G21, G22, 0.0, 0.0, 0.0, G33};
    const double I3 = det3(g);
    const double J = sqrt(I3);
    const double I1st_new = (I1 - 1.0) / J;
    double R;
    if (I1st_new <= 2) R = 1.0;
    else R = 0.5 * (I1st_new + sqrt(I1st_new * I1st_new - 4));
    R_out = R;
    J_out = J;
}

__device__ void triangular_strain_dev(const V3* O, const V3* F, double& R_out, double& J_out) {   // reg_tools.cpp:698-743
    const V3 NO = tri_normal(O[0], O[1], O[2]), NF = tri_normal(F[0], F[1], F[2]);
    V3 e1, e2, f1, f2;
    tangents(NO, e1, e2);
    tangents(NF, f1, f2);
    double TR[9] = {e1.x, e2.x, NO.x, e1.y, e2.y, NO.y, e1.z, e2.z, NO.z};     // form_matrix_from_points, point.cpp:77-95
    double TR2[9] = {f1.x, f2.x, NF.x, f1.y, f2.y, NF.y, f1.z, f2.z, NF.z};
    if (det3(TR) < 0) {
        double t;
        t = TR[0]; TR[0] = TR[1]; TR[1] = t; t = TR[3]; TR[3] = TR[4]; TR[4] = t; t = TR[6]; TR[6] = TR[7]; TR[7] = t;
    }
    if (det3(TR) < 0) {   // sic: the reference tests TRANS a second time (reg_tools.cpp:722)
        double t;
        t = TR2[0]; TR2[0] = TR2[1]; TR2[1] = t; t = TR2[3]; TR2[3] = TR2[4]; TR2[4] = t; t = TR2[6]; TR2[6] = TR2[7]; TR2[7] = t;
    }
    double A[9], B[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            A[3 * i + j] = O[i].x * TR[j] + O[i].y * TR[3 + j] + O[i].z * TR[6 + j];
            B[3 * i + j] = F[i].x * TR2[j] + F[i].y * TR2[3 + j] + F[i].z * TR2[6 + j];
        }
    triangle_strain_dev(A, B, R_out, J_out);
}

// ------------------------------------------------------------------------------------------
// HO patches: group the source vertices by their nearest CP-grid triangle, ascending source id
// ------------------------------------------------------------------------------------------
__global__ void k_count_keys(int n, const int* __restrict__ key, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && key[i] >= 0) atomicAdd(cnt + key[i], 1);
}
__global__ void k_fill_keys(int n, const int* __restrict__ key, const int* __restrict__ ptr, int* __restrict__ cursor, int* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && key[i] >= 0) ids[ptr[key[i]] + atomicAdd(cursor + key[i], 1)] = i;
}
__global__ void k_sort_segments(int nkeys, const int* __restrict__ ptr, int* __restrict__ ids, int* __restrict__ max_len) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int b = ptr[k], e = ptr[k + 1];
    for (int i = b + 1; i < e; ++i) {
        const int v = ids[i];
        int j = i - 1;
        while (j >= b && ids[j] > v) { ids[j + 1] = ids[j]; --j; }
        ids[j + 1] = v;
    }
    atomicMax(max_len, e - b);
}

// ------------------------------------------------------------------------------------------
// triplet cost kernel
// ------------------------------------------------------------------------------------------
struct TripletArgs {
    TreeView tree;
    int kind, simmeasure, ncp, nsrc, D, cfw_rows, n, max_patch;
    double percentile;
    const int* triplets;      // [ntrip][3]
    const int* req_t;         // [n] or NULL (then request r = 8 * triplet + combo, Fusion.h:181-196)
    const int* req_la; const int* req_lb; const int* req_lc;
    const int* labeling;      // [ncp] (combo mode)
    int label;                // candidate label (combo mode)
    const double* labels;     // [L][3]
    const double* rot;        // [ncp][9]
    const double* cp_xyz;     // [ncp][3] current control points (_CPgrid)
    const double* orig_xyz;   // [ncp][3] _ORIG
    const double* src_xyz; const int* prow; const int* pmem;
    const double* src_feat; const double* ref_feat; const double* cfw; const double* absw;
    double lambda, mu, kappa, k_exp, rexp;
    double* aux;              // [n][2]: R, J of the strain energy. J = -1 (a square root is never negative): out[r] is already final.
                              // A NaN pair is a legitimate result (det(F^T F) rounding below zero, reg_tools.cpp:585-587) and propagates.
    double fold_value;        // cost of a folded triangle: FOLDING * lambda (cpp:152) or FOLDING (DiscreteGroupCostFunction.cpp:40)
    double* out;              // [n]
    int* err;
    int dev_pow;              // 1: finish the cost here with host_pow (aux unused), 0: write aux, the host finishes
    int group;                // 1: gMSM form `subcorr * lambda * W^rexp` with FIX_NAN (DiscreteGroupCostFunction.cpp:49-51), 0: `likelihood + lambda * W^rexp`
    int fixnan;
    double subcorr;
    PowTables pw;
    // regoption 4 / 5 (anatomical strain): rmode >= 4
    int rmode, an_max_u, an_max_f;
    size_t slice_bytes;       // per-warp shared-memory slice: [HO part | anatomical part]
    TreeView an_tree;         // anattree over _TARGEThi
    const double* an_src; const int* an_src_tri; const double* an_tgt;
    const int* an_face_ptr; const int* an_face_ids; const int* an_face_local; const int* an_uv_ptr; const int* an_uv_ids;
    const int* an_bary_ptr; const int* an_bary_key; const double* an_bary_w;
    double* aux_anat;         // [n][2 * an_max_f] (R, J) per face when the host finishes the costs
};

// W of one deformed triangle (reg_tools.cpp:596-597)
__host__ __device__ __forceinline__ double strain_energy(double R, double J, double MU, double KAPPA, double k_exp, const PowTables* pw) {
#ifdef __CUDA_ARCH__
    const double Rshared = host_pow(R, k_exp, *pw), Jshared = host_pow(J, k_exp, *pw);
#else
    const double Rshared = pw ? host_pow(R, k_exp, *pw) : std::pow(R, k_exp), Jshared = pw ? host_pow(J, k_exp, *pw) : std::pow(J, k_exp);
#endif
    return 0.5 * (MU * (Rshared + 1.0 / Rshared - 2) + KAPPA * (Jshared + 1.0 / Jshared - 2));
}

// DiscreteCostFunction.cpp:187 / DiscreteGroupCostFunction.cpp:49-51 from the strain energy W, the reference's expression order
__host__ __device__ __forceinline__ double finish_from_energy(double likelihood, double W, double rexp, double lambda, int group, int fixnan,
                                                              double subcorr, const PowTables* pw) {
#ifdef __CUDA_ARCH__
    if (group) {
        if (fixnan && W != W) return 1e7;       // FIX_NAN
        return subcorr * lambda * host_pow(W, rexp, *pw);
    }
    return likelihood + lambda * host_pow(W, rexp, *pw);
#else
    if (group) {
        if (fixnan && W != W) return 1e7;
        return subcorr * lambda * (pw ? host_pow(W, rexp, *pw) : std::pow(W, rexp));
    }
    return likelihood + lambda * (pw ? host_pow(W, rexp, *pw) : std::pow(W, rexp));
#endif
}
__host__ __device__ __forceinline__ double finish_cost(double likelihood, double R, double J, double MU, double KAPPA, double k_exp, double rexp,
                                                       double lambda, int group, int fixnan, double subcorr, const PowTables* pw) {
    return finish_from_energy(likelihood, strain_energy(R, J, MU, KAPPA, k_exp, pw), rexp, lambda, group, fixnan, subcorr, pw);
}

__device__ __forceinline__ void triplet_labels(const TripletArgs& a, int r, int& t, int* lab) {
    if (a.req_t) {
        t = a.req_t[r];
        lab[0] = a.req_la[r]; lab[1] = a.req_lb[r]; lab[2] = a.req_lc[r];
    } else {   // Fusion.h:181-196: request r = 8 * triplet + combination
        t = r >> 3;
        const int combo = r & 7;
        for (int k = 0; k < 3; ++k) lab[k] = (combo >> (2 - k)) & 1 ? a.label : a.labeling[a.triplets[3 * (size_t)t + k]];
    }
}

// strain-only requests (no HO likelihood: the Univariate / Multivariate / Patchwise unary kinds and the gMSM triplet term): one THREAD per
// request -- a warp per request leaves 31 lanes idle here
__global__ void __launch_bounds__(128) k_strain_costs(TripletArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    int t, lab[3];
    triplet_labels(a, r, t, lab);
    V3 def[3], cur[3], org[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int id = a.triplets[3 * (size_t)t + k];
        const double* lb = a.labels + 3 * (size_t)lab[k];
        def[k] = mat_apply(a.rot + 9 * (size_t)id, V3{lb[0], lb[1], lb[2]});
        cur[k] = load_pt(a.cp_xyz, id);
        org[k] = load_pt(a.orig_xyz, id);
    }
    if (vdot(tri_normal(def[0], def[1], def[2]), tri_normal(cur[0], cur[1], cur[2])) < 0.0) {
        a.out[r] = a.fold_value;
        if (!a.dev_pow) { a.aux[2 * (size_t)r] = 0.0; a.aux[2 * (size_t)r + 1] = -1.0; }
        return;
    }
    double R, J;
    triangular_strain_dev(org, def, R, J);
    if (a.dev_pow) {
        a.out[r] = finish_cost(0.0, R, J, a.mu, a.kappa, a.k_exp, a.rexp, a.lambda, a.group, a.fixnan, a.subcorr, &a.pw);
    } else {
        a.out[r] = 0.0;
        a.aux[2 * (size_t)r] = R;
        a.aux[2 * (size_t)r + 1] = J;
    }
}

constexpr int kTripWarps = 4;

template <int G>
__global__ void __launch_bounds__(kTripWarps * 32) k_triplet_costs(TripletArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kTripWarps + warp;
    if (r >= a.n) return;
    // per-warp slice: w[max_patch][3] doubles | sim[max_patch] doubles | idx[max_patch][3] ints | anatomical part (regoption 4/5)
    const size_t slice = (size_t)a.max_patch * (4 * sizeof(double) + 3 * sizeof(int));
    unsigned char* base = smem_raw + (size_t)warp * a.slice_bytes;
    double* s_an = reinterpret_cast<double*>(base + ((slice + 15) & ~(size_t)15));   // [an_max_u][3] deformed vertices | [an_max_f] face energies
    double* s_w = reinterpret_cast<double*>(base);
    double* s_sim = s_w + 3 * (size_t)a.max_patch;
    int* s_idx = reinterpret_cast<int*>(s_sim + a.max_patch);

    int t, lab[3];
    triplet_labels(a, r, t, lab);
    V3 def[3], cur[3], org[3];
    int ids[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ids[k] = a.triplets[3 * (size_t)t + k];
        const double* lb = a.labels + 3 * (size_t)lab[k];
        def[k] = mat_apply(a.rot + 9 * (size_t)ids[k], V3{lb[0], lb[1], lb[2]});        // (*ROTATIONS)[node] * _labels[label], cpp:140-142
        cur[k] = load_pt(a.cp_xyz, ids[k]);
        org[k] = load_pt(a.orig_xyz, ids[k]);
    }
    // only estimate the cost if it does not cause folding (cpp:152)
    if (vdot(tri_normal(def[0], def[1], def[2]), tri_normal(cur[0], cur[1], cur[2])) < 0.0) {
        if (lane == 0) {
            a.out[r] = a.fold_value;
            if (!a.dev_pow) { a.aux[2 * (size_t)r] = 0.0; a.aux[2 * (size_t)r + 1] = -1.0; }   // J = -1: out[r] is final as is
        }
        return;
    }
    double likelihood = 0.0;
    if (a.kind >= MSMGPU_COST_HO_UNIVARIATE) {   // warp-uniform
        const int p0 = a.prow[t], P = a.prow[t + 1] - p0;
        const int gl = lane % G;
        int bad = 0;
        for (int i0 = 0; i0 < P; i0 += 32 / G) {
            const int i = i0 + lane / G;
            const bool active = i < P;
            V3 tmp{0, 0, 0};
            if (active) {   // cpp:503-506: project onto the CP triangle, blend with the displaced CPs, back to the sphere
                const int sv = __ldg(a.pmem + p0 + i);
                const V3 SP = project_to_plane(load_pt(a.src_xyz, sv), cur[0], cur[1], cur[2]);
                double w[3];
                bary_weights_raw(SP, cur[0], cur[1], cur[2], w);
                const V3 x = vscale(def[0], w[0]), y = vscale(def[1], w[1]), z = vscale(def[2], w[2]);   // triangle.cpp:169
                tmp = V3{x.x + y.x + z.x, x.y + y.y + z.y, x.z + y.z + z.z};
                tmp = vnormalized(tmp);
                tmp = vscale(tmp, kRad);
            }
            int st;
            const int tt = nearest_triangle<G>(a.tree, tmp, active, gl, st);
            if (active && gl == 0) {
                if (tt < 0) {
                    bad = 1;
                } else {
                    const double* v = a.tree.rec[tt].v;
                    double w[3];
                    bary_weights_raw(tmp, V3{v[0], v[1], v[2]}, V3{v[3], v[4], v[5]}, V3{v[6], v[7], v[8]}, w);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        s_w[3 * i + j] = w[j];
                        s_idx[3 * i + j] = __ldg(a.tree.tri + 3 * (size_t)tt + j);
                    }
                }
            }
        }
        if (__any_sync(kFull, bad)) {
            if (lane == 0) {
                a.out[r] = nan("");
                if (!a.dev_pow) { a.aux[2 * (size_t)r] = 0.0; a.aux[2 * (size_t)r + 1] = -1.0; }
                *a.err = 1;
            }
            return;
        }
        __syncwarp();
        const int D = a.D;
        const double* __restrict__ rf = a.ref_feat;
        const double* __restrict__ sf = a.src_feat;
        auto tgt = [&](int i, int d) -> double {
            return s_w[3 * i] * __ldg(rf + (size_t)s_idx[3 * i] * D + d) + s_w[3 * i + 1] * __ldg(rf + (size_t)s_idx[3 * i + 1] * D + d) +
                   s_w[3 * i + 2] * __ldg(rf + (size_t)s_idx[3 * i + 2] * D + d);
        };
        auto srcv = [&](int i) -> int { return __ldg(a.pmem + p0 + i); };
        const int cr = a.cfw_rows;
        double cost = 0.0;
        if (a.kind == MSMGPU_COST_HO_UNIVARIATE) {   // cpp:522-531
            for (int i = lane; i < P; i += 32) s_sim[i] = tgt(i, 0);
            __syncwarp();
            if (lane == 0)
                cost = sim_for_min(a.simmeasure, P, [&](int i) { return __ldg(sf + (size_t)srcv(i) * D); }, [&](int i) { return s_sim[i]; },
                                   [&](int i) { return cr >= 1 ? __ldg(a.cfw + (size_t)srcv(i) * cr) : 1.0; }, a.percentile);
        } else {   // cpp:601-618
            for (int i = lane; i < P; i += 32) {
                const int sv = srcv(i);
                s_sim[i] = sim_for_min(a.simmeasure, D, [&](int d) { return __ldg(sf + (size_t)sv * D + d); }, [&](int d) { return tgt(i, d); },
                                       [&](int d) { return cr >= d + 1 ? __ldg(a.cfw + (size_t)sv * cr + d) : 1.0; }, a.percentile);
            }
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < P; ++i) cost += s_sim[i];
                if (P > 0) cost /= P;
            }
        }
        if (lane == 0) likelihood = (__ldg(a.absw + ids[0]) + __ldg(a.absw + ids[1]) + __ldg(a.absw + ids[2])) / 3.0 * cost;
    }
    if (a.rmode >= 4) {   // regoption 4/5 (cpp:169-181): mean strain energy of the triplet's anatomical faces, each deformed by deform_anatomy
        const int u0 = a.an_uv_ptr[t], U = a.an_uv_ptr[t + 1] - u0;
        const int f0 = a.an_face_ptr[t], F = a.an_face_ptr[t + 1] - f0;
        for (int ub = 0; ub < U; ub += 32) {   // cpp:245-301, once per distinct vertex (`moved` / `transformed`)
            const int u = ub + lane;
            const bool active = u < U;
            V3 np{0, 0, 0};
            if (active) {
                const int tindex = __ldg(a.an_uv_ids + u0 + u);
                // `vertex[it.first] * it.second` over the ascending keys; a key that is not a node of this triplet reads a zero Point
                for (int e = __ldg(a.an_bary_ptr + tindex); e < __ldg(a.an_bary_ptr + tindex + 1); ++e) {
                    const int key = __ldg(a.an_bary_key + e);
                    const double w = __ldg(a.an_bary_w + e);
                    V3 v{0, 0, 0};
#pragma unroll
                    for (int k = 0; k < 3; ++k) if (ids[k] == key) v = def[k];
                    const V3 c = vscale(v, w);
                    np = V3{np.x + c.x, np.y + c.y, np.z + c.z};
                }
            }
            int st;
            const int tt = nearest_triangle<1>(a.an_tree, np, active, 0, st);
            if (active) {
                V3 res{0, 0, 0};
                if (tt < 0) {   // the reference carries on with a zero triangle: weights 0/0 (cpp:268-274)
                    res = V3{nan(""), nan(""), nan("")};
                } else {
                    int idx[3];
                    double w[3];
                    const int ne = sorted_weights(a.an_tree, tt, np, idx, w);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (j < ne) {
                            const V3 c = vscale(load_pt(a.an_tgt, idx[j]), w[j]);
                            res = V3{res.x + c.x, res.y + c.y, res.z + c.z};
                        }
                }
                s_an[3 * u] = res.x; s_an[3 * u + 1] = res.y; s_an[3 * u + 2] = res.z;
            }
        }
        __syncwarp();
        double* s_e = s_an + 3 * (size_t)a.an_max_u;
        for (int f = lane; f < F; f += 32) {
            const int face = __ldg(a.an_face_ids + f0 + f);
            V3 O[3], Fv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                O[i] = load_pt(a.an_src, __ldg(a.an_src_tri + 3 * (size_t)face + i));
                const int lu = __ldg(a.an_face_local + 3 * (size_t)(f0 + f) + i);
                Fv[i] = V3{s_an[3 * lu], s_an[3 * lu + 1], s_an[3 * lu + 2]};
            }
            double R, J;
            triangular_strain_dev(O, Fv, R, J);
            if (a.dev_pow) {
                s_e[f] = strain_energy(R, J, a.mu, a.kappa, a.k_exp, &a.pw);
            } else {
                a.aux_anat[2 * ((size_t)r * a.an_max_f + f)] = R;
                a.aux_anat[2 * ((size_t)r * a.an_max_f + f) + 1] = J;
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (a.dev_pow) {
                double W = 0.0;
                for (int f = 0; f < F; ++f) W += s_e[f];
                W = W / (double)F;
                a.out[r] = finish_from_energy(likelihood, W, a.rexp, a.lambda, 0, 0, 1.0, &a.pw);
            } else {
                a.out[r] = likelihood;
                a.aux[2 * (size_t)r] = 0.0; a.aux[2 * (size_t)r + 1] = 0.0;   // J != -1: not final
            }
        }
        return;
    }
    if (lane == 0) {   // regoption 2/3 (cpp:158-166), then cpp:187
        double R, J;
        triangular_strain_dev(org, def, R, J);
        if (a.dev_pow) {
            a.out[r] = finish_cost(likelihood, R, J, a.mu, a.kappa, a.k_exp, a.rexp, a.lambda, a.group, a.fixnan, a.subcorr, &a.pw);
        } else {
            a.out[r] = likelihood;         // + lambda * pow(W(R, J), rexp), added by finish_on_host()
            a.aux[2 * (size_t)r] = R;
            a.aux[2 * (size_t)r + 1] = J;
        }
    }
}

template <int G>
static msmgpu_status launch_triplet_g(const TripletArgs& a, size_t smem, cudaStream_t s) {
    if (smem > 48 * 1024) MSM_CUDA(cudaFuncSetAttribute(k_triplet_costs<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_triplet_costs<G><<<(unsigned)((a.n + kTripWarps - 1) / kTripWarps), kTripWarps * 32, smem, s>>>(a);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

template <typename T>
static msmgpu_status up(DevBuf<T>& b, const T* host, size_t n, cudaStream_t s) {
    MSM_CUDA(b.alloc(n, s));
    if (n) MSM_CUDA(cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return MSMGPU_OK;
}

// the host finish of the path without device pow: reg_tools.cpp:596-597 and DiscreteCostFunction.cpp:187 with the host libm
// (this translation unit is built with -ffp-contract=off)
static void finish_on_host(int n, const double* aux, double* out, const msmgpu_reg_params* prm, int group, int fixnan, double subcorr) {
    const double MU = prm->shear_modulus, KAPPA = prm->bulk_modulus, k_exp = prm->k_exponent, rexp = prm->exponent, lambda = prm->lambda;
#pragma omp parallel for schedule(static) if (n > 4096)
    for (int r = 0; r < n; ++r) {
        const double R = aux[2 * (size_t)r], J = aux[2 * (size_t)r + 1];
        if (J == -1.0) continue;                                // folded (or failed query): already final
        out[r] = finish_cost(out[r], R, J, MU, KAPPA, k_exp, rexp, lambda, group, fixnan, subcorr, nullptr);
    }
}

// regoption 4/5 without device pow: W_f per face, their sequential mean, then cpp:187 -- all with the host libm
static void finish_on_host_anat(int n, const double* aux, const double* aux_anat, int max_f, const std::vector<int>& face_ptr, const int32_t* h_req_t,
                                double* out, const msmgpu_reg_params* prm) {
    const double MU = prm->shear_modulus, KAPPA = prm->bulk_modulus, k_exp = prm->k_exponent, rexp = prm->exponent, lambda = prm->lambda;
#pragma omp parallel for schedule(static) if (n > 1024)
    for (int r = 0; r < n; ++r) {
        if (aux[2 * (size_t)r + 1] == -1.0) continue;           // folded (or failed HO query): already final
        const int t = h_req_t ? h_req_t[r] : r >> 3;
        const int F = face_ptr[t + 1] - face_ptr[t];
        double W = 0.0;
        for (int f = 0; f < F; ++f)
            W += strain_energy(aux_anat[2 * ((size_t)r * max_f + f)], aux_anat[2 * ((size_t)r * max_f + f) + 1], MU, KAPPA, k_exp, nullptr);
        W = W / (double)F;
        out[r] = finish_from_energy(out[r], W, rexp, lambda, 0, 0, 1.0, nullptr);
    }
}

// launches the request kernel for `a` (everything but out / aux / err / pow / shared-memory layout set by the caller) and brings the costs to `out`
// d_user != NULL: the costs stay on the device in the caller's buffer (`out` is then only scratch for the host finish when the device
// pow is disabled)
static msmgpu_status run_requests(msmgpu_ctx* ctx, TripletArgs& a, bool ho, const msmgpu_reg_params* prm, double* out,
                                  const msmgpu_costfn::Anat* anat = nullptr, const int32_t* h_req_t = nullptr, double* d_user = nullptr) {
    cudaStream_t s = ctx->stream;
    const int n = a.n;
    const DevicePow& dp = device_pow(ctx->device);
    a.dev_pow = dp.enabled ? 1 : 0;
    a.pw = dp.t;
    a.rmode = prm->rmode;
    DevBuf<double> d_out, d_aux, d_aux_anat;
    DevBuf<int> d_err;
    if (!d_user) MSM_CUDA(d_out.alloc((size_t)n, s));
    if (!a.dev_pow) MSM_CUDA(d_aux.alloc(2 * (size_t)n, s));
    MSM_CUDA(d_err.alloc(1, s));
    MSM_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));
    a.out = d_user ? d_user : d_out.p; a.aux = d_aux.p; a.err = d_err.p;
    const bool an = a.rmode >= 4;
    if (an) {
        a.an_max_u = anat->max_u; a.an_max_f = anat->max_f;
        a.an_tree = anat->tree->view();
        a.an_src = anat->asource_xyz.p; a.an_src_tri = anat->asource_tri.p; a.an_tgt = anat->atarget_xyz.p;
        a.an_face_ptr = anat->face_ptr.p; a.an_face_ids = anat->face_ids.p; a.an_face_local = anat->face_local.p;
        a.an_uv_ptr = anat->uv_ptr.p; a.an_uv_ids = anat->uv_ids.p;
        a.an_bary_ptr = anat->bary_ptr.p; a.an_bary_key = anat->bary_key.p; a.an_bary_w = anat->bary_w.p;
        if (!a.dev_pow) {
            MSM_CUDA(d_aux_anat.alloc(2 * (size_t)n * anat->max_f, s));
            a.aux_anat = d_aux_anat.p;
        }
    }
    if (!ho && !an) {
        k_strain_costs<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
        MSM_LAUNCH_CHECK();
    } else {
        const size_t ho_part = ((size_t)a.max_patch * (4 * sizeof(double) + 3 * sizeof(int)) + 15) & ~(size_t)15;
        const size_t an_part = an ? (((size_t)3 * anat->max_u + anat->max_f) * sizeof(double) + 15) & ~(size_t)15 : 0;
        a.slice_bytes = ho_part + an_part;
        const size_t smem = a.slice_bytes * kTripWarps;
        if (smem > 200 * 1024) return fail(MSMGPU_ERR_CAPACITY, "costfn_triplet: patch too large for shared memory");
        switch (ho ? query_group_width() : 1) {
            case 2: MSM_TRY(launch_triplet_g<2>(a, smem, s)); break;
            case 4: MSM_TRY(launch_triplet_g<4>(a, smem, s)); break;
            case 8: MSM_TRY(launch_triplet_g<8>(a, smem, s)); break;
            case 16: MSM_TRY(launch_triplet_g<16>(a, smem, s)); break;
            case 32: MSM_TRY(launch_triplet_g<32>(a, smem, s)); break;
            default: MSM_TRY(launch_triplet_g<1>(a, smem, s)); break;
        }
    }
    int h_err = 0;
    std::vector<double> aux, aux_anat;
    std::vector<double> scratch;
    if (d_user && !a.dev_pow) { scratch.resize((size_t)n); out = scratch.data(); }
    if (!d_user || !a.dev_pow) MSM_CUDA(cudaMemcpyAsync(out, a.out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (!a.dev_pow) {
        aux.resize(2 * (size_t)n);
        MSM_CUDA(cudaMemcpyAsync(aux.data(), d_aux.p, aux.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (an) {
            aux_anat.resize(2 * (size_t)n * anat->max_f);
            MSM_CUDA(cudaMemcpyAsync(aux_anat.data(), d_aux_anat.p, aux_anat.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
    }
    MSM_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    if (h_err) return status_to_error(MSMGPU_ERR_NO_TRIANGLE);
    if (!a.dev_pow) {
        if (an) finish_on_host_anat(n, aux.data(), aux_anat.data(), anat->max_f, anat->h_face_ptr, h_req_t, out, prm);
        else finish_on_host(n, aux.data(), out, prm, a.group, a.fixnan, a.subcorr);
        if (d_user) {
            MSM_CUDA(cudaMemcpyAsync(d_user, out, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
            MSM_CUDA(cudaStreamSynchronize(s));
        }
    }
    return MSMGPU_OK;
}

static msmgpu_status triplet_run(msmgpu_costfn* c, int ntrip, const int32_t* triplets, int L, const double* labels, const double* rotations,
                                 const double* orig_cp_xyz, const msmgpu_reg_params* prm, int n, const int32_t* req_t, const int32_t* req_la,
                                 const int32_t* req_lb, const int32_t* req_lc, const int32_t* labeling, int label, double* out) {
    if (!c || ntrip <= 0 || !triplets || L <= 0 || !labels || !rotations || !orig_cp_xyz || !prm || n <= 0 || !out)
        return fail(MSMGPU_ERR_INVALID, "costfn_triplet: bad arguments");
    if (c->ncp <= 0) return fail(MSMGPU_ERR_INVALID, "costfn_triplet: set the control-point grid first");
    if (prm->rmode < 2 || prm->rmode > 5)
        return fail(MSMGPU_ERR_INVALID, "DiscreteModel computeTripletCost regoption does not exist");   // cpp:183
    if (prm->rmode >= 4 && (!c->anat || c->anat->ntrip != ntrip))
        return fail(MSMGPU_ERR_INVALID, "costfn_triplet: regoption 4/5 needs msmgpu_costfn_set_anatomical for this control grid");
    const bool ho = c->kind >= MSMGPU_COST_HO_UNIVARIATE;
    if (ho && c->n_patch_rows != ntrip) return fail(MSMGPU_ERR_INVALID, "costfn_triplet: HO patches were built for a different CP-grid triangle count");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    DevBuf<int> d_trip, d_rt, d_la, d_lb, d_lc, d_labeling;
    DevBuf<double> d_labels, d_rot, d_orig;
    MSM_TRY(up(d_trip, triplets, 3 * (size_t)ntrip, s));
    MSM_TRY(up(d_labels, labels, 3 * (size_t)L, s));
    MSM_TRY(up(d_rot, rotations, 9 * (size_t)c->ncp, s));
    MSM_TRY(up(d_orig, orig_cp_xyz, 3 * (size_t)c->ncp, s));
    if (req_t) {
        MSM_TRY(up(d_rt, req_t, (size_t)n, s));
        MSM_TRY(up(d_la, req_la, (size_t)n, s));
        MSM_TRY(up(d_lb, req_lb, (size_t)n, s));
        MSM_TRY(up(d_lc, req_lc, (size_t)n, s));
    } else {
        MSM_TRY(up(d_labeling, labeling, (size_t)c->ncp, s));
    }
    TripletArgs a{};
    a.tree = c->tree->view();
    a.kind = c->kind; a.simmeasure = c->simmeasure; a.percentile = c->percentile; a.ncp = c->ncp; a.nsrc = c->nsrc; a.D = c->D; a.cfw_rows = c->cfw_rows; a.n = n;
    a.max_patch = ho ? std::max(c->max_patch, 1) : 1;
    a.triplets = d_trip.p; a.req_t = req_t ? d_rt.p : nullptr; a.req_la = d_la.p; a.req_lb = d_lb.p; a.req_lc = d_lc.p;
    a.labeling = d_labeling.p; a.label = label; a.labels = d_labels.p; a.rot = d_rot.p; a.cp_xyz = c->cp_xyz.p; a.orig_xyz = d_orig.p;
    a.src_xyz = c->src_xyz.p; a.prow = c->prow.p; a.pmem = c->pmem.p; a.src_feat = c->src_feat.p; a.ref_feat = c->ref_feat.p;
    a.cfw = c->cfw.p; a.absw = c->absw.p;
    a.lambda = prm->lambda; a.mu = prm->shear_modulus; a.kappa = prm->bulk_modulus; a.k_exp = prm->k_exponent; a.rexp = prm->exponent;
    a.fold_value = 1e7 * prm->lambda;
    a.group = 0; a.fixnan = 0; a.subcorr = 1.0;
    return run_requests(c->ctx, a, ho, prm, out, c->anat.get(), req_t);
}

} // namespace msm

// gMSM triplet term (DiscreteGroupCostFunction.cpp:26-52): the same strain energy on the per-subject control grids, no likelihood,
// scaled by subcorr = 0.1 * S; a folded triangle costs FOLDING (not FOLDING * lambda), a NaN energy costs FIX_NAN when --fixnan.
// Node ids in `triplets` are global (subject * ncp + vertex), so the per-subject grids are simply concatenated.
// The per-iteration arrays (grids, rotations, labels, triplets) are uploaded once into a plan; a label phase only sends the labeling.
struct msmgpu_triplet_plan {
    msmgpu_ctx* ctx = nullptr;
    int n_nodes = 0, L = 0, ntrip = 0;
    msm::DevBuf<int> trip, labeling;
    msm::DevBuf<double> labels, rot, orig, cp;
};

namespace msm {

static msmgpu_status plan_create(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                 const double* labels, int ntrip, const int32_t* triplets, std::unique_ptr<msmgpu_triplet_plan>& out) {
    if (!ctx || n_nodes <= 0 || !cp_xyz || !orig_xyz || !rotations || L <= 0 || !labels || ntrip <= 0 || !triplets)
        return fail(MSMGPU_ERR_INVALID, "group_triplet: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    out.reset(new msmgpu_triplet_plan());
    out->ctx = ctx; out->n_nodes = n_nodes; out->L = L; out->ntrip = ntrip;
    MSM_TRY(up(out->trip, triplets, 3 * (size_t)ntrip, s));
    MSM_TRY(up(out->labels, labels, 3 * (size_t)L, s));
    MSM_TRY(up(out->rot, rotations, 9 * (size_t)n_nodes, s));
    MSM_TRY(up(out->orig, orig_xyz, 3 * (size_t)n_nodes, s));
    MSM_TRY(up(out->cp, cp_xyz, 3 * (size_t)n_nodes, s));
    MSM_CUDA(cudaStreamSynchronize(s));   // the host arrays may go away after the call
    return MSMGPU_OK;
}

static msmgpu_status plan_run(msmgpu_triplet_plan* p, const msmgpu_reg_params* prm, double subcorr, int fixnan, int first_triplet, int n_triplets,
                              int n, const int32_t* req_t, const int32_t* req_la, const int32_t* req_lb, const int32_t* req_lc,
                              const int32_t* labeling, int label, double* out, double* d_user = nullptr) {
    if (!p || !prm || n <= 0 || (!out && !d_user)) return fail(MSMGPU_ERR_INVALID, "group_triplet: bad arguments");
    MSM_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    DevBuf<int> d_rt, d_la, d_lb, d_lc;
    if (req_t) {
        MSM_TRY(up(d_rt, req_t, (size_t)n, s));
        MSM_TRY(up(d_la, req_la, (size_t)n, s));
        MSM_TRY(up(d_lb, req_lb, (size_t)n, s));
        MSM_TRY(up(d_lc, req_lc, (size_t)n, s));
    } else {
        MSM_TRY(up(p->labeling, labeling, (size_t)p->n_nodes, s));
    }
    TripletArgs a{};
    a.kind = MSMGPU_COST_UNIVARIATE; a.simmeasure = 2; a.ncp = p->n_nodes; a.n = n; a.max_patch = 1;
    a.triplets = p->trip.p + 3 * (size_t)(req_t ? 0 : first_triplet);
    a.req_t = req_t ? d_rt.p : nullptr; a.req_la = d_la.p; a.req_lb = d_lb.p; a.req_lc = d_lc.p;
    a.labeling = p->labeling.p; a.label = label; a.labels = p->labels.p; a.rot = p->rot.p; a.cp_xyz = p->cp.p; a.orig_xyz = p->orig.p;
    a.lambda = prm->lambda; a.mu = prm->shear_modulus; a.kappa = prm->bulk_modulus; a.k_exp = prm->k_exponent; a.rexp = prm->exponent;
    a.fold_value = 1e7;
    a.group = 1; a.fixnan = fixnan; a.subcorr = subcorr;
    (void)n_triplets;
    return run_requests(p->ctx, a, false, prm, out, nullptr, nullptr, d_user);
}

} // namespace msm

using namespace msm;

extern "C" {

msmgpu_status msmgpu_group_triplet_costs(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                         const double* labels, int ntrip, const int32_t* triplets, const msmgpu_reg_params* prm, double subcorr, int fixnan,
                                         int n, const int32_t* req_triplet, const int32_t* req_la, const int32_t* req_lb, const int32_t* req_lc, double* out) {
    if (!req_triplet || !req_la || !req_lb || !req_lc) return fail(MSMGPU_ERR_INVALID, "group_triplet_costs: bad arguments");
    std::unique_ptr<msmgpu_triplet_plan> p;
    MSM_TRY(plan_create(ctx, n_nodes, cp_xyz, orig_xyz, rotations, L, labels, ntrip, triplets, p));
    return plan_run(p.get(), prm, subcorr, fixnan, 0, ntrip, n, req_triplet, req_la, req_lb, req_lc, nullptr, 0, out);
}

msmgpu_status msmgpu_group_triplet_batch(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                         const double* labels, int ntrip, const int32_t* triplets, const msmgpu_reg_params* prm, double subcorr, int fixnan,
                                         const int32_t* labeling, int label, double* out) {
    if (!labeling || label < 0 || label >= L) return fail(MSMGPU_ERR_INVALID, "group_triplet_batch: bad arguments");
    std::unique_ptr<msmgpu_triplet_plan> p;
    MSM_TRY(plan_create(ctx, n_nodes, cp_xyz, orig_xyz, rotations, L, labels, ntrip, triplets, p));
    return plan_run(p.get(), prm, subcorr, fixnan, 0, ntrip, 8 * ntrip, nullptr, nullptr, nullptr, nullptr, labeling, label, out);
}

msmgpu_status msmgpu_triplet_plan_create(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                         const double* labels, int ntrip, const int32_t* triplets, msmgpu_triplet_plan** out) {
    if (!out) return fail(MSMGPU_ERR_INVALID, "triplet_plan_create: bad arguments");
    *out = nullptr;
    std::unique_ptr<msmgpu_triplet_plan> p;
    MSM_TRY(plan_create(ctx, n_nodes, cp_xyz, orig_xyz, rotations, L, labels, ntrip, triplets, p));
    *out = p.release();
    return MSMGPU_OK;
}

void msmgpu_triplet_plan_destroy(msmgpu_triplet_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    delete p;
}

msmgpu_status msmgpu_triplet_plan_batch(msmgpu_triplet_plan* p, const msmgpu_reg_params* prm, double subcorr, int fixnan, int first_triplet,
                                        int n_triplets, const int32_t* labeling, int label, double* out) {
    if (!p || !labeling || label < 0 || label >= p->L || first_triplet < 0 || n_triplets <= 0 || first_triplet + (long long)n_triplets > p->ntrip)
        return fail(MSMGPU_ERR_INVALID, "triplet_plan_batch: bad arguments");
    return plan_run(p, prm, subcorr, fixnan, first_triplet, n_triplets, 8 * n_triplets, nullptr, nullptr, nullptr, nullptr, labeling, label, out);
}

msmgpu_status msmgpu_triplet_plan_batch_dev(msmgpu_triplet_plan* p, const msmgpu_reg_params* prm, double subcorr, int fixnan, int first_triplet,
                                            int n_triplets, const int32_t* labeling, int label, double* d_out) {
    if (!p || !labeling || !d_out || label < 0 || label >= p->L || first_triplet < 0 || n_triplets <= 0 || first_triplet + (long long)n_triplets > p->ntrip)
        return fail(MSMGPU_ERR_INVALID, "triplet_plan_batch_dev: bad arguments");
    return plan_run(p, prm, subcorr, fixnan, first_triplet, n_triplets, 8 * n_triplets, nullptr, nullptr, nullptr, nullptr, labeling, label, nullptr, d_out);
}

msmgpu_costfn::Anat::~Anat() {
    if (tree) msmgpu_octree_destroy(tree);
    if (thi) msmgpu_mesh_destroy(thi);
}

msmgpu_status msmgpu_costfn_set_anatomical(msmgpu_costfn* c, int ntrip, const msmgpu_anatomical* A) {
    if (!c || ntrip <= 0 || !A || A->n_av <= 0 || !A->asource_xyz || A->n_at <= 0 || !A->asource_tri || A->n_hv <= 0 || !A->thi_xyz || A->n_ht <= 0 ||
        !A->thi_tri || !A->atarget_xyz || !A->face_ptr || !A->face_ids || !A->bary_ptr || !A->bary_key || !A->bary_w)
        return fail(MSMGPU_ERR_INVALID, "costfn_set_anatomical: bad arguments");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    const int n_faces = A->face_ptr[ntrip];
    if (A->face_ptr[0] != 0 || n_faces <= 0) return fail(MSMGPU_ERR_INVALID, "costfn_set_anatomical: NEARESTFACES is empty");
    // per control triangle: the distinct _aSOURCE vertices of its faces in first-seen order (what `moved` / `transformed` hold after the
    // loop of cpp:174-179) and, per face corner, the position of its vertex in that list
    std::vector<int> uv_ptr(ntrip + 1, 0), uv_ids, face_local(3 * (size_t)n_faces);
    int max_u = 0, max_f = 0;
    for (int t = 0; t < ntrip; ++t) {
        const int f0 = A->face_ptr[t], f1 = A->face_ptr[t + 1];
        if (f1 <= f0) return fail(MSMGPU_ERR_INVALID, "costfn_set_anatomical: a control triangle has no anatomical face (the reference divides by zero)");
        const size_t u0 = uv_ids.size();
        for (int f = f0; f < f1; ++f) {
            const int face = A->face_ids[f];
            if (face < 0 || face >= A->n_at) return fail(MSMGPU_ERR_INVALID, "costfn_set_anatomical: face id out of range");
            for (int i = 0; i < 3; ++i) {
                const int v = A->asource_tri[3 * (size_t)face + i];
                if (v < 0 || v >= A->n_av) return fail(MSMGPU_ERR_INVALID, "costfn_set_anatomical: vertex id out of range");
                size_t k = u0;
                while (k < uv_ids.size() && uv_ids[k] != v) ++k;
                if (k == uv_ids.size()) uv_ids.push_back(v);
                face_local[3 * (size_t)f + i] = (int)(k - u0);
            }
        }
        uv_ptr[t + 1] = (int)uv_ids.size();
        max_u = std::max(max_u, (int)(uv_ids.size() - u0));
        max_f = std::max(max_f, f1 - f0);
    }
    for (int e = 0; e < A->bary_ptr[A->n_av]; ++e)
        if (A->bary_key[e] < 0 || A->bary_key[e] >= c->ncp) return fail(MSMGPU_ERR_INVALID, "costfn_set_anatomical: weight key is not a control point (set the control grid first)");
    std::unique_ptr<msmgpu_costfn::Anat> an(new msmgpu_costfn::Anat());
    an->ntrip = ntrip; an->n_av = A->n_av; an->max_u = max_u; an->max_f = max_f;
    an->h_face_ptr.assign(A->face_ptr, A->face_ptr + ntrip + 1);
    MSM_TRY(msmgpu_mesh_create(c->ctx, A->n_hv, A->thi_xyz, A->n_ht, A->thi_tri, &an->thi));
    MSM_TRY(msmgpu_octree_build(an->thi, &an->tree));
    MSM_TRY(up(an->asource_xyz, A->asource_xyz, 3 * (size_t)A->n_av, s));
    MSM_TRY(up(an->asource_tri, A->asource_tri, 3 * (size_t)A->n_at, s));
    MSM_TRY(up(an->atarget_xyz, A->atarget_xyz, 3 * (size_t)A->n_hv, s));
    MSM_TRY(up(an->face_ptr, A->face_ptr, (size_t)ntrip + 1, s));
    MSM_TRY(up(an->face_ids, A->face_ids, (size_t)n_faces, s));
    MSM_TRY(up(an->face_local, face_local.data(), face_local.size(), s));
    MSM_TRY(up(an->uv_ptr, uv_ptr.data(), uv_ptr.size(), s));
    MSM_TRY(up(an->uv_ids, uv_ids.data(), uv_ids.size(), s));
    MSM_TRY(up(an->bary_ptr, A->bary_ptr, (size_t)A->n_av + 1, s));
    MSM_TRY(up(an->bary_key, A->bary_key, (size_t)A->bary_ptr[A->n_av], s));
    MSM_TRY(up(an->bary_w, A->bary_w, (size_t)A->bary_ptr[A->n_av], s));
    MSM_CUDA(cudaStreamSynchronize(s));
    c->anat = std::move(an);
    return MSMGPU_OK;
}

msmgpu_status msmgpu_costfn_set_cpgrid_ho(msmgpu_costfn* c, int ncp, const double* cp_xyz, int ntri, const int32_t* cp_tri,
                                          int cfw_rows, const double* cfw, const double* absw) {
    if (!c || ncp <= 0 || !cp_xyz || ntri <= 0 || !cp_tri || !absw || cfw_rows < 0 || (cfw_rows > 0 && !cfw))
        return fail(MSMGPU_ERR_INVALID, "costfn_set_cpgrid_ho: bad arguments");
    if (c->kind < MSMGPU_COST_HO_UNIVARIATE) return fail(MSMGPU_ERR_INVALID, "costfn_set_cpgrid_ho: not an HO cost function");
    MSM_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    c->ncp = ncp; c->cfw_rows = cfw_rows;
    c->h_cp.assign(cp_xyz, cp_xyz + 3 * (size_t)ncp);
    MSM_TRY(up(c->cp_xyz, cp_xyz, 3 * (size_t)ncp, s));
    MSM_TRY(up(c->absw, absw, (size_t)ncp, s));
    if (cfw_rows > 0) {
        DevBuf<double> cm;
        MSM_TRY(up(cm, cfw, (size_t)cfw_rows * c->nsrc, s));
        MSM_CUDA(c->cfw.alloc((size_t)cfw_rows * c->nsrc, s));
        MSM_TRY(launch_transpose_f64(cfw_rows, c->nsrc, cm.p, c->cfw.p, s));
        MSM_CUDA(cudaStreamSynchronize(s));
    }
    // newresampler::Octree cp_tree(_CPgrid); closest triangle of every source vertex (cpp:472-476)
    msmgpu_mesh* cpm = nullptr;
    MSM_TRY(msmgpu_mesh_create(c->ctx, ncp, cp_xyz, ntri, cp_tri, &cpm));
    std::unique_ptr<msmgpu_mesh, void (*)(msmgpu_mesh*)> mg(cpm, msmgpu_mesh_destroy);
    msmgpu_octree* cpt = nullptr;
    MSM_TRY(msmgpu_octree_build(cpm, &cpt));
    std::unique_ptr<msmgpu_octree, void (*)(msmgpu_octree*)> tg(cpt, msmgpu_octree_destroy);
    DevBuf<int> key, st, cnt, cursor, d_max;
    MSM_CUDA(key.alloc(c->nsrc, s));
    MSM_CUDA(st.alloc(c->nsrc, s));
    MSM_CUDA(cnt.alloc(ntri, s));
    MSM_CUDA(cursor.alloc(ntri, s));
    MSM_CUDA(d_max.alloc(1, s));
    MSM_TRY(launch_nearest(cpt->view(), c->nsrc, c->src_xyz.p, key.p, nullptr, st.p, s));
    int code = 0;
    MSM_TRY(first_error(st.p, c->nsrc, s, &code));
    if (code) return status_to_error(code);
    MSM_CUDA(cudaMemsetAsync(cnt.p, 0, ntri * sizeof(int), s));
    MSM_CUDA(cudaMemsetAsync(cursor.p, 0, ntri * sizeof(int), s));
    MSM_CUDA(cudaMemsetAsync(d_max.p, 0, sizeof(int), s));
    MSM_CUDA(c->prow.alloc((size_t)ntri + 1, s));
    MSM_CUDA(c->pmem.alloc((size_t)c->nsrc, s));
    const unsigned gs = (unsigned)((c->nsrc + 255) / 256), gk = (unsigned)((ntri + 255) / 256);
    k_count_keys<<<gs, 256, 0, s>>>(c->nsrc, key.p, cnt.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(cnt.p, c->prow.p, ntri, c->prow.p + ntri, s));
    k_fill_keys<<<gs, 256, 0, s>>>(c->nsrc, key.p, c->prow.p, cursor.p, c->pmem.p);
    MSM_LAUNCH_CHECK();
    k_sort_segments<<<gk, 256, 0, s>>>(ntri, c->prow.p, c->pmem.p, d_max.p);
    MSM_LAUNCH_CHECK();
    MSM_CUDA(cudaMemcpyAsync(&c->max_patch, d_max.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    c->n_patch = c->nsrc;
    c->n_patch_rows = ntri;
    c->n_cp_tri = ntri;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_costfn_triplet_costs(msmgpu_costfn* c, int ntrip, const int32_t* triplets, int L, const double* labels, const double* rotations,
                                          const double* orig_cp_xyz, const msmgpu_reg_params* prm, int n, const int32_t* req_triplet,
                                          const int32_t* req_la, const int32_t* req_lb, const int32_t* req_lc, double* out) {
    if (!req_triplet || !req_la || !req_lb || !req_lc) return fail(MSMGPU_ERR_INVALID, "costfn_triplet_costs: bad arguments");
    return triplet_run(c, ntrip, triplets, L, labels, rotations, orig_cp_xyz, prm, n, req_triplet, req_la, req_lb, req_lc, nullptr, 0, out);
}

msmgpu_status msmgpu_costfn_triplet_batch(msmgpu_costfn* c, int ntrip, const int32_t* triplets, int L, const double* labels, const double* rotations,
                                          const double* orig_cp_xyz, const msmgpu_reg_params* prm, const int32_t* labeling, int label, double* out) {
    if (!labeling || label < 0 || label >= L) return fail(MSMGPU_ERR_INVALID, "costfn_triplet_batch: bad arguments");
    return triplet_run(c, ntrip, triplets, L, labels, rotations, orig_cp_xyz, prm, 8 * ntrip, nullptr, nullptr, nullptr, nullptr, labeling, label, out);
}

} // extern "C"
