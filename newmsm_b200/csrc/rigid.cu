// AFFINE / RIGID level: Rigid_cost_function (msm-newmeshreg/src/rigid_costfunction.cpp:32-141), Neighbourhood::update
// (reg_tools.cpp:31-58), the sparse similarity columns (similarities.cpp:27-128) and calculate_tangs (reg_tools.cpp:205-266).
//
// What the reference does per cost evaluation (rigid_cost_mesh, cpp:130-141), for every source vertex i with a non-empty initial
// neighbourhood: rotate i by the Euler angles, build the tangent basis of the ROTATED source mesh at i (normal = normalised sum of the
// incident triangle normals), find the closest TARGET triangle (octree), collect the union of the one-rings of its three corners in
// first-seen order (get_all_neighbours, cpp:143-165; its `update` test is always true, so the neighbourhood of i becomes that list and
// the similarity column of i is recomputed for it), and take the Gaussian-weighted mean of the similarities in the tangent plane
// (WLS_simgradient, cpp:63-90). The cost is the sequential sum of those means.
//
// Split used here (the library's policy for libm): everything built from + - * / sqrt runs on the device in the reference's
// operation order (--fmad=false) and is bit-identical; `exp` (one per list entry) and the sequential sums that consume it run on the host
// libm inside msmgpu_rigid_cost, as do cos / sin of the rotation matrix and acos-free set-up values. One thread per source vertex: a
// cost evaluation is ~41 k independent chains of ~2 k FP64 instructions, latency-bound and two orders of magnitude below the host's
// time for the same work.
#include "common.cuh"
#include "query.cuh"

#include <algorithm>
#include <cmath>

namespace msm {

constexpr int kRigidCap = 64;      // entries per neighbour list (three one-rings: 12 - 14 on a regular mesh)

struct RigidView {
    TreeView tree;                 // TARGET octree
    const double* tgt_xyz;         // [nv_t][3]
    const int* tgt_tri;            // [nt_t][3]
    const int* tgt_inc_ptr;        // [nv_t + 1]  incident triangles of a target vertex, ascending id (mesh.h tIDbegin order)
    const int* tgt_inc;
    const double* src_xyz;         // [nv_s][3] current SOURCE
    const int* src_tri;            // [nt_s][3]
    const int* src_inc_ptr;        // [nv_s + 1]
    const int* src_inc;
    const unsigned char* has_nbh;  // [nv_s] Neighbourhood::update found at least one target vertex (nbh->nrows(i) > 0)
    const double* A;               // input data [D][nv_s]
    const double* B;               // reference data [D][nv_t]
    const double* meanA;           // [nv_s]
    const double* meanB;           // [nv_t]
    int nv_s, nv_t, D, simmeasure;
    double two_sigma2;             // 2 * min_sigma * min_sigma
};

struct Rot9 { double m[9]; };

// rotation.t() * vector (point.cpp:167 with the stand-in's plain product): row r = sum over k of R(k, r) * v(k), from zero, k ascending
__device__ __forceinline__ V3 euler_apply(const Rot9& R, const V3& v) {
    const double vv[3] = {v.x, v.y, v.z};
    double o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) sum += R.m[3 * k + r] * vv[k];
        o[r] = sum;
    }
    return V3{o[0], o[1], o[2]};
}

__device__ __forceinline__ V3 rotated_src(const RigidView& S, const Rot9& R, int v) { return euler_apply(R, load_pt(S.src_xyz, v)); }

// similarities.cpp:52-85 (corr) and 87-104 (SSD, negated by calculate_sim_column_nbh) on full matrices
__device__ double rigid_sim(const RigidView& S, int i, int j) {
    if (S.simmeasure == 1) {
        double prod = 0.0;
        for (int r = 0; r < S.D; ++r) {
            const double a = S.A[(size_t)r * S.nv_s + i], b = S.B[(size_t)r * S.nv_t + j];
            prod += (a - b) * (a - b);
        }
        return -(sqrt(prod) / S.D);
    }
    double prod = 0.0, varA = 0.0, varB = 0.0;
    const double mA = S.meanA[i], mB = S.meanB[j];
    for (int r = 0; r < S.D; ++r) {
        const double a = S.A[(size_t)r * S.nv_s + i], b = S.B[(size_t)r * S.nv_t + j];
        prod += (a - mA) * (b - mB);
        varA += (a - mA) * (a - mA);
        varB += (b - mB) * (b - mB);
    }
    if (varA == 0.0 || varB == 0.0) return 0.0;
    return prod / (sqrt(varA) * sqrt(varB));
}

// out_cnt[i] = list length (-1: vertex skipped, 0 entries possible never), out_arg[i][k] = -(d1^2 + d2^2) / (2 sigma^2) or +1 when the
// entry is skipped (d1^2 + d2^2 == 0: a weight is never positive-argument otherwise), out_sim[i][k] = sim.peek(q_k, i)
__global__ void __launch_bounds__(128) k_rigid_eval(RigidView S, Rot9 R, int* __restrict__ out_cnt, double* __restrict__ out_arg,
                                                    double* __restrict__ out_sim, int* __restrict__ out_status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < S.nv_s && S.has_nbh[min(i, S.nv_s - 1)];
    const V3 pt = active ? rotated_src(S, R, i) : V3{0, 0, 0};
    int st;
    const int t = nearest_triangle<1>(S.tree, pt, active, 0, st);      // all lanes call it together
    if (i >= S.nv_s) return;
    if (!active) { out_cnt[i] = -1; return; }
    if (t < 0) { out_cnt[i] = -1; out_status[i] = st; return; }
    // tangent basis at i on the rotated source (reg_tools.cpp:205-266, Mesh::local_normal mesh.cpp:133-141, Triangle::normal triangle.cpp:42-47)
    V3 a{0, 0, 0};
    for (int k = S.src_inc_ptr[i]; k < S.src_inc_ptr[i + 1]; ++k) {
        const int tt = S.src_inc[k];
        const V3 v0 = rotated_src(S, R, S.src_tri[3 * tt]), v1 = rotated_src(S, R, S.src_tri[3 * tt + 1]), v2 = rotated_src(S, R, S.src_tri[3 * tt + 2]);
        const V3 n = vnormalized(vcross(vsub(v2, v0), vsub(v1, v0)));
        a.x += n.x; a.y += n.y; a.z += n.z;
    }
    a = vnormalized(a);
    if (vdot(a, pt) < 0) a = vscale(a, -1.0);
    // `abs(a.X)` is the C library's int abs(int) in that translation unit: the components are truncated towards zero first
    const int ax = abs((int)a.x), ay = abs((int)a.y), az = abs((int)a.z);
    V3 e1;
    if (ax >= ay && ax >= az) {
        const double mag = sqrt(a.z * a.z + a.y * a.y);
        e1 = mag == 0 ? V3{0, 0, 1} : V3{0, -a.z / mag, a.y / mag};
    } else if (ay >= ax && ay >= az) {
        const double mag = sqrt(a.z * a.z + a.x * a.x);
        e1 = mag == 0 ? V3{0, 0, 1} : V3{-a.z / mag, 0, a.x / mag};
    } else {
        const double mag = sqrt(a.y * a.y + a.x * a.x);
        e1 = mag == 0 ? V3{1, 0, 0} : V3{-a.y / mag, a.x / mag, 0};
    }
    const V3 e2 = vnormalized(vcross(a, e1));
    // union of the one-rings of the closest triangle's corners, first-seen order (cpp:143-165)
    int q[kRigidCap];
    int nq = 0;
    bool overflow = false;
    for (int c = 0; c < 3; ++c) {
        const int n = S.tgt_tri[3 * t + c];
        for (int k = S.tgt_inc_ptr[n]; k < S.tgt_inc_ptr[n + 1]; ++k) {
            const int j = S.tgt_inc[k];
            for (int m = 0; m < 3; ++m) {
                const int v = S.tgt_tri[3 * j + m];
                bool found = false;
                for (int u = 0; u < nq; ++u) found = found || q[u] == v;
                if (!found) {
                    if (nq < kRigidCap) q[nq++] = v; else overflow = true;
                }
            }
        }
    }
    if (overflow) { out_cnt[i] = -1; out_status[i] = MSMGPU_ERR_CAPACITY; return; }
    // WLS_simgradient (cpp:63-90) up to the exp
    const V3 origin = vscale(vnormalized(vcross(e1, e2)), kRad);
    const V3 ys = vsub(pt, origin);
    const double y11 = vdot(ys, e1), y21 = vdot(ys, e2);
    for (int k = 0; k < nq; ++k) {
        const V3 xs = vsub(load_pt(S.tgt_xyz, q[k]), origin);
        const double d1 = vdot(xs, e1) - y11, d2 = vdot(xs, e2) - y21;
        const double s = d1 * d1 + d2 * d2;
        out_arg[(size_t)i * kRigidCap + k] = s > 0 ? -s / S.two_sigma2 : 1.0;
        // a neighbour with id 0 is never stored in the sparse matrix (similarities.cpp:43: `if (nbh != 0)`), so its peek() is 0
        out_sim[(size_t)i * kRigidCap + k] = q[k] != 0 ? rigid_sim(S, i, q[k]) : 0.0;
    }
    out_cnt[i] = nq;
}

// Neighbourhood::update (reg_tools.cpp:31-58): only whether a list is empty survives the first evaluation
__global__ void k_rigid_has_neighbour(int nv_s, const double* __restrict__ src_xyz, int nv_t, const double* __restrict__ tgt_unit, double cos_ang,
                                      unsigned char* __restrict__ has) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nv_s) return;
    const V3 cr = vnormalized(load_pt(src_xyz, warp));
    bool any = false;
    for (int n0 = 0; n0 < nv_t && !any; n0 += 32) {
        const int n = n0 + lane;
        const bool in = n < nv_t && vdot(load_pt(tgt_unit, n), cr) >= cos_ang;
        any = __any_sync(0xffffffffu, in);
    }
    if (lane == 0) has[warp] = any ? 1 : 0;
}

__global__ void k_unit_rows(int n, const double* __restrict__ xyz, double* __restrict__ unit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = vnormalized(load_pt(xyz, i));
    unit[3 * (size_t)i] = p.x; unit[3 * (size_t)i + 1] = p.y; unit[3 * (size_t)i + 2] = p.z;
}

}  // namespace msm

using namespace msm;

struct msmgpu_rigid {
    msmgpu_ctx* ctx = nullptr;
    msmgpu_mesh* target = nullptr;      // owned
    msmgpu_octree* tree = nullptr;      // owned
    int nv_s = 0, nt_s = 0, nv_t = 0, D = 0, simmeasure = 2;
    double min_sigma = 0.0;
    DevBuf<double> src_xyz, A, B, meanA, meanB, arg, sim;
    DevBuf<int> src_tri, src_inc_ptr, src_inc, tgt_inc_ptr, tgt_inc, cnt, status;
    DevBuf<unsigned char> has;
    // per evaluation the lists (12 - 14 of the kRigidCap slots per vertex) are compacted on the device and come to the host through
    // page-locked buffers: 8.5 MB instead of 42 MB of pageable copies at ico6
    DevBuf<double> carg, csim;
    DevBuf<int> pos_cnt, off, total;
    double *h_arg = nullptr, *h_sim = nullptr;   // page-locked, grow-only (h_cap entries)
    int* h_cnt = nullptr;                        // page-locked [nv_s]
    size_t h_cap = 0;
    ~msmgpu_rigid() {
        if (h_arg) cudaFreeHost(h_arg);
        if (h_sim) cudaFreeHost(h_sim);
        if (h_cnt) cudaFreeHost(h_cnt);
    }
};

__global__ void k_rigid_positive_counts(int n, const int* __restrict__ cnt, int* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = cnt[i] > 0 ? cnt[i] : 0;
}
__global__ void k_rigid_compact(int n, const int* __restrict__ cnt, const int* __restrict__ off, const double* __restrict__ arg,
                                const double* __restrict__ sim, double* __restrict__ carg, double* __restrict__ csim) {
    const int i = blockIdx.x, m = cnt[i];
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        carg[off[i] + k] = arg[(size_t)i * kRigidCap + k];
        csim[off[i] + k] = sim[(size_t)i * kRigidCap + k];
    }
}

static void incidence_csr(int nv, int nt, const int32_t* tri, std::vector<int>& ptr, std::vector<int>& inc) {
    ptr.assign(nv + 1, 0);
    for (int k = 0; k < 3 * nt; ++k) ++ptr[tri[k] + 1];
    for (int v = 0; v < nv; ++v) ptr[v + 1] += ptr[v];
    inc.assign(ptr[nv], 0);
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (int t = 0; t < nt; ++t)                       // ascending triangle id per vertex (Mpoint::push_triangle order, mesh.cpp)
        for (int k = 0; k < 3; ++k) inc[cur[tri[3 * t + k]]++] = t;
}

// similarities.cpp:106-126 (calc_means): univariate data -> one mean over all vertices; multivariate -> mean over the channels of a vertex
static void sim_means(int D, int n, const double* M, std::vector<double>& out) {
    out.assign(n, 0.0);
    if (D == 1) {
        double sum = 0.0;
        for (int i = 0; i < n; ++i) sum += M[i];
        for (int i = 0; i < n; ++i) out[i] = sum / n;
    } else {
        for (int i = 0; i < n; ++i) {
            double sum = 0.0;
            for (int j = 0; j < D; ++j) sum += M[(size_t)j * n + i];
            out[i] = sum / D;
        }
    }
}

extern "C" {

// Mesh::calculate_MeanVD (mesh.cpp:276-294) for a mesh whose neighbour lists were filled by push_triangle in triangle order
// (mesh.cpp:119-131: first-seen order per vertex): sequential sum over vertices, then over each vertex's neighbours. Host code.
msmgpu_status msmgpu_mean_vertex_distance(int nv, const double* xyz, int nt, const int32_t* tri, double* out) {
    if (nv <= 0 || nt <= 0 || !xyz || !tri || !out) return fail(MSMGPU_ERR_INVALID, "mean_vertex_distance: bad arguments");
    std::vector<std::vector<int>> nbr(nv);
    auto add = [&](int a, int b) { if (std::find(nbr[a].begin(), nbr[a].end(), b) == nbr[a].end()) nbr[a].push_back(b); };
    for (int k = 0; k < nt; ++k) {
        const int n0 = tri[3 * k], n1 = tri[3 * k + 1], n2 = tri[3 * k + 2];
        if (n0 < 0 || n0 >= nv || n1 < 0 || n1 >= nv || n2 < 0 || n2 >= nv) return fail(MSMGPU_ERR_INVALID, "mean_vertex_distance: vertex id out of range");
        add(n0, n1); add(n0, n2); add(n1, n0); add(n1, n2); add(n2, n0); add(n2, n1);
    }
    long k = 0;
    double kr = 0.0;
    for (int i = 0; i < nv; ++i) {
        const V3 sp{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        for (int j : nbr[i]) { ++k; kr += vnorm(vsub(V3{xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]}, sp)); }
    }
    *out = kr / (double)k;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_rigid_create(msmgpu_ctx* ctx, int nv_t, const double* tgt_xyz, int nt_t, const int32_t* tgt_tri, int nv_s, const double* src_xyz,
                                  int nt_s, const int32_t* src_tri, int D, const double* src_feat, const double* ref_feat, int simmeasure,
                                  double mean_vertex_distance, msmgpu_rigid** out) {
    if (!ctx || !out || nv_t <= 0 || nt_t <= 0 || nv_s <= 0 || nt_s <= 0 || D <= 0 || !tgt_xyz || !tgt_tri || !src_xyz || !src_tri || !src_feat || !ref_feat)
        return fail(MSMGPU_ERR_INVALID, "rigid_create: bad arguments");
    if (simmeasure != 1 && simmeasure != 2) return fail(MSMGPU_ERR_INVALID, "rigid_create: simmeasure must be 1 (SSD) or 2 (correlation)");
    *out = nullptr;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::unique_ptr<msmgpu_rigid> r(new msmgpu_rigid());
    r->ctx = ctx; r->nv_s = nv_s; r->nt_s = nt_s; r->nv_t = nv_t; r->D = D; r->simmeasure = simmeasure;
    r->min_sigma = mean_vertex_distance;                                        // min_sigma = MVD = SOURCE.calculate_MeanVD() (cpp:35)
    MSM_TRY(msmgpu_mesh_create(ctx, nv_t, tgt_xyz, nt_t, tgt_tri, &r->target));
    msmgpu_status st = msmgpu_octree_build(r->target, &r->tree);
    if (st != MSMGPU_OK) { msmgpu_mesh_destroy(r->target); return st; }
    auto up_d = [&](DevBuf<double>& b, const double* h, size_t n) -> msmgpu_status {
        MSM_CUDA(b.alloc(n, s));
        MSM_CUDA(cudaMemcpyAsync(b.p, h, n * sizeof(double), cudaMemcpyHostToDevice, s));
        return MSMGPU_OK;
    };
    auto up_i = [&](DevBuf<int>& b, const int* h, size_t n) -> msmgpu_status {
        MSM_CUDA(b.alloc(n, s));
        MSM_CUDA(cudaMemcpyAsync(b.p, h, n * sizeof(int), cudaMemcpyHostToDevice, s));
        return MSMGPU_OK;
    };
    std::vector<int> sp, si, tp, ti;
    incidence_csr(nv_s, nt_s, src_tri, sp, si);
    incidence_csr(nv_t, nt_t, tgt_tri, tp, ti);
    std::vector<double> mA, mB;
    sim_means(D, nv_s, src_feat, mA);
    sim_means(D, nv_t, ref_feat, mB);
    msmgpu_status ok = MSMGPU_OK;
    auto chk = [&](msmgpu_status x) { if (ok == MSMGPU_OK) ok = x; };
    chk(up_d(r->src_xyz, src_xyz, 3 * (size_t)nv_s));
    chk(up_i(r->src_tri, src_tri, 3 * (size_t)nt_s));
    chk(up_i(r->src_inc_ptr, sp.data(), sp.size()));
    chk(up_i(r->src_inc, si.data(), si.size()));
    chk(up_i(r->tgt_inc_ptr, tp.data(), tp.size()));
    chk(up_i(r->tgt_inc, ti.data(), ti.size()));
    chk(up_d(r->A, src_feat, (size_t)D * nv_s));
    chk(up_d(r->B, ref_feat, (size_t)D * nv_t));
    chk(up_d(r->meanA, mA.data(), mA.size()));
    chk(up_d(r->meanB, mB.data(), mB.size()));
    if (ok == MSMGPU_OK && r->has.alloc(nv_s, s) != cudaSuccess) ok = fail(MSMGPU_ERR_CUDA, "rigid_create: allocation failed");
    if (ok == MSMGPU_OK && (r->cnt.alloc(nv_s, s) != cudaSuccess || r->status.alloc(nv_s, s) != cudaSuccess ||
                            r->arg.alloc((size_t)nv_s * kRigidCap, s) != cudaSuccess || r->sim.alloc((size_t)nv_s * kRigidCap, s) != cudaSuccess))
        ok = fail(MSMGPU_ERR_CUDA, "rigid_create: allocation failed");
    if (ok != MSMGPU_OK) { msmgpu_octree_destroy(r->tree); msmgpu_mesh_destroy(r->target); return ok; }
    // Neighbourhood::update(SOURCE, TARGET, 2 asin(4 MVD / 2R)) (cpp:41, reg_tools.cpp:31-58): cos(ang) on the host libm
    const double ang = 2 * std::asin(4 * mean_vertex_distance / (2 * kRad));
    DevBuf<double> unit;
    MSM_CUDA(unit.alloc(3 * (size_t)nv_t, s));
    k_unit_rows<<<(nv_t + 255) / 256, 256, 0, s>>>(nv_t, r->target->xyz.p, unit.p);
    MSM_LAUNCH_CHECK();
    k_rigid_has_neighbour<<<(unsigned)(((size_t)nv_s * 32 + 255) / 256), 256, 0, s>>>(nv_s, r->src_xyz.p, nv_t, unit.p, std::cos(ang), r->has.p);
    MSM_LAUNCH_CHECK();
    MSM_CUDA(cudaStreamSynchronize(s));       // the host vectors above go out of scope
    MSM_CUDA(r->pos_cnt.alloc((size_t)nv_s, s));
    MSM_CUDA(r->off.alloc((size_t)nv_s, s));
    MSM_CUDA(r->total.alloc(1, s));
    MSM_CUDA(r->carg.alloc((size_t)nv_s * kRigidCap, s));
    MSM_CUDA(r->csim.alloc((size_t)nv_s * kRigidCap, s));
    MSM_CUDA(cudaHostAlloc((void**)&r->h_cnt, (size_t)nv_s * sizeof(int), cudaHostAllocDefault));
    *out = r.release();
    return MSMGPU_OK;
}

void msmgpu_rigid_destroy(msmgpu_rigid* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    msmgpu_octree_destroy(r->tree);
    msmgpu_mesh_destroy(r->target);
    delete r;
}

msmgpu_status msmgpu_rigid_cost(msmgpu_rigid* r, const double* src_xyz, double dw1, double dw2, double dw3, double* cost) {
    if (!r || !cost) return fail(MSMGPU_ERR_INVALID, "rigid_cost: bad arguments");
    msmgpu_ctx* ctx = r->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (src_xyz) MSM_CUDA(cudaMemcpyAsync(r->src_xyz.p, src_xyz, 3 * (size_t)r->nv_s * sizeof(double), cudaMemcpyHostToDevice, s));
    // euler_rotate's matrix (point.cpp:156-165), host libm
    Rot9 R;
    R.m[0] = std::cos(dw2) * std::cos(dw3);
    R.m[1] = -std::cos(dw1) * std::sin(dw3) + std::sin(dw1) * std::sin(dw2) * std::cos(dw3);
    R.m[2] = std::sin(dw1) * std::sin(dw3) + std::cos(dw1) * std::sin(dw2) * std::cos(dw3);
    R.m[3] = std::cos(dw2) * std::sin(dw3);
    R.m[4] = std::cos(dw1) * std::cos(dw3) + std::sin(dw1) * std::sin(dw2) * std::sin(dw3);
    R.m[5] = -std::sin(dw1) * std::cos(dw3) + std::cos(dw1) * std::sin(dw2) * std::sin(dw3);
    R.m[6] = -std::sin(dw2);
    R.m[7] = std::sin(dw1) * std::cos(dw2);
    R.m[8] = std::cos(dw1) * std::cos(dw2);
    RigidView V{r->tree->view(), r->target->xyz.p, r->target->tri.p, r->tgt_inc_ptr.p, r->tgt_inc.p, r->src_xyz.p, r->src_tri.p, r->src_inc_ptr.p,
                r->src_inc.p, r->has.p, r->A.p, r->B.p, r->meanA.p, r->meanB.p, r->nv_s, r->nv_t, r->D, r->simmeasure,
                2 * r->min_sigma * r->min_sigma};
    MSM_CUDA(cudaMemsetAsync(r->status.p, 0, (size_t)r->nv_s * sizeof(int), s));
    k_rigid_eval<<<(r->nv_s + 127) / 128, 128, 0, s>>>(V, R, r->cnt.p, r->arg.p, r->sim.p, r->status.p);
    MSM_LAUNCH_CHECK();
    int code = 0;
    MSM_TRY(first_error(r->status.p, (size_t)r->nv_s, s, &code));
    if (code == MSMGPU_ERR_CAPACITY) return fail(MSMGPU_ERR_CAPACITY, "rigid_cost: a neighbour list exceeds the per-vertex capacity");
    if (code) return status_to_error(code);
    const int n = r->nv_s;
    k_rigid_positive_counts<<<(n + 255) / 256, 256, 0, s>>>(n, r->cnt.p, r->pos_cnt.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(r->pos_cnt.p, r->off.p, n, r->total.p, s));
    k_rigid_compact<<<n, 32, 0, s>>>(n, r->cnt.p, r->off.p, r->arg.p, r->sim.p, r->carg.p, r->csim.p);
    MSM_LAUNCH_CHECK();
    int h_total = 0;
    MSM_CUDA(cudaMemcpyAsync(&h_total, r->total.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaMemcpyAsync(r->h_cnt, r->cnt.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    if ((size_t)h_total > r->h_cap) {
        if (r->h_arg) cudaFreeHost(r->h_arg);
        if (r->h_sim) cudaFreeHost(r->h_sim);
        r->h_arg = r->h_sim = nullptr;
        r->h_cap = 0;
        const size_t cap = (size_t)h_total + (size_t)h_total / 4 + 1024;
        MSM_CUDA(cudaHostAlloc((void**)&r->h_arg, cap * sizeof(double), cudaHostAllocDefault));
        MSM_CUDA(cudaHostAlloc((void**)&r->h_sim, cap * sizeof(double), cudaHostAllocDefault));
        r->h_cap = cap;
    }
    if (h_total > 0) {
        MSM_CUDA(cudaMemcpyAsync(r->h_arg, r->carg.p, (size_t)h_total * sizeof(double), cudaMemcpyDeviceToHost, s));
        MSM_CUDA(cudaMemcpyAsync(r->h_sim, r->csim.p, (size_t)h_total * sizeof(double), cudaMemcpyDeviceToHost, s));
        MSM_CUDA(cudaStreamSynchronize(s));
    }
    // WLS_simgradient's weights and sums (cpp:79-88) with the host libm's exp, then the sequential total (cpp:136-137)
    std::vector<size_t> start((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) start[i + 1] = start[i] + (size_t)(r->h_cnt[i] > 0 ? r->h_cnt[i] : 0);
    std::vector<double> cur((size_t)n, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const int m = r->h_cnt[i];
        if (m < 0) continue;                      // no neighbourhood: current_sim(i) keeps its initial 0
        double SUM = 0.0, JP = 0.0;
        const double *pa = r->h_arg + start[i], *ps = r->h_sim + start[i];
        for (int k = 0; k < m; ++k) {
            const double a = pa[k];
            if (a <= 0.0) {                       // (dist_1^2 + dist_2^2) > 0
                const double w = std::exp(a);
                SUM += w;
                JP += ps[k] * w;
            }
        }
        if (SUM > 0) JP /= SUM;
        cur[i] = JP;
    }
    double total = 0.0;
    for (int i = 0; i < n; ++i) total += cur[i];
    *cost = total;
    return MSMGPU_OK;
}

}  // extern "C"
