// Nearest-triangle queries on the flattened octree and the kernels fed by them:
// barycentric weights, coordinate blends (sphere_project_warp / surface_resample),
// nearest-vertex gathers and the fused query -> weights -> 128-bit row gather resampler.
//
// Reference behaviour reproduced (msm-newresampler/src): Octree::get_closest_triangle
// (octree.cpp:156-214), get_closest_vertex_ID (216-233), Triangle::calc_barycentric_weights
// (triangle.cpp:124-143), Resampler::get_barycentric_weights (resampler.cpp:142-167), the
// interpolation loop of barycentric_data_interpolation (resampler.cpp:40-52),
// sphere_project_warp (311-328), surface_resample (284-302), nearest_neighbour_interpolation
// (232-258).
//
// Execution model: a query is owned by a GROUP of G lanes (G = 1..32, compile-time). The lanes
// of a group descend the tree redundantly (same addresses -> one broadcast load), then test
// the leaf's triangles G at a time, one 128-byte TriRec line per lane, and arg-min over
// (distance, scan position) with width-G shuffles. The reference keeps the FIRST triangle of
// the scan with the strictly smallest distance, so ties are broken by scan position, which
// makes the result independent of G.
#include "query.cuh"

namespace msm {

// ------------------------------------------------------------------------------------------
// stand-alone query kernels
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256) k_nearest(TreeView T, int n, const double* __restrict__ pts,
                                                 int* __restrict__ out_tri, int* __restrict__ out_vtx, int* __restrict__ out_status) {
    const int gl = threadIdx.x % G;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = q < n;
    const V3 pt = active ? load_pt(pts, q) : V3{0, 0, 0};
    int st;
    const int t = nearest_triangle<G>(T, pt, active, gl, st);
    if (!active || gl != 0) return;
    if (out_tri) out_tri[q] = t;
    if (out_status) out_status[q] = st;
    if (out_vtx) {   // octree.cpp:216-233
        int id = -1;
        if (t >= 0) {
            double dist = DBL_MAX;
            id = 0;
            const double* v = T.rec[t].v;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double d = vnorm(vsub(pt, V3{v[3 * k], v[3 * k + 1], v[3 * k + 2]}));
                if (d < dist) { id = __ldg(T.tri + 3 * (size_t)t + k); dist = d; }
            }
        }
        out_vtx[q] = id;
    }
}

template <int G>
__global__ void __launch_bounds__(256) k_bary_weights(TreeView T, int n, const double* __restrict__ pts, int* __restrict__ out_idx,
                                                      double* __restrict__ out_w, int* __restrict__ out_ne, int* __restrict__ out_status) {
    const int gl = threadIdx.x % G;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = q < n;
    const V3 pt = active ? load_pt(pts, q) : V3{0, 0, 0};
    int st;
    const int t = nearest_triangle<G>(T, pt, active, gl, st);
    if (!active || gl != 0) return;
    int idx[3] = {-1, -1, -1};
    double w[3] = {0, 0, 0};
    int ne = 0;
    if (t >= 0) ne = sorted_weights(T, t, pt, idx, w);
#pragma unroll
    for (int j = 0; j < 3; ++j) { out_idx[3 * (size_t)q + j] = idx[j]; out_w[3 * (size_t)q + j] = w[j]; }
    if (out_ne) out_ne[q] = ne;
    if (out_status) out_status[q] = st;
}

// the same for a batch of jobs (blockIdx.y = job): different trees and/or different point sets, outputs concatenated
template <int G, int MINB, bool LAZY = false>
__global__ void __launch_bounds__(256, MINB) k_bary_weights_batch(const QueryJob* __restrict__ jobs, int* __restrict__ out_idx, double* __restrict__ out_w,
                                                            int* __restrict__ out_ne, int* __restrict__ out_status) {
    const QueryJob job = jobs[blockIdx.y];
    const int gl = threadIdx.x % G;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if ((blockIdx.x * blockDim.x) / G >= job.n) return;     // whole CTA beyond this job's points
    const bool active = k < job.n;
    const int q = (active && job.perm) ? __ldg(job.perm + k) : k;   // processing order only; outputs are indexed by the point
    const V3 pt = active ? load_pt(job.pts, q) : V3{0, 0, 0};
    int st;
    const int t = nearest_triangle<G, LAZY>(job.tree, pt, active, gl, st);
    if (!active || gl != 0) return;
    int idx[3] = {-1, -1, -1};
    double w[3] = {0, 0, 0};
    int ne = 0;
    if (t >= 0) ne = sorted_weights<LAZY>(job.tree, t, pt, idx, w);
    const size_t o = (size_t)job.out_off + q;
#pragma unroll
    for (int j = 0; j < 3; ++j) { out_idx[3 * o + j] = idx[j]; out_w[3 * o + j] = w[j]; }
    out_ne[o] = ne;
    out_status[o] = st;
}

// The same for jobs that share ONE tree and one point count (the reverse queries of a batch: every subject's vertices located in the
// target sphere's tree), with the SUBJECT on the lanes: thread g handles point perm[g / n_jobs] of subject g % n_jobs. Subjects of a
// batch share one topology and nearly the same geometry, so the 32 lanes of a warp descend to the same leaf and test the same
// candidates: node, list, cull-sphere and record loads become broadcasts (one L1 wavefront per request instead of ~2.8,
// profiles/r2_gather_summary.md: the kernel is bound by L1 wavefronts). Outputs are written at the point's own slot: same results.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_bary_weights_across(TreeView T, const int* __restrict__ perm, const double* const* __restrict__ pts,
                                                                   const int* __restrict__ out_off, int n_jobs, int n, int* __restrict__ out_idx,
                                                                   double* __restrict__ out_w, int* __restrict__ out_ne, int* __restrict__ out_status) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = g < (long long)n * n_jobs;
    const int j = active ? (int)(g % n_jobs) : 0;
    const int k = active ? (int)(g / n_jobs) : 0;
    const int q = perm ? __ldg(perm + k) : k;
    const V3 pt = active ? load_pt(pts[j], q) : V3{0, 0, 0};
    int st;
    const int t = nearest_triangle<1>(T, pt, active, 0, st);
    if (!active) return;
    int idx[3] = {-1, -1, -1};
    double w[3] = {0, 0, 0};
    int ne = 0;
    if (t >= 0) ne = sorted_weights(T, t, pt, idx, w);
    const size_t o = (size_t)__ldg(out_off + j) + q;
#pragma unroll
    for (int i = 0; i < 3; ++i) { out_idx[3 * o + i] = idx[i]; out_w[3 * o + i] = w[i]; }
    out_ne[o] = ne;
    out_status[o] = st;
}

// sphere_project_warp (resampler.cpp:311-328, reproject = 1) / surface_resample (284-302, reproject = 0):
// newPt = sum over the weight map (ascending id) of payload[id] * w, optionally normalised * 100.
template <int G>
__global__ void __launch_bounds__(256) k_blend_coords(TreeView T, int n, const double* __restrict__ pts, const double* __restrict__ payload,
                                                      double* __restrict__ out, int reproject, int* __restrict__ out_status) {
    const int gl = threadIdx.x % G;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = q < n;
    const V3 pt = active ? load_pt(pts, q) : V3{0, 0, 0};
    int st;
    const int t = nearest_triangle<G>(T, pt, active, gl, st);
    if (!active || gl != 0) return;
    V3 np{0, 0, 0};
    if (t >= 0) {
        int idx[3];
        double w[3];
        const int ne = sorted_weights(T, t, pt, idx, w);
        for (int j = 0; j < ne; ++j) {   // Point*double then += (point.cpp:198,224)
            const V3 c = load_pt(payload, idx[j]);
            const V3 s = vscale(c, w[j]);
            np.x += s.x; np.y += s.y; np.z += s.z;
        }
        if (reproject) {   // resampler.cpp:324-325
            np = vnormalized(np);
            np.x *= kRad; np.y *= kRad; np.z *= kRad;
        }
    }
    out[3 * (size_t)q] = np.x; out[3 * (size_t)q + 1] = np.y; out[3 * (size_t)q + 2] = np.z;
    if (out_status) out_status[q] = st;
}

// ------------------------------------------------------------------------------------------
// fused barycentric resampler: query -> weights (shared memory) -> 3-row gather of D floats
//
// Warp-autonomous: every warp owns 32 consecutive targets, runs their queries (32/G at a time), parks
// (ids, weights) in its own slice of shared memory and then streams the feature rows for the same 32
// targets. No block-wide barrier (ncu r1b: 29 % of the stall samples sat on it), and phase B keeps 4 slots
// = 12 independent 128-bit loads in flight per thread (r1b: 38 % of the samples waited on ONE slot's loads).
// ------------------------------------------------------------------------------------------
constexpr int kResThreads = 256;
constexpr int kResWarpTile = 32;                                   // targets per warp
constexpr int kResTile = kResWarpTile * (kResThreads / 32);        // targets per CTA

// U = gather slots in flight per thread, MINB = CTAs per SM the register allocation must allow.
// ASYNC: phase B requests its rows with 16-byte asynchronous copies (cp.async.cg, LDGSTS) into a per-warp staging slice of dynamic
// shared memory instead of loading them into registers: the bytes in flight then cost shared memory (U * 3 * 512 B per warp), not
// registers, so U can grow without spilling the FP64 geometry of phase A. Same arithmetic, same order.
template <int G, int kResUnroll, int MINB, bool ASYNC = false>
__global__ void __launch_bounds__(kResThreads, MINB) k_bary_resample_f32(const ResampleJob* __restrict__ jobs, int n, const double* __restrict__ pts,
                                                                  int D, int* __restrict__ out_status) {
    extern __shared__ float4 s_stage_all[];   // ASYNC only: [warps][U][3][32] float4
    __shared__ int s_idx_all[kResTile * 3];
    __shared__ double s_w_all[kResTile * 3];
    __shared__ int s_ne_all[kResTile];
    const ResampleJob job = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile0 = blockIdx.x * kResTile + warp * kResWarpTile;   // first target of this warp
    if (tile0 >= n) return;                                          // whole warp out of range
    int* s_idx = s_idx_all + warp * kResWarpTile * 3;
    double* s_w = s_w_all + warp * kResWarpTile * 3;
    int* s_ne = s_ne_all + warp * kResWarpTile;
    const int gl = lane % G;

    // phase A: one lane group per target (FP64, L1/L2-resident tree, cull spheres and triangle records)
#pragma unroll 1
    for (int q = lane / G; q < kResWarpTile; q += 32 / G) {
        const int k = tile0 + q;
        const bool active = k < n;
        const V3 pt = active ? load_pt(pts, k) : V3{0, 0, 0};
        int st;
        const int t = nearest_triangle<G>(job.tree, pt, active, gl, st);
        if (gl == 0) {
            int idx[3] = {-1, -1, -1};
            double w[3] = {0, 0, 0};
            int ne = 0;
            if (t >= 0) ne = sorted_weights(job.tree, t, pt, idx, w);
#pragma unroll
            for (int j = 0; j < 3; ++j) { s_idx[3 * q + j] = idx[j]; s_w[3 * q + j] = w[j]; }
            s_ne[q] = ne;
            if (active && out_status) out_status[(size_t)blockIdx.y * n + k] = st;
            if (active && job.keep_idx) {   // hand the weight map to the adaptive-weights pass of the same batch
#pragma unroll
                for (int j = 0; j < 3; ++j) { job.keep_idx[3 * (size_t)k + j] = idx[j]; job.keep_w[3 * (size_t)k + j] = w[j]; }
                job.keep_ne[k] = ne;
            }
        }
    }
    __syncwarp();

    // phase B: out[k][:] = sum_j in[idx_j][:] * w_j, accumulated in FP64 in ascending-id order like
    // resampler.cpp:46-48, rows read and written as 128-bit words. HBM-bound part.
    const float* __restrict__ fin = job.feat_in;
    float* __restrict__ fout = job.feat_out;
    const int rows = min(kResWarpTile, n - tile0);
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(fin) | reinterpret_cast<uintptr_t>(fout)) & 15) == 0) {
        const int D4 = D >> 2;
        const float4* __restrict__ in4 = reinterpret_cast<const float4*>(fin);
        float4* __restrict__ out4 = reinterpret_cast<float4*>(fout) + (size_t)tile0 * D4;   // the warp's rows are contiguous
        const int slots = rows * D4;
        if (ASYNC) {
            float4* stage = s_stage_all + (size_t)warp * kResUnroll * 3 * 32;
            for (int s0 = lane; s0 < slots; s0 += 32 * kResUnroll) {
                double w[kResUnroll][3];
                int nes[kResUnroll];
#pragma unroll
                for (int u = 0; u < kResUnroll; ++u) {
                    const int s = min(s0 + 32 * u, slots - 1);
                    const int q = s / D4, c = s - q * D4;
                    const int ne = s_ne[q];
                    nes[u] = ne;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const bool on = j < ne;
                        const int row = on ? s_idx[3 * q + j] : max(s_idx[3 * q], 0);
                        w[u][j] = on ? s_w[3 * q + j] : 0.0;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + (u * 3 + j) * 32 + lane);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(in4 + (size_t)row * D4 + c) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                // every lane reads back only what it requested itself: no cross-lane visibility is needed
#pragma unroll
                for (int u = 0; u < kResUnroll; ++u) {
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double ww = w[u][j];
                        const float4 v = j < nes[u] ? stage[(u * 3 + j) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                        a0 += (double)v.x * ww; a1 += (double)v.y * ww; a2 += (double)v.z * ww; a3 += (double)v.w * ww;
                    }
                    if (s0 + 32 * u < slots) __stcs(out4 + (s0 + 32 * u), make_float4((float)a0, (float)a1, (float)a2, (float)a3));
                }
            }
        } else {
            for (int s0 = lane; s0 < slots; s0 += 32 * kResUnroll) {
                // branch-free issue of 12 independent 128-bit loads (out-of-range slots re-read the last slot, absent
                // map entries re-read entry 0 with weight 0), conversions and FP64 sums only afterwards
                float4 f[kResUnroll][3];
                double w[kResUnroll][3];
                int nes[kResUnroll];
    #pragma unroll
                for (int u = 0; u < kResUnroll; ++u) {
                    const int s = min(s0 + 32 * u, slots - 1);
                    const int q = s / D4, c = s - q * D4;
                    const int ne = s_ne[q];
                    nes[u] = ne;
    #pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const bool on = j < ne;
                        const int row = on ? s_idx[3 * q + j] : max(s_idx[3 * q], 0);
                        w[u][j] = on ? s_w[3 * q + j] : 0.0;
                        f[u][j] = __ldg(in4 + (size_t)row * D4 + c);
                    }
                }
    #pragma unroll
                for (int u = 0; u < kResUnroll; ++u) {
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    #pragma unroll
                    for (int j = 0; j < 3; ++j) {   // an ABSENT entry adds 0; a present one with weight 0 still multiplies (NaN * 0 = NaN, resampler.cpp:46-48)
                        const double ww = w[u][j];
                        const float4 v = j < nes[u] ? f[u][j] : make_float4(0.f, 0.f, 0.f, 0.f);
                        a0 += (double)v.x * ww; a1 += (double)v.y * ww; a2 += (double)v.z * ww; a3 += (double)v.w * ww;
                    }
                    if (s0 + 32 * u < slots) __stcs(out4 + (s0 + 32 * u), make_float4((float)a0, (float)a1, (float)a2, (float)a3));
                }
            }
        }
    } else {
        const int slots = rows * D;
        for (int s = lane; s < slots; s += 32) {
            const int q = s / D, c = s - q * D;
            const int ne = s_ne[q];
            double a = 0.0;
            for (int j = 0; j < ne; ++j) a += (double)__ldg(fin + (size_t)s_idx[3 * q + j] * D + c) * s_w[3 * q + j];
            fout[(size_t)(tile0 + q) * D + c] = (float)a;
        }
    }
}

// nearest_neighbour_interpolation (resampler.cpp:232-258): out[d][i] = in[d][vertex_i], channel-major doubles
__global__ void k_gather_channels_f64(int n, int nv, int D, const int* __restrict__ vtx, const double* __restrict__ in, double* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)n * D) return;
    const int d = (int)(i / n), k = (int)(i - (size_t)d * n);
    const int v = vtx[k];
    out[i] = v >= 0 ? in[(size_t)d * nv + v] : 0.0;
}

// ------------------------------------------------------------------------------------------
// launchers. The group width is a tuning knob (MSMGPU_QUERY_GROUP = 1,2,4,8,16,32; default 1).
// ------------------------------------------------------------------------------------------
static bool valid_group(int v) { return v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32; }

static std::atomic<int>& query_group_slot() {   // initialised once (thread-safe function-local static) from the environment
    static std::atomic<int> g{[] {
        const char* e = getenv("MSMGPU_QUERY_GROUP");
        const int v = e ? atoi(e) : 1;
        return valid_group(v) ? v : 1;
    }()};
    return g;
}

int query_group_width() { return query_group_slot().load(std::memory_order_relaxed); }

#define MSM_DISPATCH_G(G_, ...)                          \
    switch (G_) {                                        \
        case 1: { constexpr int G = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int G = 2; __VA_ARGS__; } break;   \
        case 4: { constexpr int G = 4; __VA_ARGS__; } break;   \
        case 16: { constexpr int G = 16; __VA_ARGS__; } break; \
        case 32: { constexpr int G = 32; __VA_ARGS__; } break; \
        default: { constexpr int G = 8; __VA_ARGS__; } break;  \
    }

static inline unsigned query_blocks(int n, int g) { return (unsigned)(((long long)n * g + 255) / 256); }

msmgpu_status launch_nearest(const TreeView& t, int n, const double* d_pts, int* d_tri, int* d_vertex, int* d_status, cudaStream_t s) {
    if (n <= 0) return MSMGPU_OK;
    const int g = query_group_width();
    MSM_DISPATCH_G(g, (k_nearest<G><<<query_blocks(n, G), 256, 0, s>>>(t, n, d_pts, d_tri, d_vertex, d_status)));
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status launch_bary_weights(const TreeView& t, int n, const double* d_pts, int* d_idx, double* d_w, int* d_ne, int* d_status, cudaStream_t s) {
    if (n <= 0) return MSMGPU_OK;
    const int g = query_group_width();
    MSM_DISPATCH_G(g, (k_bary_weights<G><<<query_blocks(n, G), 256, 0, s>>>(t, n, d_pts, d_idx, d_w, d_ne, d_status)));
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status launch_bary_weights_batch(const QueryJob* d_jobs, int n_jobs, int max_n, int* d_idx, double* d_w, int* d_ne, int* d_status, cudaStream_t s,
                                        bool lazy) {
    if (n_jobs <= 0 || max_n <= 0) return MSMGPU_OK;
    const int g = query_group_width();
    const int minb = tuning_get("weights_minb", "MSMGPU_WEIGHTS_MINB", 4);   // tuning knob: resident CTAs per SM (4: 64 registers, measured best)
    const dim3 grid(query_blocks(max_n, g), (unsigned)n_jobs);
    if (lazy) {   // trees without stored records (subjects of a batch job): the record values come from the corners
        const int minb_lazy = tuning_get("weights_minb_lazy", "MSMGPU_WEIGHTS_MINB_LAZY", 4);
        switch (minb_lazy) {
            case 2: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 2, true><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
            case 3: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 3, true><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
            default: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 4, true><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
        }
        MSM_LAUNCH_CHECK();
        return MSMGPU_OK;
    }
    switch (minb) {
        case 2: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 2><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
        case 3: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 3><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
        case 5: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 5><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
        default: MSM_DISPATCH_G(g, (k_bary_weights_batch<G, 4><<<grid, 256, 0, s>>>(d_jobs, d_idx, d_w, d_ne, d_status))); break;
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status launch_bary_weights_across(const TreeView& t, const int* d_perm, const double* const* d_pts, const int* d_out_off, int n_jobs, int n,
                                         int* d_idx, double* d_w, int* d_ne, int* d_status, cudaStream_t s) {
    if (n_jobs <= 0 || n <= 0) return MSMGPU_OK;
    const unsigned blocks = (unsigned)(((long long)n * n_jobs + 255) / 256);
    switch (tuning_get("across_minb", "MSMGPU_ACROSS_MINB", 4)) {   // resident CTAs per SM
        case 3: k_bary_weights_across<3><<<blocks, 256, 0, s>>>(t, d_perm, d_pts, d_out_off, n_jobs, n, d_idx, d_w, d_ne, d_status); break;
        case 5: k_bary_weights_across<5><<<blocks, 256, 0, s>>>(t, d_perm, d_pts, d_out_off, n_jobs, n, d_idx, d_w, d_ne, d_status); break;
        default: k_bary_weights_across<4><<<blocks, 256, 0, s>>>(t, d_perm, d_pts, d_out_off, n_jobs, n, d_idx, d_w, d_ne, d_status); break;
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status launch_blend_coords(const TreeView& t, int n, const double* d_pts, const double* d_payload_xyz, double* d_out, int reproject,
                                  int* d_status, cudaStream_t s) {
    if (n <= 0) return MSMGPU_OK;
    const int g = query_group_width();
    MSM_DISPATCH_G(g, (k_blend_coords<G><<<query_blocks(n, G), 256, 0, s>>>(t, n, d_pts, d_payload_xyz, d_out, reproject, d_status)));
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

template <int G, int U, int MINB>
static void launch_async(dim3 grid, const ResampleJob* d_jobs, int n, const double* d_pts, int D, int* d_status, cudaStream_t s) {
    const size_t smem = (size_t)(kResThreads / 32) * U * 3 * 32 * sizeof(float4);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_bary_resample_f32<G, U, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    k_bary_resample_f32<G, U, MINB, true><<<grid, kResThreads, smem, s>>>(d_jobs, n, d_pts, D, d_status);
}

msmgpu_status launch_bary_resample_f32(const ResampleJob* d_jobs, int n_jobs, int n, const double* d_pts, int D, int* d_status, cudaStream_t s) {
    if (n <= 0 || n_jobs <= 0) return MSMGPU_OK;
    const int g = query_group_width();
    const dim3 grid((unsigned)((n + kResTile - 1) / kResTile), (unsigned)n_jobs);
    const int variant = tuning_get("resample_variant", "MSMGPU_RESAMPLE_VARIANT", 2);
    switch (variant) {   // tuning knob (profiles/): registers per thread vs loads in flight
        case 1: MSM_DISPATCH_G(g, (k_bary_resample_f32<G, 2, 3><<<grid, kResThreads, 0, s>>>(d_jobs, n, d_pts, D, d_status))); break;
        case 2: MSM_DISPATCH_G(g, (k_bary_resample_f32<G, 2, 4><<<grid, kResThreads, 0, s>>>(d_jobs, n, d_pts, D, d_status))); break;
        case 3: MSM_DISPATCH_G(g, (k_bary_resample_f32<G, 4, 3><<<grid, kResThreads, 0, s>>>(d_jobs, n, d_pts, D, d_status))); break;
        case 5: MSM_DISPATCH_G(g, (k_bary_resample_f32<G, 2, 5><<<grid, kResThreads, 0, s>>>(d_jobs, n, d_pts, D, d_status))); break;
        // asynchronous-copy staging (cp.async): U slots x 3 rows x 512 B per warp of dynamic shared memory
        case 8: MSM_DISPATCH_G(g, (launch_async<G, 3, 4>(grid, d_jobs, n, d_pts, D, d_status, s))); break;
        case 9: MSM_DISPATCH_G(g, (launch_async<G, 4, 3>(grid, d_jobs, n, d_pts, D, d_status, s))); break;
        case 10: MSM_DISPATCH_G(g, (launch_async<G, 2, 4>(grid, d_jobs, n, d_pts, D, d_status, s))); break;
        case 11: MSM_DISPATCH_G(g, (launch_async<G, 6, 3>(grid, d_jobs, n, d_pts, D, d_status, s))); break;
        default: MSM_DISPATCH_G(g, (k_bary_resample_f32<G, 4, 2><<<grid, kResThreads, 0, s>>>(d_jobs, n, d_pts, D, d_status))); break;
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status launch_gather_channels_f64(int n, int nv, int D, const int* d_vtx, const double* d_in, double* d_out, cudaStream_t s) {
    const size_t total = (size_t)n * D;
    if (total == 0) return MSMGPU_OK;
    k_gather_channels_f64<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(n, nv, D, d_vtx, d_in, d_out);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

} // namespace msm

extern "C" msmgpu_status msmgpu_set_query_group(int lanes) {
    if (!msm::valid_group(lanes)) return msm::fail(MSMGPU_ERR_INVALID, "set_query_group: lanes must be 1, 2, 4, 8, 16 or 32");
    msm::query_group_slot().store(lanes, std::memory_order_relaxed);
    return MSMGPU_OK;
}
extern "C" int msmgpu_get_query_group(void) { return msm::query_group_width(); }
