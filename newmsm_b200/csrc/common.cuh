// Internal declarations shared by the kernels and the C-ABI layer (include/msmgpu.h).
#pragma once
#include <cuda_runtime.h>
#include <climits>
#include <cstdint>
#include <atomic>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/msmgpu.h"
#include "geom.cuh"

namespace msm {

void set_error(const std::string& msg);
msmgpu_status fail(msmgpu_status st, const std::string& msg);

#define MSM_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::msm::fail(MSMGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// after every kernel launch: count it (msmgpu_launch_count) and pick up launch errors
extern std::atomic<unsigned long long> g_launch_count;   // contexts are driven from several host threads
#define MSM_LAUNCH_CHECK()                 \
    do {                                   \
        ::msm::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
        MSM_CUDA(cudaGetLastError());      \
    } while (0)

#define MSM_TRY(expr)                          \
    do {                                       \
        msmgpu_status _s = (expr);             \
        if (_s != MSMGPU_OK) return _s;        \
    } while (0)

// Stream-ordered scratch allocation (cudaMallocAsync pool, kept warm by a high release threshold).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    bool own = true;     // false: a view of memory owned elsewhere (the caller's buffers, or a slab shared by a batch of meshes)
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s), own(o.own) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; s = o.s; own = o.own; o.p = nullptr; o.n = 0; return *this; }
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream; n = count; own = true;
        if (count == 0) { p = nullptr; return cudaSuccess; }
        return cudaMallocAsync((void**)&p, count * sizeof(T), stream);
    }
    void borrow(T* ptr, size_t count) {
        release();
        p = ptr; n = count; own = false;
    }
    void release() {
        if (p && own) cudaFreeAsync(p, s);
        p = nullptr; n = 0; own = true;
    }
};

// One int4 per octree node: x = first child (-1 for a leaf; the 8 children are contiguous in the
// reference's [i][j][k] order), y = start of the triangle list in `pairs`, z = number of triangles
// (0 for internal nodes, like Node::triangles_size() after clear_triangles(), octree.cpp:129),
// w = parent (-1 for a root).
struct Forest {
    msmgpu_ctx* ctx = nullptr;
    int n_nodes = 0;
    int n_pairs = 0;   // used length of `pairs` (all levels)
    int depth = 0;
    DevBuf<int4> nodes;
    DevBuf<int> pairs;
    DevBuf<unsigned char> node_depth;
};

struct TreeView {
    const int4* nodes;
    const int* pairs;
    const TriRec* rec;  // [nt] per-triangle query records (geom.cuh)
    const float4* cull; // [nt] conservative bounding spheres, single-precision record (geom.cuh make_cull / pack_cull)
    const int* tri;     // [nt][3]
    int root;
    const double* xyz;  // [nv][3]: with rec == nullptr (a mesh whose records were never asked for) the queries build a triangle's
                        // record values from its corners with the same expressions (query.cuh, LAZY)
};

} // namespace msm

struct msmgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t aux_stream = nullptr;      // lazily created second stream: the gather of one subject chunk overlaps the queries of the next
    cudaEvent_t aux_ev[3] = {nullptr, nullptr, nullptr};   // fork / chunk-ready / join (timing disabled)
    int* pinned = nullptr;   // 64 ints of page-locked host memory: small device -> host results inside launch sequences (level totals of the
                             // octree build) arrive by plain DMA instead of the staged, stream-draining path of pageable copies
};

struct msmgpu_octree;
struct msmgpu_mesh {
    msmgpu_ctx* ctx = nullptr;
    int nv = 0, nt = 0;
    msm::DevBuf<double> xyz;   // [nv][3]
    msm::DevBuf<int> tri;      // [nt][3]
    msm::DevBuf<msm::TriRec> rec; // [nt] one 128-byte query record per triangle (gather-free leaf scans)
    msm::DevBuf<double> area_tab;  // [nt] Triangle::area of the CURRENT coordinates (triangle.cpp:47-50), read by the vertex areas
    msm::DevBuf<uint4> qbox;   // [nt] the same box on the octree's depth-18 lattice (octree_build.cu: pack_qbox)
    msm::DevBuf<float4> cull;  // [nt] centre + r^2 of the conservative cull sphere (pack_cull)
    msm::DevBuf<float> feat;   // optional resident payload, vertex-major rows [nv][feat_D] (Mesh::pvalues, mesh.h:44)
    int feat_D = 0;
    bool tables_dirty = false;   // rec / qbox / cull / area_tab not yet computed from xyz (msm::ensure_tables batches that work)
    bool view = false;           // xyz / tri are the caller's device buffers (msmgpu_mesh_create_view_batch)
    bool lazy_rec = false;       // `rec` is filled on first use (msm::ensure_records): a subject of a batch resampling job is queried a few
                                 // 10^4 times, its 128-byte records would be 75 % of the per-triangle table traffic and mostly never read
    std::shared_ptr<msm::DevBuf<unsigned char>> slab;   // the per-triangle tables of a batch of view meshes live in one allocation
    msmgpu_mesh* area_source = nullptr;   // mesh whose geometry the cached Triangle areas belong to (msmgpu_mesh_set_area_source)
    msm::DevBuf<double> tri_area;         // optional explicit cached Triangle::area values [nt] (msmgpu_mesh_set_triangle_areas)
    // the mesh's own octree: built by the first entry point that is handed the mesh without a tree (msm::mesh_tree), kept until the
    // coordinates change (msmgpu_mesh_set_coords). Never set for views of caller buffers, whose contents the library does not track.
    msmgpu_octree* own_tree = nullptr;
    ~msmgpu_mesh();
};

struct msmgpu_octree {
    msmgpu_ctx* ctx = nullptr;   // kept here so destroy never has to look through `mesh`
    std::shared_ptr<msm::Forest> forest;
    int root = 0;
    msmgpu_mesh* mesh = nullptr;
    msm::TreeView view() const;   // records materialised if the mesh deferred them (octree_build.cu)
    msm::TreeView view_lazy() const {   // for kernels that can do without (rec may be nullptr)
        return msm::TreeView{forest->nodes.p, forest->pairs.p, mesh->rec.p, mesh->cull.p, mesh->tri.p, root, mesh->xyz.p};
    }
};

// forward barycentric weights of a batch, produced by the fused resample and consumed by the adaptive weights
struct msmgpu_fwd {
    msmgpu_ctx* ctx = nullptr;
    int S = 0, n = 0;
    bool filled = false;
    msm::DevBuf<int> idx, ne;     // [S][n][3], [S][n]
    msm::DevBuf<double> w;        // [S][n][3]
    std::vector<const msmgpu_octree*> trees;   // the trees the weights were computed in (checked by the consumer)
};

// CSR storage shared by the weight matrices of one batch (one allocation, one set of launches)
struct WeightsStore {
    msm::DevBuf<int> rowptr;   // [batch * n_rows + 1], offsets into col / val
    msm::DevBuf<int> col;      // column = source vertex id, local to its subject
    msm::DevBuf<double> val;
};

struct msmgpu_weights {
    msmgpu_ctx* ctx = nullptr;
    int n_rows = 0, n_cols = 0;
    int64_t nnz = 0;
    std::shared_ptr<WeightsStore> store;
    const int* rowptr = nullptr;   // n_rows + 1 entries (absolute offsets into store->col / val)
    int first = 0;                 // rowptr[0], cached on the host
};

namespace msm {

// ---- launchers implemented in the .cu files -------------------------------------------------
msmgpu_status mesh_refresh_tables(msmgpu_mesh* m);
// octrees of meshes handed over without one: the cached own_tree (built here when missing, all missing ones as one forest);
// views get a fresh tree that `owned` keeps alive for the caller's scope
msmgpu_status mesh_trees(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, msmgpu_octree** out, std::vector<std::unique_ptr<msmgpu_octree>>& owned);
inline msmgpu_status mesh_tree(msmgpu_mesh* m, msmgpu_octree** out, std::vector<std::unique_ptr<msmgpu_octree>>& owned) {
    return mesh_trees(m->ctx, 1, &m, out, owned);
}
msmgpu_status ensure_tables(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes);   // one launch for every dirty mesh
msmgpu_status forest_build(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, std::shared_ptr<Forest>& out, std::vector<int>& roots);

int tuning_get(const char* name, const char* env, int def);   // api.cu: knob registry (environment default, msmgpu_set_tuning)
int query_group_width();
msmgpu_status launch_nearest(const TreeView& t, int n, const double* d_pts, int* d_tri, int* d_vertex, int* d_status, cudaStream_t s);
msmgpu_status launch_bary_weights(const TreeView& t, int n, const double* d_pts, int* d_idx, double* d_w, int* d_ne, int* d_status, cudaStream_t s);
msmgpu_status launch_blend_coords(const TreeView& t, int n, const double* d_pts, const double* d_payload_xyz, double* d_out, int reproject, int* d_status, cudaStream_t s);
// row gather through bulk asynchronous copies (gather.cu), shared by the barycentric maps and the CSR rows of the adaptive weights
struct GatherJob {
    const int* rowptr;   // CSR: n_rows + 1 absolute offsets into col / val; barycentric maps: unused (three slots per row)
    const int* col;      // source row of every entry (barycentric maps: col < 0 = absent entry)
    const double* val;
    const float* in;     // [n_cols][D]
    float* out;          // [n_rows][D]
};
bool gather_bulk_supported(int D);
bool gather_bulk_enabled();
msmgpu_status launch_gather_rows_bulk(const GatherJob* d_jobs, int n_jobs, int n_rows, int D, bool bary, int device, cudaStream_t s, int max_ctas_per_sm = 0);
msmgpu_status ctx_aux(msmgpu_ctx* ctx);   // creates ctx->aux_stream / aux_ev on first use

msmgpu_status launch_gather_channels_f64(int n, int nv, int D, const int* d_vtx, const double* d_in, double* d_out, cudaStream_t s);

struct QueryJob {         // one subject of a batched barycentric-weights launch
    TreeView tree;
    const double* pts;      // [n][3]
    int n;
    int out_off;            // first output slot of this job in the concatenated outputs
    const int* perm;        // optional processing order (order.cu): thread k handles point perm[k]; outputs stay at the point's index
};
msmgpu_status morton_order(const double* d_xyz, int n, DevBuf<int>& perm, cudaStream_t s);
msmgpu_status launch_bary_weights_batch(const QueryJob* d_jobs, int n_jobs, int max_n, int* d_idx, double* d_w, int* d_ne, int* d_status, cudaStream_t s,
                                        bool lazy = false);   // lazy: every job's tree comes without records
msmgpu_status ensure_records(msmgpu_mesh* m);
// jobs that share one tree, one point count and one processing order, the subject on the lanes (query.cu: k_bary_weights_across)
msmgpu_status launch_bary_weights_across(const TreeView& t, const int* d_perm, const double* const* d_pts, const int* d_out_off, int n_jobs, int n,
                                         int* d_idx, double* d_w, int* d_ne, int* d_status, cudaStream_t s);

struct ResampleJob {      // one subject of a batched fused resample
    TreeView tree;
    const float* feat_in;   // [nv][D]
    float* feat_out;        // [n][D]
    int* keep_idx;          // optional [n][3] / [n][3] / [n]: the barycentric weight maps of this subject's targets, kept for
    double* keep_w;         // msmgpu_adaptive_weights_batch_fwd (same values get_barycentric_weights would recompute)
    int* keep_ne;
};

msmgpu_status launch_bary_resample_f32(const ResampleJob* d_jobs, int n_jobs, int n, const double* d_pts, int D, int* d_status, cudaStream_t s);

// Bitonic sort of 16 * E keys held by the 16 lanes of a half-warp (element e = r * 16 + lane16; ascending): exchanges at distance
// >= 16 stay inside the lane, shorter ones are one shuffle. `mask` names the 16 lanes of the group; both halves of a warp may run
// different E (divergent), each with its own mask.
template <int E>
__device__ __forceinline__ void half_warp_bitonic(int (&v)[E], unsigned mask) {
    const int hl = threadIdx.x & 15;
#pragma unroll
    for (int k = 2; k <= 16 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 16) {
                const int rj = j >> 4;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if (r & rj) continue;
                    const bool asc = (((r << 4) | hl) & k) == 0;
                    const int x = v[r], y = v[r | rj];
                    if ((x > y) == asc) { v[r] = y; v[r | rj] = x; }
                }
            } else {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int o = __shfl_xor_sync(mask, v[r], j, 16);
                    const bool asc = (((r << 4) | hl) & k) == 0;
                    const bool lower = (hl & j) == 0;
                    v[r] = (lower == asc) ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

msmgpu_status exclusive_scan_i32(const int* d_in, int* d_out, int n, int* d_total, cudaStream_t s);

// layout changes between the host boundary (channel-major doubles) and device rows
msmgpu_status launch_chmajor_f64_to_rows_f32(int D, int nv, const double* d_in, float* d_out, cudaStream_t s);
msmgpu_status launch_rows_f32_to_chmajor_f64(int D, int nv, const float* d_in, double* d_out, cudaStream_t s);
msmgpu_status launch_chmajor_f32_to_rows_f32(int D, int nv, const float* d_in, float* d_out, cudaStream_t s);
msmgpu_status launch_rows_f32_to_chmajor_f32(int D, int nv, const float* d_in, float* d_out, cudaStream_t s);
msmgpu_status launch_transpose_f64(int rows, int cols, const double* d_in, double* d_out, cudaStream_t s);   // [rows][cols] -> [cols][rows]

// first non-zero entry of d_status[n] -> host code (0 if none); synchronises the stream
msmgpu_status first_error(const int* d_status, size_t n, cudaStream_t s, int* host_code);
msmgpu_status status_to_error(int code);

} // namespace msm
