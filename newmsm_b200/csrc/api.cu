// C ABI (include/msmgpu.h), part 1: context, meshes, octrees, queries, barycentric resampling,
// coordinate blends. Host-pointer entry points stage through stream-ordered device buffers and
// synchronise; `_dev` entry points only enqueue work on the context stream.
#include "common.cuh"

#include <climits>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>

namespace msm {

// Tuning knobs (no effect on results): initialised from the environment on first use, changeable through msmgpu_set_tuning.
static std::mutex g_tuning_mutex;
static std::map<std::string, int>& tuning_map() { static std::map<std::string, int> m; return m; }
int tuning_get(const char* name, const char* env, int def) {
    std::lock_guard<std::mutex> g(g_tuning_mutex);
    auto& m = tuning_map();
    auto it = m.find(name);
    if (it != m.end()) return it->second;
    const char* e = env ? getenv(env) : nullptr;
    const int v = e ? atoi(e) : def;
    m[name] = v;
    return v;
}

static thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launch_count{0};

void set_error(const std::string& msg) { g_last_error = msg; }
msmgpu_status fail(msmgpu_status st, const std::string& msg) {
    g_last_error = msg;
    return st;
}

msmgpu_status ctx_aux(msmgpu_ctx* ctx) {
    if (ctx->aux_stream) return MSMGPU_OK;
    MSM_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : ctx->aux_ev) MSM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return MSMGPU_OK;
}

msmgpu_status status_to_error(int code) {
    switch (code) {
        case MSMGPU_OK: return MSMGPU_OK;
        case MSMGPU_ERR_OUT_OF_BOX: return fail(MSMGPU_ERR_OUT_OF_BOX, "Point is not in the bounding box of the mesh");   // octree.cpp:158
        case MSMGPU_ERR_NO_TRIANGLE: return fail(MSMGPU_ERR_NO_TRIANGLE, "Error in octree. No closest triangle found for the point.");   // octree.cpp:211
        default: return fail((msmgpu_status)code, "query failed");
    }
}

__global__ void k_first_error(const int* __restrict__ st, size_t n, unsigned long long* __restrict__ first) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n && st[i] != 0) atomicMin(first, (unsigned long long)i);
}

msmgpu_status first_error(const int* d_status, size_t n, cudaStream_t s, int* host_code) {
    *host_code = 0;
    if (n == 0) { MSM_CUDA(cudaStreamSynchronize(s)); return MSMGPU_OK; }
    DevBuf<unsigned long long> first;
    MSM_CUDA(first.alloc(1, s));
    MSM_CUDA(cudaMemsetAsync(first.p, 0xff, sizeof(unsigned long long), s));
    k_first_error<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_status, n, first.p);
    MSM_LAUNCH_CHECK();
    unsigned long long h = 0;
    MSM_CUDA(cudaMemcpyAsync(&h, first.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    if (h != ~0ull) {
        MSM_CUDA(cudaMemcpyAsync(host_code, d_status + h, sizeof(int), cudaMemcpyDeviceToHost, s));
        MSM_CUDA(cudaStreamSynchronize(s));
    }
    return MSMGPU_OK;
}

// ---- layout kernels (32x32 shared-memory tiles, coalesced on both sides) ----------------------
template <typename TI, typename TO>
__global__ void k_transpose(int rows, int cols, const TI* __restrict__ in, TO* __restrict__ out) {
    __shared__ TO tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = (TO)in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

template <typename TI, typename TO>
static msmgpu_status transpose(int rows, int cols, const TI* in, TO* out, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return MSMGPU_OK;
    const dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    k_transpose<TI, TO><<<grid, block, 0, s>>>(rows, cols, in, out);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status launch_chmajor_f64_to_rows_f32(int D, int nv, const double* d_in, float* d_out, cudaStream_t s) { return transpose<double, float>(D, nv, d_in, d_out, s); }
msmgpu_status launch_rows_f32_to_chmajor_f64(int D, int nv, const float* d_in, double* d_out, cudaStream_t s) { return transpose<float, double>(nv, D, d_in, d_out, s); }
msmgpu_status launch_chmajor_f32_to_rows_f32(int D, int nv, const float* d_in, float* d_out, cudaStream_t s) { return transpose<float, float>(D, nv, d_in, d_out, s); }
msmgpu_status launch_rows_f32_to_chmajor_f32(int D, int nv, const float* d_in, float* d_out, cudaStream_t s) { return transpose<float, float>(nv, D, d_in, d_out, s); }
msmgpu_status launch_transpose_f64(int rows, int cols, const double* d_in, double* d_out, cudaStream_t s) { return transpose<double, double>(rows, cols, d_in, d_out, s); }

// estimate_rotation_matrix (point.cpp:97-152). Evaluated on the HOST: its acos/sin/cos are the
// only operations on the path whose CUDA implementation is not bit-identical to glibc's, the
// matrices decide nearest-triangle ids downstream, and there are only O(N_cp * L) of them.
bool host_rotation_matrix(const double* ci_, const double* index_, double* R) {
    V3 ci{ci_[0], ci_[1], ci_[2]}, index{index_[0], index_[1], index_[2]};
    ci = vnormalized(ci);
    index = vnormalized(index);
    const double c = vdot(ci, index);
    const double theta = std::acos(c);
    if (theta > M_PI) return false;
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const V3 cr = vnormalized(vcross(ci, index));
    if (std::fabs(1 - c) < kEps) { std::memcpy(R, I, sizeof(I)); return true; }
    if (vnorm(cr) < kEps) { for (int i = 0; i < 9; ++i) R[i] = -I[i]; return true; }
    const double u[9] = {0, -cr.z, cr.y, cr.z, 0, -cr.x, -cr.y, cr.x, 0};
    if (std::fabs(-1 - c) < kEps) {
        const double o[3] = {cr.x, cr.y, cr.z};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[3 * i + j] = 2 * (o[i] * o[j]) - I[3 * i + j];
        return true;
    }
    const double sn = std::sin(theta), omc = 1 - std::cos(theta);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += u[3 * i + k] * u[3 * k + j];
            R[3 * i + j] = (I[3 * i + j] + u[3 * i + j] * sn) + omc * a;
        }
    return true;
}

static msmgpu_status check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(MSMGPU_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    return MSMGPU_OK;
}

template <typename T>
static msmgpu_status upload(DevBuf<T>& b, const T* host, size_t n, cudaStream_t s) {
    MSM_CUDA(b.alloc(n, s));
    if (n) MSM_CUDA(cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return MSMGPU_OK;
}

static msmgpu_status finish_queries(const int* d_status, size_t n, int* host_status, cudaStream_t s) {
    if (host_status) {
        MSM_CUDA(cudaMemcpyAsync(host_status, d_status, n * sizeof(int), cudaMemcpyDeviceToHost, s));
        MSM_CUDA(cudaStreamSynchronize(s));
        return MSMGPU_OK;
    }
    int code = 0;
    MSM_TRY(first_error(d_status, n, s, &code));
    return status_to_error(code);
}

} // namespace msm

using namespace msm;

extern "C" {

const char* msmgpu_last_error(void) { return g_last_error.c_str(); }
const char* msmgpu_version(void) { return "newmsm_b200 0.1 (sm_100a)"; }

/* debugging aid: the CUDA runtime's pending (non-sticky) error, cleared by the call; "" if none */
const char* msmgpu_debug_take_cuda_error(void) {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? "" : cudaGetErrorString(e);
}

msmgpu_status msmgpu_set_tuning(const char* name, int value) {
    if (!name) return fail(MSMGPU_ERR_INVALID, "set_tuning: name is NULL");
    std::lock_guard<std::mutex> g(g_tuning_mutex);
    tuning_map()[name] = value;
    return MSMGPU_OK;
}

unsigned long long msmgpu_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

int msmgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

msmgpu_status msmgpu_ctx_create(int device, void* stream, msmgpu_ctx** out) {
    if (!out) return fail(MSMGPU_ERR_INVALID, "ctx_create: out is NULL");
    *out = nullptr;
    MSM_TRY(check_device());
    MSM_CUDA(cudaSetDevice(device));
    auto* c = new msmgpu_ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail(MSMGPU_ERR_CUDA, cudaGetErrorString(e)); }
        c->own_stream = true;
    }
    // keep freed scratch in the pool: repeated calls (one per registration iteration) reuse it
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (cudaHostAlloc((void**)&c->pinned, 64 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); c->pinned = nullptr; }
    *out = c;
    return MSMGPU_OK;
}

void msmgpu_ctx_destroy(msmgpu_ctx* c) {
    if (!c) return;
    cudaStreamSynchronize(c->stream);
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
    for (cudaEvent_t e : c->aux_ev) if (e) cudaEventDestroy(e);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

msmgpu_status msmgpu_ctx_sync(msmgpu_ctx* c) {
    if (!c) return fail(MSMGPU_ERR_INVALID, "ctx is NULL");
    MSM_CUDA(cudaStreamSynchronize(c->stream));
    return MSMGPU_OK;
}
void* msmgpu_ctx_stream(msmgpu_ctx* c) { return c ? (void*)c->stream : nullptr; }

msmgpu_status msmgpu_device_malloc(msmgpu_ctx* c, size_t bytes, void** out) {
    if (!c || !out) return fail(MSMGPU_ERR_INVALID, "device_malloc: bad arguments");
    *out = nullptr;
    MSM_CUDA(cudaSetDevice(c->device));
    if (bytes) MSM_CUDA(cudaMalloc(out, bytes));
    return MSMGPU_OK;
}
msmgpu_status msmgpu_host_alloc(msmgpu_ctx* c, size_t bytes, void** out) {
    if (!c || !out) return fail(MSMGPU_ERR_INVALID, "host_alloc: bad arguments");
    *out = nullptr;
    MSM_CUDA(cudaSetDevice(c->device));
    if (bytes) MSM_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return MSMGPU_OK;
}
void msmgpu_host_free(msmgpu_ctx* c, void* ptr) {
    if (!c || !ptr) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFreeHost(ptr);
}
void msmgpu_device_free(msmgpu_ctx* c, void* ptr) {
    if (!c || !ptr) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(ptr);
}
msmgpu_status msmgpu_device_download(msmgpu_ctx* c, void* host_dst, const void* dev_src, size_t bytes) {
    if (!c || (bytes && (!host_dst || !dev_src))) return fail(MSMGPU_ERR_INVALID, "device_download: bad arguments");
    MSM_CUDA(cudaSetDevice(c->device));
    MSM_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    MSM_CUDA(cudaStreamSynchronize(c->stream));
    return MSMGPU_OK;
}
msmgpu_status msmgpu_device_copy_peer(msmgpu_ctx* dc, void* dst, msmgpu_ctx* sc, const void* src, size_t bytes) {
    if (!dc || !sc || (bytes && (!dst || !src))) return fail(MSMGPU_ERR_INVALID, "device_copy_peer: bad arguments");
    MSM_CUDA(cudaSetDevice(sc->device));
    MSM_CUDA(cudaStreamSynchronize(sc->stream));   // the source was produced on its own context's stream
    MSM_CUDA(cudaSetDevice(dc->device));
    MSM_CUDA(cudaMemcpyPeerAsync(dst, dc->device, src, sc->device, bytes, dc->stream));
    MSM_CUDA(cudaStreamSynchronize(dc->stream));
    return MSMGPU_OK;
}

static msmgpu_status mesh_create_impl(msmgpu_ctx* ctx, int nv, const double* xyz, int nt, const int32_t* tri, bool dev, msmgpu_mesh** out) {
    if (!ctx || !out || nv <= 0 || nt < 0 || !xyz || (nt > 0 && !tri)) return fail(MSMGPU_ERR_INVALID, "mesh_create: bad arguments");
    *out = nullptr;
    MSM_CUDA(cudaSetDevice(ctx->device));
    auto m = std::unique_ptr<msmgpu_mesh>(new msmgpu_mesh());
    m->ctx = ctx; m->nv = nv; m->nt = nt;
    cudaStream_t s = ctx->stream;
    const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    MSM_CUDA(m->xyz.alloc(3 * (size_t)nv, s));
    MSM_CUDA(m->tri.alloc(3 * (size_t)nt, s));
    MSM_CUDA(m->rec.alloc((size_t)nt, s));
    MSM_CUDA(m->area_tab.alloc((size_t)nt, s));
    MSM_CUDA(m->qbox.alloc((size_t)nt, s));
    MSM_CUDA(m->cull.alloc((size_t)nt, s));
    MSM_CUDA(cudaMemcpyAsync(m->xyz.p, xyz, 3 * (size_t)nv * sizeof(double), kind, s));
    if (nt) MSM_CUDA(cudaMemcpyAsync(m->tri.p, tri, 3 * (size_t)nt * sizeof(int), kind, s));
    if (dev) {
        m->tables_dirty = true;   // computed with the first octree build, one launch for a whole batch of meshes
    } else {
        MSM_TRY(mesh_refresh_tables(m.get()));
        MSM_CUDA(cudaStreamSynchronize(s));   // the host buffers may be released by the caller
    }
    *out = m.release();
    return MSMGPU_OK;
}

msmgpu_status msmgpu_mesh_create(msmgpu_ctx* ctx, int nv, const double* xyz, int nt, const int32_t* tri, msmgpu_mesh** out) {
    return mesh_create_impl(ctx, nv, xyz, nt, tri, false, out);
}
msmgpu_status msmgpu_mesh_create_dev(msmgpu_ctx* ctx, int nv, const double* d_xyz, int nt, const int32_t* d_tri, msmgpu_mesh** out) {
    return mesh_create_impl(ctx, nv, d_xyz, nt, d_tri, true, out);
}

msmgpu_status msmgpu_mesh_create_view_batch(msmgpu_ctx* ctx, int n, int nv, const double* const* d_xyz, int nt, const int32_t* d_tri, msmgpu_mesh** out) {
    if (!ctx || !out || n <= 0 || nv <= 0 || nt <= 0 || !d_xyz || !d_tri) return fail(MSMGPU_ERR_INVALID, "mesh_create_view_batch: bad arguments");
    for (int i = 0; i < n; ++i) {
        out[i] = nullptr;
        if (!d_xyz[i]) return fail(MSMGPU_ERR_INVALID, "mesh_create_view_batch: NULL coordinates");
    }
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // one allocation for the per-triangle tables of the whole batch: [rec | area | qbox | cull] per mesh, each 256-byte aligned
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // the 128-byte records are deferred (lazy_rec): the batch paths query such meshes a few 10^4 times each and build the record
    // values of the few surviving candidates on the fly; any other consumer materialises them on first use (msmgpu_octree::view)
    const bool lazy = tuning_get("lazy_records", "MSMGPU_LAZY_RECORDS", 1) != 0;
    const size_t b_rec = lazy ? 0 : up((size_t)nt * sizeof(TriRec)), b_area = up((size_t)nt * sizeof(double)), b_qbox = up((size_t)nt * sizeof(uint4)),
                 b_cull = up((size_t)nt * sizeof(float4)), per_mesh = b_rec + b_area + b_qbox + b_cull;
    auto slab = std::make_shared<DevBuf<unsigned char>>();
    MSM_CUDA(slab->alloc(per_mesh * (size_t)n, s));
    std::vector<std::unique_ptr<msmgpu_mesh>> made;
    for (int i = 0; i < n; ++i) {
        auto m = std::unique_ptr<msmgpu_mesh>(new msmgpu_mesh());
        m->ctx = ctx; m->nv = nv; m->nt = nt; m->view = true; m->slab = slab;
        unsigned char* base = slab->p + per_mesh * (size_t)i;
        m->xyz.borrow(const_cast<double*>(d_xyz[i]), 3 * (size_t)nv);
        m->tri.borrow(const_cast<int*>(d_tri), 3 * (size_t)nt);
        if (!lazy) m->rec.borrow(reinterpret_cast<TriRec*>(base), (size_t)nt);
        m->lazy_rec = lazy;
        m->area_tab.borrow(reinterpret_cast<double*>(base + b_rec), (size_t)nt);
        m->qbox.borrow(reinterpret_cast<uint4*>(base + b_rec + b_area), (size_t)nt);
        m->cull.borrow(reinterpret_cast<float4*>(base + b_rec + b_area + b_qbox), (size_t)nt);
        m->tables_dirty = true;
        made.push_back(std::move(m));
    }
    for (int i = 0; i < n; ++i) out[i] = made[i].release();
    return MSMGPU_OK;
}

msmgpu_status msmgpu_mesh_set_coords(msmgpu_mesh* m, const double* xyz) {
    if (!m || !xyz) return fail(MSMGPU_ERR_INVALID, "mesh_set_coords: bad arguments");
    if (m->view) return fail(MSMGPU_ERR_INVALID, "mesh_set_coords: the mesh is a view of the caller's buffers");
    MSM_CUDA(cudaSetDevice(m->ctx->device));
    MSM_CUDA(cudaMemcpyAsync(m->xyz.p, xyz, 3 * (size_t)m->nv * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    MSM_TRY(mesh_refresh_tables(m));
    MSM_CUDA(cudaStreamSynchronize(m->ctx->stream));
    delete m->own_tree;   // built for the previous coordinates
    m->own_tree = nullptr;
    return MSMGPU_OK;
}

void msmgpu_mesh_destroy(msmgpu_mesh* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    delete m;
}

msmgpu_status msmgpu_mesh_shape(msmgpu_mesh* m, int* nv, int* nt) {
    if (!m) return fail(MSMGPU_ERR_INVALID, "mesh is NULL");
    if (nv) *nv = m->nv;
    if (nt) *nt = m->nt;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_octree_build_batch(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, msmgpu_octree** out) {
    if (!ctx || n <= 0 || !meshes || !out) return fail(MSMGPU_ERR_INVALID, "octree_build_batch: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    std::shared_ptr<Forest> F;
    std::vector<int> roots;
    MSM_TRY(forest_build(ctx, n, meshes, F, roots));
    for (int i = 0; i < n; ++i) {
        auto* t = new msmgpu_octree();
        t->ctx = ctx;
        t->forest = F;
        t->root = roots[i];
        t->mesh = meshes[i];
        out[i] = t;
    }
    return MSMGPU_OK;
}

msmgpu_status msmgpu_octree_build(msmgpu_mesh* m, msmgpu_octree** out) {
    if (!m || !out) return fail(MSMGPU_ERR_INVALID, "octree_build: bad arguments");
    *out = nullptr;
    return msmgpu_octree_build_batch(m->ctx, 1, &m, out);
}

void msmgpu_octree_destroy(msmgpu_octree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    delete t;
}

// pre-order walk of one tree of the forest on host copies of the node / pair arrays
static msmgpu_status tree_walk(msmgpu_octree* t, const std::function<void(const int4&, int depth, const int* pairs)>& visit) {
    Forest& F = *t->forest;
    cudaStream_t s = F.ctx->stream;
    std::vector<int4> nodes(F.n_nodes);
    std::vector<int> pairs(F.n_pairs);
    MSM_CUDA(cudaMemcpyAsync(nodes.data(), F.nodes.p, nodes.size() * sizeof(int4), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaMemcpyAsync(pairs.data(), F.pairs.p, pairs.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    std::vector<std::pair<int, int>> stack{{t->root, 0}};
    while (!stack.empty()) {
        const auto [g, depth] = stack.back();
        stack.pop_back();
        const int4& nd = nodes[g];
        visit(nd, depth, pairs.data());
        if (nd.x >= 0)
            for (int c = 7; c >= 0; --c) stack.push_back({nd.x + c, depth + 1});
    }
    return MSMGPU_OK;
}

msmgpu_status msmgpu_octree_stats(msmgpu_octree* t, int* n_nodes, int* n_leaf_refs, int* depth) {
    if (!t) return fail(MSMGPU_ERR_INVALID, "octree is NULL");
    int nn = 0, nr = 0, dp = 0;
    MSM_TRY(tree_walk(t, [&](const int4& nd, int d, const int*) { ++nn; nr += nd.z; dp = d > dp ? d : dp; }));
    if (n_nodes) *n_nodes = nn;
    if (n_leaf_refs) *n_leaf_refs = nr;
    if (depth) *depth = dp;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_octree_dump(msmgpu_octree* t, int32_t* kinds, int32_t* counts, int32_t* tris) {
    if (!t || !kinds || !counts || !tris) return fail(MSMGPU_ERR_INVALID, "octree_dump: bad arguments");
    int i = 0, j = 0;
    return tree_walk(t, [&](const int4& nd, int, const int* pairs) {
        kinds[i] = nd.x < 0 ? 1 : 0;
        counts[i] = nd.z;
        ++i;
        for (int k = 0; k < nd.z; ++k) tris[j++] = pairs[nd.y + k];
    });
}

msmgpu_status msmgpu_nearest_triangle_dev(msmgpu_octree* t, int n, const double* d_pts, int32_t* d_tri, int32_t* d_vertex, int32_t* d_status) {
    if (!t || n < 0 || (n > 0 && !d_pts)) return fail(MSMGPU_ERR_INVALID, "nearest_triangle_dev: bad arguments");
    MSM_CUDA(cudaSetDevice(t->mesh->ctx->device));
    return launch_nearest(t->view(), n, d_pts, d_tri, d_vertex, d_status, t->mesh->ctx->stream);
}

msmgpu_status msmgpu_nearest_triangle(msmgpu_octree* t, int n, const double* pts, int32_t* out_tri, int32_t* out_vertex, int32_t* status) {
    if (!t || n < 0 || (n > 0 && !pts)) return fail(MSMGPU_ERR_INVALID, "nearest_triangle: bad arguments");
    if (n == 0) return MSMGPU_OK;
    MSM_CUDA(cudaSetDevice(t->mesh->ctx->device));
    cudaStream_t s = t->mesh->ctx->stream;
    DevBuf<double> d_pts;
    DevBuf<int> d_tri, d_vtx, d_st;
    MSM_TRY(upload(d_pts, pts, 3 * (size_t)n, s));
    MSM_CUDA(d_tri.alloc(n, s));
    MSM_CUDA(d_vtx.alloc(n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_nearest(t->view(), n, d_pts.p, d_tri.p, d_vtx.p, d_st.p, s));
    if (out_tri) MSM_CUDA(cudaMemcpyAsync(out_tri, d_tri.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (out_vertex) MSM_CUDA(cudaMemcpyAsync(out_vertex, d_vtx.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, status, s);
}

msmgpu_status msmgpu_bary_weights_dev(msmgpu_octree* t, int n, const double* d_pts, int32_t* d_idx, double* d_w, int32_t* d_n_entries, int32_t* d_status) {
    if (!t || n < 0 || (n > 0 && (!d_pts || !d_idx || !d_w))) return fail(MSMGPU_ERR_INVALID, "bary_weights_dev: bad arguments");
    MSM_CUDA(cudaSetDevice(t->mesh->ctx->device));
    return launch_bary_weights(t->view(), n, d_pts, d_idx, d_w, d_n_entries, d_status, t->mesh->ctx->stream);
}

msmgpu_status msmgpu_bary_weights(msmgpu_octree* t, int n, const double* pts, int32_t* idx, double* w, int32_t* n_entries) {
    if (!t || n < 0 || (n > 0 && (!pts || !idx || !w))) return fail(MSMGPU_ERR_INVALID, "bary_weights: bad arguments");
    if (n == 0) return MSMGPU_OK;
    MSM_CUDA(cudaSetDevice(t->mesh->ctx->device));
    cudaStream_t s = t->mesh->ctx->stream;
    DevBuf<double> d_pts, d_w;
    DevBuf<int> d_idx, d_ne, d_st;
    MSM_TRY(upload(d_pts, pts, 3 * (size_t)n, s));
    MSM_CUDA(d_idx.alloc(3 * (size_t)n, s));
    MSM_CUDA(d_w.alloc(3 * (size_t)n, s));
    MSM_CUDA(d_ne.alloc(n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_bary_weights(t->view(), n, d_pts.p, d_idx.p, d_w.p, d_ne.p, d_st.p, s));
    MSM_CUDA(cudaMemcpyAsync(idx, d_idx.p, 3 * (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaMemcpyAsync(w, d_w.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (n_entries) MSM_CUDA(cudaMemcpyAsync(n_entries, d_ne.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, nullptr, s);
}

msmgpu_status msmgpu_fwd_create(msmgpu_ctx* ctx, int n_subjects, int n, msmgpu_fwd** out) {
    if (!ctx || n_subjects <= 0 || n <= 0 || !out) return fail(MSMGPU_ERR_INVALID, "fwd_create: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<msmgpu_fwd> f(new msmgpu_fwd());
    f->ctx = ctx; f->S = n_subjects; f->n = n;
    MSM_CUDA(f->idx.alloc(3 * (size_t)n_subjects * n, ctx->stream));
    MSM_CUDA(f->w.alloc(3 * (size_t)n_subjects * n, ctx->stream));
    MSM_CUDA(f->ne.alloc((size_t)n_subjects * n, ctx->stream));
    *out = f.release();
    return MSMGPU_OK;
}

void msmgpu_fwd_destroy(msmgpu_fwd* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    delete f;
}

msmgpu_status msmgpu_bary_resample_batch_f32_dev(msmgpu_ctx* ctx, int n_subjects, msmgpu_octree* const* trees, int n, const double* d_pts,
                                                 int D, const float* const* d_feat_in, float* const* d_feat_out, int32_t* d_status) {
    return msmgpu_bary_resample_batch_f32_dev_keep(ctx, n_subjects, trees, n, d_pts, D, d_feat_in, d_feat_out, d_status, nullptr);
}

msmgpu_status msmgpu_bary_resample_batch_f32_dev_keep(msmgpu_ctx* ctx, int n_subjects, msmgpu_octree* const* trees, int n, const double* d_pts,
                                                      int D, const float* const* d_feat_in, float* const* d_feat_out, int32_t* d_status,
                                                      msmgpu_fwd* keep) {
    if (!ctx || n_subjects <= 0 || !trees || n < 0 || D <= 0 || !d_feat_in || !d_feat_out || (n > 0 && !d_pts))
        return fail(MSMGPU_ERR_INVALID, "bary_resample_batch: bad arguments");
    if (keep && (keep->ctx != ctx || keep->S != n_subjects || keep->n != n)) return fail(MSMGPU_ERR_INVALID, "bary_resample_batch: weight store of another shape");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::vector<ResampleJob> jobs(n_subjects);
    if (keep) { keep->trees.assign(trees, trees + n_subjects); keep->filled = true; }
    for (int i = 0; i < n_subjects; ++i)
        if (!trees[i] || trees[i]->mesh->ctx != ctx) return fail(MSMGPU_ERR_INVALID, "bary_resample_batch: tree from another context");
    if (n == 0) return MSMGPU_OK;
    // Rows of >= 128 bytes: two kernels. (1) the queries + weight maps of all subjects (k_bary_weights_batch) into the forward store
    // (the caller's `keep`, or a scratch one), (2) the bulk-copy row gather (gather.cu) over those maps. The fused kernel below is
    // kept for short rows and as the A/B reference (MSMGPU_GATHER=0).
    bool aligned = true;
    for (int i = 0; i < n_subjects; ++i) aligned = aligned && ((reinterpret_cast<uintptr_t>(d_feat_in[i]) | reinterpret_cast<uintptr_t>(d_feat_out[i])) & 15) == 0;
    // "gather": 1 (default) = two kernels, queries + bulk-copy row gather; 0 = the fused register-path kernel below.
    const int mode = tuning_get("gather", "MSMGPU_GATHER", 1);
    if (aligned && gather_bulk_supported(D) && mode != 0) {
        DevBuf<int> t_idx, t_ne, t_st;
        DevBuf<double> t_w;
        int *p_idx, *p_ne;
        double* p_w;
        const size_t tot = (size_t)n_subjects * n;
        if (keep) { p_idx = keep->idx.p; p_w = keep->w.p; p_ne = keep->ne.p; }
        else {
            MSM_CUDA(t_idx.alloc(3 * tot, s)); MSM_CUDA(t_w.alloc(3 * tot, s)); MSM_CUDA(t_ne.alloc(tot, s));
            p_idx = t_idx.p; p_w = t_w.p; p_ne = t_ne.p;
        }
        int* p_st = d_status;
        if (!p_st) { MSM_CUDA(t_st.alloc(tot, s)); p_st = t_st.p; }
        std::vector<QueryJob> qj(n_subjects);
        std::vector<GatherJob> gj(n_subjects);
        DevBuf<int> perm;
        // forward queries (few points per tree, trees streamed from HBM) measured slightly SLOWER in Morton order (profiles/r2b): opt-in
        if (tuning_get("query_order", "MSMGPU_QUERY_ORDER", 1) >= 2) MSM_TRY(morton_order(d_pts, n, perm, s));
        bool lazy = true;   // every tree without stored records -> the LAZY query kernel; a mixed batch materialises the missing ones
        for (int i = 0; i < n_subjects; ++i) lazy = lazy && !trees[i]->mesh->rec.p;
        for (int i = 0; i < n_subjects; ++i) {
            qj[i] = QueryJob{lazy ? trees[i]->view_lazy() : trees[i]->view(), d_pts, n, (int)((size_t)i * n), perm.p};
            gj[i] = GatherJob{nullptr, p_idx + 3 * (size_t)i * n, p_w + 3 * (size_t)i * n, d_feat_in[i], d_feat_out[i]};
        }
        if (tot > 0x7fffffffull / 3) return fail(MSMGPU_ERR_CAPACITY, "bary_resample_batch: batch too large");
        DevBuf<QueryJob> d_qj;
        DevBuf<GatherJob> d_gj;
        MSM_CUDA(d_qj.alloc(n_subjects, s));
        MSM_CUDA(d_gj.alloc(n_subjects, s));
        MSM_CUDA(cudaMemcpyAsync(d_qj.p, qj.data(), qj.size() * sizeof(QueryJob), cudaMemcpyHostToDevice, s));
        MSM_CUDA(cudaMemcpyAsync(d_gj.p, gj.data(), gj.size() * sizeof(GatherJob), cudaMemcpyHostToDevice, s));
        // Subject chunks: the gather of chunk k runs on the context's second stream while the queries of chunk k+1 run on the first
        // (queries are latency-bound on the octree, the gather is bandwidth-bound on the feature rows). The call stays asynchronous
        // on the context's stream: the second stream is forked from it and joined back before returning.
        // Measured (profiles/r2c_tune_gather_first_version.txt): no gain from 2 - 16 chunks — the query kernel's CTAs occupy every SM
        // before the gather of the previous chunk is admitted — so one chunk is the default.
        int chunks = std::max(1, std::min(tuning_get("bary_chunks", "MSMGPU_BARY_CHUNKS", 1), n_subjects));
        const int cap = tuning_get("gather_ctas_per_sm", "MSMGPU_GATHER_CTAS_PER_SM", 0);
        if (chunks == 1) {
            MSM_TRY(launch_bary_weights_batch(d_qj.p, n_subjects, n, p_idx, p_w, p_ne, p_st, s, lazy));
            return launch_gather_rows_bulk(d_gj.p, n_subjects, n, D, true, ctx->device, s, cap);
        }
        MSM_TRY(ctx_aux(ctx));
        MSM_CUDA(cudaEventRecord(ctx->aux_ev[0], s));                      // fork: the job tables are uploaded, earlier work is done
        MSM_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
        for (int c = 0; c < chunks; ++c) {
            const int b = (int)((long long)n_subjects * c / chunks), e = (int)((long long)n_subjects * (c + 1) / chunks);
            if (e <= b) continue;
            MSM_TRY(launch_bary_weights_batch(d_qj.p + b, e - b, n, p_idx, p_w, p_ne, p_st, s, lazy));
            MSM_CUDA(cudaEventRecord(ctx->aux_ev[1], s));
            MSM_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[1], 0));
            MSM_TRY(launch_gather_rows_bulk(d_gj.p + b, e - b, n, D, true, ctx->device, ctx->aux_stream, cap));
        }
        MSM_CUDA(cudaEventRecord(ctx->aux_ev[2], ctx->aux_stream));         // join
        MSM_CUDA(cudaStreamWaitEvent(s, ctx->aux_ev[2], 0));
        return MSMGPU_OK;
    }
    for (int i = 0; i < n_subjects; ++i) {
        jobs[i] = ResampleJob{trees[i]->view(), d_feat_in[i], d_feat_out[i], keep ? keep->idx.p + 3 * (size_t)i * n : nullptr,
                              keep ? keep->w.p + 3 * (size_t)i * n : nullptr, keep ? keep->ne.p + (size_t)i * n : nullptr};
    }
    DevBuf<ResampleJob> d_jobs;
    MSM_CUDA(d_jobs.alloc(n_subjects, s));
    // pageable source: the runtime stages it before returning, so `jobs` may go out of scope
    MSM_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(ResampleJob), cudaMemcpyHostToDevice, s));
    return launch_bary_resample_f32(d_jobs.p, n_subjects, n, d_pts, D, d_status, s);
}

msmgpu_status msmgpu_fwd_apply_batch_f32_dev(msmgpu_ctx* ctx, const msmgpu_fwd* fwd, int D, const float* const* d_feat_in, float* const* d_feat_out) {
    if (!ctx || !fwd || fwd->ctx != ctx || !fwd->filled || D <= 0 || !d_feat_in || !d_feat_out) return fail(MSMGPU_ERR_INVALID, "fwd_apply_batch: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int S = fwd->S, n = fwd->n;
    bool aligned = true;
    for (int i = 0; i < S; ++i) aligned = aligned && ((reinterpret_cast<uintptr_t>(d_feat_in[i]) | reinterpret_cast<uintptr_t>(d_feat_out[i])) & 15) == 0;
    if (!(aligned && gather_bulk_supported(D))) return fail(MSMGPU_ERR_INVALID, "fwd_apply_batch: rows must be 16-byte aligned multiples of 16 bytes, 128 B .. 2 KB");
    std::vector<GatherJob> gj(S);
    for (int i = 0; i < S; ++i) gj[i] = GatherJob{nullptr, fwd->idx.p + 3 * (size_t)i * n, fwd->w.p + 3 * (size_t)i * n, d_feat_in[i], d_feat_out[i]};
    DevBuf<GatherJob> d_gj;
    MSM_CUDA(d_gj.alloc(S, s));
    MSM_CUDA(cudaMemcpyAsync(d_gj.p, gj.data(), gj.size() * sizeof(GatherJob), cudaMemcpyHostToDevice, s));
    return launch_gather_rows_bulk(d_gj.p, S, n, D, true, ctx->device, s);
}

msmgpu_status msmgpu_bary_resample_f32_dev(msmgpu_octree* t, int n, const double* d_pts, int D, const float* d_feat_in, float* d_feat_out, int32_t* d_status) {
    if (!t) return fail(MSMGPU_ERR_INVALID, "bary_resample_f32_dev: tree is NULL");
    return msmgpu_bary_resample_batch_f32_dev(t->mesh->ctx, 1, &t, n, d_pts, D, &d_feat_in, &d_feat_out, d_status);
}

msmgpu_status msmgpu_bary_resample(msmgpu_mesh* in_mesh, int n, const double* pts, int D, const double* feat_in, double* feat_out) {
    if (!in_mesh || n <= 0 || D <= 0 || !pts || !feat_in || !feat_out) return fail(MSMGPU_ERR_INVALID, "bary_resample: bad arguments");
    msmgpu_ctx* ctx = in_mesh->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    msmgpu_octree* t = nullptr;
    std::vector<std::unique_ptr<msmgpu_octree>> guard;
    MSM_TRY(mesh_tree(in_mesh, &t, guard));
    const int nv = in_mesh->nv;
    DevBuf<double> d_pts, d_cm_in, d_cm_out;
    DevBuf<float> d_rows_in, d_rows_out;
    DevBuf<int> d_st;
    MSM_TRY(upload(d_pts, pts, 3 * (size_t)n, s));
    MSM_TRY(upload(d_cm_in, feat_in, (size_t)D * nv, s));
    MSM_CUDA(d_rows_in.alloc((size_t)D * nv, s));
    MSM_CUDA(d_rows_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_cm_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_chmajor_f64_to_rows_f32(D, nv, d_cm_in.p, d_rows_in.p, s));
    MSM_TRY(msmgpu_bary_resample_f32_dev(t, n, d_pts.p, D, d_rows_in.p, d_rows_out.p, d_st.p));
    MSM_TRY(launch_rows_f32_to_chmajor_f64(D, n, d_rows_out.p, d_cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, d_cm_out.p, (size_t)D * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, nullptr, s);
}

// FP32 payload on host buffers (channel-major floats, what GIFTI stores): Octree(in) is supplied by the caller
msmgpu_status msmgpu_bary_resample_f32(msmgpu_octree* t, int n, const double* pts, int D, const float* feat_in, float* feat_out) {
    if (!t || n <= 0 || D <= 0 || !pts || !feat_in || !feat_out) return fail(MSMGPU_ERR_INVALID, "bary_resample_f32: bad arguments");
    msmgpu_ctx* ctx = t->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int nv = t->mesh->nv;
    DevBuf<double> d_pts;
    DevBuf<float> d_cm_in, d_rows_in, d_rows_out, d_cm_out;
    DevBuf<int> d_st;
    MSM_TRY(upload(d_pts, pts, 3 * (size_t)n, s));
    MSM_TRY(upload(d_cm_in, feat_in, (size_t)D * nv, s));
    MSM_CUDA(d_rows_in.alloc((size_t)D * nv, s));
    MSM_CUDA(d_rows_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_cm_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_chmajor_f32_to_rows_f32(D, nv, d_cm_in.p, d_rows_in.p, s));
    MSM_TRY(msmgpu_bary_resample_f32_dev(t, n, d_pts.p, D, d_rows_in.p, d_rows_out.p, d_st.p));
    MSM_TRY(launch_rows_f32_to_chmajor_f32(D, n, d_rows_out.p, d_cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, d_cm_out.p, (size_t)D * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, nullptr, s);
}

// Mesh::pvalues as a device-resident payload (mesh.h:44; set_pvalues, mesh.cpp:206): one upload serves every later resample
msmgpu_status msmgpu_mesh_set_features_f32(msmgpu_mesh* m, int D, const float* feat_cm) {
    if (!m || D <= 0 || !feat_cm) return fail(MSMGPU_ERR_INVALID, "mesh_set_features_f32: bad arguments");
    MSM_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t s = m->ctx->stream;
    DevBuf<float> cm;
    MSM_TRY(upload(cm, feat_cm, (size_t)D * m->nv, s));
    MSM_CUDA(m->feat.alloc((size_t)D * m->nv, s));
    MSM_TRY(launch_chmajor_f32_to_rows_f32(D, m->nv, cm.p, m->feat.p, s));
    m->feat_D = D;
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

msmgpu_status msmgpu_mesh_bary_resample_f32(msmgpu_octree* t, int n, const double* pts, float* feat_out) {
    if (!t || n <= 0 || !pts || !feat_out) return fail(MSMGPU_ERR_INVALID, "mesh_bary_resample_f32: bad arguments");
    if (t->mesh->feat_D <= 0) return fail(MSMGPU_ERR_INVALID, "mesh_bary_resample_f32: the mesh has no resident features (msmgpu_mesh_set_features_f32)");
    msmgpu_ctx* ctx = t->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int D = t->mesh->feat_D;
    DevBuf<double> d_pts;
    DevBuf<float> d_rows_out, d_cm_out;
    DevBuf<int> d_st;
    MSM_TRY(upload(d_pts, pts, 3 * (size_t)n, s));
    MSM_CUDA(d_rows_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_cm_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(msmgpu_bary_resample_f32_dev(t, n, d_pts.p, D, t->mesh->feat.p, d_rows_out.p, d_st.p));
    MSM_TRY(launch_rows_f32_to_chmajor_f32(D, n, d_rows_out.p, d_cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, d_cm_out.p, (size_t)D * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, nullptr, s);
}

static msmgpu_status blend_impl(msmgpu_mesh* mesh, const double* payload, int n, const double* q, double* out, int reproject) {
    if (!mesh || !payload || n <= 0 || !q || !out) return fail(MSMGPU_ERR_INVALID, "blend: bad arguments");
    msmgpu_ctx* ctx = mesh->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    msmgpu_octree* t = nullptr;
    std::vector<std::unique_ptr<msmgpu_octree>> guard;
    MSM_TRY(mesh_tree(mesh, &t, guard));
    DevBuf<double> d_q, d_pay, d_out;
    DevBuf<int> d_st;
    MSM_TRY(upload(d_q, q, 3 * (size_t)n, s));
    MSM_TRY(upload(d_pay, payload, 3 * (size_t)mesh->nv, s));
    MSM_CUDA(d_out.alloc(3 * (size_t)n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_blend_coords(t->view(), n, d_q.p, d_pay.p, d_out.p, reproject, d_st.p, s));
    MSM_CUDA(cudaMemcpyAsync(out, d_out.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, nullptr, s);
}

msmgpu_status msmgpu_sphere_project_warp(msmgpu_mesh* from_mesh, const double* to_xyz, int n, const double* sphere_xyz, double* out_xyz) {
    return blend_impl(from_mesh, to_xyz, n, sphere_xyz, out_xyz, 1);
}
msmgpu_status msmgpu_surface_resample(msmgpu_mesh* sph_mesh, const double* anat_xyz, int n, const double* low_xyz, double* out_xyz) {
    return blend_impl(sph_mesh, anat_xyz, n, low_xyz, out_xyz, 0);
}

msmgpu_status msmgpu_nn_resample(msmgpu_mesh* in_mesh, int n, const double* low_xyz, int D, const double* feat_in, double* feat_out) {
    if (!in_mesh || n <= 0 || D <= 0 || !low_xyz || !feat_in || !feat_out) return fail(MSMGPU_ERR_INVALID, "nn_resample: bad arguments");
    msmgpu_ctx* ctx = in_mesh->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    msmgpu_octree* t = nullptr;
    std::vector<std::unique_ptr<msmgpu_octree>> guard;
    MSM_TRY(mesh_tree(in_mesh, &t, guard));
    DevBuf<double> d_q, d_in, d_out;
    DevBuf<int> d_vtx, d_st;
    MSM_TRY(upload(d_q, low_xyz, 3 * (size_t)n, s));
    MSM_TRY(upload(d_in, feat_in, (size_t)D * in_mesh->nv, s));
    MSM_CUDA(d_out.alloc((size_t)D * n, s));
    MSM_CUDA(d_vtx.alloc(n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_nearest(t->view(), n, d_q.p, nullptr, d_vtx.p, d_st.p, s));
    MSM_TRY(launch_gather_channels_f64(n, in_mesh->nv, D, d_vtx.p, d_in.p, d_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, d_out.p, (size_t)D * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    return finish_queries(d_st.p, n, nullptr, s);
}

// nearest_neighbour_interpolation with an exclusion mask (resampler.cpp:232-258): a target whose closest source vertex is masked out
// keeps 0 in every channel and in the new mask; the others copy the vertex's values and its mask value
msmgpu_status msmgpu_nn_resample_excl(msmgpu_mesh* in_mesh, int n, const double* low_xyz, int D, const double* feat_in, const double* excl_in,
                                      double* feat_out, double* excl_out) {
    if (!in_mesh || n <= 0 || D <= 0 || !low_xyz || !feat_in || !excl_in || !feat_out || !excl_out) return fail(MSMGPU_ERR_INVALID, "nn_resample_excl: bad arguments");
    msmgpu_ctx* ctx = in_mesh->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    msmgpu_octree* t = nullptr;
    std::vector<std::unique_ptr<msmgpu_octree>> guard;
    MSM_TRY(mesh_tree(in_mesh, &t, guard));
    DevBuf<double> d_q;
    DevBuf<int> d_vtx, d_st;
    MSM_TRY(upload(d_q, low_xyz, 3 * (size_t)n, s));
    MSM_CUDA(d_vtx.alloc(n, s));
    MSM_CUDA(d_st.alloc(n, s));
    MSM_TRY(launch_nearest(t->view(), n, d_q.p, nullptr, d_vtx.p, d_st.p, s));
    std::vector<int> vtx((size_t)n);
    MSM_CUDA(cudaMemcpyAsync(vtx.data(), d_vtx.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_TRY(finish_queries(d_st.p, n, nullptr, s));
    const int nv = in_mesh->nv;
    for (int i = 0; i < n; ++i) {
        const int cv = vtx[i];
        const bool on = cv >= 0 && excl_in[cv] != 0;
        excl_out[i] = on ? excl_in[cv] : 0.0;
        for (int d = 0; d < D; ++d) feat_out[(size_t)d * n + i] = on ? feat_in[(size_t)d * nv + cv] : 0.0;
    }
    return MSMGPU_OK;
}

msmgpu_status msmgpu_rotation_matrices(msmgpu_ctx*, int n, const double* ci, const double* index, double* R) {
    if (n < 0 || (n > 0 && (!ci || !index || !R))) return fail(MSMGPU_ERR_INVALID, "rotation_matrices: bad arguments");
    bool ok = true;
#pragma omp parallel for reduction(&& : ok)
    for (int i = 0; i < n; ++i) ok = host_rotation_matrix(ci + 3 * (size_t)i, index + 3 * (size_t)i, R + 9 * (size_t)i) && ok;
    if (!ok) return fail(MSMGPU_ERR_INVALID, "rotation angle is greater than 90 degrees");   // point.cpp:109-110
    return MSMGPU_OK;
}

} // extern "C"
