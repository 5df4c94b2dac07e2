// Adaptive-barycentric resampling weights (Workbench ADAP_BARY_AREA as implemented by
// Resampler::get_adaptive_barycentric_weights, msm-newresampler/src/resampler.cpp:72-140, without
// an exclusion mask), vertex areas (mesh.cpp:1275-1283) and the CSR apply
// (resampler.cpp:40-52).
//
// The reference builds std::map rows and accumulates in map order. Every floating-point sum in
// it is a sequential sum over a SHORT list whose order is "ascending integer key", so the device
// version is: emit (key, id, value) triples in parallel, bucket them by key (count -> scan ->
// atomic fill), sort each short bucket by id, and let one thread walk its bucket in order.
// That makes the result independent of thread scheduling and equal, bit for bit, to the
// single-threaded reference (whose own OpenMP version races on `correction`).
#include "common.cuh"

#include <algorithm>
#include <climits>
#include <type_traits>

namespace msm {

// ------------------------------------------------------------------------------------------
// bucket-by-key machinery
// ------------------------------------------------------------------------------------------
struct Buckets {
    DevBuf<int> ptr;      // [nkeys+1]
    DevBuf<int> id;       // [total]
    DevBuf<double> val;   // [total]
    int total = 0;
};

// segment s with off[s] <= i < off[s+1]  (off has n_seg + 1 entries)
__device__ __forceinline__ int find_segment(const int* __restrict__ off, int n_seg, int i) {
    int lo = 0, hi = n_seg;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Emitters: n_src() sources; load(i) gathers what the source's triples share, then count / key / id / val per triple.
// All of them cover a BATCH of subjects: vertex keys of subject s are shifted by in_off[s], target keys by s * n_low.
struct EmitVertexTriangles {   // key = vertex, id = triangle, val = cached triangle area (triangle.cpp:47-50)
    const int* const* tri; const TriRec* const* rec; const double* const* area; const int* tri_off; const int* key_off; int S; int total;
    struct Src { int t, koff; const int* tri; double area; };
    __device__ int n_src() const { return total; }
    __device__ Src load(int i) const {
        const int s = S == 1 ? 0 : find_segment(tri_off, S, i);
        const int t = i - tri_off[s];
        if (area[s]) return Src{t, key_off[s], tri[s] + 3 * (size_t)t, area[s][t]};   // the caller's cached values
        const double* v = rec[s][t].v;
        return Src{t, key_off[s], tri[s] + 3 * (size_t)t, tri_area_cached(V3{v[0], v[1], v[2]}, V3{v[3], v[4], v[5]}, V3{v[6], v[7], v[8]})};
    }
    __device__ int count(const Src&) const { return 3; }
    __device__ int key(const Src& x, int j) const { return x.koff + x.tri[j]; }
    __device__ int id(const Src& x, int) const { return x.t; }
    __device__ double val(const Src& x, int) const { return x.area; }
};
struct EmitReverse {   // resampler.cpp:91-95: reverse_reorder[target][source] = weight
    const int* ridx; const double* rw; const int* rne; const int* in_off; int S; int n_low; int total;
    struct Src { int o, local, kbase; };
    __device__ int n_src() const { return total; }
    __device__ Src load(int o) const {
        const int s = S == 1 ? 0 : find_segment(in_off, S, o);
        return Src{o, o - in_off[s], s * n_low};
    }
    __device__ int count(const Src& x) const { return rne[x.o]; }
    __device__ int key(const Src& x, int j) const { return x.kbase + ridx[3 * (size_t)x.o + j]; }
    __device__ int id(const Src& x, int) const { return x.local; }
    __device__ double val(const Src& x, int j) const { return rw[3 * (size_t)x.o + j]; }
};
struct EmitCsrColumns {   // key = column (source vertex), id = row (target), val = stored value; rows taken from the FORWARD lists only
    const int* rowptr; const int* col; const double* v; const int* in_off; const int* fne; const int* rr_ptr; int n_low; int total;
    struct Src { int r, b, e, koff; };
    __device__ int n_src() const { return total; }
    __device__ Src load(int r) const {
        const bool reverse_row = rr_ptr[r + 1] - rr_ptr[r] > fne[r];      // such rows are summed straight from the reverse maps (k_area_ratio)
        const int b = rowptr[r];
        return Src{r, b, reverse_row ? b : rowptr[r + 1], in_off[r / n_low]};
    }
    __device__ int count(const Src& x) const { return x.e - x.b; }
    __device__ int key(const Src& x, int j) const { return x.koff + col[x.b + j]; }
    __device__ int id(const Src& x, int) const { return x.r; }
    __device__ double val(const Src& x, int j) const { return v[x.b + j]; }
};

template <class E>
__global__ void k_bucket_count(E e, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n_src()) return;
    const typename E::Src x = e.load(i);
    const int c = e.count(x);
    for (int j = 0; j < c; ++j) atomicAdd(cnt + e.key(x, j), 1);
}
template <class E>
__global__ void k_bucket_fill(E e, const int* __restrict__ ptr, int* __restrict__ cursor, int* __restrict__ id, double* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n_src()) return;
    const typename E::Src x = e.load(i);
    const int c = e.count(x);
    for (int j = 0; j < c; ++j) {
        const int k = e.key(x, j);
        const int pos = ptr[k] + atomicAdd(cursor + k, 1);
        id[pos] = e.id(x, j);
        val[pos] = e.val(x, j);
    }
}
// Sort each bucket by id. Buckets hold a handful of entries (vertex valence, or the ~3 N_s / N_t sources that fall into one
// target's triangles, ~15): buckets of up to 32 entries are sorted in REGISTERS by a bitonic network on packed keys
// (id << 5 | position; ids below 2^26), the values are then moved out of place in the sorted order (val_in -> val_out, no hazard,
// nothing spilled to local memory). Longer buckets (or larger ids) are copied and sorted by insertion in place.
template <int N>
__device__ __forceinline__ void bitonic_sort_keys(int (&key)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const int a = key[i], b = key[l];
                    const bool sw = up ? a > b : a < b;
                    key[i] = sw ? b : a;
                    key[l] = sw ? a : b;
                }
            }
        }
    }
}
template <int N>
__device__ __forceinline__ void sort_bucket_regs(int b, int len, int* __restrict__ id, const double* __restrict__ val_in, double* __restrict__ val_out) {
    int key[N];
#pragma unroll
    for (int i = 0; i < N; ++i) key[i] = i < len ? ((id[b + i] << 5) | i) : INT_MAX;
    bitonic_sort_keys<N>(key);
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (i < len) {
            id[b + i] = key[i] >> 5;
            val_out[b + i] = val_in[b + (key[i] & 31)];
        }
}
__global__ void __launch_bounds__(256) k_bucket_sort(int nkeys, const int* __restrict__ ptr, int* __restrict__ id, const double* __restrict__ val_in,
                                                     double* __restrict__ val_out, int packable) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int b = ptr[k], e = ptr[k + 1], len = e - b;
    if (len <= 0) return;
    if (len == 1) { val_out[b] = val_in[b]; return; }
    if (packable && len <= 32) {
        if (len <= 4) sort_bucket_regs<4>(b, len, id, val_in, val_out);
        else if (len <= 8) sort_bucket_regs<8>(b, len, id, val_in, val_out);
        else if (len <= 16) sort_bucket_regs<16>(b, len, id, val_in, val_out);
        else sort_bucket_regs<32>(b, len, id, val_in, val_out);
        return;
    }
    for (int i = b; i < e; ++i) val_out[i] = val_in[i];
    for (int i = b + 1; i < e; ++i) {
        const int ki = id[i];
        const double vi = val_out[i];
        int j = i - 1;
        while (j >= b && id[j] > ki) { id[j + 1] = id[j]; val_out[j + 1] = val_out[j]; --j; }
        id[j + 1] = ki; val_out[j + 1] = vi;
    }
}

// `capacity` >= the number of triples that will be emitted (an upper bound avoids a host round trip)
template <class E>
static msmgpu_status bucketize(const E& e, int n_src, int nkeys, size_t capacity, Buckets& B, cudaStream_t s, int max_id = INT_MAX) {
    DevBuf<int> cnt, cursor;
    DevBuf<double> unsorted;   // values in arrival order; B.val receives them in sorted order
    MSM_CUDA(cnt.alloc(nkeys, s));
    MSM_CUDA(cursor.alloc(nkeys, s));
    MSM_CUDA(B.ptr.alloc((size_t)nkeys + 1, s));
    MSM_CUDA(B.id.alloc(capacity, s));
    MSM_CUDA(B.val.alloc(capacity, s));
    MSM_CUDA(cudaMemsetAsync(cnt.p, 0, nkeys * sizeof(int), s));
    MSM_CUDA(cudaMemsetAsync(cursor.p, 0, nkeys * sizeof(int), s));
    if (n_src <= 0 || nkeys <= 0) {
        MSM_CUDA(cudaMemsetAsync(B.ptr.p, 0, ((size_t)nkeys + 1) * sizeof(int), s));
        return MSMGPU_OK;
    }
    const unsigned gs = (unsigned)((n_src + 255) / 256), gk = (unsigned)((nkeys + 255) / 256);
    k_bucket_count<E><<<gs, 256, 0, s>>>(e, cnt.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(cnt.p, B.ptr.p, nkeys, B.ptr.p + nkeys, s));
    MSM_CUDA(unsorted.alloc(capacity, s));
    k_bucket_fill<E><<<gs, 256, 0, s>>>(e, B.ptr.p, cursor.p, B.id.p, unsorted.p);
    MSM_LAUNCH_CHECK();
    k_bucket_sort<<<gk, 256, 0, s>>>(nkeys, B.ptr.p, B.id.p, unsorted.p, B.val.p, max_id < (1 << 26) ? 1 : 0);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// vertex areas: mean of the adjacent triangles' areas, summed in push order = ascending triangle id
// (mesh.cpp:112-118, 1275-1283)
// ------------------------------------------------------------------------------------------
__global__ void k_bucket_mean(int nkeys, const int* __restrict__ ptr, const double* __restrict__ val, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int b = ptr[k], e = ptr[k + 1];
    double sum = 0.0;
    for (int i = b; i < e; ++i) sum += val[i];
    out[k] = sum / (double)(e - b);
}
__global__ void k_bucket_sum(int nkeys, const int* __restrict__ ptr, const double* __restrict__ val, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    double sum = 0.0;
    for (int i = ptr[k]; i < ptr[k + 1]; ++i) sum += val[i];
    out[k] = sum;
}

// Fast path for the vertex areas: a vertex of a surface mesh has a handful of incident triangles, so instead of the generic
// count / scan / fill / sort pipeline each vertex owns a fixed row of kIncSlots slots. One pass over the triangles stores the
// (cached) area per triangle and appends the triangle id to its three vertices' rows (atomic cursor, arrival order); a second
// pass sorts each row by triangle id in registers and sums the areas in that order = mesh.cpp:1275-1283 over tIDbegin..tIDend,
// which push_triangle fills in ascending triangle id (mesh.cpp:112-118). A vertex with more than kIncSlots triangles raises a flag
// and the batch falls back to the generic buckets.
constexpr int kIncSlots = 8;
__global__ void k_incidence_fill(EmitVertexTriangles e, int* __restrict__ cnt, int* __restrict__ inc, double* __restrict__ tri_area,
                                 int* __restrict__ overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n_src()) return;
    const EmitVertexTriangles::Src x = e.load(i);
    tri_area[i] = x.area;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int v = e.key(x, j);
        const int slot = atomicAdd(cnt + v, 1);
        if (slot < kIncSlots) inc[(size_t)v * kIncSlots + slot] = i;   // global triangle index: ascending within a mesh like the local id
        else *overflow = 1;
    }
}
__global__ void k_incidence_mean(int nkeys, const int* __restrict__ cnt, const int* __restrict__ inc, const double* __restrict__ tri_area,
                                 double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nkeys) return;
    const int n = min(cnt[v], kIncSlots);
    const int4 a = reinterpret_cast<const int4*>(inc)[2 * (size_t)v], b = reinterpret_cast<const int4*>(inc)[2 * (size_t)v + 1];
    int id[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i >= n) id[i] = INT_MAX;
#pragma unroll
    for (int round = 0; round < 8; ++round) {
#pragma unroll
        for (int i = round & 1; i + 1 < 8; i += 2)
            if (id[i] > id[i + 1]) { const int t = id[i]; id[i] = id[i + 1]; id[i + 1] = t; }
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < n) sum += __ldg(tri_area + id[i]);
    out[v] = sum / (double)n;
}

// Subjects that share one topology (msmgpu_mesh_create_view_batch) share the vertex -> triangles incidence: it is built once, sorted
// once, and every subject only gathers its own cached areas through it (same ascending-id sum as k_incidence_mean).
__global__ void k_incidence_sort(int nkeys, const int* __restrict__ cnt, int* __restrict__ inc) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nkeys) return;
    const int n = min(cnt[v], kIncSlots);
    int4* row = reinterpret_cast<int4*>(inc) + 2 * (size_t)v;
    const int4 a = row[0], b = row[1];
    int id[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i >= n) id[i] = INT_MAX;
#pragma unroll
    for (int round = 0; round < 8; ++round) {
#pragma unroll
        for (int i = round & 1; i + 1 < 8; i += 2)
            if (id[i] > id[i + 1]) { const int t = id[i]; id[i] = id[i + 1]; id[i + 1] = t; }
    }
    row[0] = make_int4(id[0], id[1], id[2], id[3]);
    row[1] = make_int4(id[4], id[5], id[6], id[7]);
}
__global__ void k_incidence_mean_shared(int nv, const int* __restrict__ cnt, const int* __restrict__ inc, const double* const* __restrict__ area,
                                        const int* __restrict__ key_off, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int n = min(__ldg(cnt + v), kIncSlots);
    const int4 a = __ldg(reinterpret_cast<const int4*>(inc) + 2 * (size_t)v), b = __ldg(reinterpret_cast<const int4*>(inc) + 2 * (size_t)v + 1);
    const int id[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const double* __restrict__ ar = area[blockIdx.y];
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < n) sum += __ldg(ar + id[i]);
    out[key_off[blockIdx.y] + v] = sum / (double)n;
}

// vertex areas of a batch of meshes, concatenated: d_out[key_off[s] + v]
static msmgpu_status vertex_areas_batch(msmgpu_ctx* ctx, int S, msmgpu_mesh* const* meshes, const std::vector<int>& key_off, double* d_out) {
    cudaStream_t s = ctx->stream;
    // Triangle::area is cached when a Triangle is constructed and survives Mesh::set_coord (triangle.cpp:31,39): a mesh that was
    // copied and then moved keeps the areas of the geometry it was copied from -> read the records of that mesh instead
    std::vector<msmgpu_mesh*> src(S);
    for (int i = 0; i < S; ++i) src[i] = meshes[i]->area_source ? meshes[i]->area_source : meshes[i];
    meshes = src.data();
    MSM_TRY(ensure_tables(ctx, S, meshes));
    std::vector<const int*> h_tri(S);
    std::vector<const TriRec*> h_rec(S);
    std::vector<const double*> h_area(S);
    std::vector<int> h_toff(S + 1, 0);
    for (int i = 0; i < S; ++i) {
        h_tri[i] = meshes[i]->tri.p;
        h_rec[i] = meshes[i]->rec.p;
        h_area[i] = meshes[i]->tri_area.p ? meshes[i]->tri_area.p : meshes[i]->area_tab.p;   // explicit cached values, else the table of the current geometry
        if (!h_area[i] && !meshes[i]->rec.p) { MSM_TRY(ensure_records(meshes[i])); h_rec[i] = meshes[i]->rec.p; }
        h_toff[i + 1] = h_toff[i] + meshes[i]->nt;
    }
    DevBuf<const int*> d_tri;
    DevBuf<const TriRec*> d_rec;
    DevBuf<const double*> d_area;
    DevBuf<int> d_toff, d_koff;
    MSM_CUDA(d_tri.alloc(S, s));
    MSM_CUDA(d_rec.alloc(S, s));
    MSM_CUDA(d_area.alloc(S, s));
    MSM_CUDA(cudaMemcpyAsync(d_area.p, h_area.data(), S * sizeof(double*), cudaMemcpyHostToDevice, s));
    MSM_CUDA(d_toff.alloc(S + 1, s));
    MSM_CUDA(d_koff.alloc(S + 1, s));
    MSM_CUDA(cudaMemcpyAsync(d_tri.p, h_tri.data(), S * sizeof(int*), cudaMemcpyHostToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(d_rec.p, h_rec.data(), S * sizeof(TriRec*), cudaMemcpyHostToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(d_toff.p, h_toff.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(d_koff.p, key_off.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    const int total_t = h_toff[S], nkeys = key_off[S];
    const EmitVertexTriangles emit{d_tri.p, d_rec.p, d_area.p, d_toff.p, d_koff.p, S, total_t};
    static const bool fast = getenv("MSMGPU_GENERIC_VERTEX_AREAS") == nullptr;
    bool shared = fast && S > 1 && meshes[0]->nt > 0 && meshes[0]->nv > 0 && getenv("MSMGPU_NO_SHARED_INCIDENCE") == nullptr;
    for (int i = 0; shared && i < S; ++i)
        shared = h_tri[i] == h_tri[0] && meshes[i]->nt == meshes[0]->nt && meshes[i]->nv == meshes[0]->nv && h_area[i] != nullptr &&
                 key_off[i + 1] - key_off[i] == meshes[0]->nv;
    if (shared) {
        const int nv = meshes[0]->nv, nt = meshes[0]->nt;
        const EmitVertexTriangles one{d_tri.p, d_rec.p, d_area.p, d_toff.p, d_toff.p /* key offset 0 */, 1, nt};
        DevBuf<int> cnt, inc, ovf;
        DevBuf<double> scratch;
        MSM_CUDA(cnt.alloc((size_t)nv, s));
        MSM_CUDA(inc.alloc((size_t)nv * kIncSlots, s));
        MSM_CUDA(ovf.alloc(1, s));
        MSM_CUDA(scratch.alloc((size_t)nt, s));
        MSM_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nv * sizeof(int), s));
        MSM_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), s));
        k_incidence_fill<<<(nt + 255) / 256, 256, 0, s>>>(one, cnt.p, inc.p, scratch.p, ovf.p);
        MSM_LAUNCH_CHECK();
        k_incidence_sort<<<(nv + 255) / 256, 256, 0, s>>>(nv, cnt.p, inc.p);
        MSM_LAUNCH_CHECK();
        k_incidence_mean_shared<<<dim3((unsigned)((nv + 255) / 256), (unsigned)S), 256, 0, s>>>(nv, cnt.p, inc.p, d_area.p, d_koff.p, d_out);
        MSM_LAUNCH_CHECK();
        int h_ovf = 0;
        MSM_CUDA(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        MSM_CUDA(cudaStreamSynchronize(s));
        if (!h_ovf) return MSMGPU_OK;      // otherwise some vertex has more than kIncSlots triangles: the per-mesh paths below
    }
    if (fast && total_t > 0 && nkeys > 0) {
        DevBuf<int> cnt, inc, ovf;
        DevBuf<double> tarea;
        MSM_CUDA(cnt.alloc((size_t)nkeys, s));
        MSM_CUDA(inc.alloc((size_t)nkeys * kIncSlots, s));
        MSM_CUDA(ovf.alloc(1, s));
        MSM_CUDA(tarea.alloc((size_t)total_t, s));
        MSM_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nkeys * sizeof(int), s));
        MSM_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), s));
        k_incidence_fill<<<(total_t + 255) / 256, 256, 0, s>>>(emit, cnt.p, inc.p, tarea.p, ovf.p);
        MSM_LAUNCH_CHECK();
        k_incidence_mean<<<(nkeys + 255) / 256, 256, 0, s>>>(nkeys, cnt.p, inc.p, tarea.p, d_out);
        MSM_LAUNCH_CHECK();
        int h_ovf = 0;
        MSM_CUDA(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        MSM_CUDA(cudaStreamSynchronize(s));
        if (!h_ovf) return MSMGPU_OK;      // otherwise some vertex has more than kIncSlots triangles: generic path below
    }
    Buckets B;
    MSM_TRY(bucketize(emit, total_t, nkeys, 3 * (size_t)total_t, B, s, total_t));   // ids = triangle ids
    k_bucket_mean<<<(nkeys + 255) / 256, 256, 0, s>>>(nkeys, B.ptr.p, B.val.p, d_out);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;   // (pageable H2D copies are staged before cudaMemcpyAsync returns, the host tables may go)
}

msmgpu_status vertex_areas_dev(msmgpu_mesh* m, double* d_out) {
    return vertex_areas_batch(m->ctx, 1, &m, std::vector<int>{0, m->nv}, d_out);
}

// ------------------------------------------------------------------------------------------
// adaptive weights (batched over subjects; rows are global: r = s * n_low + target)
// ------------------------------------------------------------------------------------------
// resampler.cpp:105-109: keep the forward list unless the transposed reverse list has MORE entries
__global__ void k_row_lengths(int n_rows, const int* __restrict__ fne, const int* __restrict__ rr_ptr, int* __restrict__ len) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rows) return;
    const int r = rr_ptr[n + 1] - rr_ptr[n];
    len[n] = r <= fne[n] ? fne[n] : r;
}
// rows in ascending column order, multiplied by the target vertex area (resampler.cpp:111-113)
__global__ void __launch_bounds__(256) k_fill_rows(int n_rows, int n_low, const int* __restrict__ rowptr, const int* __restrict__ fidx,
                                                   const double* __restrict__ fw, const int* __restrict__ fne, const int* __restrict__ rr_ptr,
                                                   const int* __restrict__ rr_id, const double* __restrict__ rr_val,
                                                   const double* __restrict__ new_area, int* __restrict__ col, double* __restrict__ val) {
    // half a warp per row (rows hold ~15 entries): consecutive lanes copy consecutive entries, so the reads of the transposed
    // reverse lists and the writes of the row are contiguous (one thread per row walked its row with a stride of the row length)
    const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int hl = threadIdx.x & 15;
    if (n >= n_rows) return;
    const int rb = rr_ptr[n], r = rr_ptr[n + 1] - rb;
    const int o = rowptr[n];
    const double a = new_area[n % n_low];
    const int nf = fne[n];
    if (r <= nf) {
        if (hl < nf) { col[o + hl] = fidx[3 * (size_t)n + hl]; val[o + hl] = fw[3 * (size_t)n + hl] * a; }
    } else {
        for (int j = hl; j < r; j += 16) { col[o + j] = rr_id[rb + j]; val[o + j] = rr_val[rb + j] * a; }
    }
}
// correction[src] = sum over the targets whose row holds src, in ascending target order, of the area-scaled weight
// (resampler.cpp:111-116); ratio[src] = oldArea[src] / correction[src] (resampler.cpp:120-123).
// A row is either the target's forward list or its transposed reverse list (resampler.cpp:105-109). The reverse-list rows that hold
// src are exactly the <= 3 targets of src's OWN reverse map (already in ascending target order), so their terms need no transpose;
// only the forward-list rows go through the column buckets (ptr / id / val, ascending row). The two ascending sequences are merged.
__global__ void k_area_ratio(int nkeys, int S, int n_low, const int* __restrict__ in_off, const int* __restrict__ ridx,
                             const double* __restrict__ rw, const int* __restrict__ rne, const int* __restrict__ fne,
                             const int* __restrict__ rr_ptr, const double* __restrict__ new_area, const int* __restrict__ ptr,
                             const int* __restrict__ id, const double* __restrict__ val, const double* __restrict__ old_area,
                             double* __restrict__ ratio) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int sub = S == 1 ? 0 : find_segment(in_off, S, k);
    const int row0 = sub * n_low;
    int an = 0, arow[3];
    double aval[3];
    const int ne = rne[k];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        arow[j] = INT_MAX; aval[j] = 0.0;
        if (j < ne) {
            const int t = ridx[3 * (size_t)k + j], n = row0 + t;
            if (rr_ptr[n + 1] - rr_ptr[n] > fne[n]) { arow[an] = n; aval[an] = rw[3 * (size_t)k + j] * new_area[t]; ++an; }
        }
    }
    double sum = 0.0;
    int ia = 0, ib = ptr[k];
    const int eb = ptr[k + 1];
    while (ia < an || ib < eb) {
        const int ra = ia < an ? arow[ia] : INT_MAX, rb = ib < eb ? id[ib] : INT_MAX;
        if (ra < rb) { sum += aval[ia]; ++ia; }
        else { sum += val[ib]; ++ib; }
    }
    ratio[k] = old_area[k] / sum;
}
// resampler.cpp:120-137: w *= oldArea[src] / correction[src]; then each row normalised to sum 1.
// The row sum is the reference's sequential sum (column order).
__global__ void __launch_bounds__(256) k_finish_rows(int n_rows, int n_low, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                     double* __restrict__ val, const int* __restrict__ in_off, const double* __restrict__ ratio) {
    // half a warp per row (rows hold ~15 entries): the 16 lanes fetch a chunk of the row, the row sum is taken in column order by
    // lane-ordered broadcasts inside the half-warp
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int hl = threadIdx.x & 15;
    const unsigned mask = 0xffffu << (threadIdx.x & 16);
    if (n >= n_rows) return;          // (whole half-warps leave together)
    const int b = rowptr[n], e = rowptr[n + 1];
    const int koff = in_off[n / n_low];
    double ws = 0.0;
    for (int i0 = b; i0 < e; i0 += 16) {
        const int i = i0 + hl;
        double v = 0.0;
        if (i < e) {
            v = val[i] * __ldg(ratio + koff + col[i]);
            val[i] = v;
        }
        const int m = min(16, e - i0);
        for (int k = 0; k < m; ++k) ws += __shfl_sync(mask, v, k, 16);
    }
    if (ws != 0.0)
        for (int i = b + hl; i < e; i += 16) val[i] /= ws;
}

// Builds the S matrices into one shared store; out[s] become views of it.
// Exclusion masks (resampler.cpp:100, 121): a target whose CLOSEST SOURCE VERTEX is masked out gets no row and adds nothing to the
// correction sums. Both follow from dropping its forward list and every reverse-map entry that points at it before the rows are built.
__global__ void k_mask_forward(int n_rows, const unsigned char* __restrict__ active, int* __restrict__ fne) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < n_rows && !active[n]) fne[n] = 0;
}
__global__ void k_mask_reverse(int nkeys, int S, int n_low, const int* __restrict__ in_off, const unsigned char* __restrict__ active,
                               int* __restrict__ ridx, double* __restrict__ rw, int* __restrict__ rne) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int row0 = (S == 1 ? 0 : find_segment(in_off, S, k)) * n_low;
    const int ne = rne[k];
    int m = 0;
    for (int j = 0; j < ne; ++j) {            // stable compaction of the (ascending) entries that survive
        const int t = ridx[3 * (size_t)k + j];
        const double w = rw[3 * (size_t)k + j];
        if (active[row0 + t]) { ridx[3 * (size_t)k + m] = t; rw[3 * (size_t)k + m] = w; ++m; }
    }
    for (int j = m; j < 3; ++j) { ridx[3 * (size_t)k + j] = -1; rw[3 * (size_t)k + j] = 0.0; }
    rne[k] = m;
}

msmgpu_status adaptive_weights_build_batch(msmgpu_ctx* ctx, int S, msmgpu_mesh* const* in_meshes, msmgpu_octree* const* in_trees,
                                           msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, msmgpu_weights** out, const msmgpu_fwd* fwd,
                                           const unsigned char* d_active = nullptr) {
    cudaStream_t s = ctx->stream;
    const int n_low = low_mesh->nv;
    std::vector<int> in_off(S + 1, 0);
    int max_nv = 0;
    for (int i = 0; i < S; ++i) {
        in_off[i + 1] = in_off[i] + in_meshes[i]->nv;
        max_nv = std::max(max_nv, in_meshes[i]->nv);
    }
    const long long NV = in_off[S], NL = (long long)S * n_low;
    if (3 * (NV + NL) > 0x7fffffffll) return fail(MSMGPU_ERR_CAPACITY, "adaptive_weights: batch too large for 32-bit offsets");

    // forward: targets located in each input mesh; reverse: input vertices located in the target mesh (resampler.cpp:74-78)
    // processing order of the queries (order.cu): one Morton permutation of the first input mesh serves every subject with the same
    // vertex count (a batch shares one topology and nearly the same geometry); one of the target vertices serves the forward queries
    DevBuf<int> perm_rev, perm_fwd;
    const int ordered = tuning_get("query_order", "MSMGPU_QUERY_ORDER", 1);   // 0 off, 1 reverse queries (default), 2 forward queries too
    if (ordered >= 1) MSM_TRY(morton_order(in_meshes[0]->xyz.p, in_meshes[0]->nv, perm_rev, s));
    if (ordered >= 2 && !fwd) MSM_TRY(morton_order(low_mesh->xyz.p, n_low, perm_fwd, s));
    std::vector<QueryJob> jobs(2 * (size_t)S);
    bool lazy_fwd = !fwd;   // forward queries in trees without stored records -> the LAZY query kernel (a mixed batch materialises them)
    for (int i = 0; lazy_fwd && i < S; ++i) lazy_fwd = !in_trees[i]->mesh->rec.p;
    for (int i = 0; i < S; ++i) {
        jobs[i] = QueryJob{(fwd || lazy_fwd) ? in_trees[i]->view_lazy() : in_trees[i]->view(), low_mesh->xyz.p, n_low, i * n_low, perm_fwd.p};
        jobs[S + i] = QueryJob{low_tree->view(), in_meshes[i]->xyz.p, in_meshes[i]->nv, in_off[i],
                               in_meshes[i]->nv == in_meshes[0]->nv ? perm_rev.p : nullptr};
    }
    DevBuf<QueryJob> d_jobs;
    DevBuf<int> d_in_off;
    MSM_CUDA(d_jobs.alloc(jobs.size(), s));
    MSM_CUDA(d_in_off.alloc(S + 1, s));
    MSM_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(QueryJob), cudaMemcpyHostToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(d_in_off.p, in_off.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    // forward weights: computed here, or taken from the fused barycentric resample of the same batch (identical values)
    DevBuf<int> fidx_own, fne_own, ridx, rne, st;
    DevBuf<double> fw_own, rw;
    struct { const int* p; } fidx, fne;
    struct { const double* p; } fw;
    if (fwd) {
        fidx.p = fwd->idx.p; fw.p = fwd->w.p; fne.p = fwd->ne.p;
    } else {
        MSM_CUDA(fidx_own.alloc(3 * (size_t)NL, s));
        MSM_CUDA(fw_own.alloc(3 * (size_t)NL, s));
        MSM_CUDA(fne_own.alloc((size_t)NL, s));
        fidx.p = fidx_own.p; fw.p = fw_own.p; fne.p = fne_own.p;
    }
    MSM_CUDA(ridx.alloc(3 * (size_t)NV, s));
    MSM_CUDA(rw.alloc(3 * (size_t)NV, s));
    MSM_CUDA(rne.alloc((size_t)NV, s));
    MSM_CUDA(st.alloc((size_t)(NL + NV), s));
    if (fwd) MSM_CUDA(cudaMemsetAsync(st.p, 0, (size_t)NL * sizeof(int), s));   // the fused kernel already reported the forward statuses
    else MSM_TRY(launch_bary_weights_batch(d_jobs.p, S, n_low, fidx_own.p, fw_own.p, fne_own.p, st.p, s, lazy_fwd));
    // reverse queries: with enough subjects of one size (one topology, nearly one geometry) the subject goes on the lanes
    bool across = S >= 8 && tuning_get("reverse_across", "MSMGPU_REVERSE_ACROSS", 1) != 0;
    for (int i = 0; across && i < S; ++i) across = in_meshes[i]->nv == in_meshes[0]->nv;
    if (across) {
        std::vector<const double*> h_pts(S);
        for (int i = 0; i < S; ++i) h_pts[i] = in_meshes[i]->xyz.p;
        DevBuf<const double*> d_pts;
        MSM_CUDA(d_pts.alloc(S, s));
        MSM_CUDA(cudaMemcpyAsync(d_pts.p, h_pts.data(), S * sizeof(double*), cudaMemcpyHostToDevice, s));   // pageable: staged before return
        MSM_TRY(launch_bary_weights_across(low_tree->view(), perm_rev.p, d_pts.p, d_in_off.p, S, in_meshes[0]->nv, ridx.p, rw.p, rne.p, st.p + NL, s));
    } else
        MSM_TRY(launch_bary_weights_batch(d_jobs.p + S, S, max_nv, ridx.p, rw.p, rne.p, st.p + NL, s));
    int code = 0;
    MSM_TRY(first_error(st.p, (size_t)(NL + NV), s, &code));   // synchronises: `jobs` / `in_off` may now go
    if (code) return status_to_error(code);
    DevBuf<int> fne_masked;
    if (d_active) {
        MSM_CUDA(fne_masked.alloc((size_t)NL, s));
        MSM_CUDA(cudaMemcpyAsync(fne_masked.p, fne.p, (size_t)NL * sizeof(int), cudaMemcpyDeviceToDevice, s));
        k_mask_forward<<<(unsigned)((NL + 255) / 256), 256, 0, s>>>((int)NL, d_active, fne_masked.p);
        MSM_LAUNCH_CHECK();
        fne.p = fne_masked.p;
        k_mask_reverse<<<(unsigned)((NV + 255) / 256), 256, 0, s>>>((int)NV, S, n_low, d_in_off.p, d_active, ridx.p, rw.p, rne.p);
        MSM_LAUNCH_CHECK();
    }

    DevBuf<double> new_area, old_area, correction;
    MSM_CUDA(new_area.alloc(n_low, s));
    MSM_CUDA(old_area.alloc((size_t)NV, s));
    MSM_CUDA(correction.alloc((size_t)NV, s));
    MSM_TRY(vertex_areas_dev(low_mesh, new_area.p));
    MSM_TRY(vertex_areas_batch(ctx, S, in_meshes, in_off, old_area.p));

    Buckets rr;   // reverse weights regrouped by (subject, target)
    MSM_TRY(bucketize(EmitReverse{ridx.p, rw.p, rne.p, d_in_off.p, S, n_low, (int)NV}, (int)NV, (int)NL, 3 * (size_t)NV, rr, s, max_nv));   // ids = source vertices

    auto store = std::make_shared<WeightsStore>();
    DevBuf<int> len;
    MSM_CUDA(len.alloc((size_t)NL, s));
    MSM_CUDA(store->rowptr.alloc((size_t)NL + 1, s));
    const unsigned gl = (unsigned)((NL + 255) / 256);
    k_row_lengths<<<gl, 256, 0, s>>>((int)NL, fne.p, rr.ptr.p, len.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(len.p, store->rowptr.p, (int)NL, store->rowptr.p + NL, s));
    // nnz <= 3 NL + 3 NV (each row is either its <= 3 forward entries or its share of the <= 3 NV reverse entries)
    const size_t cap = 3 * (size_t)(NL + NV);
    MSM_CUDA(store->col.alloc(cap, s));
    MSM_CUDA(store->val.alloc(cap, s));
    k_fill_rows<<<(unsigned)((16 * NL + 255) / 256), 256, 0, s>>>((int)NL, n_low, store->rowptr.p, fidx.p, fw.p, fne.p, rr.ptr.p, rr.id.p, rr.val.p, new_area.p, store->col.p,
                                   store->val.p);
    MSM_LAUNCH_CHECK();

    // correction[src] = sum over targets (ascending) of the area-scaled weights (resampler.cpp:114-116)
    Buckets cols;
    MSM_TRY(bucketize(EmitCsrColumns{store->rowptr.p, store->col.p, store->val.p, d_in_off.p, fne.p, rr.ptr.p, n_low, (int)NL}, (int)NL, (int)NV,
                      3 * (size_t)NL, cols, s, (int)std::min<long long>(NL, INT_MAX)));   // ids = global rows
    k_area_ratio<<<(unsigned)((NV + 255) / 256), 256, 0, s>>>((int)NV, S, n_low, d_in_off.p, ridx.p, rw.p, rne.p, fne.p, rr.ptr.p, new_area.p, cols.ptr.p,
                                                              cols.id.p, cols.val.p, old_area.p, correction.p);   // correction := ratio
    MSM_LAUNCH_CHECK();
    k_finish_rows<<<(unsigned)((NL * 16 + 255) / 256), 256, 0, s>>>((int)NL, n_low, store->rowptr.p, store->col.p, store->val.p, d_in_off.p, correction.p);
    MSM_LAUNCH_CHECK();

    // per-subject views: rowptr offsets of the subject boundaries
    std::vector<int> bounds(S + 1);
    MSM_CUDA(cudaMemcpy2DAsync(bounds.data(), sizeof(int), store->rowptr.p, (size_t)n_low * sizeof(int), sizeof(int), S + 1, cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < S; ++i) {
        auto* W = new msmgpu_weights();
        W->ctx = ctx;
        W->n_rows = n_low;
        W->n_cols = in_meshes[i]->nv;
        W->store = store;
        W->rowptr = store->rowptr.p + (size_t)i * n_low;
        W->first = bounds[i];
        W->nnz = bounds[i + 1] - bounds[i];
        out[i] = W;
    }
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// CSR apply: out[r][:] = sum over row r (ascending column) of in[col][:] * val, FP64 accumulation in
// the reference's order (resampler.cpp:46-48). One thread per (row, 16-byte chunk): the threads of a
// row read consecutive chunks of the same source rows (coalesced 128-bit loads), col/val are
// broadcast loads. The batched form covers all subjects of a store in one launch.
// ------------------------------------------------------------------------------------------
struct ApplyJob {
    const int* rowptr;   // n_rows + 1 absolute offsets
    const void* in;      // [n_cols][D]
    void* out;           // [n_rows][D]
    const double* excl;  // optional [n_cols]: entries whose source vertex has excl == 0 are skipped (resampler.cpp:46-47, 62-63)
};

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_csr_apply_f32x4(const ApplyJob* __restrict__ jobs, int n_rows, const int* __restrict__ col,
                                                         const double* __restrict__ val, int D4) {
    const ApplyJob job = jobs[blockIdx.y];
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int r = (int)(slot / D4);
    if (r >= n_rows) return;
    const int c = (int)(slot - (long long)r * D4);
    const float4* __restrict__ in = static_cast<const float4*>(job.in);
    const int b = __ldg(job.rowptr + r), e = __ldg(job.rowptr + r + 1);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    // four entries per step: their column ids, weights and source chunks are requested together (one exposed load latency per
    // four entries instead of two per entry); the sums still run in column order, padded slots are skipped, not added as zeros
    for (int i = b; i < e; i += 4) {
        int cc[4];
        double ww[4];
        float4 f[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = min(i + u, e - 1);
            cc[u] = __ldg(col + k);
            ww[u] = __ldg(val + k);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) f[u] = __ldg(in + (size_t)cc[u] * D4 + c);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i + u < e) {
                const double w = ww[u];
                a0 += (double)f[u].x * w; a1 += (double)f[u].y * w; a2 += (double)f[u].z * w; a3 += (double)f[u].w * w;
            }
        }
    }
    __stcs(static_cast<float4*>(job.out) + (size_t)r * D4 + c, make_float4((float)a0, (float)a1, (float)a2, (float)a3));
}

template <typename T>
__global__ void __launch_bounds__(256) k_csr_apply(const ApplyJob* __restrict__ jobs, int n_rows, const int* __restrict__ col,
                                                   const double* __restrict__ val, int D) {
    const ApplyJob job = jobs[blockIdx.y];
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int r = (int)(slot / D);
    if (r >= n_rows) return;
    const int c = (int)(slot - (long long)r * D);
    const T* __restrict__ in = static_cast<const T*>(job.in);
    double a = 0.0;
    for (int i = __ldg(job.rowptr + r); i < __ldg(job.rowptr + r + 1); ++i) {
        const int cc = __ldg(col + i);
        if (job.excl && __ldg(job.excl + cc) == 0) continue;
        a += (double)__ldg(in + (size_t)cc * D + c) * __ldg(val + i);
    }
    static_cast<T*>(job.out)[(size_t)r * D + c] = (T)a;
}

// all weights must share one store (one batch) and have the same n_rows
template <typename T>
static msmgpu_status csr_apply_batch(msmgpu_ctx* ctx, int n, msmgpu_weights* const* Ws, int D, const T* const* d_in, T* const* d_out,
                                     const double* const* d_excl = nullptr) {
    cudaStream_t s = ctx->stream;
    if (n <= 0 || D <= 0) return MSMGPU_OK;
    const WeightsStore* store = Ws[0]->store.get();
    const int n_rows = Ws[0]->n_rows;
    bool vec = std::is_same<T, float>::value && (D & 3) == 0;
    std::vector<ApplyJob> jobs(n);
    for (int i = 0; i < n; ++i) {
        if (Ws[i]->store.get() != store || Ws[i]->n_rows != n_rows) return fail(MSMGPU_ERR_INVALID, "weights_apply_batch: weights from different batches");
        jobs[i] = ApplyJob{Ws[i]->rowptr, d_in[i], d_out[i], d_excl ? d_excl[i] : nullptr};
        vec = vec && !jobs[i].excl && ((reinterpret_cast<uintptr_t>(d_in[i]) | reinterpret_cast<uintptr_t>(d_out[i])) & 15) == 0;
    }
    if (n_rows == 0) return MSMGPU_OK;
    // The bulk-copy gather (gather.cu) is opt-in for CSR rows ("gather_csr"): a row of the adaptive matrix has ~15 entries and every
    // source row is re-read ~3 times from L2, so the kernel moves 12.6 GB through the copy engine, whose measured rate for 400-byte rows
    // (~16 B/clk/SM, profiles/r2a_tune_gather.txt) makes it 1.6x SLOWER than the 128-bit register loads below, which go through L1.
    if (vec && gather_bulk_supported(D) && tuning_get("gather_csr", "MSMGPU_GATHER_CSR", 0) != 0) {
        std::vector<GatherJob> gj(n);
        for (int i = 0; i < n; ++i)
            gj[i] = GatherJob{Ws[i]->rowptr, store->col.p, store->val.p, reinterpret_cast<const float*>(d_in[i]), reinterpret_cast<float*>(d_out[i])};
        DevBuf<GatherJob> d_gj;
        MSM_CUDA(d_gj.alloc(n, s));
        MSM_CUDA(cudaMemcpyAsync(d_gj.p, gj.data(), n * sizeof(GatherJob), cudaMemcpyHostToDevice, s));   // pageable: staged before return
        return launch_gather_rows_bulk(d_gj.p, n, n_rows, D, false, ctx->device, s);
    }
    DevBuf<ApplyJob> d_jobs;
    MSM_CUDA(d_jobs.alloc(n, s));
    MSM_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), n * sizeof(ApplyJob), cudaMemcpyHostToDevice, s));   // pageable: staged before return
    if (vec) {
        const int D4 = D >> 2;
        const long long slots = (long long)n_rows * D4;
        const dim3 grid((unsigned)((slots + 255) / 256), (unsigned)n);
        switch (tuning_get("apply_minb", "MSMGPU_APPLY_MINB", 5)) {   // resident CTAs per SM (5 = the compiler's 48 registers)
            case 6: k_csr_apply_f32x4<6><<<grid, 256, 0, s>>>(d_jobs.p, n_rows, store->col.p, store->val.p, D4); break;
            case 8: k_csr_apply_f32x4<8><<<grid, 256, 0, s>>>(d_jobs.p, n_rows, store->col.p, store->val.p, D4); break;
            default: k_csr_apply_f32x4<5><<<grid, 256, 0, s>>>(d_jobs.p, n_rows, store->col.p, store->val.p, D4); break;
        }
    } else {
        const long long slots = (long long)n_rows * D;
        k_csr_apply<T><<<dim3((unsigned)((slots + 255) / 256), (unsigned)n), 256, 0, s>>>(d_jobs.p, n_rows, store->col.p, store->val.p, D);
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status csr_apply_f32(msmgpu_weights* W, int D, const float* d_in, float* d_out) { return csr_apply_batch<float>(W->ctx, 1, &W, D, &d_in, &d_out); }
msmgpu_status csr_apply_f64(msmgpu_weights* W, int D, const double* d_in, double* d_out) { return csr_apply_batch<double>(W->ctx, 1, &W, D, &d_in, &d_out); }

} // namespace msm

using namespace msm;

extern "C" {

msmgpu_status msmgpu_mesh_set_area_source(msmgpu_mesh* m, msmgpu_mesh* area_mesh) {
    if (!m || (area_mesh && (area_mesh->ctx != m->ctx || area_mesh->nv != m->nv || area_mesh->nt != m->nt || area_mesh->area_source)))
        return fail(MSMGPU_ERR_INVALID, "mesh_set_area_source: meshes must share the context and the topology");
    m->area_source = area_mesh == m ? nullptr : area_mesh;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_mesh_set_triangle_areas(msmgpu_mesh* m, const double* areas) {
    if (!m) return fail(MSMGPU_ERR_INVALID, "mesh_set_triangle_areas: mesh is NULL");
    MSM_CUDA(cudaSetDevice(m->ctx->device));
    if (!areas) { m->tri_area.release(); return MSMGPU_OK; }
    MSM_CUDA(m->tri_area.alloc((size_t)m->nt, m->ctx->stream));
    MSM_CUDA(cudaMemcpyAsync(m->tri_area.p, areas, (size_t)m->nt * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    MSM_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return MSMGPU_OK;
}

msmgpu_status msmgpu_mesh_vertex_areas(msmgpu_mesh* m, double* out) {
    if (!m || !out) return fail(MSMGPU_ERR_INVALID, "mesh_vertex_areas: bad arguments");
    MSM_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t s = m->ctx->stream;
    DevBuf<double> d;
    MSM_CUDA(d.alloc(m->nv, s));
    MSM_TRY(vertex_areas_dev(m, d.p));
    MSM_CUDA(cudaMemcpyAsync(out, d.p, m->nv * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

msmgpu_status msmgpu_adaptive_weights_batch_fwd(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* in_meshes, msmgpu_octree* const* in_trees,
                                                msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, const msmgpu_fwd* fwd, msmgpu_weights** out) {
    if (!ctx || n <= 0 || !in_meshes || !low_mesh || !out || low_mesh->ctx != ctx) return fail(MSMGPU_ERR_INVALID, "adaptive_weights_batch: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    for (int i = 0; i < n; ++i) out[i] = nullptr;
    // missing trees are built as one forest
    std::vector<msmgpu_mesh*> need;
    for (int i = 0; i < n; ++i) {
        if (!in_meshes[i] || in_meshes[i]->ctx != ctx) return fail(MSMGPU_ERR_INVALID, "adaptive_weights_batch: mesh from another context");
        if (!in_trees || !in_trees[i]) need.push_back(in_meshes[i]);
        else if (in_trees[i]->mesh != in_meshes[i]) return fail(MSMGPU_ERR_INVALID, "adaptive_weights: tree/mesh mismatch");
    }
    if (!low_tree) need.push_back(low_mesh);
    else if (low_tree->mesh != low_mesh) return fail(MSMGPU_ERR_INVALID, "adaptive_weights: tree/mesh mismatch");
    std::vector<msmgpu_octree*> built(need.size(), nullptr);
    std::vector<std::unique_ptr<msmgpu_octree>> owned;
    if (!need.empty()) MSM_TRY(mesh_trees(ctx, (int)need.size(), need.data(), built.data(), owned));   // the meshes' own (cached) trees
    std::vector<msmgpu_octree*> trees(n);
    size_t k = 0;
    for (int i = 0; i < n; ++i) trees[i] = (in_trees && in_trees[i]) ? in_trees[i] : built[k++];
    if (!low_tree) low_tree = built[k++];
    if (fwd) {
        if (fwd->ctx != ctx || !fwd->filled || fwd->S != n || fwd->n != low_mesh->nv) return fail(MSMGPU_ERR_INVALID, "adaptive_weights_batch_fwd: weight store does not match the batch");
        for (int i = 0; i < n; ++i)
            if (fwd->trees[i] != trees[i]) return fail(MSMGPU_ERR_INVALID, "adaptive_weights_batch_fwd: weights were computed in other trees");
    }
    return adaptive_weights_build_batch(ctx, n, in_meshes, trees.data(), low_mesh, low_tree, out, fwd);
}

msmgpu_status msmgpu_adaptive_weights_batch(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* in_meshes, msmgpu_octree* const* in_trees,
                                            msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, msmgpu_weights** out) {
    return msmgpu_adaptive_weights_batch_fwd(ctx, n, in_meshes, in_trees, low_mesh, low_tree, nullptr, out);
}

msmgpu_status msmgpu_adaptive_weights_ex(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree,
                                         msmgpu_weights** out) {
    if (!in_mesh || !low_mesh || !out || in_mesh->ctx != low_mesh->ctx) return fail(MSMGPU_ERR_INVALID, "adaptive_weights: bad arguments");
    return msmgpu_adaptive_weights_batch(in_mesh->ctx, 1, &in_mesh, &in_tree, low_mesh, low_tree, out);
}

msmgpu_status msmgpu_adaptive_weights(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, msmgpu_weights** out) {
    return msmgpu_adaptive_weights_ex(in_mesh, nullptr, low_mesh, nullptr, out);
}

msmgpu_status msmgpu_weights_shape(msmgpu_weights* w, int* n_rows, int* n_cols, int64_t* nnz) {
    if (!w) return fail(MSMGPU_ERR_INVALID, "weights is NULL");
    if (n_rows) *n_rows = w->n_rows;
    if (n_cols) *n_cols = w->n_cols;
    if (nnz) *nnz = w->nnz;
    return MSMGPU_OK;
}

msmgpu_status msmgpu_weights_export(msmgpu_weights* w, int32_t* rowptr, int32_t* col, double* val) {
    if (!w) return fail(MSMGPU_ERR_INVALID, "weights is NULL");
    MSM_CUDA(cudaSetDevice(w->ctx->device));
    cudaStream_t s = w->ctx->stream;
    if (rowptr) MSM_CUDA(cudaMemcpyAsync(rowptr, w->rowptr, ((size_t)w->n_rows + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (col && w->nnz) MSM_CUDA(cudaMemcpyAsync(col, w->store->col.p + w->first, (size_t)w->nnz * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (val && w->nnz) MSM_CUDA(cudaMemcpyAsync(val, w->store->val.p + w->first, (size_t)w->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    if (rowptr)   // offsets are absolute inside the batch store: make them local to this matrix
        for (int i = 0; i <= w->n_rows; ++i) rowptr[i] -= w->first;
    return MSMGPU_OK;
}

void msmgpu_weights_destroy(msmgpu_weights* w) {
    if (!w) return;
    cudaSetDevice(w->ctx->device);
    delete w;
}

msmgpu_status msmgpu_weights_apply_f32_dev(msmgpu_weights* w, int D, const float* d_in, float* d_out) {
    if (!w || D <= 0 || !d_in || !d_out) return fail(MSMGPU_ERR_INVALID, "weights_apply: bad arguments");
    MSM_CUDA(cudaSetDevice(w->ctx->device));
    return csr_apply_f32(w, D, d_in, d_out);
}

msmgpu_status msmgpu_weights_apply_batch_f32_dev(msmgpu_ctx* ctx, int n, msmgpu_weights* const* ws, int D, const float* const* d_in, float* const* d_out) {
    if (!ctx || n <= 0 || !ws || D <= 0 || !d_in || !d_out) return fail(MSMGPU_ERR_INVALID, "weights_apply_batch: bad arguments");
    for (int i = 0; i < n; ++i)
        if (!ws[i] || !d_in[i] || !d_out[i]) return fail(MSMGPU_ERR_INVALID, "weights_apply_batch: NULL entry");
    MSM_CUDA(cudaSetDevice(ctx->device));
    return csr_apply_batch<float>(ctx, n, ws, D, d_in, d_out);
}

msmgpu_status msmgpu_weights_apply_batch_f64_dev(msmgpu_ctx* ctx, int n, msmgpu_weights* const* ws, int D, const double* const* d_in, double* const* d_out) {
    if (!ctx || n <= 0 || !ws || D <= 0 || !d_in || !d_out) return fail(MSMGPU_ERR_INVALID, "weights_apply_batch_f64: bad arguments");
    for (int i = 0; i < n; ++i)
        if (!ws[i] || !d_in[i] || !d_out[i]) return fail(MSMGPU_ERR_INVALID, "weights_apply_batch_f64: NULL entry");
    MSM_CUDA(cudaSetDevice(ctx->device));
    return csr_apply_batch<double>(ctx, n, ws, D, d_in, d_out);
}

// metric_resample (resampler.cpp:304-309) on host buffers, FP64 payload: channel-major in/out like Mesh::pvalues
msmgpu_status msmgpu_metric_resample(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, int D, const double* feat_in, double* feat_out) {
    if (!in_mesh || !low_mesh || D <= 0 || !feat_in || !feat_out) return fail(MSMGPU_ERR_INVALID, "metric_resample: bad arguments");
    msmgpu_weights* W = nullptr;
    MSM_TRY(msmgpu_adaptive_weights(in_mesh, low_mesh, &W));
    std::unique_ptr<msmgpu_weights> guard(W);
    cudaStream_t s = in_mesh->ctx->stream;
    const int nv = in_mesh->nv, nl = low_mesh->nv;
    DevBuf<double> cm_in, rows_in, rows_out, cm_out;
    MSM_CUDA(cm_in.alloc((size_t)D * nv, s));
    MSM_CUDA(rows_in.alloc((size_t)D * nv, s));
    MSM_CUDA(rows_out.alloc((size_t)D * nl, s));
    MSM_CUDA(cm_out.alloc((size_t)D * nl, s));
    MSM_CUDA(cudaMemcpyAsync(cm_in.p, feat_in, (size_t)D * nv * sizeof(double), cudaMemcpyHostToDevice, s));
    MSM_TRY(launch_transpose_f64(D, nv, cm_in.p, rows_in.p, s));
    MSM_TRY(csr_apply_f64(W, D, rows_in.p, rows_out.p));
    MSM_TRY(launch_transpose_f64(nl, D, rows_out.p, cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, cm_out.p, (size_t)D * nl * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

__global__ void k_active_from_excl(int n, const int* __restrict__ vtx, const double* __restrict__ excl, unsigned char* __restrict__ active) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) active[i] = (vtx[i] >= 0 && excl[vtx[i]] != 0) ? 1 : 0;
}

// get_adaptive_barycentric_weights with an exclusion mask (resampler.cpp:72-140 with EXCL): d_excl = device copy of EXCL's values
static msmgpu_status adaptive_weights_excl_dev(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, const double* d_excl, msmgpu_weights** out) {
    msmgpu_ctx* ctx = in_mesh->ctx;
    cudaStream_t s = ctx->stream;
    msmgpu_mesh* ms[2] = {in_mesh, low_mesh};
    msmgpu_octree* ts[2] = {nullptr, nullptr};
    std::vector<std::unique_ptr<msmgpu_octree>> owned;
    MSM_TRY(mesh_trees(ctx, 2, ms, ts, owned));
    // active targets: EXCL(octreeSearch_in.get_closest_vertex_ID(target)) != 0 (resampler.cpp:100)
    const int nl = low_mesh->nv;
    DevBuf<int> tri, vtx, st;
    DevBuf<unsigned char> active;
    MSM_CUDA(tri.alloc(nl, s)); MSM_CUDA(vtx.alloc(nl, s)); MSM_CUDA(st.alloc(nl, s)); MSM_CUDA(active.alloc(nl, s));
    MSM_TRY(launch_nearest(ts[0]->view(), nl, low_mesh->xyz.p, tri.p, vtx.p, st.p, s));
    int code = 0;
    MSM_TRY(first_error(st.p, (size_t)nl, s, &code));
    if (code) return status_to_error(code);
    k_active_from_excl<<<(nl + 255) / 256, 256, 0, s>>>(nl, vtx.p, d_excl, active.p);
    MSM_LAUNCH_CHECK();
    return adaptive_weights_build_batch(ctx, 1, &in_mesh, &ts[0], low_mesh, ts[1], out, nullptr, active.p);
}

msmgpu_status msmgpu_adaptive_weights_excl(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, const double* excl, msmgpu_weights** out) {
    if (!in_mesh || !low_mesh || !excl || !out || in_mesh->ctx != low_mesh->ctx) return fail(MSMGPU_ERR_INVALID, "adaptive_weights_excl: bad arguments");
    MSM_CUDA(cudaSetDevice(in_mesh->ctx->device));
    cudaStream_t s = in_mesh->ctx->stream;
    DevBuf<double> d_excl;
    MSM_CUDA(d_excl.alloc(in_mesh->nv, s));
    MSM_CUDA(cudaMemcpyAsync(d_excl.p, excl, (size_t)in_mesh->nv * sizeof(double), cudaMemcpyHostToDevice, s));
    MSM_TRY(adaptive_weights_excl_dev(in_mesh, low_mesh, d_excl.p, out));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

// barycentric_data_interpolation / metric_resample with EXCL (resampler.cpp:30-70): masked weights, masked sums, and the mask itself
// resampled with the same weights (excl_out replaces *EXCL, cpp:66)
msmgpu_status msmgpu_metric_resample_excl(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, int D, const double* feat_in, const double* excl_in,
                                          double* feat_out, double* excl_out) {
    if (!in_mesh || !low_mesh || D <= 0 || !feat_in || !excl_in || !feat_out || !excl_out || in_mesh->ctx != low_mesh->ctx)
        return fail(MSMGPU_ERR_INVALID, "metric_resample_excl: bad arguments");
    msmgpu_ctx* ctx = in_mesh->ctx;
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int nv = in_mesh->nv, nl = low_mesh->nv;
    DevBuf<double> d_excl, cm_in, rows_in, rows_out, cm_out, excl_rows;
    MSM_CUDA(d_excl.alloc(nv, s));
    MSM_CUDA(cudaMemcpyAsync(d_excl.p, excl_in, (size_t)nv * sizeof(double), cudaMemcpyHostToDevice, s));
    msmgpu_weights* W = nullptr;
    MSM_TRY(adaptive_weights_excl_dev(in_mesh, low_mesh, d_excl.p, &W));
    std::unique_ptr<msmgpu_weights> guard(W);
    MSM_CUDA(cm_in.alloc((size_t)D * nv, s));
    MSM_CUDA(rows_in.alloc((size_t)D * nv, s));
    MSM_CUDA(rows_out.alloc((size_t)D * nl, s));
    MSM_CUDA(cm_out.alloc((size_t)D * nl, s));
    MSM_CUDA(excl_rows.alloc(nl, s));
    MSM_CUDA(cudaMemcpyAsync(cm_in.p, feat_in, (size_t)D * nv * sizeof(double), cudaMemcpyHostToDevice, s));
    MSM_TRY(launch_transpose_f64(D, nv, cm_in.p, rows_in.p, s));
    const double* in1 = rows_in.p; double* out1 = rows_out.p; const double* ex = d_excl.p;
    MSM_TRY(csr_apply_batch<double>(ctx, 1, &W, D, &in1, &out1, &ex));
    const double* in2 = d_excl.p; double* out2 = excl_rows.p;
    MSM_TRY(csr_apply_batch<double>(ctx, 1, &W, 1, &in2, &out2, &ex));
    MSM_TRY(launch_transpose_f64(nl, D, rows_out.p, cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, cm_out.p, (size_t)D * nl * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaMemcpyAsync(excl_out, excl_rows.p, (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

// metric_resample of the mesh's RESIDENT features (msmgpu_mesh_set_features_f32): no feature upload
msmgpu_status msmgpu_mesh_metric_resample_f32(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, float* feat_out) {
    if (!in_mesh || !low_mesh || !feat_out) return fail(MSMGPU_ERR_INVALID, "mesh_metric_resample_f32: bad arguments");
    if (in_mesh->feat_D <= 0) return fail(MSMGPU_ERR_INVALID, "mesh_metric_resample_f32: the mesh has no resident features (msmgpu_mesh_set_features_f32)");
    msmgpu_weights* W = nullptr;
    MSM_TRY(msmgpu_adaptive_weights_ex(in_mesh, in_tree, low_mesh, low_tree, &W));
    std::unique_ptr<msmgpu_weights> guard(W);
    cudaStream_t s = in_mesh->ctx->stream;
    const int D = in_mesh->feat_D, nl = low_mesh->nv;
    DevBuf<float> rows_out, cm_out;
    MSM_CUDA(rows_out.alloc((size_t)D * nl, s));
    MSM_CUDA(cm_out.alloc((size_t)D * nl, s));
    MSM_TRY(csr_apply_f32(W, D, in_mesh->feat.p, rows_out.p));
    MSM_TRY(launch_rows_f32_to_chmajor_f32(D, nl, rows_out.p, cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, cm_out.p, (size_t)D * nl * sizeof(float), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

// FP32 payload (what GIFTI stores, mesh.cpp:625): channel-major host floats, optional pre-built trees
msmgpu_status msmgpu_metric_resample_f32(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree,
                                         int D, const float* feat_in, float* feat_out) {
    if (!in_mesh || !low_mesh || D <= 0 || !feat_in || !feat_out) return fail(MSMGPU_ERR_INVALID, "metric_resample_f32: bad arguments");
    msmgpu_weights* W = nullptr;
    MSM_TRY(msmgpu_adaptive_weights_ex(in_mesh, in_tree, low_mesh, low_tree, &W));
    std::unique_ptr<msmgpu_weights> guard(W);
    cudaStream_t s = in_mesh->ctx->stream;
    const int nv = in_mesh->nv, nl = low_mesh->nv;
    DevBuf<float> cm_in, rows_in, rows_out, cm_out;
    MSM_CUDA(cm_in.alloc((size_t)D * nv, s));
    MSM_CUDA(rows_in.alloc((size_t)D * nv, s));
    MSM_CUDA(rows_out.alloc((size_t)D * nl, s));
    MSM_CUDA(cm_out.alloc((size_t)D * nl, s));
    MSM_CUDA(cudaMemcpyAsync(cm_in.p, feat_in, (size_t)D * nv * sizeof(float), cudaMemcpyHostToDevice, s));
    MSM_TRY(launch_chmajor_f32_to_rows_f32(D, nv, cm_in.p, rows_in.p, s));
    MSM_TRY(csr_apply_f32(W, D, rows_in.p, rows_out.p));
    MSM_TRY(launch_rows_f32_to_chmajor_f32(D, nl, rows_out.p, cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, cm_out.p, (size_t)D * nl * sizeof(float), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

} // extern "C"
