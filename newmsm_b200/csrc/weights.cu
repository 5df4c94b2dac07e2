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

// emitters: n_src() sources, each emitting count(i) triples (key, id, val)
struct EmitVertexTriangles {   // key = vertex, id = triangle, val = cached triangle area (triangle.cpp:47-50)
    const int* tri; const TriRec* rec; int nt;
    __device__ int n_src() const { return nt; }
    __device__ int count(int) const { return 3; }
    __device__ int key(int t, int j) const { return tri[3 * (size_t)t + j]; }
    __device__ int id(int t, int) const { return t; }
    __device__ double val(int t, int) const {
        const double* v = rec[t].v;
        return tri_area_cached(V3{v[0], v[1], v[2]}, V3{v[3], v[4], v[5]}, V3{v[6], v[7], v[8]});
    }
};
struct EmitReverse {   // resampler.cpp:91-95: reverse_reorder[target][source] = weight
    const int* ridx; const double* rw; const int* rne; int n;
    __device__ int n_src() const { return n; }
    __device__ int count(int o) const { return rne[o]; }
    __device__ int key(int o, int j) const { return ridx[3 * (size_t)o + j]; }
    __device__ int id(int o, int) const { return o; }
    __device__ double val(int o, int j) const { return rw[3 * (size_t)o + j]; }
};
struct EmitCsrColumns {   // key = column (source vertex), id = row (target), val = stored value
    const int* rowptr; const int* col; const double* v; int n_rows;
    __device__ int n_src() const { return n_rows; }
    __device__ int count(int r) const { return rowptr[r + 1] - rowptr[r]; }
    __device__ int key(int r, int j) const { return col[rowptr[r] + j]; }
    __device__ int id(int r, int) const { return r; }
    __device__ double val(int r, int j) const { return v[rowptr[r] + j]; }
};

template <class E>
__global__ void k_bucket_count(E e, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n_src()) return;
    const int c = e.count(i);
    for (int j = 0; j < c; ++j) atomicAdd(cnt + e.key(i, j), 1);
}
template <class E>
__global__ void k_bucket_fill(E e, const int* __restrict__ ptr, int* __restrict__ cursor, int* __restrict__ id, double* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n_src()) return;
    const int c = e.count(i);
    for (int j = 0; j < c; ++j) {
        const int k = e.key(i, j);
        const int pos = ptr[k] + atomicAdd(cursor + k, 1);
        id[pos] = e.id(i, j);
        val[pos] = e.val(i, j);
    }
}
// insertion sort of each bucket by id (buckets hold a handful of entries: vertex valence, or the
// ~3 N_s / N_t sources that fall into one target's triangles)
__global__ void k_bucket_sort(int nkeys, const int* __restrict__ ptr, int* __restrict__ id, double* __restrict__ val) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int b = ptr[k], e = ptr[k + 1];
    for (int i = b + 1; i < e; ++i) {
        const int ki = id[i];
        const double vi = val[i];
        int j = i - 1;
        while (j >= b && id[j] > ki) { id[j + 1] = id[j]; val[j + 1] = val[j]; --j; }
        id[j + 1] = ki; val[j + 1] = vi;
    }
}

template <class E>
static msmgpu_status bucketize(const E& e, int n_src, int nkeys, Buckets& B, cudaStream_t s) {
    DevBuf<int> cnt, cursor, d_total;
    MSM_CUDA(cnt.alloc(nkeys, s));
    MSM_CUDA(cursor.alloc(nkeys, s));
    MSM_CUDA(d_total.alloc(1, s));
    MSM_CUDA(B.ptr.alloc((size_t)nkeys + 1, s));
    MSM_CUDA(cudaMemsetAsync(cnt.p, 0, nkeys * sizeof(int), s));
    MSM_CUDA(cudaMemsetAsync(cursor.p, 0, nkeys * sizeof(int), s));
    const unsigned gs = (unsigned)((n_src + 255) / 256), gk = (unsigned)((nkeys + 255) / 256);
    if (n_src > 0) k_bucket_count<E><<<gs, 256, 0, s>>>(e, cnt.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(cnt.p, B.ptr.p, nkeys, d_total.p, s));
    MSM_CUDA(cudaMemcpyAsync(B.ptr.p + nkeys, d_total.p, sizeof(int), cudaMemcpyDeviceToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(&B.total, d_total.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    MSM_CUDA(B.id.alloc((size_t)B.total, s));
    MSM_CUDA(B.val.alloc((size_t)B.total, s));
    if (B.total > 0) {
        k_bucket_fill<E><<<gs, 256, 0, s>>>(e, B.ptr.p, cursor.p, B.id.p, B.val.p);
        MSM_LAUNCH_CHECK();
        k_bucket_sort<<<gk, 256, 0, s>>>(nkeys, B.ptr.p, B.id.p, B.val.p);
        MSM_LAUNCH_CHECK();
    }
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

msmgpu_status vertex_areas_dev(msmgpu_mesh* m, double* d_out) {
    cudaStream_t s = m->ctx->stream;
    Buckets B;
    MSM_TRY(bucketize(EmitVertexTriangles{m->tri.p, m->rec.p, m->nt}, m->nt, m->nv, B, s));
    k_bucket_mean<<<(m->nv + 255) / 256, 256, 0, s>>>(m->nv, B.ptr.p, B.val.p, d_out);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// adaptive weights
// ------------------------------------------------------------------------------------------
// resampler.cpp:105-109: keep the forward list unless the transposed reverse list has MORE entries
__global__ void k_row_lengths(int n_low, const int* __restrict__ fne, const int* __restrict__ rr_ptr, int* __restrict__ len) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_low) return;
    const int r = rr_ptr[n + 1] - rr_ptr[n];
    len[n] = r <= fne[n] ? fne[n] : r;
}
// rows in ascending column order, multiplied by the target vertex area (resampler.cpp:111-113)
__global__ void k_fill_rows(int n_low, const int* __restrict__ rowptr, const int* __restrict__ fidx, const double* __restrict__ fw,
                            const int* __restrict__ fne, const int* __restrict__ rr_ptr, const int* __restrict__ rr_id,
                            const double* __restrict__ rr_val, const double* __restrict__ new_area, int* __restrict__ col, double* __restrict__ val) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_low) return;
    const int rb = rr_ptr[n], r = rr_ptr[n + 1] - rb;
    const int o = rowptr[n];
    const double a = new_area[n];
    if (r <= fne[n]) {
        for (int j = 0; j < fne[n]; ++j) { col[o + j] = fidx[3 * (size_t)n + j]; val[o + j] = fw[3 * (size_t)n + j] * a; }
    } else {
        for (int j = 0; j < r; ++j) { col[o + j] = rr_id[rb + j]; val[o + j] = rr_val[rb + j] * a; }
    }
}
// resampler.cpp:120-137: w *= oldArea[src] / correction[src]; then each row normalised to sum 1
__global__ void k_finish_rows(int n_low, const int* __restrict__ rowptr, const int* __restrict__ col, double* __restrict__ val,
                              const double* __restrict__ old_area, const double* __restrict__ correction) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_low) return;
    const int b = rowptr[n], e = rowptr[n + 1];
    double ws = 0.0;
    for (int i = b; i < e; ++i) {
        const int c = col[i];
        const double v = val[i] * (old_area[c] / correction[c]);
        val[i] = v;
        ws += v;
    }
    if (ws != 0.0)
        for (int i = b; i < e; ++i) val[i] /= ws;
}

msmgpu_status adaptive_weights_build(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree,
                                     msmgpu_weights* W) {
    msmgpu_ctx* ctx = in_mesh->ctx;
    cudaStream_t s = ctx->stream;
    const int nv_in = in_mesh->nv, nv_low = low_mesh->nv;
    // forward: targets located in the input mesh; reverse: input vertices located in the target mesh (resampler.cpp:74-78)
    DevBuf<int> fidx, fne, ridx, rne, st;
    DevBuf<double> fw, rw;
    MSM_CUDA(fidx.alloc(3 * (size_t)nv_low, s));
    MSM_CUDA(fw.alloc(3 * (size_t)nv_low, s));
    MSM_CUDA(fne.alloc(nv_low, s));
    MSM_CUDA(ridx.alloc(3 * (size_t)nv_in, s));
    MSM_CUDA(rw.alloc(3 * (size_t)nv_in, s));
    MSM_CUDA(rne.alloc(nv_in, s));
    MSM_CUDA(st.alloc((size_t)nv_low + nv_in, s));
    MSM_TRY(launch_bary_weights(in_tree->view(), nv_low, low_mesh->xyz.p, fidx.p, fw.p, fne.p, st.p, s));
    MSM_TRY(launch_bary_weights(low_tree->view(), nv_in, in_mesh->xyz.p, ridx.p, rw.p, rne.p, st.p + nv_low, s));
    int code = 0;
    MSM_TRY(first_error(st.p, (size_t)nv_low + nv_in, s, &code));
    if (code) return status_to_error(code);

    DevBuf<double> new_area, old_area, correction;
    MSM_CUDA(new_area.alloc(nv_low, s));
    MSM_CUDA(old_area.alloc(nv_in, s));
    MSM_CUDA(correction.alloc(nv_in, s));
    MSM_TRY(vertex_areas_dev(low_mesh, new_area.p));
    MSM_TRY(vertex_areas_dev(in_mesh, old_area.p));

    Buckets rr;   // reverse weights regrouped by target
    MSM_TRY(bucketize(EmitReverse{ridx.p, rw.p, rne.p, nv_in}, nv_in, nv_low, rr, s));

    DevBuf<int> len, d_nnz;
    MSM_CUDA(len.alloc(nv_low, s));
    MSM_CUDA(d_nnz.alloc(1, s));
    MSM_CUDA(W->rowptr.alloc((size_t)nv_low + 1, s));
    const unsigned gl = (unsigned)((nv_low + 255) / 256);
    k_row_lengths<<<gl, 256, 0, s>>>(nv_low, fne.p, rr.ptr.p, len.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(len.p, W->rowptr.p, nv_low, d_nnz.p, s));
    MSM_CUDA(cudaMemcpyAsync(W->rowptr.p + nv_low, d_nnz.p, sizeof(int), cudaMemcpyDeviceToDevice, s));
    int nnz = 0;
    MSM_CUDA(cudaMemcpyAsync(&nnz, d_nnz.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    MSM_CUDA(W->col.alloc((size_t)nnz, s));
    MSM_CUDA(W->val.alloc((size_t)nnz, s));
    k_fill_rows<<<gl, 256, 0, s>>>(nv_low, W->rowptr.p, fidx.p, fw.p, fne.p, rr.ptr.p, rr.id.p, rr.val.p, new_area.p, W->col.p, W->val.p);
    MSM_LAUNCH_CHECK();

    // correction[src] = sum over targets (ascending) of the area-scaled weights (resampler.cpp:114-116)
    Buckets cols;
    MSM_TRY(bucketize(EmitCsrColumns{W->rowptr.p, W->col.p, W->val.p, nv_low}, nv_low, nv_in, cols, s));
    k_bucket_sum<<<(nv_in + 255) / 256, 256, 0, s>>>(nv_in, cols.ptr.p, cols.val.p, correction.p);
    MSM_LAUNCH_CHECK();
    k_finish_rows<<<gl, 256, 0, s>>>(nv_low, W->rowptr.p, W->col.p, W->val.p, old_area.p, correction.p);
    MSM_LAUNCH_CHECK();
    W->ctx = ctx;
    W->n_rows = nv_low;
    W->n_cols = nv_in;
    W->nnz = nnz;
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// CSR apply: out[r][:] = sum over row r (ascending column) of in[col][:] * val, FP64 accumulation in
// the reference's order (resampler.cpp:46-48). One thread per (row, 16-byte chunk): the threads of a
// row read consecutive chunks of the same source rows (coalesced 128-bit loads), col/val are
// broadcast loads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_csr_apply_f32x4(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                         const double* __restrict__ val, int D4, const float4* __restrict__ in,
                                                         float4* __restrict__ out) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int r = (int)(slot / D4);
    if (r >= n_rows) return;
    const int c = (int)(slot - (long long)r * D4);
    const int b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int i = b; i < e; ++i) {
        const float4 f = __ldg(in + (size_t)__ldg(col + i) * D4 + c);
        const double w = __ldg(val + i);
        a0 += (double)f.x * w; a1 += (double)f.y * w; a2 += (double)f.z * w; a3 += (double)f.w * w;
    }
    __stcs(out + (size_t)r * D4 + c, make_float4((float)a0, (float)a1, (float)a2, (float)a3));
}

template <typename T>
__global__ void __launch_bounds__(256) k_csr_apply(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                   const double* __restrict__ val, int D, const T* __restrict__ in, T* __restrict__ out) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int r = (int)(slot / D);
    if (r >= n_rows) return;
    const int c = (int)(slot - (long long)r * D);
    double a = 0.0;
    for (int i = __ldg(rowptr + r); i < __ldg(rowptr + r + 1); ++i) a += (double)__ldg(in + (size_t)__ldg(col + i) * D + c) * __ldg(val + i);
    out[(size_t)r * D + c] = (T)a;
}

msmgpu_status csr_apply_f32(const msmgpu_weights* W, int D, const float* d_in, float* d_out, cudaStream_t s) {
    if (W->n_rows == 0 || D == 0) return MSMGPU_OK;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0) {
        const int D4 = D >> 2;
        const long long slots = (long long)W->n_rows * D4;
        k_csr_apply_f32x4<<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(W->n_rows, W->rowptr.p, W->col.p, W->val.p, D4,
                                                                         reinterpret_cast<const float4*>(d_in), reinterpret_cast<float4*>(d_out));
    } else {
        const long long slots = (long long)W->n_rows * D;
        k_csr_apply<float><<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(W->n_rows, W->rowptr.p, W->col.p, W->val.p, D, d_in, d_out);
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status csr_apply_f64(const msmgpu_weights* W, int D, const double* d_in, double* d_out, cudaStream_t s) {
    if (W->n_rows == 0 || D == 0) return MSMGPU_OK;
    const long long slots = (long long)W->n_rows * D;
    k_csr_apply<double><<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(W->n_rows, W->rowptr.p, W->col.p, W->val.p, D, d_in, d_out);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

} // namespace msm

using namespace msm;

extern "C" {

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

msmgpu_status msmgpu_adaptive_weights_ex(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree,
                                         msmgpu_weights** out) {
    if (!in_mesh || !low_mesh || !out || in_mesh->ctx != low_mesh->ctx) return fail(MSMGPU_ERR_INVALID, "adaptive_weights: bad arguments");
    *out = nullptr;
    MSM_CUDA(cudaSetDevice(in_mesh->ctx->device));
    std::unique_ptr<msmgpu_octree> own_in, own_low;
    if (!in_tree || !low_tree) {   // both missing trees are built as one forest
        msmgpu_mesh* ms[2];
        msmgpu_octree* ts[2];
        int k = 0;
        if (!in_tree) ms[k++] = in_mesh;
        if (!low_tree) ms[k++] = low_mesh;
        MSM_TRY(msmgpu_octree_build_batch(in_mesh->ctx, k, ms, ts));
        k = 0;
        if (!in_tree) { own_in.reset(ts[k++]); in_tree = own_in.get(); }
        if (!low_tree) { own_low.reset(ts[k++]); low_tree = own_low.get(); }
    }
    if (in_tree->mesh != in_mesh || low_tree->mesh != low_mesh) return fail(MSMGPU_ERR_INVALID, "adaptive_weights: tree/mesh mismatch");
    auto W = std::unique_ptr<msmgpu_weights>(new msmgpu_weights());
    MSM_TRY(adaptive_weights_build(in_mesh, in_tree, low_mesh, low_tree, W.get()));
    MSM_CUDA(cudaStreamSynchronize(in_mesh->ctx->stream));
    *out = W.release();
    return MSMGPU_OK;
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
    if (rowptr) MSM_CUDA(cudaMemcpyAsync(rowptr, w->rowptr.p, ((size_t)w->n_rows + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (col && w->nnz) MSM_CUDA(cudaMemcpyAsync(col, w->col.p, (size_t)w->nnz * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (val && w->nnz) MSM_CUDA(cudaMemcpyAsync(val, w->val.p, (size_t)w->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
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
    return csr_apply_f32(w, D, d_in, d_out, w->ctx->stream);
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
    MSM_TRY(csr_apply_f64(W, D, rows_in.p, rows_out.p, s));
    MSM_TRY(launch_transpose_f64(nl, D, rows_out.p, cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, cm_out.p, (size_t)D * nl * sizeof(double), cudaMemcpyDeviceToHost, s));
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
    MSM_TRY(csr_apply_f32(W, D, rows_in.p, rows_out.p, s));
    MSM_TRY(launch_rows_f32_to_chmajor_f32(D, nl, rows_out.p, cm_out.p, s));
    MSM_CUDA(cudaMemcpyAsync(feat_out, cm_out.p, (size_t)D * nl * sizeof(float), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

} // extern "C"
