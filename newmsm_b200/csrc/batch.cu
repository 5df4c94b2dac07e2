// Batch resampling from HOST buffers as one pipelined call: the per-subject loop of an HCP-style job (BASELINE configs[1]) —
// Octree(subject) + barycentric_data_interpolation through get_barycentric_weights (msm-newresampler/src/resampler.cpp:40-52, 142-167)
// and metric_resample (resampler.cpp:304, adaptive barycentric weights 72-140) of S subjects onto one target sphere.
//
// Why an entry point of its own: through the per-subject host calls (msmgpu_mesh_create, msmgpu_mesh_set_features_f32,
// msmgpu_octree_build, msmgpu_mesh_*_resample_f32) every subject pays five stream synchronisations and the job is bound by PCIe
// (65.5 MB of FP32 features per subject against 0.2 ms of device work). Here the subjects go through in chunks on three streams:
// chunk k+1 is uploaded (copy-in stream) while chunk k is computed with the BATCHED device kernels (the context's stream: one forest
// build, one query launch, one gather launch, one adaptive-weights batch and one CSR apply per chunk) and chunk k-1 is downloaded
// (copy-out stream). The H2D engine never waits for the host, so the call runs at the rate of the link.
// Results are those of the per-subject calls bit for bit (same kernels, same inputs): tests/test_gpu_parity.py.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace msm {

struct BatchStage {
    DevBuf<double> xyz;                         // [C][nv][3]
    DevBuf<float> feat_cm, feat_rows;           // [C][D][nv] as uploaded, [C][nv][D] rows
    DevBuf<float> outb_rows, outa_rows;         // [C][n_low][D]
    DevBuf<float> outb_cm, outa_cm;             // [C][D][n_low]
    DevBuf<int> status;                         // [C][n_low]
    cudaEvent_t in_done = nullptr, compute_done = nullptr, out_done = nullptr;
    bool used = false;
};

struct BatchPipeline {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<BatchStage> stages;
    ~BatchPipeline() {
        for (BatchStage& st : stages)
            for (cudaEvent_t e : {st.in_done, st.compute_done, st.out_done})
                if (e) cudaEventDestroy(e);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

}  // namespace msm

using namespace msm;

extern "C" msmgpu_status msmgpu_resample_batch_host_f32(msmgpu_ctx* ctx, int n_subjects, int nv, const double* const* xyz, int nt, const int32_t* tri,
                                                        int n_low, const double* low_xyz, int n_low_tri, const int32_t* low_tri, int D,
                                                        const float* const* feat_cm, float* const* out_bary_cm, float* const* out_adaptive_cm,
                                                        int chunk) {
    if (!ctx || n_subjects <= 0 || nv <= 0 || !xyz || nt <= 0 || !tri || n_low <= 0 || !low_xyz || n_low_tri <= 0 || !low_tri || D <= 0 || !feat_cm ||
        (!out_bary_cm && !out_adaptive_cm))
        return fail(MSMGPU_ERR_INVALID, "resample_batch_host_f32: bad arguments");
    for (int s_ = 0; s_ < n_subjects; ++s_)
        if (!xyz[s_] || !feat_cm[s_] || (out_bary_cm && !out_bary_cm[s_]) || (out_adaptive_cm && !out_adaptive_cm[s_]))
            return fail(MSMGPU_ERR_INVALID, "resample_batch_host_f32: NULL subject buffer");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t sc = ctx->stream;
    const bool want_b = out_bary_cm != nullptr, want_a = out_adaptive_cm != nullptr;
    const int C = std::max(1, std::min(chunk > 0 ? chunk : tuning_get("batch_chunk", "MSMGPU_BATCH_CHUNK", 4), n_subjects));
    const int NB = 3;   // stage buffers: upload of k+1, compute of k, download of k-1

    // the target sphere: one mesh, one tree, shared by every chunk
    msmgpu_mesh* low = nullptr;
    msmgpu_octree* low_tree = nullptr;
    MSM_TRY(msmgpu_mesh_create(ctx, n_low, low_xyz, n_low_tri, low_tri, &low));
    struct LowGuard { msmgpu_mesh* m; msmgpu_octree** t; ~LowGuard() { if (*t) msmgpu_octree_destroy(*t); msmgpu_mesh_destroy(m); } } low_guard{low, &low_tree};
    MSM_TRY(msmgpu_octree_build(low, &low_tree));
    DevBuf<int> d_tri;
    MSM_CUDA(d_tri.alloc(3 * (size_t)nt, sc));
    MSM_CUDA(cudaMemcpyAsync(d_tri.p, tri, 3 * (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, sc));

    BatchPipeline P;
    MSM_CUDA(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
    MSM_CUDA(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
    P.stages.resize(NB);
    const size_t fin = (size_t)D * nv, fout = (size_t)D * n_low;
    for (BatchStage& st : P.stages) {
        MSM_CUDA(st.xyz.alloc((size_t)C * 3 * nv, sc));
        MSM_CUDA(st.feat_cm.alloc((size_t)C * fin, sc));
        MSM_CUDA(st.feat_rows.alloc((size_t)C * fin, sc));
        if (want_b) { MSM_CUDA(st.outb_rows.alloc((size_t)C * fout, sc)); MSM_CUDA(st.outb_cm.alloc((size_t)C * fout, sc)); }
        if (want_a) { MSM_CUDA(st.outa_rows.alloc((size_t)C * fout, sc)); MSM_CUDA(st.outa_cm.alloc((size_t)C * fout, sc)); }
        MSM_CUDA(st.status.alloc((size_t)C * n_low, sc));
        MSM_CUDA(cudaEventCreateWithFlags(&st.in_done, cudaEventDisableTiming));
        MSM_CUDA(cudaEventCreateWithFlags(&st.compute_done, cudaEventDisableTiming));
        MSM_CUDA(cudaEventCreateWithFlags(&st.out_done, cudaEventDisableTiming));
    }
    MSM_CUDA(cudaStreamSynchronize(sc));   // the stage buffers exist before the other streams touch them

    const int n_chunks = (n_subjects + C - 1) / C;
    auto first_of = [&](int k) { return k * C; };
    auto count_of = [&](int k) { return std::min(C, n_subjects - k * C); };

    auto upload = [&](int k) -> msmgpu_status {
        BatchStage& st = P.stages[k % NB];
        if (st.used) MSM_CUDA(cudaStreamWaitEvent(P.s_in, st.compute_done, 0));   // the chunk that used these input buffers has been computed
        for (int j = 0; j < count_of(k); ++j) {
            const int sb = first_of(k) + j;
            MSM_CUDA(cudaMemcpyAsync(st.xyz.p + (size_t)j * 3 * nv, xyz[sb], 3 * (size_t)nv * sizeof(double), cudaMemcpyHostToDevice, P.s_in));
            MSM_CUDA(cudaMemcpyAsync(st.feat_cm.p + (size_t)j * fin, feat_cm[sb], fin * sizeof(float), cudaMemcpyHostToDevice, P.s_in));
        }
        MSM_CUDA(cudaEventRecord(st.in_done, P.s_in));
        return MSMGPU_OK;
    };

    auto compute = [&](int k) -> msmgpu_status {
        BatchStage& st = P.stages[k % NB];
        const int n = count_of(k);
        MSM_CUDA(cudaStreamWaitEvent(sc, st.in_done, 0));
        if (st.used) MSM_CUDA(cudaStreamWaitEvent(sc, st.out_done, 0));            // the previous results of these output buffers have left
        std::vector<const double*> xp(n);
        std::vector<const float*> fp(n);
        std::vector<float*> ob(n), oa(n);
        for (int j = 0; j < n; ++j) {
            xp[j] = st.xyz.p + (size_t)j * 3 * nv;
            fp[j] = st.feat_rows.p + (size_t)j * fin;
            ob[j] = want_b ? st.outb_rows.p + (size_t)j * fout : nullptr;
            oa[j] = want_a ? st.outa_rows.p + (size_t)j * fout : nullptr;
            MSM_TRY(launch_chmajor_f32_to_rows_f32(D, nv, st.feat_cm.p + (size_t)j * fin, st.feat_rows.p + (size_t)j * fin, sc));
        }
        std::vector<msmgpu_mesh*> meshes(n, nullptr);
        std::vector<msmgpu_octree*> trees(n, nullptr);
        std::vector<msmgpu_weights*> ws(n, nullptr);
        msmgpu_fwd* fwd = nullptr;
        struct Cleanup {
            std::vector<msmgpu_mesh*>& m; std::vector<msmgpu_octree*>& t; std::vector<msmgpu_weights*>& w; msmgpu_fwd*& f;
            ~Cleanup() {
                for (auto* x : w) if (x) msmgpu_weights_destroy(x);
                if (f) msmgpu_fwd_destroy(f);
                for (auto* x : t) if (x) msmgpu_octree_destroy(x);
                for (auto* x : m) if (x) msmgpu_mesh_destroy(x);
            }
        } cleanup{meshes, trees, ws, fwd};
        MSM_TRY(msmgpu_mesh_create_view_batch(ctx, n, nv, xp.data(), nt, d_tri.p, meshes.data()));
        MSM_TRY(msmgpu_octree_build_batch(ctx, n, meshes.data(), trees.data()));
        MSM_CUDA(cudaMemsetAsync(st.status.p, 0, (size_t)n * n_low * sizeof(int), sc));
        if (want_b && want_a) {   // both methods: the forward weight maps are computed once (msmgpu_fwd)
            MSM_TRY(msmgpu_fwd_create(ctx, n, n_low, &fwd));
            MSM_TRY(msmgpu_bary_resample_batch_f32_dev_keep(ctx, n, trees.data(), n_low, low->xyz.p, D, fp.data(), ob.data(), st.status.p, fwd));
            MSM_TRY(msmgpu_adaptive_weights_batch_fwd(ctx, n, meshes.data(), trees.data(), low, low_tree, fwd, ws.data()));
        } else if (want_b) {
            MSM_TRY(msmgpu_bary_resample_batch_f32_dev(ctx, n, trees.data(), n_low, low->xyz.p, D, fp.data(), ob.data(), st.status.p));
        } else {
            MSM_TRY(msmgpu_adaptive_weights_batch(ctx, n, meshes.data(), trees.data(), low, low_tree, ws.data()));
        }
        if (want_b) {
            int code = 0;
            MSM_TRY(first_error(st.status.p, (size_t)n * n_low, sc, &code));
            if (code) return status_to_error(code);
        }
        if (want_a) MSM_TRY(msmgpu_weights_apply_batch_f32_dev(ctx, n, ws.data(), D, fp.data(), oa.data()));
        for (int j = 0; j < n; ++j) {
            if (want_b) MSM_TRY(launch_rows_f32_to_chmajor_f32(D, n_low, ob[j], st.outb_cm.p + (size_t)j * fout, sc));
            if (want_a) MSM_TRY(launch_rows_f32_to_chmajor_f32(D, n_low, oa[j], st.outa_cm.p + (size_t)j * fout, sc));
        }
        MSM_CUDA(cudaEventRecord(st.compute_done, sc));
        MSM_CUDA(cudaStreamSynchronize(sc));   // the chunk's handles are released below; their memory is stream-ordered on sc
        st.used = true;
        return MSMGPU_OK;
    };

    auto download = [&](int k) -> msmgpu_status {
        BatchStage& st = P.stages[k % NB];
        MSM_CUDA(cudaStreamWaitEvent(P.s_out, st.compute_done, 0));
        for (int j = 0; j < count_of(k); ++j) {
            const int sb = first_of(k) + j;
            if (want_b) MSM_CUDA(cudaMemcpyAsync(out_bary_cm[sb], st.outb_cm.p + (size_t)j * fout, fout * sizeof(float), cudaMemcpyDeviceToHost, P.s_out));
            if (want_a) MSM_CUDA(cudaMemcpyAsync(out_adaptive_cm[sb], st.outa_cm.p + (size_t)j * fout, fout * sizeof(float), cudaMemcpyDeviceToHost, P.s_out));
        }
        MSM_CUDA(cudaEventRecord(st.out_done, P.s_out));
        return MSMGPU_OK;
    };

    msmgpu_status result = MSMGPU_OK;
    auto run = [&]() -> msmgpu_status {
        MSM_TRY(upload(0));
        for (int k = 0; k < n_chunks; ++k) {
            if (k + 1 < n_chunks) MSM_TRY(upload(k + 1));   // queued before the host blocks inside compute(k)
            MSM_TRY(compute(k));
            MSM_TRY(download(k));
        }
        return MSMGPU_OK;
    };
    result = run();
    // every stream drains before the stage buffers go out of scope (also on the error path)
    cudaStreamSynchronize(P.s_in);
    cudaStreamSynchronize(P.s_out);
    cudaStreamSynchronize(sc);
    if (result == MSMGPU_OK) MSM_CUDA(cudaGetLastError());
    return result;
}
