// Neighbourhood search of newresampler::smooth_data (msm-newresampler/src/resampler.cpp:169-230) on the device.
//
// For every target i the reference scans ALL vertices n of sphLow (O(V^2), 1.7e9 pair tests at ico6) and keeps those with
//     unit(sphLow[n]) . ref_i >= cos(ang),   ref_i = unit(sphLow[closest_vertex_i]),   ang = 4 asin(sigma / 2R)
// together with the chord |ref_i - unit(sphLow[n])| (resampler.cpp:186-200). Both are +, -, *, /, sqrt on doubles in a fixed order, so
// they are bit-identical on the device (--fmad=false); the Gaussian weights that follow need asin / exp and stay on the host libm,
// on the short lists only (DESIGN.md §4.3 policy). One warp per target: 32 candidates per step, ballot-ordered output = ascending n.
#include "common.cuh"
#include "query.cuh"

#include <cmath>

namespace msm {

__global__ void k_unit_points(int n, const double* __restrict__ xyz, double* __restrict__ unit) {   // Point::normalize, point.cpp:26-34
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = vnormalized(load_pt(xyz, i));
    unit[3 * (size_t)i] = p.x; unit[3 * (size_t)i + 1] = p.y; unit[3 * (size_t)i + 2] = p.z;
}

constexpr int kSmoothWarps = 8;

// FILL = false: count[i] = list length. FILL = true: members / chords written at rowptr[i] in ascending n.
template <bool FILL>
__global__ void __launch_bounds__(kSmoothWarps * 32) k_smooth_neighbours(int n, const double* __restrict__ unit, const int* __restrict__ closest,
                                                                         double cos_ang, int* __restrict__ count, const int* __restrict__ rowptr,
                                                                         int* __restrict__ members, double* __restrict__ chords) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kSmoothWarps + warp;
    if (i >= n) return;
    const V3 ref = load_pt(unit, closest[i]);
    int total = 0;
    int base = FILL ? rowptr[i] : 0;
    for (int n0 = 0; n0 < n; n0 += 32) {
        const int c = n0 + lane;
        bool in = false;
        V3 a{0, 0, 0};
        if (c < n) {
            a = load_pt(unit, c);
            in = vdot(a, ref) >= cos_ang;                       // (actual | ref) >= cos(ang)
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (FILL) {
            if (in) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                members[pos] = c;
                chords[pos] = vnorm(vsub(ref, a));              // (ref - actual).norm()
            }
            base += __popc(m);
        } else {
            total += __popc(m);
        }
    }
    if (!FILL && lane == 0) count[i] = total;
}

}  // namespace msm

using namespace msm;

extern "C" msmgpu_status msmgpu_smooth_neighbourhoods(msmgpu_ctx* ctx, int n, const double* low_xyz, const int32_t* closest, double cos_ang,
                                                      int32_t* rowptr, int64_t cap, int32_t* members, double* chords) {
    if (!ctx || n <= 0 || !low_xyz || !closest || !rowptr) return fail(MSMGPU_ERR_INVALID, "smooth_neighbourhoods: bad arguments");
    // closest[] indexes low_xyz (the reference reads sphLow through ids that came from `orig`, resampler.cpp:185-186: valid only
    // when both meshes have the same vertices); an id outside [0, n) would be an out-of-bounds read on the device
    for (int i = 0; i < n; ++i)
        if (closest[i] < 0 || closest[i] >= n) return fail(MSMGPU_ERR_INVALID, "smooth_neighbourhoods: closest[] holds a vertex id outside the mesh");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevBuf<double> d_xyz, d_unit, d_ch;
    DevBuf<int> d_closest, d_cnt, d_mem;
    MSM_CUDA(d_xyz.alloc(3 * (size_t)n, s));
    MSM_CUDA(d_unit.alloc(3 * (size_t)n, s));
    MSM_CUDA(d_closest.alloc((size_t)n, s));
    MSM_CUDA(d_cnt.alloc((size_t)n + 1, s));
    MSM_CUDA(cudaMemcpyAsync(d_xyz.p, low_xyz, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(d_closest.p, closest, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    k_unit_points<<<(n + 255) / 256, 256, 0, s>>>(n, d_xyz.p, d_unit.p);
    MSM_LAUNCH_CHECK();
    const unsigned grid = (unsigned)((n + kSmoothWarps - 1) / kSmoothWarps);
    k_smooth_neighbours<false><<<grid, kSmoothWarps * 32, 0, s>>>(n, d_unit.p, d_closest.p, cos_ang, d_cnt.p, nullptr, nullptr, nullptr);
    MSM_LAUNCH_CHECK();
    std::vector<int> cnt((size_t)n);
    MSM_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    int64_t total = 0;
    for (int i = 0; i < n; ++i) { rowptr[i] = (int32_t)total; total += cnt[i]; }
    if (total > 0x7fffffffll) return fail(MSMGPU_ERR_CAPACITY, "smooth_neighbourhoods: more than 2^31 neighbour pairs");
    rowptr[n] = (int32_t)total;
    if (!members || !chords || cap < total) return MSMGPU_OK;   // sizing call: rowptr[n] tells the caller what to allocate
    MSM_CUDA(d_mem.alloc((size_t)total, s));
    MSM_CUDA(d_ch.alloc((size_t)total, s));
    MSM_CUDA(cudaMemcpyAsync(d_cnt.p, rowptr, ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    k_smooth_neighbours<true><<<grid, kSmoothWarps * 32, 0, s>>>(n, d_unit.p, d_closest.p, cos_ang, nullptr, d_cnt.p, d_mem.p, d_ch.p);
    MSM_LAUNCH_CHECK();
    MSM_CUDA(cudaMemcpyAsync(members, d_mem.p, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaMemcpyAsync(chords, d_ch.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

// newresampler::smooth_data (resampler.cpp:169-230) as one call: the O(V^2) neighbourhood scan on the device (above), the Gaussian
// weights (asin, exp, sqrt on the host libm) and the reference's sequential sums on the host over the short lists, with the reference's
// expressions and order, so values are the reference's bit for bit. With an exclusion mask (excl != NULL, resampler.cpp:201-225): a
// target whose closest vertex has EXCL <= 0 keeps zeros; otherwise every weight is multiplied by the neighbour's mask value and the new
// mask is the ratio of the masked to the unmasked weight sum. feat_cm [D][n_feat] = orig's pvalues (indexed by the ids of low_xyz's
// vertices, like the reference does), out_cm [D][n], excl_out [n] (written only with a mask).
extern "C" msmgpu_status msmgpu_smooth_data(msmgpu_ctx* ctx, int n, const double* low_xyz, const int32_t* closest, double sigma, int D, int n_feat,
                                            const double* feat_cm, int n_excl, const double* excl, double* out_cm, double* excl_out) {
    if (!ctx || n <= 0 || D <= 0 || !low_xyz || !closest || !feat_cm || !out_cm || n_feat < n || (excl && (!excl_out || n_excl < n)))
        return fail(MSMGPU_ERR_INVALID, "smooth_data: bad arguments");
    if (excl)
        for (int i = 0; i < n; ++i)
            if (closest[i] < 0 || closest[i] >= n_excl) return fail(MSMGPU_ERR_INVALID, "smooth_data: closest[] holds a vertex id outside the mask");
    const double ang = 4 * std::asin(sigma / (2 * kRad));
    std::vector<int32_t> rowptr((size_t)n + 1);
    MSM_TRY(msmgpu_smooth_neighbourhoods(ctx, n, low_xyz, closest, std::cos(ang), rowptr.data(), 0, nullptr, nullptr));
    std::vector<int32_t> members((size_t)rowptr[n]);
    std::vector<double> chords((size_t)rowptr[n]);
    MSM_TRY(msmgpu_smooth_neighbourhoods(ctx, n, low_xyz, closest, std::cos(ang), rowptr.data(), rowptr[n], members.data(), chords.data()));
    for (size_t k = 0; k < (size_t)D * n; ++k) out_cm[k] = 0.0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (excl) excl_out[i] = 0.0;
        if (excl && !(excl[closest[i]] > 0)) continue;
        double SUM = 0.0, excl_sum = 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {   // resampler.cpp:203-214, neighbours in ascending id
            const double geodesic_dist = 2 * kRad * std::asin(chords[e] / (2 * kRad));
            double weight = (1 / std::sqrt(2 * M_PI * sigma * sigma)) * std::exp(-(geodesic_dist * geodesic_dist) / (2 * sigma * sigma));
            excl_sum += weight;
            if (excl) weight = excl[members[e]] * weight;
            SUM += weight;
            for (int d = 0; d < D; ++d) out_cm[(size_t)d * n + i] += feat_cm[(size_t)d * n_feat + members[e]] * weight;
        }
        if (excl_sum != 0.0 && excl) excl_out[i] = SUM / excl_sum;
        for (int d = 0; d < D; ++d)
            if (SUM != 0.0) out_cm[(size_t)d * n + i] /= SUM;
    }
    return MSMGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// variance_normalise (msm-newmeshreg/src/reg_tools.cpp:804-844): per channel, over the vertices the exclusion mask keeps, Welford's
// running mean / variance in vertex order, then (x - mean) / sqrt(var). The recurrence is sequential by definition (every step divides
// by the running count), so one thread walks one channel (IEEE + - * / only: the reference's bits); the rescaling is elementwise.
// ---------------------------------------------------------------------------------------------------------------------------------
namespace msm {

__global__ void k_welford_channels(int D, int n, const double* __restrict__ data, const double* __restrict__ excl, double* __restrict__ mean_var) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double* x = data + (size_t)d * n;
    double mean = 0.0, var = 0.0;
    unsigned int j = 0;   // the reference's loop counter over the compacted vector (cpp:820-825)
    for (int i = 0; i < n; ++i) {
        if (excl && !(excl[i] > 0.0)) continue;   // cpp:811
        const double v = x[i];
        const double delta = v - mean;
        mean += delta / (double)(j + 1u);
        var += delta * (v - mean);
        ++j;
    }
    var /= (double)((size_t)j - (size_t)1);   // `_data[i].size() - 1` in size_t (cpp:827): 2^64 - 1 for an empty channel, 0 for one value
    mean_var[2 * d] = mean;
    mean_var[2 * d + 1] = var;
}

__global__ void k_rescale_channels(int D, int n, double* __restrict__ data, const double* __restrict__ excl, const double* __restrict__ mean_var) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)D * n) return;
    const int d = (int)(i / n), v = (int)(i % n);
    if (excl && !(excl[v] > 0.0)) return;         // excluded vertices keep their values (cpp:841)
    const double mean = mean_var[2 * d], var = mean_var[2 * d + 1];
    double x = data[i] - mean;
    if (var > 0.0) x /= sqrt(var);               // cpp:831-832
    data[i] = x;
}

}  // namespace msm

extern "C" msmgpu_status msmgpu_variance_normalise(msmgpu_ctx* ctx, int D, int n, double* data_cm, const double* excl) {
    using namespace msm;
    if (!ctx || D <= 0 || n <= 0 || !data_cm) return fail(MSMGPU_ERR_INVALID, "variance_normalise: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevBuf<double> d_data, d_excl, d_mv;
    MSM_CUDA(d_data.alloc((size_t)D * n, s));
    MSM_CUDA(d_mv.alloc(2 * (size_t)D, s));
    MSM_CUDA(cudaMemcpyAsync(d_data.p, data_cm, (size_t)D * n * sizeof(double), cudaMemcpyHostToDevice, s));
    if (excl) {
        MSM_CUDA(d_excl.alloc((size_t)n, s));
        MSM_CUDA(cudaMemcpyAsync(d_excl.p, excl, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    k_welford_channels<<<(unsigned)((D + 31) / 32), 32, 0, s>>>(D, n, d_data.p, excl ? d_excl.p : nullptr, d_mv.p);
    MSM_LAUNCH_CHECK();
    k_rescale_channels<<<(unsigned)(((size_t)D * n + 255) / 256), 256, 0, s>>>(D, n, d_data.p, excl ? d_excl.p : nullptr, d_mv.p);
    MSM_LAUNCH_CHECK();
    MSM_CUDA(cudaMemcpyAsync(data_cm, d_data.p, (size_t)D * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}
