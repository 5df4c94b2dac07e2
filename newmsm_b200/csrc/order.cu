// Spatially coherent processing order for batches of point queries.
//
// The nearest-triangle query (query.cuh) is bound by L1 tag lookups of per-lane loads (node, leaf list, cull spheres, 128-byte
// triangle records; profiles/r2e_ncu_summary.md): lanes whose points fall into the same octree leaf turn 32 loads into one
// broadcast. Mesh vertex order is only partly coherent (icosphere subdivision appends every level's new vertices after all older
// ones), so the queries of a batch are processed along a Morton curve: 10 bits per axis on the octree's root cube, one radix sort
// of (code, index) per batch — the subjects of a batch share one topology and nearly the same geometry, so one permutation
// serves all of them. The permutation only decides WHICH thread handles a point; every output is written at the point's own index,
// so results do not depend on it.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace msm {

__device__ __forceinline__ unsigned spread10(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void k_morton_keys(int n, const double* __restrict__ xyz, unsigned* __restrict__ key, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sc = 1024.0 / (2.0 * kBounds);
    auto q = [&](double v) { const double t = (v + kBounds) * sc; return (unsigned)(t < 0.0 ? 0.0 : (t > 1023.0 ? 1023.0 : t)); };
    key[i] = (spread10(q(xyz[3 * (size_t)i])) << 2) | (spread10(q(xyz[3 * (size_t)i + 1])) << 1) | spread10(q(xyz[3 * (size_t)i + 2]));
    idx[i] = i;
}

// perm[k] = index of the k-th point along the curve (stream-ordered; perm is allocated here)
msmgpu_status morton_order(const double* d_xyz, int n, DevBuf<int>& perm, cudaStream_t s) {
    DevBuf<unsigned> k_in, k_out;
    DevBuf<int> i_in;
    MSM_CUDA(k_in.alloc(n, s));
    MSM_CUDA(k_out.alloc(n, s));
    MSM_CUDA(i_in.alloc(n, s));
    MSM_CUDA(perm.alloc(n, s));
    if (n == 0) return MSMGPU_OK;
    k_morton_keys<<<(n + 255) / 256, 256, 0, s>>>(n, d_xyz, k_in.p, i_in.p);
    MSM_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    MSM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, i_in.p, perm.p, n, 0, 30, s));
    DevBuf<unsigned char> tmp;
    MSM_CUDA(tmp.alloc(tmp_bytes, s));
    MSM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, i_in.p, perm.p, n, 0, 30, s));
    g_launch_count.fetch_add(3, std::memory_order_relaxed);   // the sort's own kernels (histogram, onesweep passes)
    return MSMGPU_OK;
}

}  // namespace msm
