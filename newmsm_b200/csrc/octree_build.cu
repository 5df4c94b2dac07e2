// Device-side construction of the reference's octree (msm-newresampler/src/octree.cpp:31-141,
// node.cpp:67-120) for a whole batch of meshes at once ("forest").
//
// The reference inserts triangles one at a time and splits a leaf when it holds n >= 50
// triangles AND  num_split > 0 && total_size < 3 n  (octree.cpp:69-102), where
// split_size_i = 8 >> #{axes on which the triangle's AABB lies on one side of the midpoint}.
// Two observations make this order-dependent procedure data-parallel without changing its result:
//   (1) every node always receives its triangles in ascending id order, whatever the moment
//       its ancestors split, so a node's final content is the ordered list L of all triangles
//       whose AABB touches its closed box (can_contain, node.cpp:112-120);
//   (2) total_size < 3n  <=>  sum_{i<n} (split_size_i - 3) < 0, and a negative sum already
//       implies num_split > 0. A node is therefore internal iff some prefix of L of length
//       n >= 50 has a negative running sum of (split_size_i - 3).
// So the tree is built level by level: one CTA per node scans its list (block-wide prefix sum),
// decides, counts the triangles going to each of the 8 children, and after a device-wide scan
// of those counts a second pass scatters the ids with a stable (order-preserving) ballot rank.
// Node boxes are dyadic fractions of [-101,101] and are exact in FP64, so midpoints equal the
// reference's (lo+hi)/2 bit for bit.
#include "common.cuh"

namespace msm {

// ------------------------------------------------------------------------------------------
// exclusive scan (int32), reduce-then-scan over 2048-element tiles, recursive on tile sums
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? smem_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nwarp) smem_warp[lane] = winc - w;   // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;           // block total
    }
    __syncthreads();
    const int res = inc - v + smem_warp[warp];
    block_total = smem_warp[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const int* __restrict__ in, int n, int* __restrict__ sums) {
    __shared__ int sw[33];
    const int base = blockIdx.x * kScanTile;
    int acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int i = base + k * kScanThreads + threadIdx.x;
        if (i < n) acc += in[i];
    }
    int total;
    block_exclusive_scan(acc, sw, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int n,
                                                             const int* __restrict__ tile_off, int* __restrict__ total_out) {
    __shared__ int sw[33];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        acc += v[k];
    }
    int total;
    int off = block_exclusive_scan(acc, sw, total) + (tile_off ? tile_off[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = off;
        off += v[k];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = off;
}

msmgpu_status exclusive_scan_i32(const int* d_in, int* d_out, int n, int* d_total, cudaStream_t s) {
    if (n <= 0) {
        if (d_total) MSM_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), s));
        return MSMGPU_OK;
    }
    const int tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        k_scan_tiles<<<1, kScanThreads, 0, s>>>(d_in, d_out, n, nullptr, d_total);
        MSM_LAUNCH_CHECK();
        return MSMGPU_OK;
    }
    DevBuf<int> sums, offs;
    MSM_CUDA(sums.alloc(tiles, s));
    MSM_CUDA(offs.alloc(tiles, s));
    k_scan_tile_sums<<<tiles, kScanThreads, 0, s>>>(d_in, n, sums.p);
    MSM_LAUNCH_CHECK();
    MSM_TRY(exclusive_scan_i32(sums.p, offs.p, tiles, nullptr, s));
    k_scan_tiles<<<tiles, kScanThreads, 0, s>>>(d_in, d_out, n, offs.p, d_total);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// per-triangle tables
// ------------------------------------------------------------------------------------------
__global__ void k_mesh_tables(int nt, const double* __restrict__ xyz, const int* __restrict__ tri,
                              TriRec* __restrict__ rec, double* __restrict__ aabb, double* __restrict__ cull) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    double lo[3], hi[3], cv[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int v = tri[3 * t + k];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double c = xyz[3 * (size_t)v + a];
            cv[3 * k + a] = c;
            if (k == 0) { lo[a] = c; hi[a] = c; }
            else { // octree.cpp:52-58
                if (c < lo[a]) lo[a] = c;
                if (c > hi[a]) hi[a] = c;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { aabb[6 * (size_t)t + a] = lo[a]; aabb[6 * (size_t)t + 3 + a] = hi[a]; }
    TriRec r;
    make_trirec(V3{cv[0], cv[1], cv[2]}, V3{cv[3], cv[4], cv[5]}, V3{cv[6], cv[7], cv[8]}, r);
    rec[t] = r;
    double c4[4];
    make_cull(V3{cv[0], cv[1], cv[2]}, V3{cv[3], cv[4], cv[5]}, V3{cv[6], cv[7], cv[8]}, c4);
    reinterpret_cast<double2*>(cull)[2 * (size_t)t] = make_double2(c4[0], c4[1]);
    reinterpret_cast<double2*>(cull)[2 * (size_t)t + 1] = make_double2(c4[2], c4[3]);
}

msmgpu_status mesh_refresh_tables(msmgpu_mesh* m) {
    if (m->nt == 0) return MSMGPU_OK;
    k_mesh_tables<<<(m->nt + 255) / 256, 256, 0, m->ctx->stream>>>(m->nt, m->xyz.p, m->tri.p, m->rec.p, m->aabb.p, m->cull.p);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// level-synchronous build
// ------------------------------------------------------------------------------------------
struct BuildNode {          // build-time only
    double lo[3];           // lower corner (exact dyadic)
    int mesh;               // index into the per-mesh AABB pointer table
    int depth;
};

__device__ __forceinline__ void classify(const double* __restrict__ bb, const double* lo, double half,
                                         int& a_val, unsigned& child_mask) {
    // bb = {lo xyz, hi xyz} of the triangle; node box = [lo, lo+2*half] per axis, mid = lo+half
    int split_size = 8;
    unsigned in0[3], in1[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double b0 = lo[d], b1 = lo[d] + half, b2 = lo[d] + (half + half);
        const double tlo = bb[d], thi = bb[3 + d];
        // node.cpp:79-89 containing_oct: strict '<' against the midpoint, on both corners
        if ((tlo < b1) == (thi < b1)) split_size >>= 1;
        // node.cpp:112-120 can_contain for the lower / upper child (closed intervals)
        in0[d] = !(thi < b0 || tlo > b1);
        in1[d] = !(thi < b1 || tlo > b2);
    }
    a_val = split_size - 3;
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const unsigned x = (c & 4) ? in1[0] : in0[0];
        const unsigned y = (c & 2) ? in1[1] : in0[1];
        const unsigned z = (c & 1) ? in1[2] : in0[2];
        m |= (x & y & z) << c;
    }
    child_mask = m;
}

// One CTA per node of the current level: split decision + per-child triangle counts.
__global__ void k_decide_count(int node_begin, const int4* __restrict__ nodes, const BuildNode* __restrict__ bn,
                               const int* __restrict__ pairs, const double* const* __restrict__ mesh_aabb,
                               double root_half, int* __restrict__ split_flag, int* __restrict__ child_cnt) {
    __shared__ int sw[33];
    __shared__ int s_cnt[8];
    __shared__ int s_found;
    const int li = blockIdx.x;
    const int g = node_begin + li;
    const int4 nd = nodes[g];
    const int cnt = nd.z;
    if (cnt < kMaxTriangles) {       // octree.cpp:69: the test only runs once a leaf holds >= 50
        if (threadIdx.x == 0) split_flag[li] = 0;
        if (threadIdx.x < 8) child_cnt[li * 8 + threadIdx.x] = 0;
        return;
    }
    const BuildNode b = bn[g];
    const double half = ldexp(root_half, -b.depth);
    const double* __restrict__ aabb = mesh_aabb[b.mesh];
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_found = 0;
    __syncthreads();
    int carry = 0, found = 0;
    int my_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int base = 0; base < cnt; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int a = 0; unsigned mask = 0;
        if (i < cnt) {
            const int t = pairs[nd.y + i];
            classify(aabb + 6 * (size_t)t, b.lo, half, a, mask);
        }
        int total;
        const int excl = block_exclusive_scan(a, sw, total);
        const int incl = carry + excl + a;          // running sum over the first (i+1) triangles
        if (i < cnt && i + 1 >= kMaxTriangles && incl < 0) found = 1;
        carry += total;
#pragma unroll
        for (int c = 0; c < 8; ++c) my_cnt[c] += (mask >> c) & 1u;
    }
    // reduce child counts and the decision
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int v = my_cnt[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[c], v);
    }
    if (__any_sync(0xffffffffu, found) && (threadIdx.x & 31) == 0) s_found = 1;
    __syncthreads();
    const int split = s_found;
    if (threadIdx.x == 0) split_flag[li] = split;
    if (threadIdx.x < 8) child_cnt[li * 8 + threadIdx.x] = split ? s_cnt[threadIdx.x] : 0;
}

// One thread per node of the level: create the 8 children of every splitting node.
__global__ void k_make_children(int node_begin, int n_level, int4* __restrict__ nodes, BuildNode* __restrict__ bn,
                                unsigned char* __restrict__ node_depth,
                                const int* __restrict__ split_flag, const int* __restrict__ split_rank,
                                const int* __restrict__ child_cnt, const int* __restrict__ child_off,
                                int next_node_begin, int next_pair_base, double root_half, int node_cap) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_level) return;
    if (!split_flag[li]) return;
    const int g = node_begin + li;
    const int first = next_node_begin + 8 * split_rank[li];
    if (first + 8 > node_cap) return;   // capacity checked on the host from the scan totals
    const BuildNode b = bn[g];
    const double half = ldexp(root_half, -b.depth);
    int4 nd = nodes[g];
    nd.x = first;
    nd.z = 0;                           // clear_triangles(), octree.cpp:129
    nodes[g] = nd;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        BuildNode cb;
        cb.lo[0] = b.lo[0] + ((c & 4) ? half : 0.0);   // node.cpp:99-106: child box = [b[o], b[o+1]]
        cb.lo[1] = b.lo[1] + ((c & 2) ? half : 0.0);
        cb.lo[2] = b.lo[2] + ((c & 1) ? half : 0.0);
        cb.mesh = b.mesh;
        cb.depth = b.depth + 1;
        bn[first + c] = cb;
        node_depth[first + c] = (unsigned char)(b.depth + 1);
        nodes[first + c] = make_int4(-1, next_pair_base + child_off[li * 8 + c], child_cnt[li * 8 + c], g);
    }
}

// One CTA per node: stable scatter of the triangle ids into the children's lists.
__global__ void k_scatter(int node_begin, const int4* __restrict__ nodes, const BuildNode* __restrict__ bn,
                          int* __restrict__ pairs, const double* const* __restrict__ mesh_aabb, double root_half,
                          const int* __restrict__ split_flag, const int* __restrict__ list_start,
                          const int* __restrict__ list_count, const int* __restrict__ child_off, int next_pair_base) {
    __shared__ int s_warp[32][8];
    __shared__ int s_carry[8];
    const int li = blockIdx.x;
    if (!split_flag[li]) return;
    const int g = node_begin + li;
    const BuildNode b = bn[g];
    const double half = ldexp(root_half, -b.depth);
    const double* __restrict__ aabb = mesh_aabb[b.mesh];
    const int start = list_start[li], cnt = list_count[li];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x < 8) s_carry[threadIdx.x] = next_pair_base + child_off[li * 8 + threadIdx.x];
    __syncthreads();
    for (int base = 0; base < cnt; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int a; unsigned mask = 0; int t = -1;
        if (i < cnt) {
            t = pairs[start + i];
            classify(aabb + 6 * (size_t)t, b.lo, half, a, mask);
        }
        int rank[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned bal = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
            rank[c] = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) s_warp[warp][c] = __popc(bal);
        }
        __syncthreads();
        if (threadIdx.x < 8) {          // exclusive prefix over warps for child threadIdx.x, then advance the carry
            int run = s_carry[threadIdx.x];
            for (int w = 0; w < nwarp; ++w) { const int v = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = run; run += v; }
            s_carry[threadIdx.x] = run;
        }
        __syncthreads();
        if (t >= 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if ((mask >> c) & 1u) pairs[s_warp[warp][c] + rank[c]] = t;
        }
        __syncthreads();
    }
}

__global__ void k_save_lists(int node_begin, int n_level, const int4* __restrict__ nodes, int* __restrict__ list_start, int* __restrict__ list_count) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_level) return;
    const int4 nd = nodes[node_begin + li];
    list_start[li] = nd.y;
    list_count[li] = nd.z;
}

__global__ void k_init_roots(int n, int4* nodes, BuildNode* bn, unsigned char* node_depth, int* pairs,
                             const int* __restrict__ root_pair_off, const int* __restrict__ mesh_nt) {
    // root r: list = all triangles of mesh r in id order (octree.cpp:42-61 pushes every triangle into the root)
    const int r = blockIdx.y;
    const int nt = mesh_nt[r];
    const int off = root_pair_off[r];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) pairs[off + i] = i;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        nodes[r] = make_int4(-1, off, nt, -1);
        BuildNode b;
        b.lo[0] = b.lo[1] = b.lo[2] = -kBounds;   // octree.cpp:33-37
        b.mesh = r;
        b.depth = 0;
        bn[r] = b;
        node_depth[r] = 0;
    }
}

msmgpu_status forest_build(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, std::shared_ptr<Forest>& out, std::vector<int>& roots) {
    cudaStream_t s = ctx->stream;
    if (n <= 0) return fail(MSMGPU_ERR_INVALID, "forest_build: no meshes");
    long long total_t = 0;
    std::vector<int> h_nt(n), h_off(n);
    std::vector<const double*> h_aabb(n);
    for (int i = 0; i < n; ++i) {
        if (!meshes[i] || meshes[i]->ctx != ctx) return fail(MSMGPU_ERR_INVALID, "forest_build: mesh from another context");
        h_nt[i] = meshes[i]->nt;
        h_off[i] = (int)total_t;
        h_aabb[i] = meshes[i]->aabb.p;
        total_t += meshes[i]->nt;
    }
    // capacities: the sum of list lengths over all levels is ~10x the triangle count on sphere meshes
    // (leaf duplication ~2.2x); nodes ~0.2 per triangle. Grown and retried on overflow.
    long long pair_cap = 24 * total_t + 4096ll * n;
    long long node_cap = total_t + 4096ll * n;
    const double root_half = kBounds;   // half width of the root cube

    for (int attempt = 0; attempt < 4; ++attempt) {
        if (pair_cap > 0x7fffffffll || node_cap > 0x7fffffffll) return fail(MSMGPU_ERR_CAPACITY, "forest_build: batch too large for 32-bit offsets");
        auto F = std::make_shared<Forest>();
        F->ctx = ctx;
        DevBuf<BuildNode> bn;
        DevBuf<const double*> d_aabb;
        DevBuf<int> d_nt, d_off;
        MSM_CUDA(F->nodes.alloc(node_cap, s));
        MSM_CUDA(F->pairs.alloc(pair_cap, s));
        MSM_CUDA(F->node_depth.alloc(node_cap, s));
        MSM_CUDA(bn.alloc(node_cap, s));
        MSM_CUDA(d_aabb.alloc(n, s));
        MSM_CUDA(d_nt.alloc(n, s));
        MSM_CUDA(d_off.alloc(n, s));
        MSM_CUDA(cudaMemcpyAsync(d_aabb.p, h_aabb.data(), n * sizeof(double*), cudaMemcpyHostToDevice, s));
        MSM_CUDA(cudaMemcpyAsync(d_nt.p, h_nt.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        MSM_CUDA(cudaMemcpyAsync(d_off.p, h_off.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        {
            int max_nt = 1;
            for (int v : h_nt) max_nt = v > max_nt ? v : max_nt;
            dim3 grid((unsigned)std::min((max_nt + 255) / 256, 1024), (unsigned)n);
            k_init_roots<<<grid, 256, 0, s>>>(n, F->nodes.p, bn.p, F->node_depth.p, F->pairs.p, d_off.p, d_nt.p);
            MSM_LAUNCH_CHECK();
        }
        int node_begin = 0, n_level = n;
        int n_nodes = n;
        long long n_pairs = total_t;
        int depth = 0;
        bool overflow = false;
        DevBuf<int> split_flag, split_rank, child_cnt, child_off, list_start, list_count, totals;
        MSM_CUDA(totals.alloc(2, s));
        while (n_level > 0) {
            MSM_CUDA(split_flag.alloc(n_level, s));
            MSM_CUDA(split_rank.alloc(n_level, s));
            MSM_CUDA(child_cnt.alloc((size_t)n_level * 8, s));
            MSM_CUDA(child_off.alloc((size_t)n_level * 8, s));
            MSM_CUDA(list_start.alloc(n_level, s));
            MSM_CUDA(list_count.alloc(n_level, s));
            // wide CTAs while lists are long (top levels), narrow ones for the many small deep nodes
            const long long avg = (n_pairs - (long long)F->n_pairs) / std::max(n_level, 1);
            (void)avg;
            const int tb = depth <= 2 ? 1024 : (depth <= 4 ? 256 : 128);
            k_decide_count<<<n_level, tb, 0, s>>>(node_begin, F->nodes.p, bn.p, F->pairs.p, d_aabb.p, root_half, split_flag.p, child_cnt.p);
            MSM_LAUNCH_CHECK();
            MSM_TRY(exclusive_scan_i32(split_flag.p, split_rank.p, n_level, totals.p, s));
            MSM_TRY(exclusive_scan_i32(child_cnt.p, child_off.p, n_level * 8, totals.p + 1, s));
            int h_tot[2];
            MSM_CUDA(cudaMemcpyAsync(h_tot, totals.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            MSM_CUDA(cudaStreamSynchronize(s));
            const int n_split = h_tot[0];
            const int new_pairs = h_tot[1];
            if (n_split == 0) break;
            if ((long long)n_nodes + 8ll * n_split > node_cap || n_pairs + new_pairs > pair_cap) { overflow = true; break; }
            k_save_lists<<<(n_level + 255) / 256, 256, 0, s>>>(node_begin, n_level, F->nodes.p, list_start.p, list_count.p);
            MSM_LAUNCH_CHECK();
            k_make_children<<<(n_level + 255) / 256, 256, 0, s>>>(node_begin, n_level, F->nodes.p, bn.p, F->node_depth.p, split_flag.p,
                                                                 split_rank.p, child_cnt.p, child_off.p, n_nodes, (int)n_pairs, root_half,
                                                                 (int)node_cap);
            MSM_LAUNCH_CHECK();
            k_scatter<<<n_level, tb, 0, s>>>(node_begin, F->nodes.p, bn.p, F->pairs.p, d_aabb.p, root_half, split_flag.p, list_start.p,
                                             list_count.p, child_off.p, (int)n_pairs);
            MSM_LAUNCH_CHECK();
            node_begin = n_nodes;
            n_level = 8 * n_split;
            n_nodes += n_level;
            n_pairs += new_pairs;
            ++depth;
            if (depth > 40) return fail(MSMGPU_ERR_CAPACITY, "forest_build: depth limit (degenerate mesh?)");
        }
        if (overflow) { pair_cap *= 2; node_cap *= 2; continue; }
        F->n_nodes = n_nodes;
        F->n_pairs = (int)n_pairs;
        F->depth = depth;
        roots.resize(n);
        for (int i = 0; i < n; ++i) roots[i] = i;
        out = F;
        return MSMGPU_OK;
    }
    return fail(MSMGPU_ERR_CAPACITY, "forest_build: capacity retries exhausted");
}

} // namespace msm
