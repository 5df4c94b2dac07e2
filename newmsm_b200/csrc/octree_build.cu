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
//
// Three constructions share those invariants and give the same tree (tests/test_gpu_parity.py: test_octree_top_phase_*):
//   * the TOP PHASE (k_top_*): the first D0 levels of a dense mesh in one pass over its triangles, from the number of triangles that
//     touch each cell of the 8^d lattices — a sufficient form of the split rule that needs no list order (note above k_top_count);
//   * k_level_fused: one kernel per level for levels with short lists (a warp per node, the prefix rule by warp scan);
//   * the chunked level pass (k_chunk_stats / k_node_combine / k_scatter_chunk): teams of threads per list chunk, for long lists
//     and for levels below the depth-18 lattice.
#include "common.cuh"

#include <algorithm>
#include <chrono>
#include <climits>

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

// second kernel of the two-launch form: every CTA first adds up the sums of the tiles before its own (at most kScanTile of them,
// kScanItems per thread) instead of reading them from a third, recursive launch
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles_from_sums(const int* __restrict__ in, int* __restrict__ out, int n,
                                                                       const int* __restrict__ tile_sums, int* __restrict__ total_out) {
    __shared__ int sw[33];
    int before = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += kScanThreads) before += tile_sums[t];
    int tile_off;
    block_exclusive_scan(before, sw, tile_off);   // tile_off = block total = sum of the sums of tiles 0 .. blockIdx.x - 1
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        acc += v[k];
    }
    int total;
    int off = block_exclusive_scan(acc, sw, total) + tile_off;
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
    k_scan_tile_sums<<<tiles, kScanThreads, 0, s>>>(d_in, n, sums.p);
    MSM_LAUNCH_CHECK();
    if (tiles <= kScanTile) {   // up to 4 M elements: two launches
        k_scan_tiles_from_sums<<<tiles, kScanThreads, 0, s>>>(d_in, d_out, n, sums.p, d_total);
        MSM_LAUNCH_CHECK();
        return MSMGPU_OK;
    }
    MSM_CUDA(offs.alloc(tiles, s));
    MSM_TRY(exclusive_scan_i32(sums.p, offs.p, tiles, nullptr, s));
    k_scan_tiles<<<tiles, kScanThreads, 0, s>>>(d_in, d_out, n, offs.p, d_total);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// Triangle boxes on the octree's own lattice. Every bound the build compares a triangle's AABB with (node lower corner, midpoint,
// upper corner; node.cpp:79-120) is a line g(k) = -101 + k * 202 / 2^18 of the depth-18 lattice of the root cube, exactly
// representable in double. For a coordinate x let q = max{k : g(k) <= x} (-1 below the cube, 2^18 at or above its upper face) and
// e = (g(q) == x). Then  x < g(Q) <=> q < Q  and  x > g(Q) <=> q > Q or (q == Q and not e): the FP64 comparisons of `classify`
// become integer comparisons with identical outcomes, and the 48-byte AABB gather of the level passes shrinks to 16 bytes
// (lower corners: q + 1 and e in 20 bits; upper corners only ever appear on the left of '<': q + 1 in 19 bits). Nodes deeper than
// 17 levels (cells below 1.5e-3 on a radius-100 sphere) use the double-precision path.
// ------------------------------------------------------------------------------------------
constexpr int kGridBits = 18;
constexpr int kGridMaxDepth = kGridBits - 1;
__host__ __device__ __forceinline__ double grid_line(int k) { return -kBounds + (double)k * (2.0 * kBounds / (double)(1 << kGridBits)); }
__device__ __forceinline__ int grid_floor(double x, bool& exact) {
    exact = false;
    if (!(x >= -kBounds)) return -1;
    if (x >= kBounds) { exact = x == kBounds; return 1 << kGridBits; }
    int k = (int)((x + kBounds) * ((double)(1 << kGridBits) / (2.0 * kBounds)));   // estimate only: corrected against the exact lines below
    k = max(0, min(k, (1 << kGridBits) - 1));
    while (grid_line(k) > x) --k;          // the estimate is off by at most one line
    while (grid_line(k + 1) <= x) ++k;
    exact = grid_line(k) == x;
    return k;
}
__device__ __forceinline__ uint4 pack_qbox(const double* lo, const double* hi) {
    unsigned l[3], h[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        bool e, eh;
        const int ql = grid_floor(lo[a], e), qh = grid_floor(hi[a], eh);
        l[a] = ((unsigned)(ql + 1) << 1) | (e ? 1u : 0u);   // 20 bits
        h[a] = (unsigned)(qh + 1);                           // 19 bits
    }
    return make_uint4(l[0] | ((h[0] & 0xfffu) << 20), l[1] | ((h[1] & 0xfffu) << 20), l[2] | ((h[2] & 0xfffu) << 20),
                      (h[0] >> 12) | ((h[1] >> 12) << 7) | ((h[2] >> 12) << 14));
}

// ------------------------------------------------------------------------------------------
// per-triangle tables
// ------------------------------------------------------------------------------------------
struct TableJob {
    const double* xyz; const int* tri; TriRec* rec; double* area; float4* cull; uint4* qbox; int nt;
};

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_mesh_tables(const TableJob* __restrict__ jobs) {
    // the 128-byte records of a warp's 32 triangles are contiguous (4 KB): they are transposed through shared memory so that every
    // store instruction of the warp writes 256 consecutive bytes instead of 32 lines at a 128-byte stride
    __shared__ double s_rec[8][32 * 17];
    const TableJob job = jobs[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = t < job.nt;
    double* mine = s_rec[threadIdx.x >> 5];
    if (valid) {
        double lo[3], hi[3], cv[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = job.tri[3 * (size_t)t + k];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double c = job.xyz[3 * (size_t)v + a];
                cv[3 * k + a] = c;
                if (k == 0) { lo[a] = c; hi[a] = c; }
                else { // octree.cpp:52-58
                    if (c < lo[a]) lo[a] = c;
                    if (c > hi[a]) hi[a] = c;
                }
            }
        }
        if (job.qbox) job.qbox[t] = pack_qbox(lo, hi);   // the FP64 box itself is not stored: below the lattice (depth > 17) k_chunk_stats rebuilds it from the record
        if (job.area) job.area[t] = tri_area_cached(V3{cv[0], cv[1], cv[2]}, V3{cv[3], cv[4], cv[5]}, V3{cv[6], cv[7], cv[8]});   // Triangle::area of this geometry
        if (job.rec) {   // (uniform over the launch's job: a mesh either stores its records or defers them, msmgpu_mesh::lazy_rec)
            TriRec r;
            make_trirec(V3{cv[0], cv[1], cv[2]}, V3{cv[3], cv[4], cv[5]}, V3{cv[6], cv[7], cv[8]}, r);
            static_assert(sizeof(TriRec) == 16 * sizeof(double), "TriRec is 16 doubles");
            const double rp[16] = {r.v[0], r.v[1], r.v[2], r.v[3], r.v[4], r.v[5], r.v[6], r.v[7], r.v[8], r.s3[0], r.s3[1], r.s3[2], r.d, r.e12, r.e13, r.e23};
#pragma unroll
            for (int f = 0; f < 16; ++f) mine[lane * 17 + f] = rp[f];
        }
        if (job.cull) {
            double c4[4];
            make_cull(V3{cv[0], cv[1], cv[2]}, V3{cv[3], cv[4], cv[5]}, V3{cv[6], cv[7], cv[8]}, c4);
            job.cull[t] = pack_cull(c4);
        }
    }
    __syncwarp();
    const int t0 = t - lane;   // first triangle of this warp
    if (job.rec && t0 < job.nt) {
        double* out = reinterpret_cast<double*>(job.rec + t0);
        const int n_valid = min(32, job.nt - t0);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int e = k * 32 + lane, rec = e >> 4;
            if (rec < n_valid) out[e] = mine[rec * 17 + (e & 15)];
        }
    }
}

// per-triangle tables of every mesh whose coordinates changed since they were last computed, in ONE launch
msmgpu_status ensure_tables(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes) {
    std::vector<TableJob> jobs;
    int max_nt = 0;
    for (int i = 0; i < n; ++i) {
        msmgpu_mesh* m = meshes[i];
        if (!m->tables_dirty || m->nt == 0) { m->tables_dirty = false; continue; }
        jobs.push_back(TableJob{m->xyz.p, m->tri.p, m->rec.p, m->area_tab.p, m->cull.p, m->qbox.p, m->nt});
        max_nt = std::max(max_nt, m->nt);
        m->tables_dirty = false;
    }
    if (jobs.empty()) return MSMGPU_OK;
    cudaStream_t s = ctx->stream;
    DevBuf<TableJob> d_jobs;
    MSM_CUDA(d_jobs.alloc(jobs.size(), s));
    MSM_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(TableJob), cudaMemcpyHostToDevice, s));   // pageable: staged before return
    switch (tuning_get("tables_minb", "MSMGPU_TABLES_MINB", 5)) {   // resident CTAs per SM; measured 3 / 4 / 5 / 6: tables + forest 3.44 / 3.35 / 3.32 / 3.56 ms (ncu: FP64 pipe 51 % at 35 % occupancy with the compiler's 72 registers)
        case 4: k_mesh_tables<4><<<dim3((unsigned)((max_nt + 255) / 256), (unsigned)jobs.size()), 256, 0, s>>>(d_jobs.p); break;
        case 3: k_mesh_tables<3><<<dim3((unsigned)((max_nt + 255) / 256), (unsigned)jobs.size()), 256, 0, s>>>(d_jobs.p); break;
        case 6: k_mesh_tables<6><<<dim3((unsigned)((max_nt + 255) / 256), (unsigned)jobs.size()), 256, 0, s>>>(d_jobs.p); break;
        default: k_mesh_tables<5><<<dim3((unsigned)((max_nt + 255) / 256), (unsigned)jobs.size()), 256, 0, s>>>(d_jobs.p); break;
    }
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

// the 128-byte query records of a mesh that deferred them (msmgpu_mesh::lazy_rec), for the consumers that read them
msmgpu_status ensure_records(msmgpu_mesh* m) {
    if (m->rec.p || m->nt == 0) return MSMGPU_OK;
    cudaStream_t s = m->ctx->stream;
    MSM_CUDA(m->rec.alloc((size_t)m->nt, s));
    const TableJob job{m->xyz.p, m->tri.p, m->rec.p, nullptr, nullptr, nullptr, m->nt};
    DevBuf<TableJob> d_job;
    MSM_CUDA(d_job.alloc(1, s));
    MSM_CUDA(cudaMemcpyAsync(d_job.p, &job, sizeof(TableJob), cudaMemcpyHostToDevice, s));   // pageable: staged before return
    k_mesh_tables<5><<<dim3((unsigned)((m->nt + 255) / 256), 1u), 256, 0, s>>>(d_job.p);
    MSM_LAUNCH_CHECK();
    return MSMGPU_OK;
}

msmgpu_status mesh_refresh_tables(msmgpu_mesh* m) {
    m->tables_dirty = true;
    return ensure_tables(m->ctx, 1, &m);
}

msmgpu_status mesh_trees(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, msmgpu_octree** out, std::vector<std::unique_ptr<msmgpu_octree>>& owned) {
    std::vector<msmgpu_mesh*> need;
    for (int i = 0; i < n; ++i)
        if (meshes[i]->view || !meshes[i]->own_tree) {
            bool seen = false;
            for (msmgpu_mesh* m : need) seen = seen || m == meshes[i];
            if (!seen) need.push_back(meshes[i]);
        }
    std::vector<msmgpu_octree*> built(need.size(), nullptr);
    if (!need.empty()) MSM_TRY(msmgpu_octree_build_batch(ctx, (int)need.size(), need.data(), built.data()));
    for (size_t k = 0; k < need.size(); ++k) {
        if (need[k]->view) owned.emplace_back(built[k]);
        else need[k]->own_tree = built[k];
    }
    for (int i = 0; i < n; ++i) {
        if (!meshes[i]->view) { out[i] = meshes[i]->own_tree; continue; }
        for (size_t k = 0; k < need.size(); ++k) if (need[k] == meshes[i]) out[i] = built[k];
    }
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

// `classify` on the lattice (see the note on triangle boxes above): identical decisions, integer compares
__device__ __forceinline__ void classify_q(const uint4 qb, const int* L0, int Hh, int& a_val, unsigned& child_mask) {
    int split_size = 8;
    unsigned in0[3], in1[3];
    const unsigned w[3] = {qb.x, qb.y, qb.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int B0 = L0[d], B1 = B0 + Hh, B2 = B1 + Hh;
        const unsigned l20 = w[d] & 0xfffffu;
        const int qlo = (int)(l20 >> 1) - 1;
        const bool elo = l20 & 1u;
        const int qhi = (int)((w[d] >> 20) | (((qb.w >> (7 * d)) & 0x7fu) << 12)) - 1;
        const bool lo_lt_b1 = qlo < B1, hi_lt_b1 = qhi < B1;
        if (lo_lt_b1 == hi_lt_b1) split_size >>= 1;
        const bool lo_gt_b1 = qlo > B1 || (qlo == B1 && !elo);
        const bool lo_gt_b2 = qlo > B2 || (qlo == B2 && !elo);
        in0[d] = !(qhi < B0 || lo_gt_b1);
        in1[d] = !(hi_lt_b1 || lo_gt_b2);
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

// ------------------------------------------------------------------------------------------
// Chunked level pass. A node's list is cut into chunks of K = TPC * IPT triangles; one TEAM of TPC threads
// (a warp, or a whole CTA) owns one (node, chunk), every thread IPT consecutive list positions. That keeps the
// top levels (few nodes, 10^5-triangle lists) and the bottom levels (10^5 nodes, short lists) equally busy.
//   k_chunk_stats   per chunk: sum of a = split_size - 3, the smallest inclusive prefix of a over the chunk's
//                   positions p with p + 1 >= 50 (chunk-relative), and the 8 child counts
//   k_node_combine  per node: walks its chunks in order -> split decision (some prefix of length >= 50 has a
//                   negative running sum), child totals, and each chunk's exclusive child offsets
//   k_scatter_chunk per chunk: stable scatter (list order preserved) into the children's lists
// ------------------------------------------------------------------------------------------
constexpr int kStatInts = 10;   // sum_a, min_prefix, cnt[8]

// node of slot li of the current level: levels made by the level passes are contiguous, the first level after the top phase is a list
__device__ __forceinline__ int level_node(const int* __restrict__ node_map, int node_begin, int li) { return node_map ? __ldg(node_map + li) : node_begin + li; }

template <int TPC>
__device__ __forceinline__ int team_lane() { return TPC == 32 ? (threadIdx.x & 31) : threadIdx.x; }

// exclusive scan + total over a team (warp shuffles for TPC = 32, shared memory otherwise)
template <int TPC, typename T>
__device__ __forceinline__ T team_exclusive_scan(T v, T* smem /* >= 33 entries when TPC > 32 */, T& total) {
    const int lane = threadIdx.x & 31;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (TPC == 32) {
        total = __shfl_sync(0xffffffffu, inc, 31);
        return inc - v;
    } else {
        const int warp = threadIdx.x >> 5, nwarp = TPC >> 5;
        if (lane == 31) smem[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const T w = lane < nwarp ? smem[lane] : T(0);
            T winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const T t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            if (lane < nwarp) smem[lane] = winc - w;
            if (lane == 31) smem[32] = winc;
        }
        __syncthreads();
        const T res = inc - v + smem[warp];
        total = smem[32];
        __syncthreads();
        return res;
    }
}

template <int TPC>
__device__ __forceinline__ int team_min(int v, int* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (TPC == 32) return v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = TPC >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    int r = smem[0];
    for (int w = 1; w < nwarp; ++w) r = min(r, smem[w]);
    __syncthreads();
    return r;
}

template <int TPC, int IPT>
__global__ void __launch_bounds__(TPC == 32 ? 256 : TPC) k_chunk_stats(int node_begin, int n_level, int max_chunks, const int4* __restrict__ nodes,
                                                                        const BuildNode* __restrict__ bn, const int* __restrict__ pairs,
                                                                        const TriRec* const* __restrict__ mesh_rec,
                                                                        const uint4* const* __restrict__ mesh_qbox, double root_half,
                                                                        int* __restrict__ stats, unsigned char* __restrict__ pmask, int level_base,
                                                                        const int* __restrict__ live, int n_live, const int* __restrict__ node_map) {
    constexpr int K = TPC * IPT;
    constexpr int TEAMS = TPC == 32 ? 8 : 1;
    __shared__ int s_scan[TPC == 32 ? 1 : 33];
    __shared__ int s_red[TPC == 32 ? 1 : 32];
    const int slot = blockIdx.x * TEAMS + (TPC == 32 ? (threadIdx.x >> 5) : 0);   // live nodes on grid.x (no 65535 limit)
    const int chunk = blockIdx.y;
    const bool node_ok = slot < n_live;
    const int li = node_ok ? live[slot] : 0;    // only nodes that can still split (>= 50 triangles) get a team; the others are final leaves
    int4 nd = make_int4(0, 0, 0, 0);
    const int gnode = node_ok ? level_node(node_map, node_begin, li) : 0;
    if (node_ok) nd = nodes[gnode];
    const int cnt = nd.z;
    const bool busy = node_ok && cnt >= kMaxTriangles && chunk * K < cnt;   // octree.cpp:69: the test only runs from 50 triangles on
    if (TPC == 32) { if (!busy) return; }                                  // warp-uniform
    else if (!busy) return;                                                // CTA-uniform
    const BuildNode b = bn[gnode];
    const double half = ldexp(root_half, -b.depth);
    const TriRec* __restrict__ rec = mesh_rec[b.mesh];
    const uint4* __restrict__ qbox = mesh_qbox[b.mesh];
    const bool on_grid = b.depth <= kGridMaxDepth;          // uniform over the team
    int L0[3] = {0, 0, 0}, Hh = 0;
    if (on_grid) {
        const double h = 2.0 * kBounds / (double)(1 << kGridBits);
#pragma unroll
        for (int d = 0; d < 3; ++d) L0[d] = (int)((b.lo[d] + kBounds) / h);   // exact: the corner is a lattice line
        Hh = 1 << (kGridMaxDepth - b.depth);
    }
    const int tl = team_lane<TPC>();
    const int p0 = chunk * K + tl * IPT;
    int a[IPT];
    int local_sum = 0;
    unsigned long long lo4 = 0, hi4 = 0;   // per-child counts of this thread's entries, 16 bits per child (a chunk holds K <= 8192 < 2^16 entries)
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        a[k] = 0;
        const int p = p0 + k;
        if (p < cnt) {
            unsigned mask;
            const int tid = __ldg(pairs + nd.y + p);
            if (on_grid) classify_q(__ldg(qbox + tid), L0, Hh, a[k], mask);
            else {   // deeper than the lattice: the triangle's FP64 box from its record, same min / max as k_mesh_tables (octree.cpp:46-59)
                const double* v = rec[tid].v;
                double bb[6];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double lo = v[d], hi = v[d];
                    if (v[3 + d] < lo) lo = v[3 + d];
                    if (v[3 + d] > hi) hi = v[3 + d];
                    if (v[6 + d] < lo) lo = v[6 + d];
                    if (v[6 + d] > hi) hi = v[6 + d];
                    bb[d] = lo; bb[3 + d] = hi;
                }
                classify(bb, b.lo, half, a[k], mask);
            }
            pmask[(size_t)(nd.y - level_base) + p] = (unsigned char)mask;   // kept for k_scatter_chunk: the 48-byte AABB is gathered once per level, not twice
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                lo4 += (unsigned long long)((mask >> c) & 1u) << (16 * c);
                hi4 += (unsigned long long)((mask >> (4 + c)) & 1u) << (16 * c);
            }
        }
        local_sum += a[k];
    }
    int total;
    const int excl = team_exclusive_scan<TPC, int>(local_sum, s_scan, total);
    int run = excl, mn = INT_MAX;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        run += a[k];
        const int p = p0 + k;
        if (p < cnt && p + 1 >= kMaxTriangles) mn = min(mn, run);
    }
    mn = team_min<TPC>(mn, s_red);
    int* out = stats + ((size_t)li * max_chunks + chunk) * kStatInts;
    // the eight child counts of the chunk: packed warp reduction, then one combine over the CTA's warps (a single barrier)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo4 += __shfl_xor_sync(0xffffffffu, lo4, o);
        hi4 += __shfl_xor_sync(0xffffffffu, hi4, o);
    }
    if (TPC == 32) {
        if (tl < 8) out[2 + tl] = (int)(((tl < 4 ? lo4 : hi4) >> (16 * (tl & 3))) & 0xffffull);
    } else {
        __shared__ unsigned long long s_part[2 * (TPC / 32)];
        const int warp = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { s_part[2 * warp] = lo4; s_part[2 * warp + 1] = hi4; }
        __syncthreads();
        if (threadIdx.x < 8) {
            unsigned long long acc = 0;
            for (int w = 0; w < TPC / 32; ++w) acc += s_part[2 * w + (threadIdx.x < 4 ? 0 : 1)];
            out[2 + threadIdx.x] = (int)((acc >> (16 * (threadIdx.x & 3))) & 0xffffull);
        }
    }
    if (tl == 0) { out[0] = total; out[1] = mn; }
}

// one thread per node of the level
__global__ void k_node_combine(int node_begin, int n_level, int max_chunks, int K, const int4* __restrict__ nodes, int* __restrict__ stats,
                               int* __restrict__ split_flag, int* __restrict__ child_cnt, int* __restrict__ max_child_cnt, int* __restrict__ n_live_next,
                               const int* __restrict__ node_map) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_level) return;
    const int cnt = nodes[level_node(node_map, node_begin, li)].z;
    int tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int split = 0;
    if (cnt >= kMaxTriangles) {
        const int nch = (cnt + K - 1) / K;
        long long run = 0;
        for (int ch = 0; ch < nch; ++ch) {
            int* st = stats + ((size_t)li * max_chunks + ch) * kStatInts;
            if (st[1] != INT_MAX && run + st[1] < 0) split = 1;
            run += st[0];
#pragma unroll
            for (int c = 0; c < 8; ++c) { const int v = st[2 + c]; st[2 + c] = tot[c]; tot[c] += v; }   // counts -> exclusive offsets
        }
    }
    split_flag[li] = split;
    int mx = 0, nl = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int v = split ? tot[c] : 0;
        child_cnt[li * 8 + c] = v;
        mx = max(mx, v);
        nl += v >= kMaxTriangles;
    }
    if (mx > 0) atomicMax(max_child_cnt, mx);   // longest list of the next level (picks its team width)
    if (nl > 0) atomicAdd(n_live_next, nl);     // children that may split again = the teams of the next level
}

// One thread per node of the level: create the 8 children of every splitting node.
__global__ void k_make_children(int node_begin, int n_level, int4* __restrict__ nodes, BuildNode* __restrict__ bn,
                                unsigned char* __restrict__ node_depth,
                                const int* __restrict__ split_flag, const int* __restrict__ split_rank,
                                const int* __restrict__ child_cnt, const int* __restrict__ child_off,
                                int next_node_begin, int next_pair_base, double root_half, int node_cap,
                                int* __restrict__ live_next, int* __restrict__ live_cursor, const int* __restrict__ node_map) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_level) return;
    if (!split_flag[li]) return;
    const int g = level_node(node_map, node_begin, li);
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
    // work list of the next level: its nodes with >= 50 triangles (any order: every per-node result is stored by node index)
    int nl = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) nl += child_cnt[li * 8 + c] >= kMaxTriangles;
    if (nl > 0) {
        int o = atomicAdd(live_cursor, nl);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (child_cnt[li * 8 + c] >= kMaxTriangles) live_next[o++] = 8 * split_rank[li] + c;
    }
}

template <int TPC, int IPT>
__global__ void __launch_bounds__(TPC == 32 ? 256 : TPC) k_scatter_chunk(int node_begin, int n_level, int max_chunks, const BuildNode* __restrict__ bn,
                                                                          int* __restrict__ pairs, const unsigned char* __restrict__ pmask, int level_base,
                                                                          const int* __restrict__ split_flag,
                                                                          const int* __restrict__ list_start, const int* __restrict__ list_count,
                                                                          const int* __restrict__ child_off, const int* __restrict__ stats,
                                                                          int next_pair_base, const int* __restrict__ live, int n_live) {
    constexpr int K = TPC * IPT;
    constexpr int TEAMS = TPC == 32 ? 8 : 1;
    __shared__ unsigned long long s_scan[TPC == 32 ? 1 : 33];
    const int slot = blockIdx.x * TEAMS + (TPC == 32 ? (threadIdx.x >> 5) : 0);   // live nodes on grid.x (no 65535 limit)
    const int chunk = blockIdx.y;
    if (slot >= n_live) return;
    const int li = live[slot];
    if (!split_flag[li]) return;
    const int cnt = list_count[li];
    if (chunk * K >= cnt) return;
    const int start = list_start[li];
    const int tl = team_lane<TPC>();
    const int p0 = chunk * K + tl * IPT;
    int tri[IPT];
    unsigned mask[IPT];
    unsigned long long lo4 = 0, hi4 = 0;   // per-thread child counts, 16 bits per child (K <= 8192 < 2^16)
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        mask[k] = 0;
        tri[k] = -1;
        const int p = p0 + k;
        if (p < cnt) {
            tri[k] = pairs[start + p];
            mask[k] = pmask[(size_t)(start - level_base) + p];   // child-overlap mask computed by k_chunk_stats for this list position
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                lo4 += (unsigned long long)((mask[k] >> c) & 1u) << (16 * c);
                hi4 += (unsigned long long)((mask[k] >> (4 + c)) & 1u) << (16 * c);
            }
        }
    }
    unsigned long long tot;
    const unsigned long long elo = team_exclusive_scan<TPC, unsigned long long>(lo4, s_scan, tot);
    const unsigned long long ehi = team_exclusive_scan<TPC, unsigned long long>(hi4, s_scan, tot);
    const int* st = stats + ((size_t)li * max_chunks + chunk) * kStatInts;
    int pos[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const unsigned long long e = c < 4 ? elo : ehi;
        pos[c] = next_pair_base + child_off[li * 8 + c] + st[2 + c] + (int)((e >> (16 * (c & 3))) & 0xffffull);
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        if (tri[k] < 0) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if ((mask[k] >> c) & 1u) pairs[pos[c]++] = tri[k];
    }
}


__global__ void k_save_lists(int node_begin, int n_level, const int4* __restrict__ nodes, int* __restrict__ list_start, int* __restrict__ list_count,
                             const int* __restrict__ node_map) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_level) return;
    const int4 nd = nodes[level_node(node_map, node_begin, li)];
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

// ------------------------------------------------------------------------------------------
// Top phase: the first D0 levels of a dense mesh in ONE pass over its triangles instead of D0 level passes.
//
// A node splits when it holds n >= 50 triangles and total_size < 3 n (octree.cpp:69-102, evaluated at every insertion from the
// 50th on). Let m_t be the number of the node's 8 child cubes that triangle t TOUCHES (closed boxes, can_contain). Per axis a
// triangle that does not straddle the midpoint touches one half -- or both when its lower corner lies exactly on the midpoint --
// so m_t >= split_size_t and  sum over the children of their triangle counts = sum_t m_t >= total_size.  Hence
//        count(node) >= 50  and  sum_c count(child c) < 3 count(node)
// is SUFFICIENT for the split (the whole list is one of the prefixes the reference tests), and it needs no order and nothing but
// the number of triangles touching each cell of the 8^d lattices of the root cube, d <= D0:
//   k_top_count     every triangle adds 1 to the cells it touches at every depth (32-bit reds; depths 1..3 through block-local
//                   shared-memory counters, which every block hits)
//   k_top_decide    per dense cell: the sufficient condition -> top bit of its counter
//   k_top_children  depth by depth over the dense cells: a decided cell gets its 8 children (node records, boxes, list space
//                   from one atomic cursor per warp)
//   k_top_fill      every triangle appends its id to the list of each touched cell that exists and was not split
//   k_top_sort_*    the lists are put in ascending id order = the order the reference's sequential insertion produces
// A cell with count >= 50 that fails the sufficient test is simply not decided here: it keeps its list and goes, like the depth-D0
// cells, on the work list of the exact level passes below, which evaluate the prefix rule. Decisions are therefore the reference's
// in every case; only the node numbering differs (atomic cursors), which no result depends on. D0 is chosen per mesh from its
// triangle count (cells of depth D0-1 hold >= 64 triangles on average on a sphere); knob "build_top" (MSMGPU_BUILD_TOP): -1 auto,
// 0 off, k force D0 = k.
// ------------------------------------------------------------------------------------------
constexpr int kTopMaxDepth = 6;
constexpr int kTopTrisPerBlock = 1024;   // k_top_count: 256 threads x 4 triangles
constexpr int kTopNone = -0x40000000;    // nid entries at or below this: the cell has no node
__host__ __device__ __forceinline__ long long top_cells_before(int d) { return ((1ll << (3 * d)) - 1) / 7; }   // cells of depths < d

struct TopJob { const uint4* qbox; int nt; int d0; const int* perm; };   // perm: processing order of the triangles (optional)

__device__ __forceinline__ unsigned top_spread(unsigned v) {
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8) | ((v & 32u) << 10);
}
__device__ __forceinline__ unsigned top_cell(int ix, int iy, int iz) {   // child digit per level = 4 x + 2 y + z (classify's numbering)
    return (top_spread((unsigned)ix) << 2) | (top_spread((unsigned)iy) << 1) | top_spread((unsigned)iz);
}
struct TopBox { int qlo[3], qhi[3]; bool elo[3]; };
__device__ __forceinline__ TopBox top_unpack(const uint4 qb) {
    TopBox b;
    const unsigned w[3] = {qb.x, qb.y, qb.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const unsigned l20 = w[d] & 0xfffffu;
        b.qlo[d] = (int)(l20 >> 1) - 1;
        b.elo[d] = l20 & 1u;
        b.qhi[d] = (int)((w[d] >> 20) | (((qb.w >> (7 * d)) & 0x7fu) << 12)) - 1;
    }
    return b;
}
// cells [lo, hi] of depth d whose closed interval the triangle's [tlo, thi] touches on one axis (node.cpp:112-120 on the lattice)
__device__ __forceinline__ void top_range(const TopBox& b, int axis, int d, int& lo, int& hi) {
    const int sh = kGridBits - d;
    hi = min(b.qhi[axis] >> sh, (1 << d) - 1);
    const int q = b.qlo[axis];
    if (q < 0) lo = 0;
    else {
        int k = q >> sh;
        if (b.elo[axis] && (q & ((1 << sh) - 1)) == 0) --k;   // the lower corner lies exactly on the line: the cell below touches it too
        lo = max(k, 0);
    }
}

// The touched cells of ALL depths <= d0 from the ranges at depth d0: on every axis the range at depth d is the range at depth d0
// shifted right by (d0 - d) (the "corner exactly on a line" correction of top_range survives the shift: a lattice line of depth d is
// one of depth d0), and a dilated coordinate (bit b at position 3b) shifts by 3 (d0 - d). Common case: at most two cells per axis.
struct TopSpan {
    unsigned lo[3], hi[3];   // dilated coordinates at depth d0, already moved to their bit lane (x << 2, y << 1, z)
    bool narrow;             // every axis touches one or two cells at depth d0
    bool empty;              // the triangle lies outside the root cube on some axis
};
__device__ __forceinline__ TopSpan top_span(const TopBox& b, int d0) {
    TopSpan sp;
    sp.narrow = true;
    sp.empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int lo, hi;
        top_range(b, a, d0, lo, hi);
        if (hi < lo) sp.empty = true;
        if (hi - lo > 1) sp.narrow = false;
        sp.lo[a] = top_spread((unsigned)max(lo, 0)) << (2 - a);
        sp.hi[a] = top_spread((unsigned)max(hi, 0)) << (2 - a);
    }
    return sp;
}
// The up-to-8 cells of depth d (= d0 - up) a narrow span touches, as 8 slots (x, y, z each low / high): slot c exists iff every axis
// it takes the high cell on really has a second cell.
struct TopCells { unsigned x0, x1, y0, y1, z0, z1; bool dx, dy, dz; };
__device__ __forceinline__ TopCells top_cells(const TopSpan& sp, int up) {
    const int s3 = 3 * up;   // (a dilated coordinate shifted by a multiple of 3 stays in its bit lane)
    TopCells c;
    c.x0 = sp.lo[0] >> s3; c.x1 = sp.hi[0] >> s3;
    c.y0 = sp.lo[1] >> s3; c.y1 = sp.hi[1] >> s3;
    c.z0 = sp.lo[2] >> s3; c.z1 = sp.hi[2] >> s3;
    c.dx = c.x1 != c.x0; c.dy = c.y1 != c.y0; c.dz = c.z1 != c.z0;
    return c;
}
__device__ __forceinline__ bool top_slot(const TopCells& c, int slot, unsigned& cell) {
    cell = ((slot & 4) ? c.x1 : c.x0) | ((slot & 2) ? c.y1 : c.y0) | ((slot & 1) ? c.z1 : c.z0);
    return (!(slot & 4) || c.dx) && (!(slot & 2) || c.dy) && (!(slot & 1) || c.dz);
}
template <typename F>
__device__ __forceinline__ void top_for_cells(const TopCells& c, F&& f) {
    f(c.x0 | c.y0 | c.z0);
    if (c.dz) f(c.x0 | c.y0 | c.z1);
    if (c.dy) { f(c.x0 | c.y1 | c.z0); if (c.dz) f(c.x0 | c.y1 | c.z1); }
    if (c.dx) {
        f(c.x1 | c.y0 | c.z0);
        if (c.dz) f(c.x1 | c.y0 | c.z1);
        if (c.dy) { f(c.x1 | c.y1 | c.z0); if (c.dz) f(c.x1 | c.y1 | c.z1); }
    }
}
// Warp-aggregated "+1 per (lane, cell)": the lanes of a warp that name the same cell in the same slot send ONE atomic. All 32 lanes
// call this together; `on` = this lane takes part. With the triangles processed along a space-filling curve (TopJob::perm) a warp's
// 32 triangles fall into a handful of cells, so the number of atomics drops by an order of magnitude.
template <typename F>
__device__ __forceinline__ void top_warp_add(bool on, const TopCells& c, F&& add /* (cell, count) by the group's leader */) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int slot = 0; slot < 8; ++slot) {
        unsigned cell;
        const bool has = on && top_slot(c, slot, cell);
        const unsigned any = __ballot_sync(0xffffffffu, has);
        if (any == 0u) continue;   // warp-uniform
        if (has) {
            const unsigned grp = __match_any_sync(any, cell);
            if ((int)__ffs(grp) - 1 == lane) add(cell, __popc(grp));
        }
    }
}

__global__ void __launch_bounds__(256) k_top_count(const TopJob* __restrict__ jobs, unsigned* __restrict__ cnt, long long cells_per_mesh, int agg) {
    __shared__ unsigned s_cnt[8 + 64 + 512];   // depths 1..3
    const TopJob job = jobs[blockIdx.y];
    const int t0 = blockIdx.x * kTopTrisPerBlock;
    if (t0 >= job.nt || job.d0 <= 0) return;
    for (int i = threadIdx.x; i < 8 + 64 + 512; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    unsigned* g = cnt + (size_t)blockIdx.y * cells_per_mesh;
    for (int k0 = t0; k0 < min(t0 + kTopTrisPerBlock, job.nt); k0 += blockDim.x) {   // (uniform trip count: the warp stays converged)
        const int k = k0 + threadIdx.x;
        const bool valid = k < min(t0 + kTopTrisPerBlock, job.nt);
        const int t = valid ? (job.perm ? __ldg(job.perm + k) : k) : 0;
        TopBox b{};
        TopSpan sp{};
        sp.empty = true; sp.narrow = true;
        if (valid) { b = top_unpack(__ldg(job.qbox + t)); sp = top_span(b, job.d0); }
        const bool fast = valid && !sp.empty && sp.narrow;   // (an empty span lies outside the root cube: only the root holds it)
        int base = (int)top_cells_before(job.d0);
        for (int d = job.d0; d >= 1; base = (base - 1) >> 3, --d) {   // (8^d - 1) / 7 -> (8^(d-1) - 1) / 7
            const TopCells c = top_cells(sp, job.d0 - d);
            if (!agg) {
                if (fast) {
                    if (d <= 3) top_for_cells(c, [&](unsigned cell) { atomicAdd(&s_cnt[base - 1 + cell], 1u); });
                    else top_for_cells(c, [&](unsigned cell) { atomicAdd(g + base + cell, 1u); });
                }
            } else if (agg == 2) {   // the lower-corner cell (every lane has one) aggregated over the warp, the straddlers' other cells one by one
                const unsigned on = __ballot_sync(0xffffffffu, fast);
                if (fast) {
                    const unsigned c0 = c.x0 | c.y0 | c.z0;
                    const unsigned grp = __match_any_sync(on, c0);
                    if ((int)__ffs(grp) - 1 == (int)(threadIdx.x & 31)) {
                        if (d <= 3) atomicAdd(&s_cnt[base - 1 + c0], (unsigned)__popc(grp));
                        else atomicAdd(g + base + c0, (unsigned)__popc(grp));
                    }
                    if (c.dx | c.dy | c.dz) {
                        bool first = true;
                        if (d <= 3) top_for_cells(c, [&](unsigned cell) { if (!first) atomicAdd(&s_cnt[base - 1 + cell], 1u); first = false; });
                        else top_for_cells(c, [&](unsigned cell) { if (!first) atomicAdd(g + base + cell, 1u); first = false; });
                    }
                }
            } else if (d <= 3) top_warp_add(fast, c, [&](unsigned cell, int n) { atomicAdd(&s_cnt[base - 1 + cell], (unsigned)n); });
            else top_warp_add(fast, c, [&](unsigned cell, int n) { atomicAdd(g + base + cell, (unsigned)n); });
        }
        if (valid && !sp.empty && !sp.narrow) {
            for (int d = 1; d <= job.d0; ++d) {   // a triangle wider than a cell of depth d0 (coarse meshes at a forced depth)
                int lo[3], hi[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) top_range(b, a, d, lo[a], hi[a]);
                for (int ix = lo[0]; ix <= hi[0]; ++ix)
                    for (int iy = lo[1]; iy <= hi[1]; ++iy)
                        for (int iz = lo[2]; iz <= hi[2]; ++iz) {
                            const unsigned cell = top_cell(ix, iy, iz);
                            if (d <= 3) atomicAdd(&s_cnt[top_cells_before(d) - 1 + cell], 1u);
                            else atomicAdd(g + top_cells_before(d) + cell, 1u);
                        }
            }
        }
    }
    __syncthreads();
    const int n_sm = (int)top_cells_before(min(job.d0, 3) + 1) - 1;
    for (int i = threadIdx.x; i < n_sm; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(g + 1 + i, s_cnt[i]);
}

// the sufficient condition, for every dense cell of the depths below the mesh's D0 -> top bit of the counter
__global__ void k_top_decide(int n, int dmax, const TopJob* __restrict__ jobs, unsigned* __restrict__ cnt, long long cells_per_mesh) {
    const long long per = top_cells_before(dmax);   // cells of depths 0 .. dmax-1
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= per * n) return;
    const int m = (int)(gi / per);
    const long long c = gi % per;
    int d = 0;
    while (top_cells_before(d + 1) <= c) ++d;
    const TopJob job = jobs[m];
    if (d >= job.d0) return;
    unsigned* g = cnt + (size_t)m * cells_per_mesh;
    const unsigned cell = (unsigned)(c - top_cells_before(d));
    const unsigned own = d == 0 ? (unsigned)job.nt : g[c];   // the root receives every triangle (octree.cpp:42-61)
    if (d == 0) g[0] = own;
    if (own < (unsigned)kMaxTriangles) return;
    const unsigned* ch = g + top_cells_before(d + 1) + ((size_t)cell << 3);
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += ch[k] & 0x7fffffffu;   // (a child's own flag may already be set)
    if (sum < 3ull * own) g[c] = own | 0x80000000u;
}

struct TopCursors { unsigned long long pairs; int nodes, live, max_live_cnt, max_list, overflow, pad; };

// nid entry of a dense cell: <= kTopNone = no node; -2 - id = node id, split by the top phase; >= 0: a node that keeps its list --
// its id at the depths below the mesh's D0 (write cursor in `fillc`), and at depth D0 directly the write cursor of its list
__device__ __forceinline__ int top_nid_split(int id) { return -2 - id; }

// one atomic per warp on a cursor every lane advances: returns this lane's start
__device__ __forceinline__ int warp_reserve(int* cursor, int want) {
    const int lane = threadIdx.x & 31;
    int inc = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + inc - want;
}
__device__ __forceinline__ unsigned long long warp_reserve64(unsigned long long* cursor, int want) {
    const int lane = threadIdx.x & 31;
    int inc = want;   // a warp's lists stay far below 2^31 entries
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned long long base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + (unsigned long long)(inc - want);
}

// roots: node m = root of mesh m
__global__ void k_top_roots(int n, const TopJob* __restrict__ jobs, const unsigned* __restrict__ cnt, long long cells_per_mesh,
                            int* __restrict__ nid, int4* __restrict__ nodes, BuildNode* __restrict__ bn, unsigned char* __restrict__ node_depth,
                            TopCursors* __restrict__ cur, int* __restrict__ live, int live_cap) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const TopJob job = jobs[m];
    BuildNode b;
    b.lo[0] = b.lo[1] = b.lo[2] = -kBounds;
    b.mesh = m;
    b.depth = 0;
    bn[m] = b;
    node_depth[m] = 0;
    if (job.d0 > 0 && (cnt[(size_t)m * cells_per_mesh] >> 31)) {
        nid[(size_t)m * cells_per_mesh] = top_nid_split(m);
        nodes[m] = make_int4(-1, 0, 0, -1);
        return;
    }
    nid[(size_t)m * cells_per_mesh] = kTopNone;   // the root keeps the identity list (k_top_root_lists); nothing is appended to it
    const int off = (int)atomicAdd(&cur->pairs, (unsigned long long)job.nt);
    nodes[m] = make_int4(-1, off, job.nt, -1);
    if (job.nt >= kMaxTriangles) {
        const int o = atomicAdd(&cur->live, 1);
        if (o < live_cap) live[o] = m; else cur->overflow = 1;
        atomicMax(&cur->max_live_cnt, job.nt);
    }
}
__global__ void k_top_root_lists(const int* __restrict__ nid, long long cells_per_mesh, const int4* __restrict__ nodes, int* __restrict__ pairs) {
    if (nid[(size_t)blockIdx.y * cells_per_mesh] > kTopNone) return;   // the root was split
    const int4 nd = nodes[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nd.z; i += gridDim.x * blockDim.x) pairs[nd.y + i] = i;
}

// one thread per dense cell of depth d: the 8 children of every cell the top phase splits
__global__ void __launch_bounds__(256) k_top_children(int n, int d, const TopJob* __restrict__ jobs, const unsigned* __restrict__ cnt,
                                                      long long cells_per_mesh, int* __restrict__ nid, int* __restrict__ fillc, long long fillc_per_mesh,
                                                      int4* __restrict__ nodes, BuildNode* __restrict__ bn, unsigned char* __restrict__ node_depth,
                                                      TopCursors* __restrict__ cur, int* __restrict__ live, int live_cap, int node_cap, double root_half) {
    const long long per = 1ll << (3 * d);
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = gi < per * n;
    const int m = in_range ? (int)(gi / per) : 0;
    const unsigned cell = in_range ? (unsigned)(gi % per) : 0u;
    const size_t mb = (size_t)m * cells_per_mesh;
    const int code = in_range ? nid[mb + top_cells_before(d) + cell] : kTopNone;
    const bool split = code <= -2 && code > kTopNone;
    const int id = split ? -2 - code : -1;
    const int d0 = split ? jobs[m].d0 : 0;
    const size_t cb = mb + top_cells_before(d + 1) + ((size_t)cell << 3);
    unsigned v[8];
    int list_total = 0, n_live = 0, max_list = 0;
    if (split) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            v[c] = cnt[cb + c];
            if (v[c] >> 31) continue;   // split in its turn (only below D0 - 1: k_top_decide)
            const int k = (int)v[c];
            list_total += k;
            n_live += k >= kMaxTriangles;
            max_list = max(max_list, k);
        }
    }
    // per-warp reservations: node records, list space, work-list slots
    const int first = warp_reserve(&cur->nodes, split ? 8 : 0);
    unsigned long long off = warp_reserve64(&cur->pairs, list_total);
    int lslot = warp_reserve(&cur->live, n_live);
    const int wmax = __reduce_max_sync(0xffffffffu, max_list);
    if ((threadIdx.x & 31) == 0 && wmax > 1) {
        atomicMax(&cur->max_list, wmax);
        if (wmax >= kMaxTriangles) atomicMax(&cur->max_live_cnt, wmax);
    }
    if (!split) return;
    if (first + 8 > node_cap || off + (unsigned long long)list_total > 0x7fffffffull) { cur->overflow = 1; return; }
    const int parent = nodes[id].w;
    nodes[id] = make_int4(first, 0, 0, parent);
    const BuildNode b = bn[id];
    const double half = ldexp(root_half, -d);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        BuildNode ch;
        ch.lo[0] = b.lo[0] + ((c & 4) ? half : 0.0);   // node.cpp:99-106
        ch.lo[1] = b.lo[1] + ((c & 2) ? half : 0.0);
        ch.lo[2] = b.lo[2] + ((c & 1) ? half : 0.0);
        ch.mesh = m;
        ch.depth = d + 1;
        bn[first + c] = ch;
        node_depth[first + c] = (unsigned char)(d + 1);
        if (v[c] >> 31) {
            nid[cb + c] = top_nid_split(first + c);
            nodes[first + c] = make_int4(-1, 0, 0, id);
            continue;
        }
        const int k = (int)v[c];
        nodes[first + c] = make_int4(-1, (int)off, k, id);
        if (d + 1 == d0) nid[cb + c] = (int)off;   // depth D0: the entry is the list's write cursor
        else {
            nid[cb + c] = first + c;
            fillc[(size_t)m * fillc_per_mesh + top_cells_before(d + 1) + ((size_t)cell << 3) + c] = (int)off;
        }
        off += (unsigned long long)k;
        if (k >= kMaxTriangles) {
            if (lslot < live_cap) live[lslot] = first + c; else cur->overflow = 1;
            ++lslot;
        }
    }
}

__global__ void __launch_bounds__(256) k_top_fill(const TopJob* __restrict__ jobs, int* __restrict__ nid, long long cells_per_mesh,
                                                  int* __restrict__ fillc, long long fillc_per_mesh, int* __restrict__ pairs, int agg) {
    const TopJob job = jobs[blockIdx.y];
    if ((long long)blockIdx.x * blockDim.x >= job.nt || job.d0 <= 0) return;   // whole CTA beyond this mesh
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < job.nt;
    const int t = valid ? (job.perm ? __ldg(job.perm + k) : k) : 0;
    int* __restrict__ ids = nid + (size_t)blockIdx.y * cells_per_mesh;
    TopBox b{};
    TopSpan sp{};
    sp.empty = true; sp.narrow = true;
    if (valid) { b = top_unpack(__ldg(job.qbox + t)); sp = top_span(b, job.d0); }
    if (!valid || sp.empty) sp.empty = true;   // outside the root cube: no cell below the root holds it
    // the common case first: every touched cell of depth D0 exists (all its ancestors were split); its nid entry is the list's write
    // cursor. The lanes of a warp that append to the same list reserve their places with one atomic (slot by slot, top_warp_add).
    bool missing = false;
    int* __restrict__ cur0 = ids + top_cells_before(job.d0);
    const int lane = threadIdx.x & 31;
    {
        const bool fast = valid && !sp.empty && sp.narrow;
        const TopCells c = top_cells(sp, 0);
        if (!agg) {
            if (fast) top_for_cells(c, [&](unsigned cell) {
                const int pos = atomicAdd(cur0 + cell, 1);
                if (pos >= 0) pairs[pos] = t;
                else missing = true;
            });
        } else if (agg == 2) {
            const unsigned on = __ballot_sync(0xffffffffu, fast);
            if (fast) {
                const unsigned c0 = c.x0 | c.y0 | c.z0;
                const unsigned grp = __match_any_sync(on, c0);
                const int leader = (int)__ffs(grp) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(cur0 + c0, __popc(grp));
                pos = __shfl_sync(grp, pos, leader);
                if (pos >= 0) pairs[pos + __popc(grp & ((1u << lane) - 1u))] = t;
                else missing = true;
                if (c.dx | c.dy | c.dz) {
                    bool first = true;
                    top_for_cells(c, [&](unsigned cell) {
                        if (!first) {
                            const int p2 = atomicAdd(cur0 + cell, 1);
                            if (p2 >= 0) pairs[p2] = t;
                            else missing = true;
                        }
                        first = false;
                    });
                }
            }
        } else
#pragma unroll
        for (int slot = 0; slot < 8; ++slot) {
            unsigned cell;
            const bool has = fast && top_slot(c, slot, cell);
            const unsigned any = __ballot_sync(0xffffffffu, has);
            if (any == 0u) continue;   // warp-uniform
            if (has) {
                const unsigned grp = __match_any_sync(any, cell);
                const int leader = (int)__ffs(grp) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(cur0 + cell, __popc(grp));
                pos = __shfl_sync(grp, pos, leader);
                if (pos >= 0) pairs[pos + __popc(grp & ((1u << lane) - 1u))] = t;
                else missing = true;   // (the entry of a cell without node stays far below zero)
            }
        }
    }
    if (valid && !sp.empty && !sp.narrow) {
        int lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) top_range(b, a, job.d0, lo[a], hi[a]);
        for (int ix = lo[0]; ix <= hi[0]; ++ix)
            for (int iy = lo[1]; iy <= hi[1]; ++iy)
                for (int iz = lo[2]; iz <= hi[2]; ++iz) {
                    const int pos = atomicAdd(cur0 + top_cell(ix, iy, iz), 1);
                    if (pos >= 0) pairs[pos] = t;
                    else missing = true;
                }
    }
    if (!missing) return;
    int* __restrict__ fc = fillc + (size_t)blockIdx.y * fillc_per_mesh;
    for (int d = 1; d < job.d0; ++d) {   // some ancestor was not split: the list-keeping cells above
        int lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) top_range(b, a, d, lo[a], hi[a]);
        bool any = false;
        for (int ix = lo[0]; ix <= hi[0]; ++ix)
            for (int iy = lo[1]; iy <= hi[1]; ++iy)
                for (int iz = lo[2]; iz <= hi[2]; ++iz) {
                    const long long ci = top_cells_before(d) + top_cell(ix, iy, iz);
                    const int id = ids[ci];
                    if (id <= kTopNone) continue;
                    any = true;
                    if (id >= 0) pairs[atomicAdd(fc + ci, 1)] = t;
                }
        if (!any) break;   // none of the touched cells exists: their descendants do not exist either
    }
}

// ascending id order inside every list the top phase filled = the order the reference's sequential insertion produces.
// Half a warp per node (lists hold ~20 ids), shuffle network of common.cuh on E registers per lane.
template <int E>
__device__ __forceinline__ void half_warp_sort_list(int* __restrict__ list, int cnt, unsigned mask) {
    const int hl = threadIdx.x & 15;
    int v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) { const int e = (r << 4) | hl; v[r] = e < cnt ? list[e] : INT_MAX; }
    half_warp_bitonic<E>(v, mask);
#pragma unroll
    for (int r = 0; r < E; ++r) { const int e = (r << 4) | hl; if (e < cnt) list[e] = v[r]; }
}
__global__ void __launch_bounds__(256) k_top_sort_warp(int first_node, int n_nodes, const int4* __restrict__ nodes, int* __restrict__ pairs) {
    const int id = first_node + blockIdx.x * 16 + (threadIdx.x >> 4);
    const unsigned mask = 0xffffu << (threadIdx.x & 16);
    if (id >= n_nodes) return;       // (whole half-warps leave together)
    const int4 nd = nodes[id];
    if (nd.x != -1 || nd.z <= 1 || nd.z > 256) return;
    int* list = pairs + nd.y;
    if (nd.z <= 16) half_warp_sort_list<1>(list, nd.z, mask);
    else if (nd.z <= 32) half_warp_sort_list<2>(list, nd.z, mask);
    else if (nd.z <= 64) half_warp_sort_list<4>(list, nd.z, mask);
    else if (nd.z <= 128) half_warp_sort_list<8>(list, nd.z, mask);
    else half_warp_sort_list<16>(list, nd.z, mask);
}
constexpr int kTopSortCap = 4096;
__global__ void __launch_bounds__(256) k_top_sort_block(int first_node, int n_nodes, const int4* __restrict__ nodes, int* __restrict__ pairs) {
    __shared__ int buf[kTopSortCap];
    const int id = first_node + blockIdx.x;
    if (id >= n_nodes) return;
    const int4 nd = nodes[id];
    if (nd.x != -1 || nd.z <= 256 || nd.z > kTopSortCap) return;
    int N = 512;
    while (N < nd.z) N <<= 1;
    for (int i = threadIdx.x; i < N; i += blockDim.x) buf[i] = i < nd.z ? pairs[nd.y + i] : INT_MAX;
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const int x = buf[i], y = buf[p];
                    if (((i & k) == 0) == (x > y)) { buf[i] = y; buf[p] = x; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < nd.z; i += blockDim.x) pairs[nd.y + i] = buf[i];
}
__global__ void k_tri_points(int nt, const double* __restrict__ xyz, const int* __restrict__ tri, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int v = tri[3 * (size_t)t];   // the first corner stands for the triangle: only coherence of the order matters
#pragma unroll
    for (int a = 0; a < 3; ++a) out[3 * (size_t)t + a] = xyz[3 * (size_t)v + a];
}
__global__ void k_iota(int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// ------------------------------------------------------------------------------------------
// One level pass in ONE kernel, for levels whose lists are short (the levels below the top phase: <= a few hundred triangles per
// node): a warp owns a node, walks its list in chunks of 32 — lattice classification, running sum of a = split_size - 3 by warp
// scan, the reference's prefix rule (some prefix of length >= 50 with a negative sum, octree.cpp:69-102), per-child counts by
// ballots — and, if the node splits, takes 8 node records and the children's list space from atomic cursors and scatters the ids
// in list order (ballot rank). That replaces k_chunk_stats, k_node_combine, two device scans, k_save_lists, k_make_children and
// k_scatter_chunk (12 launches and their buffers) of the chunked level pass, which stays for long lists and for levels below the
// lattice. Same decisions, same lists; node numbering by cursor.
// ------------------------------------------------------------------------------------------
struct FuseCursors { unsigned long long pairs; int nodes, live, max_cnt, overflow; };
constexpr int kFuseMaxList = 2048;

__global__ void __launch_bounds__(256) k_level_fused(int node_begin, const int* __restrict__ live, int n_live, const int* __restrict__ node_map,
                                                     int4* __restrict__ nodes, BuildNode* __restrict__ bn, unsigned char* __restrict__ node_depth,
                                                     int* __restrict__ pairs, const uint4* const* __restrict__ mesh_qbox,
                                                     unsigned char* __restrict__ pmask, int level_base, double root_half, FuseCursors* __restrict__ cur,
                                                     int* __restrict__ next_nodes, int next_cap, int node_cap, long long pair_cap) {
    const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= n_live) return;                         // warp-uniform
    const int li = live ? live[slot] : slot;
    const int g = level_node(node_map, node_begin, li);
    const int4 nd = nodes[g];
    const int cnt = nd.z;
    if (nd.x >= 0 || cnt < kMaxTriangles) return;       // octree.cpp:69: the test only runs from 50 triangles on
    const BuildNode b = bn[g];
    const uint4* __restrict__ qbox = mesh_qbox[b.mesh];
    int L0[3];
    const double h = 2.0 * kBounds / (double)(1 << kGridBits);
#pragma unroll
    for (int d = 0; d < 3; ++d) L0[d] = (int)((b.lo[d] + kBounds) / h);   // exact: the corner is a lattice line
    const int Hh = 1 << (kGridMaxDepth - b.depth);
    unsigned char* __restrict__ pm = pmask + (size_t)(nd.y - level_base);
    int run = 0, ccnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool split = false;
    for (int p0 = 0; p0 < cnt; p0 += 32) {
        const int p = p0 + lane;
        int a = 0;
        unsigned mask = 0;
        if (p < cnt) {
            classify_q(__ldg(qbox + __ldg(pairs + nd.y + p)), L0, Hh, a, mask);
            pm[p] = (unsigned char)mask;
        }
        int inc = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (p < cnt && p + 1 >= kMaxTriangles && run + inc < 0) split = true;
        run += __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
        for (int c = 0; c < 8; ++c) ccnt[c] += __popc(__ballot_sync(0xffffffffu, (mask >> c) & 1u));
    }
    if (!__any_sync(0xffffffffu, split)) return;
    int total = 0, n_next = 0, cmax = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { total += ccnt[c]; n_next += ccnt[c] >= kMaxTriangles; cmax = max(cmax, ccnt[c]); }
    int first = 0, lslot = 0;
    unsigned long long off = 0;
    if (lane == 0) {
        first = atomicAdd(&cur->nodes, 8);
        off = atomicAdd(&cur->pairs, (unsigned long long)total);
        if (n_next) lslot = atomicAdd(&cur->live, n_next);
        atomicMax(&cur->max_cnt, cmax);
        if (first + 8 > node_cap || off + (unsigned long long)total > (unsigned long long)pair_cap || lslot + n_next > next_cap) cur->overflow = 1;
    }
    first = __shfl_sync(0xffffffffu, first, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    lslot = __shfl_sync(0xffffffffu, lslot, 0);
    if (first + 8 > node_cap || off + (unsigned long long)total > (unsigned long long)pair_cap || lslot + n_next > next_cap) return;
    int coff[8];
    {
        int acc = (int)off;
#pragma unroll
        for (int c = 0; c < 8; ++c) { coff[c] = acc; acc += ccnt[c]; }
    }
    if (lane == 0) nodes[g] = make_int4(first, nd.y, 0, nd.w);   // clear_triangles(), octree.cpp:129
    {
        const double half = ldexp(root_half, -b.depth);
        int ls = lslot;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (lane == c) {
                BuildNode cb;
                cb.lo[0] = b.lo[0] + ((c & 4) ? half : 0.0);   // node.cpp:99-106
                cb.lo[1] = b.lo[1] + ((c & 2) ? half : 0.0);
                cb.lo[2] = b.lo[2] + ((c & 1) ? half : 0.0);
                cb.mesh = b.mesh;
                cb.depth = b.depth + 1;
                bn[first + c] = cb;
                node_depth[first + c] = (unsigned char)(b.depth + 1);
                nodes[first + c] = make_int4(-1, coff[c], ccnt[c], g);
                if (ccnt[c] >= kMaxTriangles) next_nodes[ls] = first + c;
            }
            ls += ccnt[c] >= kMaxTriangles;
        }
    }
    const unsigned lt = (1u << lane) - 1u;
    for (int p0 = 0; p0 < cnt; p0 += 32) {
        const int p = p0 + lane;
        const unsigned mask = p < cnt ? pm[p] : 0u;
        const int tid = p < cnt ? pairs[nd.y + p] : 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned bal = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
            if ((mask >> c) & 1u) pairs[coff[c] + __popc(bal & lt)] = tid;
            coff[c] += __popc(bal);
        }
    }
}

static int top_depth_for(int nt, int knob) {
    if (knob == 0) return 0;
    if (knob > 0) return std::min(knob, kTopMaxDepth);
    int d0 = 0;
    while (d0 < kTopMaxDepth && (1ll << (2 * d0)) * 245 <= nt) ++d0;   // 4^(D0-1) <= nt / 245: cells of depth D0-1 hold >= 64 triangles on a sphere
    return d0;
}

msmgpu_status forest_build(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, std::shared_ptr<Forest>& out, std::vector<int>& roots) {
    cudaStream_t s = ctx->stream;
    if (n <= 0) return fail(MSMGPU_ERR_INVALID, "forest_build: no meshes");
    for (int i = 0; i < n; ++i)
        if (!meshes[i] || meshes[i]->ctx != ctx) return fail(MSMGPU_ERR_INVALID, "forest_build: mesh from another context");
    MSM_TRY(ensure_tables(ctx, n, meshes));
    long long total_t = 0;
    std::vector<int> h_nt(n), h_off(n);
    std::vector<const TriRec*> h_rec(n);
    std::vector<const uint4*> h_qbox(n);
    for (int i = 0; i < n; ++i) {
        if (!meshes[i] || meshes[i]->ctx != ctx) return fail(MSMGPU_ERR_INVALID, "forest_build: mesh from another context");
        h_nt[i] = meshes[i]->nt;
        h_off[i] = (int)total_t;
        h_rec[i] = meshes[i]->rec.p;
        h_qbox[i] = meshes[i]->qbox.p;
        total_t += meshes[i]->nt;
    }
    // capacities: the sum of list lengths over all levels is ~10x the triangle count on sphere meshes
    // (leaf duplication ~2.2x); nodes ~0.2 per triangle. Grown and retried on overflow.
    long long pair_cap = 24 * total_t + 4096ll * n;
    long long node_cap = total_t + 4096ll * n;
    const double root_half = kBounds;   // half width of the root cube

    bool top_disabled = false;
    for (int attempt = 0; attempt < 5; ++attempt) {
        if (pair_cap > 0x7fffffffll || node_cap > 0x7fffffffll) return fail(MSMGPU_ERR_CAPACITY, "forest_build: batch too large for 32-bit offsets");
        auto F = std::make_shared<Forest>();
        F->ctx = ctx;
        DevBuf<BuildNode> bn;
        DevBuf<const TriRec*> d_rec;
        DevBuf<const uint4*> d_qbox;
        DevBuf<int> d_nt, d_off;
        MSM_CUDA(F->nodes.alloc(node_cap, s));
        MSM_CUDA(F->pairs.alloc(pair_cap + 4, s));   // + 4: the queries read leaf lists as aligned 128-bit words (query.cuh)
        DevBuf<unsigned char> pmask;   // per list position of the CURRENT level: which of the 8 children the triangle goes to
        long long pmask_cap = 3 * total_t + 4096;
        MSM_CUDA(pmask.alloc((size_t)pmask_cap, s));
        int level_base = 0;            // the current level's lists occupy pairs[level_base, level_base + level_entries)
        long long level_entries = total_t;
        MSM_CUDA(F->node_depth.alloc(node_cap, s));
        MSM_CUDA(bn.alloc(node_cap, s));
        MSM_CUDA(d_rec.alloc(n, s));
        MSM_CUDA(d_qbox.alloc(n, s));
        MSM_CUDA(cudaMemcpyAsync(d_qbox.p, h_qbox.data(), n * sizeof(uint4*), cudaMemcpyHostToDevice, s));
        MSM_CUDA(d_nt.alloc(n, s));
        MSM_CUDA(d_off.alloc(n, s));
        MSM_CUDA(cudaMemcpyAsync(d_rec.p, h_rec.data(), n * sizeof(TriRec*), cudaMemcpyHostToDevice, s));
        MSM_CUDA(cudaMemcpyAsync(d_nt.p, h_nt.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        MSM_CUDA(cudaMemcpyAsync(d_off.p, h_off.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        int node_begin = 0, n_level = n;
        int n_nodes = n;
        long long n_pairs = total_t;
        int depth = 0;
        bool overflow = false;
        DevBuf<int> split_flag, split_rank, child_cnt, child_off, list_start, list_count, totals, stats, live, live_next;
        MSM_CUDA(totals.alloc(5, s));   // [0] splits, [1] new list entries, [2] longest child list, [3] live children, [4] live-list cursor
        int n_live = n;                 // every root is a candidate
        int level_max_cnt = 1;
        for (int v : h_nt) level_max_cnt = std::max(level_max_cnt, v);
        int max_nt = 1;
        for (int v : h_nt) max_nt = v > max_nt ? v : max_nt;
        // ---- top phase (see the note above k_top_count): the first levels of the dense meshes without level passes ----
        std::vector<TopJob> top_jobs(n);
        int top_dmax = 0;
        {
            const int knob = top_disabled ? 0 : tuning_get("build_top", "MSMGPU_BUILD_TOP", -1);
            for (int i = 0; i < n; ++i) {
                top_jobs[i] = TopJob{h_qbox[i], h_nt[i], top_depth_for(h_nt[i], knob), nullptr};
                top_dmax = std::max(top_dmax, top_jobs[i].d0);
            }
        }
        // Optional processing order of the triangles (knob "build_top_order", off by default; results do not depend on it: the lists
        // are sorted afterwards). Meshes that share one topology share nearly the same geometry, so one Morton order of the first
        // mesh's triangles makes the 32 triangles of a warp land in a handful of lattice cells. MEASURED (profiles/t4_top_phase.md):
        // slower — the box reads become gathers and the full 8-slot warp aggregation (knob "build_top_agg" = 1) costs more
        // instructions than the atomics it saves; what pays is aggregating the lower-corner cell only (= 2, the default), in id order.
        std::vector<DevBuf<int>> top_perms;
        const int top_agg = tuning_get("build_top_agg", "MSMGPU_BUILD_TOP_AGG", 2);
        if (top_dmax > 0 && tuning_get("build_top_order", "MSMGPU_BUILD_TOP_ORDER", 0) != 0) {
            std::vector<char> done(n, 0);
            for (int i = 0; i < n; ++i) {
                if (done[i] || top_jobs[i].d0 <= 0) continue;
                std::vector<int> same;
                for (int j = i; j < n; ++j)
                    if (!done[j] && top_jobs[j].d0 > 0 && meshes[j]->tri.p == meshes[i]->tri.p && h_nt[j] == h_nt[i] && meshes[j]->nv == meshes[i]->nv) same.push_back(j);
                for (int j : same) done[j] = 1;
                if ((int)same.size() < 4) continue;   // the sort costs more than it saves on a few meshes
                DevBuf<double> pts;
                MSM_CUDA(pts.alloc(3 * (size_t)h_nt[i], s));
                k_tri_points<<<(h_nt[i] + 255) / 256, 256, 0, s>>>(h_nt[i], meshes[i]->xyz.p, meshes[i]->tri.p, pts.p);
                MSM_LAUNCH_CHECK();
                top_perms.emplace_back();
                MSM_TRY(morton_order(pts.p, h_nt[i], top_perms.back(), s));
                for (int j : same) top_jobs[j].perm = top_perms.back().p;
            }
        }
        bool top_done = false;
        DevBuf<int> top_nodes;          // nodes of the first level after the top phase (node_map of the level kernels)
        const int* node_map = nullptr;
        if (top_dmax > 0) {
            const long long cells_per_mesh = top_cells_before(top_dmax + 1);
            const long long fillc_per_mesh = top_cells_before(top_dmax);       // write cursors of list-keeping cells above their mesh's D0
            const size_t n_cells = (size_t)n * (size_t)cells_per_mesh;
            const int live_cap = (int)std::min<long long>(pair_cap / kMaxTriangles + n, 0x3fffffffll);
            if (node_cap >= -(long long)kTopNone - 2) return fail(MSMGPU_ERR_CAPACITY, "forest_build: batch too large for the top phase's node codes");
            DevBuf<TopJob> d_jobs;
            DevBuf<unsigned> cnt;
            DevBuf<int> nid, fillc;
            DevBuf<TopCursors> cur;
            MSM_CUDA(d_jobs.alloc(n, s));
            MSM_CUDA(cudaMemcpyAsync(d_jobs.p, top_jobs.data(), n * sizeof(TopJob), cudaMemcpyHostToDevice, s));   // pageable: staged before return
            MSM_CUDA(cnt.alloc(n_cells, s));
            MSM_CUDA(nid.alloc(n_cells, s));
            MSM_CUDA(fillc.alloc((size_t)n * (size_t)fillc_per_mesh, s));
            MSM_CUDA(cur.alloc(1, s));
            MSM_CUDA(top_nodes.alloc((size_t)live_cap, s));
            MSM_CUDA(cudaMemsetAsync(cnt.p, 0, n_cells * sizeof(unsigned), s));
            MSM_CUDA(cudaMemsetAsync(nid.p, 0x80, n_cells * sizeof(int), s));   // 0x80808080 < kTopNone, and stays there under k_top_fill's increments
            TopCursors h_cur{};
            h_cur.nodes = n;   // the roots are nodes 0 .. n-1
            MSM_CUDA(cudaMemcpyAsync(cur.p, &h_cur, sizeof(TopCursors), cudaMemcpyHostToDevice, s));
            k_top_count<<<dim3((unsigned)((max_nt + kTopTrisPerBlock - 1) / kTopTrisPerBlock), (unsigned)n), 256, 0, s>>>(d_jobs.p, cnt.p, cells_per_mesh, top_agg);
            MSM_LAUNCH_CHECK();
            {
                const long long threads = (long long)n * top_cells_before(top_dmax);
                k_top_decide<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(n, top_dmax, d_jobs.p, cnt.p, cells_per_mesh);
                MSM_LAUNCH_CHECK();
            }
            k_top_roots<<<(n + 127) / 128, 128, 0, s>>>(n, d_jobs.p, cnt.p, cells_per_mesh, nid.p, F->nodes.p, bn.p, F->node_depth.p, cur.p, top_nodes.p, live_cap);
            MSM_LAUNCH_CHECK();
            for (int d = 0; d < top_dmax; ++d) {
                const long long threads = (long long)n << (3 * d);
                k_top_children<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(n, d, d_jobs.p, cnt.p, cells_per_mesh, nid.p, fillc.p, fillc_per_mesh, F->nodes.p,
                                                                                 bn.p, F->node_depth.p, cur.p, top_nodes.p, live_cap, (int)node_cap, root_half);
                MSM_LAUNCH_CHECK();
            }
            TopCursors* h_res = reinterpret_cast<TopCursors*>(ctx->pinned ? (void*)ctx->pinned : (void*)&h_cur);
            MSM_CUDA(cudaMemcpyAsync(h_res, cur.p, sizeof(TopCursors), cudaMemcpyDeviceToHost, s));
            MSM_CUDA(cudaStreamSynchronize(s));
            const TopCursors r = *h_res;
            if (r.overflow || (long long)r.pairs > pair_cap || r.nodes > node_cap || r.live > live_cap) { pair_cap *= 2; node_cap *= 2; continue; }
            if (r.max_list > kTopSortCap) { top_disabled = true; continue; }   // a list longer than the sort kernels take: exact level passes from the root
            // roots that keep their list (meshes too small to split, D0 = 0, or undecided): the identity, octree.cpp:42-61
            k_top_root_lists<<<dim3((unsigned)std::min((max_nt + 255) / 256, 64), (unsigned)n), 256, 0, s>>>(nid.p, cells_per_mesh, F->nodes.p, F->pairs.p);
            MSM_LAUNCH_CHECK();
            k_top_fill<<<dim3((unsigned)((max_nt + 255) / 256), (unsigned)n), 256, 0, s>>>(d_jobs.p, nid.p, cells_per_mesh, fillc.p, fillc_per_mesh, F->pairs.p, top_agg);
            MSM_LAUNCH_CHECK();
            if (r.nodes > n) {
                k_top_sort_warp<<<(unsigned)((r.nodes - n + 15) / 16), 256, 0, s>>>(n, r.nodes, F->nodes.p, F->pairs.p);
                MSM_LAUNCH_CHECK();
                if (r.max_list > 256) {
                    k_top_sort_block<<<(unsigned)(r.nodes - n), 256, 0, s>>>(n, r.nodes, F->nodes.p, F->pairs.p);
                    MSM_LAUNCH_CHECK();
                }
            }
            // the exact level passes continue on ONE level made of the nodes that still hold >= 50 triangles (depth-D0 cells and
            // undecided ones, whatever their depth: BuildNode carries it), addressed through node_map
            node_map = top_nodes.p;
            node_begin = 0;
            n_level = r.live;
            n_nodes = r.nodes;
            n_pairs = (long long)r.pairs;
            level_entries = n_pairs;
            level_base = 0;
            n_live = r.live;
            level_max_cnt = std::max(1, r.max_live_cnt);
            depth = top_dmax;
            MSM_CUDA(live.alloc((size_t)std::max(n_live, 1), s));
            if (n_live > 0) {
                k_iota<<<(n_live + 255) / 256, 256, 0, s>>>(n_live, live.p);
                MSM_LAUNCH_CHECK();
            }
            top_done = true;
        }
        if (!top_done) {
            dim3 grid((unsigned)std::min((max_nt + 255) / 256, 1024), (unsigned)n);
            k_init_roots<<<grid, 256, 0, s>>>(n, F->nodes.p, bn.p, F->node_depth.p, F->pairs.p, d_off.p, d_nt.p);
            MSM_LAUNCH_CHECK();
            MSM_CUDA(live.alloc(n, s));
            std::vector<int> ident(n);
            for (int i = 0; i < n; ++i) ident[i] = i;
            MSM_CUDA(cudaMemcpyAsync(live.p, ident.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));   // pageable: staged before return
        }
        static const bool timing = getenv("MSMGPU_BUILD_TIMING") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double t_enq = 0, t_wait = 0, t_alloc = 0;
        std::vector<cudaEvent_t> evs;
        auto mark = [&] { if (timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); evs.push_back(e); } };
        DevBuf<int> fuse_nodes[2];
        DevBuf<FuseCursors> fuse_cur;
        int fuse_flip = 0;
        bool live_valid = true;          // `live` holds the work list of the current level (false: the level is node_map[0 .. n_live) itself)
        const bool fuse_knob = tuning_get("build_fused_levels", "MSMGPU_BUILD_FUSED_LEVELS", 1) != 0;
        while (n_level > 0) {
            if (fuse_knob && n_live > 0 && level_max_cnt <= kFuseMaxList && depth + 1 <= kGridMaxDepth) {
                // short lists: the whole level pass in one kernel (k_level_fused)
                if (n_pairs > pmask_cap) {
                    pmask_cap = n_pairs + n_pairs / 4;
                    MSM_CUDA(pmask.alloc((size_t)pmask_cap, s));
                }
                const int pm_base = node_map ? 0 : level_base;   // (after the top phase / a fused level the lists lie anywhere in `pairs`)
                if (!fuse_cur.p) MSM_CUDA(fuse_cur.alloc(1, s));
                const int next_cap = (int)std::min<long long>((pair_cap - n_pairs) / kMaxTriangles + 8, 0x3fffffffll);
                DevBuf<int>& nxt = fuse_nodes[fuse_flip];
                MSM_CUDA(nxt.alloc((size_t)std::max(next_cap, 1), s));
                FuseCursors hc{};
                hc.pairs = (unsigned long long)n_pairs;
                hc.nodes = n_nodes;
                MSM_CUDA(cudaMemcpyAsync(fuse_cur.p, &hc, sizeof(FuseCursors), cudaMemcpyHostToDevice, s));   // pageable: staged before return
                k_level_fused<<<(unsigned)((n_live + 7) / 8), 256, 0, s>>>(node_begin, live_valid ? live.p : nullptr, n_live, node_map, F->nodes.p, bn.p,
                                                                          F->node_depth.p, F->pairs.p, d_qbox.p, pmask.p, pm_base, root_half, fuse_cur.p,
                                                                          nxt.p, next_cap, (int)node_cap, pair_cap);
                MSM_LAUNCH_CHECK();
                FuseCursors* hr = reinterpret_cast<FuseCursors*>(ctx->pinned ? (void*)ctx->pinned : (void*)&hc);
                MSM_CUDA(cudaMemcpyAsync(hr, fuse_cur.p, sizeof(FuseCursors), cudaMemcpyDeviceToHost, s));
                MSM_CUDA(cudaStreamSynchronize(s));
                const FuseCursors r = *hr;
                if (r.overflow) { overflow = true; break; }
                if (r.nodes == n_nodes) break;            // nothing split: done
                n_nodes = r.nodes;
                n_pairs = (long long)r.pairs;
                node_map = nxt.p;                         // the next level = the children that still hold >= 50 triangles
                fuse_flip ^= 1;
                live_valid = false;
                node_begin = 0;
                n_level = n_live = r.live;
                level_base = 0;
                level_entries = n_pairs;
                level_max_cnt = std::max(1, r.max_cnt);
                ++depth;
                if (depth > 40) return fail(MSMGPU_ERR_CAPACITY, "forest_build: depth limit (degenerate mesh?)");
                continue;
            }
            if (!live_valid) {   // back to the chunked pass after fused levels: its work list is the identity over node_map
                MSM_CUDA(live.alloc((size_t)std::max(n_live, 1), s));
                if (n_live > 0) { k_iota<<<(n_live + 255) / 256, 256, 0, s>>>(n_live, live.p); MSM_LAUNCH_CHECK(); }
                live_valid = true;
            }
            const double t0 = timing ? now() : 0;
            MSM_CUDA(split_flag.alloc(n_level, s));
            MSM_CUDA(split_rank.alloc(n_level, s));
            MSM_CUDA(child_cnt.alloc((size_t)n_level * 8, s));
            MSM_CUDA(child_off.alloc((size_t)n_level * 8, s));
            MSM_CUDA(list_start.alloc(n_level, s));
            MSM_CUDA(list_count.alloc(n_level, s));
            // team width by the longest list of the level: whole CTAs at the top, warps at the bottom
            if (timing) t_alloc += now() - t0;
            if (level_entries > pmask_cap) {
                pmask_cap = level_entries + level_entries / 4;
                MSM_CUDA(pmask.alloc((size_t)pmask_cap, s));
            }
            const int max_cnt = level_max_cnt;
            int K, max_chunks;
            if (max_cnt > 16384) K = 1024 * 8; else if (max_cnt > 512) K = 256 * 4; else K = 32 * 4;
            max_chunks = std::max(1, (max_cnt + K - 1) / K);
            MSM_CUDA(stats.alloc((size_t)n_level * max_chunks * kStatInts, s));
            if (max_chunks > 65535) return fail(MSMGPU_ERR_CAPACITY, "forest_build: list too long for the chunk grid");
            const dim3 g_cta((unsigned)std::max(n_live, 1), (unsigned)max_chunks), g_warp((unsigned)((std::max(n_live, 1) + 7) / 8), (unsigned)max_chunks);
            if (depth > kGridMaxDepth) {   // below the lattice k_chunk_stats rebuilds the FP64 boxes from the records: deferred ones are needed now
                bool changed = false;
                for (int i = 0; i < n; ++i)
                    if (!meshes[i]->rec.p && meshes[i]->nt > 0) { MSM_TRY(ensure_records(meshes[i])); h_rec[i] = meshes[i]->rec.p; changed = true; }
                if (changed) MSM_CUDA(cudaMemcpyAsync(d_rec.p, h_rec.data(), n * sizeof(TriRec*), cudaMemcpyHostToDevice, s));
            }
            mark();
            if (n_live == 0) {}   // nothing can split: k_node_combine clears the flags and the loop ends
            else if (K == 8192)
                k_chunk_stats<1024, 8><<<g_cta, 1024, 0, s>>>(node_begin, n_level, max_chunks, F->nodes.p, bn.p, F->pairs.p, d_rec.p, d_qbox.p, root_half, stats.p, pmask.p, level_base, live.p, n_live, node_map);
            else if (K == 1024)
                k_chunk_stats<256, 4><<<g_cta, 256, 0, s>>>(node_begin, n_level, max_chunks, F->nodes.p, bn.p, F->pairs.p, d_rec.p, d_qbox.p, root_half, stats.p, pmask.p, level_base, live.p, n_live, node_map);
            else
                k_chunk_stats<32, 4><<<g_warp, 256, 0, s>>>(node_begin, n_level, max_chunks, F->nodes.p, bn.p, F->pairs.p, d_rec.p, d_qbox.p, root_half, stats.p, pmask.p, level_base, live.p, n_live, node_map);
            MSM_LAUNCH_CHECK();
            mark();
            MSM_CUDA(cudaMemsetAsync(totals.p + 2, 0, 3 * sizeof(int), s));
            k_node_combine<<<(n_level + 255) / 256, 256, 0, s>>>(node_begin, n_level, max_chunks, K, F->nodes.p, stats.p, split_flag.p, child_cnt.p,
                                                                 totals.p + 2, totals.p + 3, node_map);
            MSM_LAUNCH_CHECK();
            MSM_TRY(exclusive_scan_i32(split_flag.p, split_rank.p, n_level, totals.p, s));
            MSM_TRY(exclusive_scan_i32(child_cnt.p, child_off.p, n_level * 8, totals.p + 1, s));
            int h_stack[4];
            int* h_tot = ctx->pinned ? ctx->pinned : h_stack;
            mark();
            const double t1 = timing ? now() : 0;
            MSM_CUDA(cudaMemcpyAsync(h_tot, totals.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
            MSM_CUDA(cudaStreamSynchronize(s));
            if (timing) { t_enq += t1 - t0; t_wait += now() - t1; }
            const int n_split = h_tot[0];
            const int new_pairs = h_tot[1];
            if (n_split == 0) break;
            if ((long long)n_nodes + 8ll * n_split > node_cap || n_pairs + new_pairs > pair_cap) { overflow = true; break; }
            k_save_lists<<<(n_level + 255) / 256, 256, 0, s>>>(node_begin, n_level, F->nodes.p, list_start.p, list_count.p, node_map);
            MSM_LAUNCH_CHECK();
            MSM_CUDA(live_next.alloc((size_t)std::max(h_tot[3], 1), s));
            k_make_children<<<(n_level + 255) / 256, 256, 0, s>>>(node_begin, n_level, F->nodes.p, bn.p, F->node_depth.p, split_flag.p,
                                                                 split_rank.p, child_cnt.p, child_off.p, n_nodes, (int)n_pairs, root_half,
                                                                 (int)node_cap, live_next.p, totals.p + 4, node_map);
            MSM_LAUNCH_CHECK();
            if (K == 8192)
                k_scatter_chunk<1024, 8><<<g_cta, 1024, 0, s>>>(node_begin, n_level, max_chunks, bn.p, F->pairs.p, pmask.p, level_base, split_flag.p,
                                                               list_start.p, list_count.p, child_off.p, stats.p, (int)n_pairs, live.p, n_live);
            else if (K == 1024)
                k_scatter_chunk<256, 4><<<g_cta, 256, 0, s>>>(node_begin, n_level, max_chunks, bn.p, F->pairs.p, pmask.p, level_base, split_flag.p,
                                                             list_start.p, list_count.p, child_off.p, stats.p, (int)n_pairs, live.p, n_live);
            else
                k_scatter_chunk<32, 4><<<g_warp, 256, 0, s>>>(node_begin, n_level, max_chunks, bn.p, F->pairs.p, pmask.p, level_base, split_flag.p,
                                                             list_start.p, list_count.p, child_off.p, stats.p, (int)n_pairs, live.p, n_live);
            MSM_LAUNCH_CHECK();
            mark();
            level_max_cnt = h_tot[2];
            std::swap(live, live_next);       // (stream-ordered: the old list is released after the kernels that read it)
            node_map = nullptr;               // the children just made are contiguous
            n_live = h_tot[3];
            level_base = (int)n_pairs;        // the children's lists were appended at the old end of `pairs`
            level_entries = new_pairs;
            node_begin = n_nodes;
            n_level = 8 * n_split;
            n_nodes += n_level;
            n_pairs += new_pairs;
            ++depth;
            if (depth > 40) return fail(MSMGPU_ERR_CAPACITY, "forest_build: depth limit (degenerate mesh?)");
        }
        if (timing) {
            cudaStreamSynchronize(s);
            // events per level: before the chunk statistics, after them, before the totals come to the host, after the scatter
            double ph[4] = {0, 0, 0, 0};   // statistics | combine + scans | totals round trip + children + scatter | between levels
            for (size_t i = 0; i + 1 < evs.size(); ++i) {
                float ms = 0;
                cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
                ph[i % 4] += ms * 1e3;
            }
            fprintf(stderr, "[msmgpu build] device time by phase: chunk stats %.0f us, combine + scans %.0f us, children + scatter %.0f us, between levels %.0f us\n",
                    ph[0], ph[1], ph[2], ph[3]);
            for (cudaEvent_t e : evs) cudaEventDestroy(e);
        }
        if (timing) fprintf(stderr, "[msmgpu build] %d meshes, depth %d: host enqueue %.0f us (of which scratch allocation %.0f us), waiting for the level totals %.0f us\n", n, depth, t_enq, t_alloc, t_wait);
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

msmgpu_mesh::~msmgpu_mesh() { delete own_tree; }

msm::TreeView msmgpu_octree::view() const {
    if (!mesh->rec.p && mesh->nt > 0 && msm::ensure_records(mesh) != MSMGPU_OK)
        fprintf(stderr, "[msmgpu] could not materialise the triangle records of a mesh: %s\n", msmgpu_last_error());
    return view_lazy();
}
