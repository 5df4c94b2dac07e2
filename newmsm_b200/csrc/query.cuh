// Device-side nearest-triangle query on the flattened octree, shared by query.cu and cost.cu.
// See query.cu for the execution model (groups of G lanes per query).
#pragma once
#include "common.cuh"

#include <climits>

namespace msm {

constexpr unsigned kFull = 0xffffffffu;

template <int G>
__device__ __forceinline__ void group_argmin(double& d, int& pos, int& t) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(kFull, d, o, G);
        const int op = __shfl_xor_sync(kFull, pos, o, G);
        const int ot = __shfl_xor_sync(kFull, t, o, G);
        if (od < d || (od == d && op < pos)) { d = od; pos = op; t = ot; }
    }
}

// A triangle's query record, stored (TriRec) or, LAZY, built from its corners with the expressions make_trirec stores
// (project_to_plane / boundary_distance of geom.cuh are those expressions; --fmad=false: same bits)
template <bool LAZY>
__device__ __forceinline__ void tri_corners(const TreeView& T, int t, V3& v1, V3& v2, V3& v3) {
    if (LAZY) {
        const int i0 = __ldg(T.tri + 3 * (size_t)t), i1 = __ldg(T.tri + 3 * (size_t)t + 1), i2 = __ldg(T.tri + 3 * (size_t)t + 2);
        v1 = V3{__ldg(T.xyz + 3 * (size_t)i0), __ldg(T.xyz + 3 * (size_t)i0 + 1), __ldg(T.xyz + 3 * (size_t)i0 + 2)};
        v2 = V3{__ldg(T.xyz + 3 * (size_t)i1), __ldg(T.xyz + 3 * (size_t)i1 + 1), __ldg(T.xyz + 3 * (size_t)i1 + 2)};
        v3 = V3{__ldg(T.xyz + 3 * (size_t)i2), __ldg(T.xyz + 3 * (size_t)i2 + 1), __ldg(T.xyz + 3 * (size_t)i2 + 2)};
    } else {
        const double* v = T.rec[t].v;
        v1 = V3{v[0], v[1], v[2]}; v2 = V3{v[3], v[4], v[5]}; v3 = V3{v[6], v[7], v[8]};
    }
}
template <bool LAZY>
__device__ __forceinline__ bool tri_inside(const TreeView& T, const V3& pt, int t) {
    if (!LAZY) return rec_inside(pt, T.rec + t);
    V3 v1, v2, v3;
    tri_corners<true>(T, t, v1, v2, v3);
    return in_triangle(project_to_plane(pt, v1, v2, v3), v1, v2, v3);
}
template <bool LAZY>
__device__ __forceinline__ double tri_distance_inside(const TreeView& T, const V3& pt, int t) {
    if (!LAZY) return rec_distance_inside(pt, T.rec + t);
    V3 v1, v2, v3;
    tri_corners<true>(T, t, v1, v2, v3);
    return boundary_distance(project_to_plane(pt, v1, v2, v3), v1, v2, v3);
}
template <bool LAZY>
__device__ __forceinline__ double tri_distance(const TreeView& T, const V3& pt, int t) {
    if (!LAZY) return rec_distance(pt, T.rec + t);
    V3 v1, v2, v3;
    tri_corners<true>(T, t, v1, v2, v3);
    return distance_to_triangle(pt, v1, v2, v3);
}

// All 32 lanes of the warp must call this together (inactive queries pass active = false);
// the G lanes of a group pass the same point. `gl` = lane index inside the group.
// Returns the triangle id (same value in all lanes of the group) or -1; status as msmgpu_status.
template <int G, bool LAZY = false>
__device__ __forceinline__ int nearest_triangle(const TreeView& T, const V3& pt, bool active, int gl, int& status) {
    status = MSMGPU_OK;
    if (active) {   // node.cpp:67-77 on the root cube [-101,101]^3 (octree.cpp:157-158)
        if (pt.x < -kBounds || pt.x > kBounds || pt.y < -kBounds || pt.y > kBounds || pt.z < -kBounds || pt.z > kBounds) {
            status = MSMGPU_ERR_OUT_OF_BOX;
            active = false;
        }
    }
    double best_d = DBL_MAX;
    int best_pos = INT_MAX, best_t = -1, parent = -1;
    if (active) {
        int4 nd = __ldg(T.nodes + T.root);
        double lox = -kBounds, loy = -kBounds, loz = -kBounds, half = kBounds;
        // octree.cpp:165-170: the LAST child (i,j,k nesting) whose closed box contains pt wins; child
        // boxes are products of [lo,mid] / [mid,hi], so that is "upper half iff p >= mid" per axis.
        while (nd.x >= 0) {
            int c = 0;
            const double mx = lox + half, my = loy + half, mz = loz + half;
            if (pt.x >= mx) { c |= 4; lox = mx; }
            if (pt.y >= my) { c |= 2; loy = my; }
            if (pt.z >= mz) { c |= 1; loz = mz; }
            half *= 0.5;
            nd = __ldg(T.nodes + nd.x + c);
        }
        parent = nd.w;
        // octree.cpp:172-178, in two phases so that the expensive test runs on few, well-packed lanes:
        // (1) cull: one 16-byte sphere test per candidate, survivors recorded in a bit mask
        //     (bit j = this lane's j-th candidate, i = gl + j*G);
        // (2) the reference's full test (projection, 3 same-side tests, boundary distance) for survivors, in
        //     ascending scan position, so "first strictly smaller distance wins" is preserved.
        const double pp = vdot(pt, pt);
        const int* __restrict__ list = T.pairs + nd.y;
        unsigned long long keep = 0ull;
        const int mine = (nd.z - gl + G - 1) / G;             // candidates owned by this lane (may be <= 0)
        const int fast = mine < 64 ? mine : 64;
        // the scan is a chain of dependent loads (id -> cull sphere): batches of 4 candidates, with the ids of the NEXT batch
        // requested before the spheres of the current one are tested, so one load latency per batch is exposed instead of two.
        // With one lane per query the lane's candidates are contiguous in the list: they are fetched as ALIGNED 128-bit words
        // (4 ids per L1 access instead of 1; the kernel is bound by L1 wavefronts of divergent loads, profiles/r1u). The word that
        // holds the first candidate may start up to 3 ids earlier and the last one may end up to 3 ids later: both stay inside
        // the `pairs` allocation and those slots are masked out.
        if (G == 1) {
            const int lead = nd.y & 3;                                      // ids of the first word that precede the list
            const int4* __restrict__ words = reinterpret_cast<const int4*>(T.pairs + (nd.y - lead));
            const int n_words = (lead + fast + 3) >> 2;
            int4 wn = n_words > 0 ? __ldg(words) : make_int4(0, 0, 0, 0);
            for (int k = 0; k < n_words; ++k) {
                const int4 wv = wn;
                if (k + 1 < n_words) wn = __ldg(words + k + 1);
                const int jb = 4 * k - lead;                                // list position of wv.x
                const int t[4] = {wv.x, wv.y, wv.z, wv.w};
                bool ok[4];
                float4 cs[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ok[u] = jb + u >= 0 && jb + u < fast;
                    cs[u] = __ldg(T.cull + (ok[u] ? t[u] : 0));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok[u] && cull_keep_f4(cs[u], pt, pp)) keep |= 1ull << (jb + u);
            }
        } else {
            int tn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) tn[u] = u < fast ? __ldg(list + gl + u * G) : 0;
            for (int j0 = 0; j0 < fast; j0 += 4) {
                int t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = tn[u];
#pragma unroll
                for (int u = 0; u < 4; ++u) tn[u] = j0 + 4 + u < fast ? __ldg(list + gl + (j0 + 4 + u) * G) : 0;
                float4 cs[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) cs[u] = __ldg(T.cull + t[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j0 + u < fast && cull_keep_f4(cs[u], pt, pp)) keep |= 1ull << (j0 + u);
            }
        }
        // (2a) containment (projection + 3 same-side tests) for the survivors; (2b) the boundary distance only for the triangles
        // that contain the projection (one, rarely two or three on an edge / at a vertex). Splitting the loop keeps the lanes of a
        // warp converged in the expensive part instead of dragging ~6 active lanes through it once per survivor.
        unsigned long long inside = 0ull;
        while (keep) {
            const int j = __ffsll((long long)keep) - 1;
            keep &= keep - 1;
            if (tri_inside<LAZY>(T, pt, __ldg(list + gl + j * G))) inside |= 1ull << j;
        }
        while (inside) {
            const int j = __ffsll((long long)inside) - 1;
            inside &= inside - 1;
            const int i = gl + j * G;
            const int t = __ldg(list + i);
            const double d = tri_distance_inside<LAZY>(T, pt, t);
            if (d > kNotInTriangle && d < best_d) { best_d = d; best_pos = i; best_t = t; }
        }
        for (int j = 64; j < mine; ++j) {                     // oversized leaves (split refused, octree.cpp:102): no mask
            const int i = gl + j * G;
            const int t = __ldg(list + i);
            if (!cull_keep_f4(__ldg(T.cull + t), pt, pp)) continue;
            const double d = tri_distance<LAZY>(T, pt, t);
            if (d > kNotInTriangle && d < best_d) { best_d = d; best_pos = i; best_t = t; }
        }
    }
    group_argmin<G>(best_d, best_pos, best_t);
    bool need = active && best_t < 0;
    if (__any_sync(kFull, need)) {
        if (need && parent < 0) { status = MSMGPU_ERR_NO_TRIANGLE; need = false; }   // root is a leaf
        int first_child = 0;
        if (need) {   // fallback 1, octree.cpp:180-192: every leaf child of the parent, children in order
            first_child = __ldg(T.nodes + parent).x;
            int base = 0;
            for (int c = 0; c < 8; ++c) {
                const int4 ch = __ldg(T.nodes + first_child + c);
                for (int i = gl; i < ch.z; i += G) {
                    const int t = __ldg(T.pairs + ch.y + i);
                    if (!cull_keep_f4(__ldg(T.cull + t), pt, vdot(pt, pt))) continue;
                    const double d = tri_distance<LAZY>(T, pt, t);
                    if (d > kNotInTriangle && d < best_d) { best_d = d; best_pos = base + i; best_t = t; }
                }
                base += ch.z;
            }
        }
        group_argmin<G>(best_d, best_pos, best_t);
        const bool need2 = need && best_t < 0;
        if (__any_sync(kFull, need2)) {
            if (need2) {   // fallback 2, octree.cpp:194-208: triangle owning the geodesically nearest corner.
                // 2R asin(c/2R) is increasing in the chord c for c <= 2R, so the chord is compared directly. Known difference: two
                // DISTINCT chords whose asin values round to the same double tie in the reference (first corner kept) and not here
                // (smaller chord kept); asin is not evaluated on the device because it is not bit-identical to glibc's.
                int base = 0;
                for (int c = 0; c < 8; ++c) {
                    const int4 ch = __ldg(T.nodes + first_child + c);
                    for (int i = gl; i < ch.z; i += G) {
                        const int t = __ldg(T.pairs + ch.y + i);
                        V3 cv[3];
                        tri_corners<LAZY>(T, t, cv[0], cv[1], cv[2]);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const double d = vnorm(vsub(cv[k], pt));
                            // a chord longer than the diameter makes the reference's asin NaN and its `<` false: corner skipped
                            if (d <= 2.0 * kRad && d < best_d) { best_d = d; best_pos = 3 * (base + i) + k; best_t = t; }
                        }
                    }
                    base += ch.z;
                }
            }
            group_argmin<G>(best_d, best_pos, best_t);
        }
        if (need && best_t < 0) status = MSMGPU_ERR_NO_TRIANGLE;   // octree.cpp:210-211
    }
    return best_t;
}

__device__ __forceinline__ V3 load_pt(const double* __restrict__ pts, int i) {
    return V3{__ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1), __ldg(pts + 3 * (size_t)i + 2)};
}
__device__ __forceinline__ void rec_corners(const TriRec* __restrict__ r, V3& v1, V3& v2, V3& v3) {
    v1 = V3{r->v[0], r->v[1], r->v[2]};
    v2 = V3{r->v[3], r->v[4], r->v[5]};
    v3 = V3{r->v[6], r->v[7], r->v[8]};
}

// calc_barycentric_weights (triangle.cpp:124-143) -> std::map<int,double> semantics: entries in
// ascending vertex id, a repeated id keeps the LAST value assigned. Returns the entry count.
template <bool LAZY = false>
__device__ __forceinline__ int sorted_weights(const TreeView& T, int t, const V3& pt, int* idx, double* w) {
    V3 v1, v2, v3, PP;
    if (LAZY) {
        tri_corners<true>(T, t, v1, v2, v3);
        PP = project_to_plane(pt, v1, v2, v3);
    } else {
        const TriRec* r = T.rec + t;
        rec_corners(r, v1, v2, v3);
        PP = rec_project(pt, V3{r->s3[0], r->s3[1], r->s3[2]}, r->d);
    }
    const double Aa = tri_area(PP, v2, v3);
    const double Ab = tri_area(PP, v1, v3);
    const double Ac = tri_area(PP, v1, v2);
    const double A = Aa + Ab + Ac;
    const int i0 = __ldg(T.tri + 3 * (size_t)t), i1 = __ldg(T.tri + 3 * (size_t)t + 1), i2 = __ldg(T.tri + 3 * (size_t)t + 2);
    const double w0 = Aa / A, w1 = Ab / A, w2 = Ac / A;
    // std::map semantics without indexed local arrays (they would live in local memory): assignments in the order 0, 1, 2, a
    // repeated key keeps the LAST value; then a 3-element sorting network on (key, value), absent entries carry key INT_MAX
    int ka = i0, kb = INT_MAX, kc = INT_MAX;
    double va = w0, vb = 0.0, vc = 0.0;
    if (i1 == ka) va = w1; else { kb = i1; vb = w1; }
    if (i2 == ka) va = w2; else if (i2 == kb) vb = w2; else { kc = i2; vc = w2; }
    if (kb == INT_MAX && kc != INT_MAX) { kb = kc; vb = vc; kc = INT_MAX; vc = 0.0; }   // entries are packed: a, b, c
    auto cswap = [](int& k1, double& v1, int& k2, double& v2) {
        if (k2 < k1) { const int tk = k1; k1 = k2; k2 = tk; const double tv = v1; v1 = v2; v2 = tv; }
    };
    cswap(ka, va, kb, vb);
    cswap(kb, vb, kc, vc);
    cswap(ka, va, kb, vb);
    const int n = 1 + (kb != INT_MAX) + (kc != INT_MAX);
    idx[0] = ka; w[0] = va;
    idx[1] = kb != INT_MAX ? kb : -1; w[1] = kb != INT_MAX ? vb : 0.0;
    idx[2] = kc != INT_MAX ? kc : -1; w[2] = kc != INT_MAX ? vc : 0.0;
    return n;
}

} // namespace msm
