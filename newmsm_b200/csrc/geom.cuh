// FP64 geometric primitives for the sm_100a kernels.
//
// Every decision the reference makes on the nearest-triangle path is an IEEE-754 double
// comparison of values built from + - * / sqrt in a fixed order (point.cpp:26-75,
// triangle.cpp:85-157). These device functions evaluate the same expressions in the same
// association order; the translation unit is compiled with --fmad=false so nvcc never
// contracts a*b+c, and CUDA's double / and sqrt are correctly rounded, hence triangle ids
// and barycentric weights reproduce the CPU path bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>

namespace msm {

constexpr double kEps = 1e-8;          // point.h:31 EPSILON
constexpr double kRad = 100.0;         // point.h:32 RAD
constexpr double kBounds = 101.0;      // octree.h:37 MESH_BOUNDS
constexpr int kMaxTriangles = 50;      // node.h:33 MAX_TRIANGLES
constexpr double kNotInTriangle = -1.0; // octree.h:35

struct V3 { double x, y, z; };

__host__ __device__ __forceinline__ V3 vsub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ V3 vscale(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ double vdot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// point.cpp:177-182 (note the operand order of the middle component)
__host__ __device__ __forceinline__ V3 vcross(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, b.x * a.z - b.z * a.x, a.x * b.y - b.x * a.y};
}
__host__ __device__ __forceinline__ double vnorm(const V3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// point.cpp:26-34: no-op for |a| <= 1e-8
__host__ __device__ __forceinline__ V3 vnormalized(V3 a) {
    const double n = vnorm(a);
    if (n > kEps) { a.x /= n; a.y /= n; a.z /= n; }
    return a;
}
// 3x3 row-major matrix times point, point.cpp:202-208
__host__ __device__ __forceinline__ V3 mat_apply(const double* M, const V3& v) {
    return {M[0] * v.x + M[1] * v.y + M[2] * v.z,
            M[3] * v.x + M[4] * v.y + M[5] * v.z,
            M[6] * v.x + M[7] * v.y + M[8] * v.z};
}

// point.cpp:46-61 — ray origin->vb intersected with the plane of (v1,v2,v3)
__host__ __device__ __forceinline__ V3 project_to_plane(const V3& vb, const V3& v1, const V3& v2, const V3& v3) {
    const V3 s1 = vnormalized(vsub(v3, v1));
    const V3 s2 = vnormalized(vsub(v2, v1));
    const V3 s3 = vnormalized(vcross(s1, s2));
    const double si = vdot(s3, v1) / vdot(s3, vb);
    return vscale(vb, si);
}

// point.cpp:36-39
__host__ __device__ __forceinline__ bool same_side(const V3& p1, const V3& p2, const V3& a, const V3& b) {
    const V3 ba = vsub(b, a);
    return vdot(vcross(ba, vsub(p1, a)), vcross(ba, vsub(p2, a))) > -kEps;
}
// point.cpp:41-44
__host__ __device__ __forceinline__ bool in_triangle(const V3& p, const V3& a, const V3& b, const V3& c) {
    return same_side(p, a, b, c) && same_side(p, b, c, a) && same_side(p, c, a, b);
}
// point.cpp:68-75
__host__ __device__ __forceinline__ double tri_area(const V3& v0, const V3& v1, const V3& v2) {
    return 0.5 * vnorm(vcross(vsub(v1, v0), vsub(v2, v0)));
}
// triangle.cpp:47-50 (cached Triangle::area — operands the other way round)
__host__ __device__ __forceinline__ double tri_area_cached(const V3& v0, const V3& v1, const V3& v2) {
    return 0.5 * vnorm(vcross(vsub(v2, v0), vsub(v1, v0)));
}

// triangle.cpp:85-122 — distance from x0 (already in the plane) to the triangle boundary
__host__ __device__ __forceinline__ double boundary_distance(const V3& x0, const V3& x1, const V3& x2, const V3& x3) {
    double d, dmin = DBL_MAX;
    const V3 a1 = vsub(x0, x1), a2 = vsub(x0, x2), a3 = vsub(x0, x3);
    V3 u = vsub(x2, x1);
    if (vdot(a1, u) > 0 && vdot(a2, u) < 0) {
        d = vnorm(vcross(a1, a2)) / vnorm(u);
        if (d < dmin) dmin = d;
    }
    u = vsub(x3, x1);
    if (vdot(a1, u) > 0 && vdot(a3, u) < 0) {
        d = vnorm(vcross(a1, a3)) / vnorm(u);
        if (d < dmin) dmin = d;
    }
    u = vsub(x3, x2);
    if (vdot(a2, u) > 0 && vdot(a3, u) < 0) {
        d = vnorm(vcross(a2, a3)) / vnorm(u);
        if (d < dmin) dmin = d;
    }
    d = vnorm(a1); if (d < dmin) dmin = d;
    d = vnorm(a2); if (d < dmin) dmin = d;
    d = vnorm(a3); if (d < dmin) dmin = d;
    return dmin;
}

// octree.cpp:143-154
__host__ __device__ __forceinline__ double distance_to_triangle(const V3& pt, const V3& v0, const V3& v1, const V3& v2) {
    const V3 mP = project_to_plane(pt, v0, v1, v2);
    if (in_triangle(mP, v0, v1, v2)) return boundary_distance(mP, v0, v1, v2);
    return kNotInTriangle;
}

// triangle.cpp:124-143: weights for (v1,v2,v3) with the query projected into the plane first
__host__ __device__ __forceinline__ void bary_weights_projected(const V3& p, const V3& v1, const V3& v2, const V3& v3, double* w) {
    const V3 PP = project_to_plane(p, v1, v2, v3);
    const double Aa = tri_area(PP, v2, v3);
    const double Ab = tri_area(PP, v1, v3);
    const double Ac = tri_area(PP, v1, v2);
    const double A = Aa + Ab + Ac;
    w[0] = Aa / A; w[1] = Ab / A; w[2] = Ac / A;
}
// triangle.cpp:145-157: same without the projection (used by the cost functions)
__host__ __device__ __forceinline__ void bary_weights_raw(const V3& p, const V3& v1, const V3& v2, const V3& v3, double* w) {
    const double Aa = tri_area(p, v2, v3);
    const double Ab = tri_area(p, v1, v3);
    const double Ac = tri_area(p, v1, v2);
    const double A = Aa + Ab + Ac;
    w[0] = Aa / A; w[1] = Ab / A; w[2] = Ac / A;
}

// ------------------------------------------------------------------------------------------
// Per-triangle query record: one 128-byte line per triangle, so a lane testing a triangle
// touches exactly one L1/L2 line. Everything in it is a function of the triangle only and is
// evaluated with the same expressions the reference evaluates per query (point.cpp:48-57,
// triangle.cpp:91-110), so using the stored values changes no bit of any decision.
// ------------------------------------------------------------------------------------------
struct __align__(16) TriRec {
    double v[9];      // corners v1 v2 v3 (Triangle::vertices order)
    double s3[3];     // unit plane normal as project_point builds it (point.cpp:48-55)
    double d;         // s3 . v1 (point.cpp:57)
    double e12, e13, e23;   // |v2-v1|, |v3-v1|, |v3-v2| (denominators in triangle.cpp:91-110)
};
static_assert(sizeof(TriRec) == 128, "TriRec must be one 128-byte line");

__host__ __device__ __forceinline__ void make_trirec(const V3& v1, const V3& v2, const V3& v3, TriRec& r) {
    r.v[0] = v1.x; r.v[1] = v1.y; r.v[2] = v1.z;
    r.v[3] = v2.x; r.v[4] = v2.y; r.v[5] = v2.z;
    r.v[6] = v3.x; r.v[7] = v3.y; r.v[8] = v3.z;
    const V3 s1 = vnormalized(vsub(v3, v1));
    const V3 s2 = vnormalized(vsub(v2, v1));
    const V3 s3 = vnormalized(vcross(s1, s2));
    r.s3[0] = s3.x; r.s3[1] = s3.y; r.s3[2] = s3.z;
    r.d = vdot(s3, v1);
    r.e12 = vnorm(vsub(v2, v1));
    r.e13 = vnorm(vsub(v3, v1));
    r.e23 = vnorm(vsub(v3, v2));
}

// project_point (point.cpp:46-61) with the triangle-only part taken from the record
__device__ __forceinline__ V3 rec_project(const V3& pt, const V3& s3, double d) {
    const double si = d / vdot(s3, pt);
    return vscale(pt, si);
}

// triangle.cpp:85-122 with the edge lengths taken from the record
__device__ __forceinline__ double rec_boundary_distance(const V3& x0, const V3& x1, const V3& x2, const V3& x3,
                                                        double e12, double e13, double e23) {
    double d, dmin = DBL_MAX;
    const V3 a1 = vsub(x0, x1), a2 = vsub(x0, x2), a3 = vsub(x0, x3);
    V3 u = vsub(x2, x1);
    if (vdot(a1, u) > 0 && vdot(a2, u) < 0) {
        d = vnorm(vcross(a1, a2)) / e12;
        if (d < dmin) dmin = d;
    }
    u = vsub(x3, x1);
    if (vdot(a1, u) > 0 && vdot(a3, u) < 0) {
        d = vnorm(vcross(a1, a3)) / e13;
        if (d < dmin) dmin = d;
    }
    u = vsub(x3, x2);
    if (vdot(a2, u) > 0 && vdot(a3, u) < 0) {
        d = vnorm(vcross(a2, a3)) / e23;
        if (d < dmin) dmin = d;
    }
    d = vnorm(a1); if (d < dmin) dmin = d;
    d = vnorm(a2); if (d < dmin) dmin = d;
    d = vnorm(a3); if (d < dmin) dmin = d;
    return dmin;
}

// octree.cpp:143-154 for one triangle record (loaded as eight 16-byte words)
__device__ __forceinline__ double rec_distance(const V3& pt, const TriRec* __restrict__ rp) {
    const double2* q = reinterpret_cast<const double2*>(rp);
    const double2 a = __ldg(q + 0), b = __ldg(q + 1), c = __ldg(q + 2), d4 = __ldg(q + 3), e = __ldg(q + 4);
    const double2 f = __ldg(q + 5), g = __ldg(q + 6);
    const V3 v1{a.x, a.y, b.x}, v2{b.y, c.x, c.y}, v3{d4.x, d4.y, e.x};
    const V3 s3{e.y, f.x, f.y};
    const V3 mP = rec_project(pt, s3, g.x);
    if (!in_triangle(mP, v1, v2, v3)) return kNotInTriangle;
    const double2 h = __ldg(q + 7);
    return rec_boundary_distance(mP, v1, v2, v3, g.y, h.x, h.y);
}

// The two halves of rec_distance, for callers that first collect the containing triangles of all their candidates and only then
// evaluate the (expensive: 6 sqrt, 3 div) boundary distance of those few: same expressions, same bits.
__device__ __forceinline__ bool rec_inside(const V3& pt, const TriRec* __restrict__ rp) {
    const double2* q = reinterpret_cast<const double2*>(rp);
    const double2 a = __ldg(q + 0), b = __ldg(q + 1), c = __ldg(q + 2), d4 = __ldg(q + 3), e = __ldg(q + 4);
    const double2 f = __ldg(q + 5), g = __ldg(q + 6);
    const V3 v1{a.x, a.y, b.x}, v2{b.y, c.x, c.y}, v3{d4.x, d4.y, e.x};
    const V3 mP = rec_project(pt, V3{e.y, f.x, f.y}, g.x);
    return in_triangle(mP, v1, v2, v3);
}
__device__ __forceinline__ double rec_distance_inside(const V3& pt, const TriRec* __restrict__ rp) {
    const double2* q = reinterpret_cast<const double2*>(rp);
    const double2 a = __ldg(q + 0), b = __ldg(q + 1), c = __ldg(q + 2), d4 = __ldg(q + 3), e = __ldg(q + 4);
    const double2 f = __ldg(q + 5), g = __ldg(q + 6), h = __ldg(q + 7);
    const V3 v1{a.x, a.y, b.x}, v2{b.y, c.x, c.y}, v3{d4.x, d4.y, e.x};
    const V3 mP = rec_project(pt, V3{e.y, f.x, f.y}, g.x);
    return rec_boundary_distance(mP, v1, v2, v3, g.y, h.x, h.y);
}

// ------------------------------------------------------------------------------------------
// Conservative cull record per triangle: a sphere (C, r) that contains every point the reference's
// point_in_triangle (point.cpp:36-44) can accept for this triangle, so a query whose line
// {t * pt} misses the sphere can skip the triangle without evaluating it -- the skipped test would
// have returned NOT_IN_TRIANGLE, hence no decision changes.
//
// same_side accepts p iff  dot(cross(e, p - P), n2) > -1e-8  with n2 = cross(e, O - P) normal to the
// plane; the left side equals |e| |n2| h, h = signed in-plane distance of p from the edge line
// (positive towards the opposite corner O). The accepted set is therefore the prism over the
// triangle grown by h_k = 1e-8 / (|e_k| |n2_k|) across each edge. The grown triangle lies inside the
// triangle scaled about its incentre I by (rho + h) / rho (rho = inradius, h = max h_k), and the
// projected query lies in the plane up to rounding, so r = lambda * max |V_i - I| (inflated for the
// rounding of the reference's own evaluation and of the cull test) bounds it. Degenerate
// triangles (n2 = 0, accepted everywhere) get r = +inf and are never culled.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void make_cull(const V3& a, const V3& b, const V3& c, double* out4) {
    // out4 = centre, RADIUS (not squared: pack_cull adds the float-rounding slack to the radius first).
    // Only conservativeness matters here, not bit-compatibility with anything: |n2_k| = 2 * area for every edge k, so one cross
    // product serves all three (the 1e-9 factors cover the rounding differences between the three ways of computing it).
    const double la = vnorm(vsub(b, c)), lb = vnorm(vsub(c, a)), lc = vnorm(vsub(a, b));
    const double per = la + lb + lc;
    const double n2 = vnorm(vcross(vsub(b, a), vsub(c, a))) * (1.0 - 1e-9);
    const double lmax = fmax(la, fmax(lb, lc)), lmin = fmin(la, fmin(lb, lc));
    const double inf = 1.0 / 0.0;
    double r = inf;
    V3 I{0, 0, 0};
    const double pmin = lmin * n2;
    if (per > 0 && pmin > 0) {
        I = V3{(la * a.x + lb * b.x + lc * c.x) / per, (la * a.y + lb * b.y + lc * c.y) / per, (la * a.z + lb * b.z + lc * c.z) / per};
        const double rho = n2 / per;                             // 2 * area / perimeter (slightly under-estimated: larger lambda)
        const double h = (1e-8 / pmin + 1e-12 * (1.0 + lmax)) * 1.000001;
        const double lambda = 1.0 + h / rho;
        const V3 da = vsub(a, I), db = vsub(b, I), dc = vsub(c, I);
        const double reach = sqrt(fmax(vdot(da, da), fmax(vdot(db, db), vdot(dc, dc)))) * (1.0 + 1e-9);
        r = lambda * reach * (1.0 + 1e-7) + 1e-7;
        if (!(r == r)) r = inf;                                  // NaN from 0/0 -> never cull
    }
    out4[0] = I.x; out4[1] = I.y; out4[2] = I.z; out4[3] = r;
}

// The record is STORED in single precision (16 bytes: one 128-bit load per candidate, twice as many spheres per cache line):
// centre rounded to float, radius grown by the ACTUAL rounding error of the centre |C - Cf| (evaluated in double, padded) and r^2
// rounded up, so the stored sphere contains the double-precision one: still conservative.
// The test itself stays in double: keep the triangle iff the line through the origin and pt may touch the sphere,
// |C x pt|^2 <= r^2 |pt|^2.
__host__ __device__ __forceinline__ float4 pack_cull(const double* c4) {
    const float cx = (float)c4[0], cy = (float)c4[1], cz = (float)c4[2];
    const double ex = c4[0] - (double)cx, ey = c4[1] - (double)cy, ez = c4[2] - (double)cz;
    const double r = c4[3] + sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9) + 1e-9;   // +inf stays +inf (degenerate: never culled)
    const double r2 = r * r * (1.0 + 1e-6);
    float r2f = (float)r2;
    if ((double)r2f < r2) r2f = nextafterf(r2f, INFINITY);
    return make_float4(cx, cy, cz, r2f);
}
__device__ __forceinline__ bool cull_keep_f4(const float4& c, const V3& pt, double pp) {
    const double cx = (double)c.x, cy = (double)c.y, cz = (double)c.z;
    const double x = __fma_rn(cy, pt.z, -(cz * pt.y));
    const double y = __fma_rn(cz, pt.x, -(cx * pt.z));
    const double z = __fma_rn(cx, pt.y, -(cy * pt.x));
    const double d2 = __fma_rn(x, x, __fma_rn(y, y, z * z));
    return !(d2 > (double)c.w * pp);     // written so that NaN keeps the triangle
}

} // namespace msm
