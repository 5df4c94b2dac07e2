// resampler_adapter.hpp — the reference-side binding of include/msmgpu.h for msm-newresampler.
//
// Header-only. It keeps the reference's C++ interface (msm-newresampler/src/resampler.h:38-53,
// octree.h:39-59): same class / function names, argument meaning and error behaviour
// (newresampler::MeshException with the reference's messages), in namespace newresampler_gpu,
// and forwards to the C ABI. A maintainer switches a call site by changing the namespace (or with
// `namespace newresampler = newresampler_gpu;`-style aliases per function) and linking
// libmsmgpu.so; see INTEGRATION.md. `nthreads` arguments are accepted and ignored.
//
// Checked in-process against the reference's own CPU functions by integration/adapter_check.cpp.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#ifndef NEWMSM_B200_RESAMPLER_HEADER
#define NEWMSM_B200_RESAMPLER_HEADER "newresampler/resampler.h"
#endif
#include NEWMSM_B200_RESAMPLER_HEADER   // the reference's Mesh / Point / Triangle / MeshException

#include "../msmgpu.h"

namespace newresampler_gpu {

using newresampler::Mesh;
using newresampler::MeshException;
using newresampler::Point;
using newresampler::Triangle;

namespace detail {

inline void check(msmgpu_status st) {
    if (st != MSMGPU_OK) {
        static thread_local std::string msg;   // MeshException keeps the pointer (meshException.h:31)
        msg = msmgpu_last_error();
        throw MeshException(msg.c_str());
    }
}

inline msmgpu_ctx* context() {   // one context per process on $MSMGPU_DEVICE (default 0)
    static msmgpu_ctx* ctx = [] {
        const char* e = std::getenv("MSMGPU_DEVICE");
        msmgpu_ctx* c = nullptr;
        check(msmgpu_ctx_create(e ? std::atoi(e) : 0, nullptr, &c));
        return c;
    }();
    return ctx;
}

// The marshalling between the reference's Mesh (shared_ptr<Mpoint> per vertex, vector<vector<double>> pvalues: mesh.h:38-44) and the flat
// arrays of the C ABI is memory-bound host work: the loops below are spread over the host threads.
inline std::vector<double> coords_of(const Mesh& m) {
    const int n = m.nvertices();
    std::vector<double> xyz(3 * (size_t)n);
#pragma omp parallel for schedule(static) if (n > 4096)
    for (int i = 0; i < n; ++i) {
        const Point& p = m.get_coord(i);
        xyz[3 * (size_t)i] = p.X; xyz[3 * (size_t)i + 1] = p.Y; xyz[3 * (size_t)i + 2] = p.Z;
    }
    return xyz;
}

inline void flatten_pvalues(const Mesh& m, double* f) {   // channel-major [D][V] like Mesh::pvalues (mesh.h:44)
    const int D = m.get_dimension(), V = m.nvertices();
#pragma omp parallel for collapse(2) schedule(static) if ((size_t)D * V > 16384)
    for (int d = 0; d < D; ++d)
        for (int v = 0; v < V; ++v) f[(size_t)d * V + v] = m.get_pvalue(v, d);
}
inline std::vector<double> pvalues_of(const Mesh& m) {
    std::vector<double> f((size_t)m.get_dimension() * m.nvertices());
    flatten_pvalues(m, f.data());
    return f;
}

// Grow-only page-locked staging buffers (msmgpu_host_alloc), one per role, kept for the life of the process: Mesh::pvalues has to be
// flattened anyway, and flattening straight into page-locked memory lets the C ABI's copies run as DMA at link speed (pageable memory
// goes through the driver's staging copy at a fifth of that). One adapter call at a time uses them (adapter_lock()).
class Staging {
    double* p = nullptr;
    size_t cap = 0;

public:
    double* get(size_t n) {
        if (n > cap) {
            if (p) msmgpu_host_free(context(), p);
            p = nullptr; cap = 0;
            check(msmgpu_host_alloc(context(), n * sizeof(double), reinterpret_cast<void**>(&p)));
            cap = n;
        }
        return p;
    }
};
inline Staging& staging(int role) { static Staging s[2]; return s[role]; }   // 0: input channels, 1: output channels
inline std::mutex& adapter_lock() { static std::mutex m; return m; }

// result = copy of `geometry` (the copy resampler.cpp:37-38 makes too: one shared_ptr<Mpoint> with two adjacency vectors per vertex,
// 14 ms at ico6) carrying D channels produced by device_work(double* out_cm). The copy is host-only work and does not depend on the
// device's result, so the C-ABI call runs on a helper thread meanwhile.
template <class F>
inline Mesh resampled_mesh(const Mesh& geometry, int D, F&& device_work) {
    const size_t V = (size_t)geometry.nvertices();
    double* cm = staging(1).get((size_t)D * V);
    std::string err;
    struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } helper;
    helper.t = std::thread([&] {
        try { device_work(cm); } catch (const MeshException& e) { err = e.what(); if (err.empty()) err = "msmgpu call failed"; }
    });
    Mesh out = geometry;
    helper.t.join();
    if (!err.empty()) {
        static thread_local std::string msg;   // MeshException keeps the pointer (meshException.h:31)
        msg = err;
        throw MeshException(msg.c_str());
    }
    out.initialize_pvalues(0);   // clear (mesh.cpp:252-258); whole channels are appended instead of D * V bounds-checked set_pvalue calls
    for (int d = 0; d < D; ++d) out.push_pvalues(std::vector<double>(cm + (size_t)d * V, cm + (size_t)(d + 1) * V));
    return out;
}

// device copy of a reference Mesh (geometry only) with its octree, built on first use
class DeviceMesh {
public:
    DeviceMesh(int nv_, int nt_, std::vector<double>&& xyz_, std::vector<int32_t>&& tri_, std::vector<double>&& area_)
        : nv(nv_), nt(nt_), xyz(std::move(xyz_)), tri(std::move(tri_)), area(std::move(area_)) {
        check(msmgpu_mesh_create(context(), nv, xyz.data(), nt, tri.data(), &h));
        // Triangle::area is cached at construction and survives set_coord (triangle.cpp:31,39): hand over the values this Mesh
        // object actually holds, so compute_vertex_area (mesh.cpp:1275) is reproduced whatever the object's history
        if (nt > 0) check(msmgpu_mesh_set_triangle_areas(h, area.data()));
    }
    ~DeviceMesh() {
        if (tree_) msmgpu_octree_destroy(tree_);
        msmgpu_mesh_destroy(h);
    }
    DeviceMesh(const DeviceMesh&) = delete;
    DeviceMesh& operator=(const DeviceMesh&) = delete;
    msmgpu_octree* tree() {
        if (!tree_) check(msmgpu_octree_build(h, &tree_));
        return tree_;
    }
    bool holds(int nv_, int nt_, const std::vector<double>& x, const std::vector<int32_t>& t, const std::vector<double>& a) const {
        return nv_ == nv && nt_ == nt && std::memcmp(x.data(), xyz.data(), x.size() * sizeof(double)) == 0 &&
               std::memcmp(t.data(), tri.data(), t.size() * sizeof(int32_t)) == 0 && std::memcmp(a.data(), area.data(), a.size() * sizeof(double)) == 0;
    }
    msmgpu_mesh* h = nullptr;
    int nv, nt;
    std::vector<double> xyz;     // host copies: the key of the cache below
    std::vector<int32_t> tri;
    std::vector<double> area;

private:
    msmgpu_octree* tree_ = nullptr;
};

// The device copy of `m`. The reference's drivers hand the same meshes to the resampler again and again (the input spheres and the
// data grids of featurespace::initialise at every level, the control grid of every resample_weights call, mesh_registration.cpp:
// 170-222): the last few device meshes are kept, keyed by CONTENT (coordinates, faces and cached triangle areas compared bit for
// bit), so a repeated mesh costs one pass over its host arrays instead of an upload, the per-triangle tables and an octree build.
inline std::shared_ptr<DeviceMesh> device_mesh(const Mesh& m) {
    const int nv = m.nvertices(), nt = m.ntriangles();
    std::vector<double> xyz = coords_of(m), area((size_t)nt);
    std::vector<int32_t> tri(3 * (size_t)nt);
#pragma omp parallel for schedule(static) if (nt > 4096)
    for (int t = 0; t < nt; ++t) {
        for (int k = 0; k < 3; ++k) tri[3 * (size_t)t + k] = m.get_triangle_vertexID(t, k);
        area[t] = m.get_triangle_area(t);
    }
    static std::mutex mu;
    static std::list<std::shared_ptr<DeviceMesh>> cache;
    constexpr size_t kKeep = 6;
    std::lock_guard<std::mutex> g(mu);
    for (auto it = cache.begin(); it != cache.end(); ++it)
        if ((*it)->holds(nv, nt, xyz, tri, area)) {
            cache.splice(cache.begin(), cache, it);
            return cache.front();
        }
    cache.push_front(std::make_shared<DeviceMesh>(nv, nt, std::move(xyz), std::move(tri), std::move(area)));
    if (cache.size() > kKeep) cache.pop_back();
    return cache.front();
}

inline Mesh with_pvalues(const Mesh& geometry, int D, const std::vector<double>& cm) {   // like resampler.cpp:37-38, 54-57
    Mesh out = geometry;
    out.initialize_pvalues(0);   // clear (mesh.cpp:252-258); whole channels are appended instead of D * V bounds-checked set_pvalue calls
    const size_t V = (size_t)out.nvertices();
    for (int d = 0; d < D; ++d) out.push_pvalues(std::vector<double>(cm.begin() + (size_t)d * V, cm.begin() + (size_t)(d + 1) * V));
    return out;
}

// an exclusion mask's values: Mesh::get_pvalue(i) of its first channel (resampler.cpp:47, 100)
inline std::vector<double> mask_of(const Mesh& excl) {
    std::vector<double> e((size_t)excl.nvertices());
    for (int i = 0; i < excl.nvertices(); ++i) e[i] = excl.get_pvalue(i);
    return e;
}
// `Mesh exclusion = sphLow; exclusion.set_pvalue(k, v)` (resampler.cpp:36, 64): a copy of the geometry carrying the new mask in channel 0
inline Mesh with_mask(const Mesh& geometry, const std::vector<double>& values) {
    Mesh out = geometry;
    for (int k = 0; k < out.nvertices(); ++k) out.set_pvalue(k, values[k]);
    return out;
}

}  // namespace detail

// octree.h:39-59. Besides the reference's per-point calls there are batched ones: a GPU launch per point
// would be absurd, so hot callers pass all their points at once.
class Octree {
    std::shared_ptr<detail::DeviceMesh> dm;   // shared with the cache: the tree lives as long as this object or the cache entry
    msmgpu_octree* h = nullptr;
    const Mesh* target;

public:
    explicit Octree(const Mesh& t) : dm(detail::device_mesh(t)), h(dm->tree()), target(&t) {}
    Octree(const Octree&) = delete;
    Octree& operator=(const Octree&) = delete;

    msmgpu_octree* handle() const { return h; }

    std::vector<int> get_closest_triangle_ids(const std::vector<Point>& pts) const {
        std::vector<double> q(3 * pts.size());
        for (size_t i = 0; i < pts.size(); ++i) { q[3 * i] = pts[i].X; q[3 * i + 1] = pts[i].Y; q[3 * i + 2] = pts[i].Z; }
        std::vector<int32_t> ids(pts.size());
        detail::check(msmgpu_nearest_triangle(h, (int)pts.size(), q.data(), ids.data(), nullptr, nullptr));   // throws like octree.cpp:158 / 211
        return std::vector<int>(ids.begin(), ids.end());
    }
    Triangle get_closest_triangle(const Point& pt) const { return target->get_triangle(get_closest_triangle_ids({pt})[0]); }

    std::vector<int> get_closest_vertex_IDs(const std::vector<Point>& pts) const {
        std::vector<double> q(3 * pts.size());
        for (size_t i = 0; i < pts.size(); ++i) { q[3 * i] = pts[i].X; q[3 * i + 1] = pts[i].Y; q[3 * i + 2] = pts[i].Z; }
        std::vector<int32_t> ids(pts.size());
        detail::check(msmgpu_nearest_triangle(h, (int)pts.size(), q.data(), nullptr, ids.data(), nullptr));
        return std::vector<int>(ids.begin(), ids.end());
    }
    int get_closest_vertex_ID(const Point& pt) const { return get_closest_vertex_IDs({pt})[0]; }
};

// resampler.h:38-44
class Resampler {
public:
    std::vector<std::map<int, double>> get_barycentric_weights(const Mesh& low, const Mesh& /*orig*/, const Octree& oct, int /*nthreads*/ = 1) {
        const int n = low.nvertices();
        const std::vector<double> q = detail::coords_of(low);
        std::vector<int32_t> idx(3 * (size_t)n), ne(n);
        std::vector<double> w(3 * (size_t)n);
        detail::check(msmgpu_bary_weights(oct.handle(), n, q.data(), idx.data(), w.data(), ne.data()));
        std::vector<std::map<int, double>> out(n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < ne[i]; ++j) out[i].emplace_hint(out[i].end(), idx[3 * (size_t)i + j], w[3 * (size_t)i + j]);
        return out;
    }

    std::vector<std::map<int, double>> get_adaptive_barycentric_weights(const Mesh& in_mesh, const Mesh& sphLow, int /*nthreads*/ = 1,
                                                                        std::shared_ptr<Mesh> EXCL = std::shared_ptr<Mesh>()) {
        const auto pa = detail::device_mesh(in_mesh), pb = detail::device_mesh(sphLow);
        detail::DeviceMesh &a = *pa, &b = *pb;
        msmgpu_weights* W = nullptr;
        if (EXCL) {   // resampler.cpp:100, 121: targets whose closest source vertex is masked out get no row
            if (EXCL->nvertices() != in_mesh.nvertices()) throw MeshException("Exclusion mask differs in nvertices from data");
            detail::check(msmgpu_adaptive_weights_excl(a.h, b.h, detail::mask_of(*EXCL).data(), &W));
        } else {
            detail::check(msmgpu_adaptive_weights(a.h, b.h, &W));
        }
        int n_rows = 0;
        int64_t nnz = 0;
        msmgpu_weights_shape(W, &n_rows, nullptr, &nnz);
        std::vector<int32_t> rowptr((size_t)n_rows + 1), col((size_t)nnz);
        std::vector<double> val((size_t)nnz);
        const msmgpu_status st = msmgpu_weights_export(W, rowptr.data(), col.data(), val.data());
        msmgpu_weights_destroy(W);
        detail::check(st);
        std::vector<std::map<int, double>> out(n_rows);
        for (int r = 0; r < n_rows; ++r)
            for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) out[r].emplace_hint(out[r].end(), col[e], val[e]);
        return out;
    }

    Mesh barycentric_data_interpolation(const Mesh& metric_in, const Mesh& sphLow, int nthreads = 1,
                                        std::shared_ptr<Mesh> EXCL = std::shared_ptr<Mesh>()) {
        if (EXCL && EXCL->nvertices() != metric_in.nvertices()) throw MeshException("Exclusion mask differs in nvertices from data");   // resampler.cpp:33-34
        const auto pa = detail::device_mesh(metric_in), pb = detail::device_mesh(sphLow);
        detail::DeviceMesh &a = *pa, &b = *pb;
        const int D = metric_in.get_dimension();
        std::lock_guard<std::mutex> g(detail::adapter_lock());
        double* fin = detail::staging(0).get((size_t)D * metric_in.nvertices());
        detail::flatten_pvalues(metric_in, fin);
        if (EXCL) {   // masked weights and sums + the mask resampled with the same weights, which replaces *EXCL (resampler.cpp:55-67)
            std::vector<double> eout((size_t)sphLow.nvertices());
            const std::vector<double> ein = detail::mask_of(*EXCL);
            Mesh out = detail::resampled_mesh(sphLow, D, [&](double* fout) {
                detail::check(msmgpu_metric_resample_excl(a.h, b.h, D, fin, ein.data(), fout, eout.data()));
            });
            *EXCL = detail::with_mask(sphLow, eout);
            return out;
        }
        return detail::resampled_mesh(sphLow, D, [&](double* fout) { detail::check(msmgpu_metric_resample(a.h, b.h, D, fin, fout)); });
    }
};

// resampler.h:46-53
inline Mesh metric_resample(const Mesh& in, const Mesh& target, int nthreads = 1, std::shared_ptr<Mesh> EXCL = std::shared_ptr<Mesh>()) {
    return Resampler().barycentric_data_interpolation(in, target, nthreads, EXCL);
}

inline Mesh surface_resample(const Mesh& anat_orig, const Mesh& sphere_orig, const Mesh& sphere_low, int /*nthreads*/ = 1) {
    // resampler.cpp:284-302: new coordinates of sphere_low's vertices = blend of anat_orig over sphere_orig's triangles
    const auto ps = detail::device_mesh(sphere_orig);
    detail::DeviceMesh& s = *ps;
    const std::vector<double> anat = detail::coords_of(anat_orig), low = detail::coords_of(sphere_low);
    std::vector<double> out(low.size());
    detail::check(msmgpu_surface_resample(s.h, anat.data(), sphere_low.nvertices(), low.data(), out.data()));
    Mesh res = sphere_low;
    for (int i = 0; i < res.nvertices(); ++i) res.set_coord(i, Point(out[3 * (size_t)i], out[3 * (size_t)i + 1], out[3 * (size_t)i + 2]));
    return res;
}

inline Mesh project_anatomical_mesh(const Mesh& orig, const Mesh& target, const Mesh& anat, int nthreads = 1) {   // resampler.cpp:260-282
    return newresampler_gpu::surface_resample(anat, orig, target, nthreads);
}

inline void sphere_project_warp(Mesh& sphere, const Mesh& from, const Mesh& to, int /*nthreads*/ = 1) {   // resampler.cpp:311-328
    const auto pf = detail::device_mesh(from);
    detail::DeviceMesh& f = *pf;
    const std::vector<double> t = detail::coords_of(to), q = detail::coords_of(sphere);
    std::vector<double> out(q.size());
    detail::check(msmgpu_sphere_project_warp(f.h, t.data(), sphere.nvertices(), q.data(), out.data()));
    for (int i = 0; i < sphere.nvertices(); ++i) sphere.set_coord(i, Point(out[3 * (size_t)i], out[3 * (size_t)i + 1], out[3 * (size_t)i + 2]));
}

inline Mesh nearest_neighbour_interpolation(Mesh& orig, const Mesh& sphLow, int nthreads = 1, std::shared_ptr<Mesh> EXCL = std::shared_ptr<Mesh>()) {
    newresampler::check_scale(orig, sphLow);   // resampler.cpp:232-258
    const auto pa = detail::device_mesh(orig);
    detail::DeviceMesh& a = *pa;
    const int D = orig.get_dimension();
    const std::vector<double> low = detail::coords_of(sphLow);
    std::lock_guard<std::mutex> g(detail::adapter_lock());
    double* fin = detail::staging(0).get((size_t)D * orig.nvertices());
    detail::flatten_pvalues(orig, fin);
    if (EXCL) {
        std::vector<double> eout((size_t)sphLow.nvertices());
        const std::vector<double> ein = detail::mask_of(*EXCL);
        Mesh out = detail::resampled_mesh(sphLow, D, [&](double* fout) {
            detail::check(msmgpu_nn_resample_excl(a.h, sphLow.nvertices(), low.data(), D, fin, ein.data(), fout, eout.data()));
        });
        *EXCL = detail::with_mask(sphLow, eout);
        return out;
    }
    return detail::resampled_mesh(sphLow, D, [&](double* fout) { detail::check(msmgpu_nn_resample(a.h, sphLow.nvertices(), low.data(), D, fin, fout)); });
}

}  // namespace newresampler_gpu

namespace newresampler_gpu {

// smooth_data (resampler.cpp:169-230): Gaussian smoothing of the per-vertex data with an O(V^2) neighbourhood scan per call
// (1.7e9 pair tests at ico6). The scan is pure IEEE arithmetic and runs on the device (msmgpu_smooth_neighbourhoods); the weights
// need asin / exp and are evaluated on the host libm, on the short lists only, with the reference's expression and summation
// order (inside msmgpu_smooth_data), so the result is the reference's bit for bit, exclusion masks included (resampler.cpp:201-225).
inline Mesh smooth_data(Mesh& orig, const Mesh& sphLow, double sigma, int /*nthreads*/ = 1, std::shared_ptr<Mesh> EXCL = std::shared_ptr<Mesh>()) {
    newresampler::check_scale(orig, sphLow);
    const int n = sphLow.nvertices(), D = orig.get_dimension();
    std::vector<Point> pts((size_t)n);
    for (int i = 0; i < n; ++i) pts[i] = sphLow.get_coord(i);
    const std::vector<int> closest = Octree(orig).get_closest_vertex_IDs(pts);   // oct_search.get_closest_vertex_ID(ci), resampler.cpp:184
    const std::vector<double> low = detail::coords_of(sphLow);
    std::vector<int32_t> c32(closest.begin(), closest.end());
    std::vector<double> eout, mask;
    if (EXCL) { mask = detail::mask_of(*EXCL); eout.resize((size_t)n); }
    std::lock_guard<std::mutex> g(detail::adapter_lock());
    double* fin = detail::staging(0).get((size_t)D * orig.nvertices());
    detail::flatten_pvalues(orig, fin);
    Mesh res = detail::resampled_mesh(sphLow, D, [&](double* out) {
        detail::check(msmgpu_smooth_data(detail::context(), n, low.data(), c32.data(), sigma, D, orig.nvertices(), fin, (int)mask.size(),
                                         EXCL ? mask.data() : nullptr, out, EXCL ? eout.data() : nullptr));
    });
    if (EXCL) *EXCL = detail::with_mask(sphLow, eout);     // resampler.cpp:226
    return res;
}

// make_mesh_from_icosa (mesh.cpp:1111-1196) without the O(V^2) duplicate-midpoint scan of retessellate (mesh.cpp:910-1008):
// a midpoint belongs to an undirected edge, so the linear search over all added points (Point== with 1e-8 tolerance,
// mesh.cpp:945-960) is an edge hash lookup. Vertex ids, face ids, face orientation and every coordinate are the ones the
// reference produces: new points are numbered in first-seen order, per old face (v0,v1,v2) in the order mid(v1,v2), mid(v0,v2),
// mid(v0,v1); the four children are (p2,p0,p1), (p1,v0,p2), (p0,v2,p1), (p2,v1,p0); all points are re-normalised after every
// level with Point::normalize's arithmetic (point.cpp:26-34). Host code: 5.6 s -> 40 ms for ico6.
inline Mesh make_mesh_from_icosa(int n) {
    const double t = 0.8506508084, o = 0.5257311121;
    std::vector<double> v = {t, o, 0, -t, o, 0, -t, -o, 0, t, -o, 0, o, 0, t, o, 0, -t, -o, 0, -t, -o, 0, t, 0, t, o, 0, -t, o, 0, -t, -o, 0, t, -o};
    enum { ZA, ZB, ZC, ZD, YA, YB, YC, YD, XA, XB, XC, XD };
    const int base[20][3] = {{YD, XA, YA}, {XB, YD, YA}, {XD, YC, YB}, {YC, XC, YB}, {ZD, YA, ZA}, {YB, ZD, ZA}, {ZB, YD, ZC}, {YC, ZB, ZC},
                             {XD, ZA, XA}, {ZB, XD, XA}, {ZD, XC, XB}, {XC, ZC, XB}, {ZA, YA, XA}, {YB, ZA, XD}, {ZD, XB, YA}, {XC, ZD, YB},
                             {ZB, XA, YD}, {XD, ZB, YC}, {XB, ZC, YD}, {ZC, XC, YC}};
    std::vector<int> f;
    for (const auto& b : base) { f.push_back(b[0]); f.push_back(b[2]); f.push_back(b[1]); }   // swap_orientation (mesh.cpp:1183-1184)
    auto normalise = [](std::vector<double>& c) {
        for (size_t i = 0; i < c.size() / 3; ++i) {
            const double x = c[3 * i], y = c[3 * i + 1], z = c[3 * i + 2];
            const double nrm = std::sqrt(x * x + y * y + z * z);
            if (nrm > 1e-8) { c[3 * i] = x / nrm; c[3 * i + 1] = y / nrm; c[3 * i + 2] = z / nrm; }
        }
    };
    for (int level = 0; level < n; ++level) {
        const int nv = (int)v.size() / 3, nt = (int)f.size() / 3;
        std::unordered_map<uint64_t, int> mid;
        mid.reserve((size_t)nt * 2);
        std::vector<int> nf;
        nf.reserve((size_t)nt * 12);
        auto midpoint = [&](int a, int b) {   // (a + b) / 2 with a, b as written in mesh.cpp:928-936
            const uint64_t key = ((uint64_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b);
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            const int id = (int)v.size() / 3;
            for (int k = 0; k < 3; ++k) v.push_back((v[3 * (size_t)a + k] + v[3 * (size_t)b + k]) / 2);
            mid.emplace(key, id);
            return id;
        };
        for (int i = 0; i < nt; ++i) {
            const int v0 = f[3 * (size_t)i], v1 = f[3 * (size_t)i + 1], v2 = f[3 * (size_t)i + 2];
            const int p0 = midpoint(v1, v2), p1 = midpoint(v0, v2), p2 = midpoint(v0, v1);
            const int kids[12] = {p2, p0, p1, p1, v0, p2, p0, v2, p1, p2, v1, p0};
            nf.insert(nf.end(), kids, kids + 12);
        }
        f.swap(nf);
        (void)nv;
        // the LAST level is normalised on the Mesh object below: the reference creates its Triangle objects (and caches their
        // areas, triangle.cpp:31) before that normalisation, and compute_vertex_area later reads those cached values
        if (level + 1 < n) normalise(v);
    }
    Mesh ret;
    const int nv = (int)v.size() / 3, nt = (int)f.size() / 3;
    for (int i = 0; i < nv; ++i) ret.push_point(std::make_shared<newresampler::Mpoint>(v[3 * (size_t)i], v[3 * (size_t)i + 1], v[3 * (size_t)i + 2], i));
    const auto& pts = ret.get_all_points();
    if (n > 0) {
        for (int i = 0; i < nt; ++i) ret.push_triangle(Triangle(pts[f[3 * (size_t)i]], pts[f[3 * (size_t)i + 1]], pts[f[3 * (size_t)i + 2]], i));
        for (auto i = ret.vbegin(); i != ret.vend(); i++) (*i)->normalize();   // mesh.cpp:1006-1007
    } else {   // the bare icosahedron keeps the adjacency lists of the un-swapped faces (pushed first, swapped after: mesh.cpp:1162-1184)
        for (int i = 0; i < nt; ++i) ret.push_triangle(Triangle(pts[f[3 * (size_t)i]], pts[f[3 * (size_t)i + 2]], pts[f[3 * (size_t)i + 1]], i));
        for (auto i = ret.tbegin(); i != ret.tend(); i++) i->swap_orientation();
    }
    ret.push_pvalues(std::vector<double>((size_t)nv, 0.0));
    return ret;
}

}  // namespace newresampler_gpu
