// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// extern "C" driver around the UNMODIFIED reference resampler library, compiled in place
// from /root/reference/libraries/msm-newresampler/src/*.cpp by oracle/Makefile into
// oracle/_ref/libref_newresampler.so.  It only calls the reference's public interface
// (newresampler::Mesh / Octree / Resampler / free functions, resampler.h:38-53,
// octree.h:39-59) and is used (a) to pin oracle/msm_oracle.cpp, (b) to generate the
// fixtures under tests/golden/, (c) as the "reference" CPU baseline of bench.py.
// Built with -fno-access-control so the tree dump can walk Octree::octree_root.
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "resampler.h"

using namespace newresampler;

namespace {

Mesh* build_mesh(int nv, const double* xyz, int nt, const int* tri) {
    Mesh tmp;
    for (int i = 0; i < nv; ++i)
        tmp.push_point(std::make_shared<Mpoint>(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], i));
    auto& pts = tmp.get_all_points();
    for (int t = 0; t < nt; ++t) {
        Triangle tr(pts[tri[3 * t]], pts[tri[3 * t + 1]], pts[tri[3 * t + 2]], t);
        tmp.push_triangle(tr);
    }
    tmp.initialize_pvalues(1);
    // copy => triangles re-created, cached areas refreshed (mesh.cpp:37-53)
    return new Mesh(tmp);
}

void dump_node(const Node* n, std::vector<int>& kinds, std::vector<int>& counts, std::vector<int>& tris) {
    // pre-order; children visited in the reference's [i][j][k] nesting
    kinds.push_back(n->is_leaf ? 1 : 0);
    counts.push_back(n->triangles_size());
    for (int i = 0; i < n->triangles_size(); ++i) tris.push_back(n->get_triangle(i).get_no());
    if (!n->is_leaf)
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) dump_node(n->children[i][j][k], kinds, counts, tris);
}

int flatten_weights(const std::vector<std::map<int, double>>& w, int* rowptr, int* col, double* val, int cap) {
    int pos = 0;
    for (size_t r = 0; r < w.size(); ++r) {
        rowptr[r] = pos;
        for (const auto& it : w[r]) {
            if (pos < cap) { col[pos] = it.first; val[pos] = it.second; }
            ++pos;
        }
    }
    rowptr[w.size()] = pos;
    return pos;
}

} // namespace

extern "C" {

void* ref_mesh_new(int nv, const double* xyz, int nt, const int* tri) {
    try { return build_mesh(nv, xyz, nt, tri); } catch (...) { return nullptr; }
}

// make_mesh_from_icosa(n) (mesh.cpp:1111) then true_rescale(rad) (mesh.cpp:1210), then copy.
void* ref_mesh_icosa(int n, double rad) {
    Mesh m = make_mesh_from_icosa(n);
    true_rescale(m, rad);
    return new Mesh(m);
}

void ref_mesh_free(void* m) { delete static_cast<Mesh*>(m); }
int ref_mesh_nvertices(void* m) { return static_cast<Mesh*>(m)->nvertices(); }
int ref_mesh_ntriangles(void* m) { return static_cast<Mesh*>(m)->ntriangles(); }

void ref_mesh_export(void* mp, double* xyz, int* tri) {
    Mesh* m = static_cast<Mesh*>(mp);
    for (int i = 0; i < m->nvertices(); ++i) {
        const Point& p = m->get_coord(i);
        xyz[3 * i] = p.X; xyz[3 * i + 1] = p.Y; xyz[3 * i + 2] = p.Z;
    }
    for (int t = 0; t < m->ntriangles(); ++t)
        for (int k = 0; k < 3; ++k) tri[3 * t + k] = m->get_triangle_vertexID(t, k);
}

// data is channel-major [D][V], the reference's own pvalues layout (mesh.h:44)
void ref_mesh_set_pvalues(void* mp, int D, const double* data) {
    Mesh* m = static_cast<Mesh*>(mp);
    m->initialize_pvalues(D);
    const int V = m->nvertices();
    for (int d = 0; d < D; ++d)
        for (int v = 0; v < V; ++v) m->set_pvalue(v, data[(size_t)d * V + v], d);
}

void ref_vertex_areas(void* mp, double* out) {
    Mesh* m = static_cast<Mesh*>(mp);
    for (int i = 0; i < m->nvertices(); ++i) out[i] = compute_vertex_area(i, *m);
}

void* ref_octree_new(void* mesh) { return new Octree(*static_cast<Mesh*>(mesh)); }
void ref_octree_free(void* t) { delete static_cast<Octree*>(t); }

// status: 0 ok, 1 = "Point is not in the bounding box" (octree.cpp:158), 2 = no triangle (octree.cpp:211)
void ref_octree_query(void* tp, int n, const double* pts, int* out_tri, int* out_vertex, int* status, int nthreads) {
    Octree* t = static_cast<Octree*>(tp);
    #pragma omp parallel for num_threads(nthreads)
    for (int i = 0; i < n; ++i) {
        Point p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        try {
            Triangle tr = t->get_closest_triangle(p);
            out_tri[i] = tr.get_no();
            if (out_vertex) out_vertex[i] = t->get_closest_vertex_ID(p);
            status[i] = 0;
        } catch (MeshException& e) {
            out_tri[i] = -1;
            if (out_vertex) out_vertex[i] = -1;
            status[i] = (std::strstr(e.what(), "bounding box") != nullptr) ? 1 : 2;
        }
    }
}

// Pre-order dump of the pointer tree. Returns number of nodes; arrays sized by caller
// (call once with caps = 0 to get sizes through n_tris_out).
int ref_octree_dump(void* tp, int* kinds, int* counts, int node_cap, int* tris, int tri_cap, int* n_tris_out) {
    Octree* t = static_cast<Octree*>(tp);
    std::vector<int> k, c, tr;
    dump_node(t->octree_root, k, c, tr);
    if ((int)k.size() <= node_cap) {
        std::memcpy(kinds, k.data(), k.size() * sizeof(int));
        std::memcpy(counts, c.data(), c.size() * sizeof(int));
    }
    if ((int)tr.size() <= tri_cap) std::memcpy(tris, tr.data(), tr.size() * sizeof(int));
    *n_tris_out = (int)tr.size();
    return (int)k.size();
}

// Resampler::get_barycentric_weights(low, orig, oct) (resampler.cpp:142): per query the
// std::map<int,double> in key order. out arrays [3*N]; n_entries[N] (<3 for degenerate ids).
int ref_bary_weights(void* low, void* orig, void* oct, int* idx, double* w, int* n_entries, int nthreads) {
    Resampler r;
    try {
        auto ws = r.get_barycentric_weights(*static_cast<Mesh*>(low), *static_cast<Mesh*>(orig),
                                            *static_cast<Octree*>(oct), nthreads);
        for (size_t k = 0; k < ws.size(); ++k) {
            int j = 0;
            for (const auto& it : ws[k]) { idx[3 * k + j] = it.first; w[3 * k + j] = it.second; ++j; }
            n_entries[k] = j;
            for (; j < 3; ++j) { idx[3 * k + j] = -1; w[3 * k + j] = 0.0; }
        }
    } catch (MeshException&) { return 1; }
    return 0;
}

// Resampler::get_adaptive_barycentric_weights (resampler.cpp:72) -> CSR. Returns nnz
// (or -1 on exception). rowptr[N_low+1]; col/val filled up to cap.
int ref_adaptive_weights(void* in, void* low, int nthreads, int* rowptr, int* col, double* val, int cap) {
    Resampler r;
    try {
        auto ws = r.get_adaptive_barycentric_weights(*static_cast<Mesh*>(in), *static_cast<Mesh*>(low), nthreads);
        return flatten_weights(ws, rowptr, col, val, cap);
    } catch (MeshException&) { return -1; }
}

// metric_resample(in, low, nthreads) (resampler.cpp:304). out is channel-major [D][N_low].
// Returns wall seconds of the call itself (steady_clock), negative on exception.
double ref_metric_resample(void* in, void* low, int nthreads, double* out) {
    Mesh* mi = static_cast<Mesh*>(in);
    Mesh* ml = static_cast<Mesh*>(low);
    try {
        auto t0 = std::chrono::steady_clock::now();
        Mesh res = metric_resample(*mi, *ml, nthreads);
        auto t1 = std::chrono::steady_clock::now();
        const int D = res.get_dimension(), N = res.nvertices();
        if (out)
            for (int d = 0; d < D; ++d)
                for (int v = 0; v < N; ++v) out[(size_t)d * N + v] = res.get_pvalue(v, d);
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (MeshException&) { return -1.0; }
}

// Plain (non-adaptive) barycentric resample assembled from the reference's own pieces:
// Octree(in) + get_barycentric_weights(low,in,oct,nthreads) + the interpolation loop of
// resampler.cpp:40-52. Returns wall seconds.
double ref_bary_resample(void* in, void* low, int nthreads, double* out) {
    Mesh* mi = static_cast<Mesh*>(in);
    Mesh* ml = static_cast<Mesh*>(low);
    Resampler r;
    try {
        auto t0 = std::chrono::steady_clock::now();
        Octree oct(*mi);
        auto ws = r.get_barycentric_weights(*ml, *mi, oct, nthreads);
        const int D = mi->get_dimension(), N = ml->nvertices();
        std::vector<double> tmp((size_t)D * N);
        for (int d = 0; d < D; ++d) {
            #pragma omp parallel for num_threads(nthreads)
            for (int k = 0; k < N; ++k) {
                double val = 0.0;
                for (const auto& it : ws[k]) val += mi->get_pvalue(it.first, d) * it.second;
                tmp[(size_t)d * N + k] = val;
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        if (out) std::memcpy(out, tmp.data(), tmp.size() * sizeof(double));
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (MeshException&) { return -1.0; }
}

// sphere_project_warp(sphere, from, to) (resampler.cpp:311); sphere modified in place, exported to out_xyz.
int ref_sphere_project_warp(void* sphere, void* from, void* to, int nthreads, double* out_xyz) {
    Mesh* s = static_cast<Mesh*>(sphere);
    try { sphere_project_warp(*s, *static_cast<Mesh*>(from), *static_cast<Mesh*>(to), nthreads); }
    catch (MeshException&) { return 1; }
    for (int i = 0; i < s->nvertices(); ++i) {
        const Point& p = s->get_coord(i);
        out_xyz[3 * i] = p.X; out_xyz[3 * i + 1] = p.Y; out_xyz[3 * i + 2] = p.Z;
    }
    return 0;
}

// surface_resample(anatOrig, sphOrig, sphLow) (resampler.cpp:284)
int ref_surface_resample(void* anat, void* sph, void* low, int nthreads, double* out_xyz) {
    try {
        Mesh r = surface_resample(*static_cast<Mesh*>(anat), *static_cast<Mesh*>(sph), *static_cast<Mesh*>(low), nthreads);
        for (int i = 0; i < r.nvertices(); ++i) {
            const Point& p = r.get_coord(i);
            out_xyz[3 * i] = p.X; out_xyz[3 * i + 1] = p.Y; out_xyz[3 * i + 2] = p.Z;
        }
    } catch (MeshException&) { return 1; }
    return 0;
}

// nearest_neighbour_interpolation(orig, low) (resampler.cpp:232). out [D][N_low]
int ref_nn_resample(void* in, void* low, int nthreads, double* out) {
    try {
        Mesh r = nearest_neighbour_interpolation(*static_cast<Mesh*>(in), *static_cast<Mesh*>(low), nthreads);
        const int D = r.get_dimension(), N = r.nvertices();
        for (int d = 0; d < D; ++d)
            for (int v = 0; v < N; ++v) out[(size_t)d * N + v] = r.get_pvalue(v, d);
    } catch (MeshException&) { return 1; }
    return 0;
}

// ---- exclusion masks: the reference's functions with EXCL (a Mesh with one channel), resampler.cpp:30-70, 169-258 ----
static std::shared_ptr<Mesh> excl_mesh(const Mesh& geometry, const double* excl) {
    auto m = std::make_shared<Mesh>(geometry);
    m->initialize_pvalues(1);
    for (int v = 0; v < m->nvertices(); ++v) m->set_pvalue(v, excl[v], 0);
    return m;
}
static void export_pvalues(const Mesh& r, double* out) {
    const int D = r.get_dimension(), N = r.nvertices();
    for (int d = 0; d < D; ++d)
        for (int v = 0; v < N; ++v) out[(size_t)d * N + v] = r.get_pvalue(v, d);
}
// metric_resample(in, low, nthreads, EXCL): out [D][N_low], excl_out [N_low] = the resampled mask that replaces *EXCL
int ref_metric_resample_excl(void* in, void* low, int nthreads, const double* excl, double* out, double* excl_out) {
    try {
        Mesh* mi = static_cast<Mesh*>(in);
        std::shared_ptr<Mesh> E = excl_mesh(*mi, excl);
        Mesh res = metric_resample(*mi, *static_cast<Mesh*>(low), nthreads, E);
        export_pvalues(res, out);
        for (int v = 0; v < E->nvertices(); ++v) excl_out[v] = E->get_pvalue(v);
    } catch (MeshException&) { return 1; }
    return 0;
}
// get_adaptive_barycentric_weights(in, low, nthreads, EXCL) as CSR
int ref_adaptive_weights_excl(void* in, void* low, int nthreads, const double* excl, int* rowptr, int* col, double* val, int cap) {
    try {
        Mesh* mi = static_cast<Mesh*>(in);
        Resampler r;
        std::vector<std::map<int, double>> w = r.get_adaptive_barycentric_weights(*mi, *static_cast<Mesh*>(low), nthreads, excl_mesh(*mi, excl));
        int pos = 0;
        for (size_t k = 0; k < w.size(); ++k) {
            rowptr[k] = pos;
            for (const auto& e : w[k]) { if (pos < cap) { col[pos] = e.first; val[pos] = e.second; } ++pos; }
        }
        rowptr[w.size()] = pos;
        return pos;
    } catch (MeshException&) { return -1; }
}
// smooth_data(orig, low, sigma, nthreads, EXCL or none): out [D][N_low]; excl NULL = no mask
int ref_smooth_data(void* orig, void* low, double sigma, int nthreads, const double* excl, double* out, double* excl_out) {
    try {
        Mesh* mo = static_cast<Mesh*>(orig);
        std::shared_ptr<Mesh> E = excl ? excl_mesh(*mo, excl) : std::shared_ptr<Mesh>();
        Mesh res = smooth_data(*mo, *static_cast<Mesh*>(low), sigma, nthreads, E);
        export_pvalues(res, out);
        if (E) for (int v = 0; v < E->nvertices(); ++v) excl_out[v] = E->get_pvalue(v);
    } catch (MeshException&) { return 1; }
    return 0;
}
int ref_nn_resample_excl(void* in, void* low, int nthreads, const double* excl, double* out, double* excl_out) {
    try {
        Mesh* mi = static_cast<Mesh*>(in);
        std::shared_ptr<Mesh> E = excl_mesh(*mi, excl);
        Mesh res = nearest_neighbour_interpolation(*mi, *static_cast<Mesh*>(low), nthreads, E);
        export_pvalues(res, out);
        for (int v = 0; v < E->nvertices(); ++v) excl_out[v] = E->get_pvalue(v);
    } catch (MeshException&) { return 1; }
    return 0;
}

// estimate_rotation_matrix(ci, index) (point.cpp:97) -> row-major 3x3
void ref_rotation_matrix(const double* ci, const double* index, double* R) {
    NEWMAT::Matrix M = estimate_rotation_matrix(Point(ci[0], ci[1], ci[2]), Point(index[0], index[1], index[2]));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = M(i + 1, j + 1);
}

} // extern "C"
