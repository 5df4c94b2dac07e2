// In-process drop-in check (TEST INFRASTRUCTURE, built by `make -C oracle adapter_check` into oracle/_ref/):
// the reference's own Mesh objects go through include/newmsm_b200/resampler_adapter.hpp to the GPU and
// the results are compared, bit for bit, with the reference's own CPU functions called on the same objects.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "newmsm_b200/resampler_adapter.hpp"

using namespace newresampler;

static int failures = 0;
#define EXPECT(cond, what)                                   \
    do {                                                     \
        if (!(cond)) { std::printf("FAIL: %s\n", what); ++failures; } \
        else std::printf("ok:   %s\n", what);                \
    } while (0)

static Mesh sphere(int level, double rx, double ry, double rz) {
    Mesh m = make_mesh_from_icosa(level);
    true_rescale(m, RAD);
    if (rx != 0 || ry != 0 || rz != 0) {
        for (int i = 0; i < m.nvertices(); ++i) {
            Point p = m.get_coord(i);
            Point q(p.X * std::cos(rz) - p.Y * std::sin(rz), p.X * std::sin(rz) + p.Y * std::cos(rz), p.Z);
            Point r(q.X * std::cos(ry) + q.Z * std::sin(ry), q.Y, -q.X * std::sin(ry) + q.Z * std::cos(ry));
            Point s(r.X, r.Y * std::cos(rx) - r.Z * std::sin(rx), r.Y * std::sin(rx) + r.Z * std::cos(rx));
            s.normalize();
            m.set_coord(i, s * RAD);
        }
    }
    return Mesh(m);   // copy => triangles rebuilt, cached areas refreshed (mesh.cpp:37-53)
}

static bool same_pvalues(const Mesh& a, const Mesh& b) {
    if (a.get_dimension() != b.get_dimension() || a.nvertices() != b.nvertices()) return false;
    for (int d = 0; d < a.get_dimension(); ++d)
        for (int v = 0; v < a.nvertices(); ++v) {
            const double x = a.get_pvalue(v, d), y = b.get_pvalue(v, d);
            if (std::memcmp(&x, &y, sizeof(double)) != 0) return false;
        }
    return true;
}
static bool same_coords(const Mesh& a, const Mesh& b) {
    for (int v = 0; v < a.nvertices(); ++v) {
        const Point &p = a.get_coord(v), &q = b.get_coord(v);
        if (p.X != q.X || p.Y != q.Y || p.Z != q.Z) return false;
    }
    return true;
}

int main() {
    try {
        Mesh in = sphere(5, 0, 0, 0), low = sphere(4, 0.013, 0.021, 0.034);
        in.initialize_pvalues(3);
        for (int d = 0; d < 3; ++d)
            for (int v = 0; v < in.nvertices(); ++v) {
                const Point& p = in.get_coord(v);
                in.set_pvalue(v, std::cos(0.05 * p.X * (d + 1) + 0.3) + std::sin(0.04 * p.Y - 0.02 * p.Z * d), d);
            }

        // metric_resample (adaptive barycentric), resampler.cpp:304
        Mesh cpu = metric_resample(in, low, 1);
        Mesh gpu = newresampler_gpu::metric_resample(in, low, 1);
        EXPECT(same_pvalues(cpu, gpu), "metric_resample: GPU == reference CPU, bit for bit");

        // adaptive weights as vector<map<int,double>>
        Resampler rc;
        newresampler_gpu::Resampler rg;
        EXPECT(rc.get_adaptive_barycentric_weights(in, low, 1) == rg.get_adaptive_barycentric_weights(in, low, 1),
               "get_adaptive_barycentric_weights: identical maps");

        // octree queries + barycentric weights
        Octree oc(in);
        newresampler_gpu::Octree og(in);
        std::vector<Point> pts;
        for (int v = 0; v < low.nvertices(); ++v) pts.push_back(low.get_coord(v));
        const std::vector<int> ids = og.get_closest_triangle_ids(pts);
        bool ok = true;
        for (size_t i = 0; i < pts.size(); ++i) ok = ok && oc.get_closest_triangle(pts[i]).get_no() == ids[i];
        EXPECT(ok, "get_closest_triangle: identical triangle ids");
        EXPECT(oc.get_closest_triangle(pts[7]).get_no() == og.get_closest_triangle(pts[7]).get_no(), "get_closest_triangle(Point): same Triangle");
        EXPECT(oc.get_closest_vertex_ID(pts[11]) == og.get_closest_vertex_ID(pts[11]), "get_closest_vertex_ID");
        EXPECT(rc.get_barycentric_weights(low, in, oc, 1) == rg.get_barycentric_weights(low, in, og, 1), "get_barycentric_weights: identical maps");

        // sphere_project_warp, resampler.cpp:311
        Mesh to = sphere(5, 0.0, 0.01, -0.02);
        Mesh s1 = low, s2 = low;
        sphere_project_warp(s1, in, to, 1);
        newresampler_gpu::sphere_project_warp(s2, in, to, 1);
        EXPECT(same_coords(s1, s2), "sphere_project_warp: identical coordinates");

        // surface_resample / nearest neighbour
        EXPECT(same_coords(surface_resample(to, in, low, 1), newresampler_gpu::surface_resample(to, in, low, 1)), "surface_resample: identical coordinates");
        EXPECT(same_pvalues(nearest_neighbour_interpolation(in, low, 1), newresampler_gpu::nearest_neighbour_interpolation(in, low, 1)),
               "nearest_neighbour_interpolation: identical values");

        // make_mesh_from_icosa (mesh.cpp:1111): identical object state, including the Triangle areas cached BEFORE the last
        // normalisation and the per-point neighbour / triangle lists (the CPU generator needs no device: checked for 0..5)
        bool gen_ok = true;
        for (int level = 0; level <= 5; ++level) {
            Mesh a = make_mesh_from_icosa(level), b = newresampler_gpu::make_mesh_from_icosa(level);
            gen_ok = gen_ok && a.nvertices() == b.nvertices() && a.ntriangles() == b.ntriangles() && same_coords(a, b) && same_pvalues(a, b);
            for (int t = 0; gen_ok && t < a.ntriangles(); ++t) {
                for (int k = 0; k < 3; ++k) gen_ok = gen_ok && a.get_triangle_vertexID(t, k) == b.get_triangle_vertexID(t, k);
                const double x = a.get_triangle_area(t), y = b.get_triangle_area(t);
                gen_ok = gen_ok && std::memcmp(&x, &y, sizeof(double)) == 0 && a.get_triangle(t).get_no() == b.get_triangle(t).get_no();
            }
            for (int v = 0; gen_ok && v < a.nvertices(); ++v) {
                gen_ok = gen_ok && std::vector<int>(a.nbegin(v), a.nend(v)) == std::vector<int>(b.nbegin(v), b.nend(v));
                gen_ok = gen_ok && std::vector<int>(a.tIDbegin(v), a.tIDend(v)) == std::vector<int>(b.tIDbegin(v), b.tIDend(v));
            }
        }
        EXPECT(gen_ok, "make_mesh_from_icosa: identical vertices, faces, cached areas and adjacency lists for ico0..ico5");

        // meshes as the reference's call sites make them: generated, rescaled, moved, NOT copied -> Triangle::area is stale
        // (triangle.cpp:31,39) and compute_vertex_area (mesh.cpp:1275) reads it. The adapter mirrors the object's cached values.
        {
            Mesh stale_low = make_mesh_from_icosa(4);
            true_rescale(stale_low, RAD);
            Mesh stale_in = in;
            Mesh moved = sphere(5, 0.004, -0.003, 0.002);
            for (int v = 0; v < stale_in.nvertices(); ++v) stale_in.set_coord(v, moved.get_coord(v));
            EXPECT(same_pvalues(metric_resample(stale_in, stale_low, 1), newresampler_gpu::metric_resample(stale_in, stale_low, 1)),
                   "metric_resample on meshes with stale cached triangle areas: GPU == reference CPU, bit for bit");
            EXPECT(!same_pvalues(metric_resample(stale_in, stale_low, 1), metric_resample(Mesh(stale_in), Mesh(stale_low), 1)),
                   "(those stale areas do change the result, i.e. the case is not vacuous)");
        }

        // smooth_data (resampler.cpp:169): device neighbourhood scan + host-libm Gaussian weights
        {
            Mesh a = in, b = in;
            EXPECT(same_pvalues(smooth_data(a, a, 4.0, 1), newresampler_gpu::smooth_data(b, b, 4.0, 1)), "smooth_data (sigma 4, D = 3): identical values");
            Mesh c = low, d = low;
            c.initialize_pvalues(1); d.initialize_pvalues(1);
            for (int v = 0; v < c.nvertices(); ++v) { c.set_pvalue(v, std::sin(0.07 * c.get_coord(v).Z), 0); d.set_pvalue(v, std::sin(0.07 * d.get_coord(v).Z), 0); }
            EXPECT(same_pvalues(smooth_data(c, c, 2.0, 1), newresampler_gpu::smooth_data(d, d, 2.0, 1)), "smooth_data (sigma 2, rotated ico4): identical values");
        }

        // exclusion masks (resampler.cpp:30-140, 169-258 with EXCL): data AND the replaced *EXCL mesh, reference vs adapter
        {
            auto mask = [&](const Mesh& g) {
                auto e = std::make_shared<Mesh>(g);
                e->initialize_pvalues(1);
                for (int v = 0; v < e->nvertices(); ++v) {
                    const double z = e->get_coord(v).Z;
                    e->set_pvalue(v, z > 55.0 ? 0.0 : (z > 30.0 ? 0.5 + 0.5 * (55.0 - z) / 25.0 : 1.0), 0);
                }
                return e;
            };
            std::shared_ptr<Mesh> e1 = mask(in), e2 = mask(in);
            EXPECT(same_pvalues(metric_resample(in, low, 1, e1), newresampler_gpu::metric_resample(in, low, 1, e2)) && same_pvalues(*e1, *e2),
                   "metric_resample with EXCL: identical data and identical resampled mask");
            e1 = mask(in); e2 = mask(in);
            EXPECT(rc.get_adaptive_barycentric_weights(in, low, 1, e1) == rg.get_adaptive_barycentric_weights(in, low, 1, e2),
                   "get_adaptive_barycentric_weights with EXCL: identical maps");
            e1 = mask(in); e2 = mask(in);
            Mesh a = in, b = in;
            EXPECT(same_pvalues(smooth_data(a, a, 4.0, 1, e1), newresampler_gpu::smooth_data(b, b, 4.0, 1, e2)) && same_pvalues(*e1, *e2),
                   "smooth_data with EXCL: identical data and identical new mask");
            e1 = mask(in); e2 = mask(in);
            EXPECT(same_pvalues(nearest_neighbour_interpolation(a, low, 1, e1), newresampler_gpu::nearest_neighbour_interpolation(b, low, 1, e2)) &&
                       same_pvalues(*e1, *e2),
                   "nearest_neighbour_interpolation with EXCL: identical data and identical new mask");
        }

        // error behaviour: the reference's exception with the reference's message (octree.cpp:158)
        bool threw = false;
        try { og.get_closest_triangle(Point(0, 0, 150)); } catch (MeshException& e) { threw = std::strstr(e.what(), "bounding box") != nullptr; }
        EXPECT(threw, "out-of-box query throws MeshException(\"Point is not in the bounding box of the mesh\")");
    } catch (std::exception& e) {
        std::printf("FAIL: unexpected exception\n");
        return 2;
    }
    std::printf(failures ? "ADAPTER CHECK FAILED (%d)\n" : "ADAPTER CHECK PASSED\n", failures);
    return failures ? 1 : 0;
}
