// TEST INFRASTRUCTURE ONLY — see msm_oracle.h. CPU restatement of the reference hot path,
// written from the reference's behaviour (file:line cited per function), own data structures.
// Compiled with -O2 -ffp-contract=off (no FMA) so FP64 decisions are IEEE-reproducible.
#include "msm_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <limits>
#include <map>
#include <numeric>
#include <vector>

namespace {

constexpr double EPS = 1e-8;   // point.h:31
constexpr double RAD = 100.0;  // point.h:32
constexpr int MAX_TRIANGLES = 50; // node.h:33
constexpr double MESH_BOUNDS = 101.0; // octree.h:37
constexpr double NOT_IN_TRIANGLE = -1.0; // octree.h:35
constexpr double DMAX = std::numeric_limits<double>::max();

struct P3 { double X, Y, Z; };

inline P3 sub(const P3& a, const P3& b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z}; }      // point.cpp:185
inline P3 mul(const P3& a, double d) { return {a.X * d, a.Y * d, a.Z * d}; }                // point.cpp:198
inline double dot(const P3& a, const P3& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }   // point.cpp:173
inline P3 cross(const P3& a, const P3& b) {                                                 // point.cpp:177
    return {a.Y * b.Z - a.Z * b.Y, b.X * a.Z - b.Z * a.X, a.X * b.Y - b.X * a.Y};
}
inline double norm(const P3& a) { return std::sqrt(a.X * a.X + a.Y * a.Y + a.Z * a.Z); }     // point.h:44
inline void normalize(P3& a) {                                                              // point.cpp:26
    double n = norm(a);
    if (n > EPS) { a.X /= n; a.Y /= n; a.Z /= n; }
}
inline P3 matvec(const double* M, const P3& v) {                                            // point.cpp:202
    return {M[0] * v.X + M[1] * v.Y + M[2] * v.Z,
            M[3] * v.X + M[4] * v.Y + M[5] * v.Z,
            M[6] * v.X + M[7] * v.Y + M[8] * v.Z};
}

// point.cpp:46-61
inline P3 project_point(const P3& vb, const P3& v1, const P3& v2, const P3& v3) {
    P3 s1 = sub(v3, v1); normalize(s1);
    P3 s2 = sub(v2, v1); normalize(s2);
    P3 s3 = cross(s1, s2); normalize(s3);
    double si = dot(s3, v1) / dot(s3, vb);
    return mul(vb, si);
}
// point.cpp:36-44
inline bool same_side(const P3& p1, const P3& p2, const P3& a, const P3& b) {
    return dot(cross(sub(b, a), sub(p1, a)), cross(sub(b, a), sub(p2, a))) > -EPS;
}
inline bool point_in_triangle(const P3& p, const P3& a, const P3& b, const P3& c) {
    return same_side(p, a, b, c) && same_side(p, b, c, a) && same_side(p, c, a, b);
}
// point.cpp:68-75
inline double compute_area(const P3& v0, const P3& v1, const P3& v2) {
    return 0.5 * norm(cross(sub(v1, v0), sub(v2, v0)));
}
// triangle.cpp:85-122
inline double dist_to_point(const P3& x0, const P3& x1, const P3& x2, const P3& x3) {
    double d, dmin = DMAX;
    P3 u = sub(x2, x1);
    if (dot(sub(x0, x1), u) > 0 && dot(sub(x0, x2), u) < 0) {
        d = norm(cross(sub(x0, x1), sub(x0, x2))) / norm(sub(x2, x1));
        if (d < dmin) dmin = d;
    }
    u = sub(x3, x1);
    if (dot(sub(x0, x1), u) > 0 && dot(sub(x0, x3), u) < 0) {
        d = norm(cross(sub(x0, x1), sub(x0, x3))) / norm(sub(x3, x1));
        if (d < dmin) dmin = d;
    }
    u = sub(x3, x2);
    if (dot(sub(x0, x2), u) > 0 && dot(sub(x0, x3), u) < 0) {
        d = norm(cross(sub(x0, x2), sub(x0, x3))) / norm(sub(x3, x2));
        if (d < dmin) dmin = d;
    }
    d = norm(sub(x0, x1)); if (d < dmin) dmin = d;
    d = norm(sub(x0, x2)); if (d < dmin) dmin = d;
    d = norm(sub(x0, x3)); if (d < dmin) dmin = d;
    return dmin;
}

struct ONode {
    double b[3][3];       // [axis][lo, mid, hi]  (node.h:42)
    bool leaf = true;
    ONode* ch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // index = i*4+j*2+k
    ONode* parent = nullptr;
    std::vector<int> tris;
    ~ONode() { for (auto* c : ch) delete c; }
};

} // namespace

struct orc_octree {
    ONode* root = nullptr;
    int nv = 0, nt = 0;
    std::vector<P3> v;
    std::vector<int> tri;
    ~orc_octree() { delete root; }
    const P3& tv(int t, int k) const { return v[tri[3 * t + k]]; }
};

namespace {

// node.cpp:67-77 (closed box)
inline bool contains_point(const ONode* n, const P3& p) {
    const double q[3] = {p.X, p.Y, p.Z};
    for (int i = 0; i < 3; ++i) {
        if (q[i] < n->b[i][0]) return false;
        if (q[i] > n->b[i][2]) return false;
    }
    return true;
}
// node.cpp:112-120 (closed AABB overlap)
inline bool can_contain(const ONode* n, const double* lo, const double* hi) {
    for (int i = 0; i < 3; ++i)
        if (hi[i] < n->b[i][0] || lo[i] > n->b[i][2]) return false;
    return true;
}
// node.cpp:91-110
void make_children(ONode* n) {
    n->leaf = false;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                ONode* c = new ONode();
                const int o[3] = {i, j, k};
                for (int a = 0; a < 3; ++a) {
                    c->b[a][0] = n->b[a][o[a]];
                    c->b[a][2] = n->b[a][o[a] + 1];
                    c->b[a][1] = (c->b[a][0] + c->b[a][2]) / 2.0;
                }
                c->parent = n;
                n->ch[i * 4 + j * 2 + k] = c;
            }
}

void tri_aabb(const orc_octree* T, int t, double* lo, double* hi) { // octree.cpp:46-59
    const P3& a = T->tv(t, 0);
    lo[0] = hi[0] = a.X; lo[1] = hi[1] = a.Y; lo[2] = hi[2] = a.Z;
    for (int k = 1; k < 3; ++k) {
        const P3& p = T->tv(t, k);
        const double q[3] = {p.X, p.Y, p.Z};
        for (int i = 0; i < 3; ++i) {
            if (q[i] < lo[i]) lo[i] = q[i];
            if (q[i] > hi[i]) hi[i] = q[i];
        }
    }
}

// octree.cpp:63-141
void add_triangle(orc_octree* T, ONode* n, int t, const double* lo, const double* hi) {
    if (n->leaf) {
        n->tris.push_back(t);
        const int num = (int)n->tris.size();
        if (num >= MAX_TRIANGLES) {
            int total_size = 0, num_split = 0;
            double tlo[3], thi[3];
            for (int i = 0; i < num; ++i) {
                tri_aabb(T, n->tris[i], tlo, thi);
                int split_size = 8;
                for (int d = 0; d < 3; ++d) // node.cpp:79-89: containing_oct uses strict '<' against the midpoint
                    if ((tlo[d] < n->b[d][1]) == (thi[d] < n->b[d][1])) split_size >>= 1;
                total_size += split_size;
                if (split_size != 8) ++num_split;
            }
            if (num_split > 0 && total_size < 3 * num) {
                make_children(n);
                for (int i = 0; i < num; ++i) {
                    tri_aabb(T, n->tris[i], tlo, thi);
                    for (int c = 0; c < 8; ++c)
                        if (can_contain(n->ch[c], tlo, thi)) add_triangle(T, n->ch[c], n->tris[i], tlo, thi);
                }
                n->tris.clear();
            }
        }
    } else {
        for (int c = 0; c < 8; ++c)
            if (can_contain(n->ch[c], lo, hi)) add_triangle(T, n->ch[c], t, lo, hi);
    }
}

// octree.cpp:143-154
inline double distance_to_triangle(const orc_octree* T, const P3& pt, int t) {
    const P3 &v0 = T->tv(t, 0), &v1 = T->tv(t, 1), &v2 = T->tv(t, 2);
    P3 mP = project_point(pt, v0, v1, v2);
    if (point_in_triangle(mP, v0, v1, v2)) return dist_to_point(mP, v0, v1, v2);
    return NOT_IN_TRIANGLE;
}

// octree.cpp:156-214. Returns triangle id or -1; *status as in the header.
int closest_triangle(const orc_octree* T, const P3& pt, int* status, int* path) {
    if (!contains_point(T->root, pt)) { *status = 1; return -1; }
    int best = -1;
    double best_d = DMAX;
    const ONode* cur = T->root;
    while (!cur->leaf) {
        // the reference's range-for keeps iterating the ORIGINAL node's children while
        // reassigning current_oct, so the LAST containing child wins
        const ONode* next = cur;
        for (int c = 0; c < 8; ++c)
            if (contains_point(cur->ch[c], pt)) next = cur->ch[c];
        cur = next;
    }
    for (int t : cur->tris) {
        double d = distance_to_triangle(T, pt, t);
        if (d > NOT_IN_TRIANGLE && d < best_d) { best = t; best_d = d; }
    }
    int used = 0;
    if (best < 0) {
        if (!cur->parent) { *status = 2; return -1; } // reference dereferences nullptr here (App. A.4)
        used = 1;
        best_d = DMAX;
        for (int c = 0; c < 8; ++c)
            for (int t : cur->parent->ch[c]->tris) {
                double d = distance_to_triangle(T, pt, t);
                if (d > NOT_IN_TRIANGLE && d < best_d) { best = t; best_d = d; }
            }
    }
    if (best < 0) {
        used = 2;
        best_d = DMAX;
        for (int c = 0; c < 8; ++c)
            for (int t : cur->parent->ch[c]->tris)
                for (int k = 0; k < 3; ++k) {
                    double d = 2 * RAD * std::asin(norm(sub(T->tv(t, k), pt)) / (2 * RAD));
                    if (d < best_d) { best = t; best_d = d; }
                }
    }
    if (best < 0) { *status = 2; return -1; }
    *status = 0;
    if (path) *path = used;
    return best;
}

// octree.cpp:216-233
int closest_vertex_of(const orc_octree* T, const P3& pt, int t) {
    double dist = DMAX;
    int id = 0;
    for (int k = 0; k < 3; ++k) {
        double d = norm(sub(pt, T->tv(t, k)));
        if (d < dist) { id = T->tri[3 * t + k]; dist = d; }
    }
    return id;
}

// triangle.cpp:124-143 -> entries in ascending key order (std::map semantics)
int bary_weights_of(const orc_octree* T, const P3& p, int t, int* idx, double* w) {
    const P3 &v1 = T->tv(t, 0), &v2 = T->tv(t, 1), &v3 = T->tv(t, 2);
    P3 PP = project_point(p, v1, v2, v3);
    double Aa = compute_area(PP, v2, v3);
    double Ab = compute_area(PP, v1, v3);
    double Ac = compute_area(PP, v1, v2);
    double A = Aa + Ab + Ac;
    std::map<int, double> m;
    m[T->tri[3 * t]] = Aa / A;
    m[T->tri[3 * t + 1]] = Ab / A;
    m[T->tri[3 * t + 2]] = Ac / A;
    int j = 0;
    for (auto& it : m) { idx[j] = it.first; w[j] = it.second; ++j; }
    int n = j;
    for (; j < 3; ++j) { idx[j] = -1; w[j] = 0.0; }
    return n;
}

// triangle.cpp:145-157 (NO projection of vref)
inline void bary_interp_weights(const P3& v1, const P3& v2, const P3& v3, const P3& vref, double* w) {
    double Aa = compute_area(vref, v2, v3);
    double Ab = compute_area(vref, v1, v3);
    double Ac = compute_area(vref, v1, v2);
    double A = Aa + Ab + Ac;
    w[0] = Aa / A; w[1] = Ab / A; w[2] = Ac / A;
}

orc_octree* build_tree(int nv, const double* xyz, int nt, const int* tri) {
    orc_octree* T = new orc_octree();
    T->nv = nv; T->nt = nt;
    T->v.resize(nv);
    for (int i = 0; i < nv; ++i) T->v[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    T->tri.assign(tri, tri + 3 * (size_t)nt);
    T->root = new ONode();
    for (int a = 0; a < 3; ++a) { // octree.cpp:31-40, node.cpp:43-57
        T->root->b[a][0] = -MESH_BOUNDS;
        T->root->b[a][2] = MESH_BOUNDS;
        T->root->b[a][1] = (T->root->b[a][0] + T->root->b[a][2]) / 2.0;
    }
    double lo[3], hi[3];
    for (int t = 0; t < nt; ++t) { // octree.cpp:42-61
        tri_aabb(T, t, lo, hi);
        add_triangle(T, T->root, t, lo, hi);
    }
    return T;
}

void dump(const ONode* n, std::vector<int>& kinds, std::vector<int>& counts, std::vector<int>& tris) {
    kinds.push_back(n->leaf ? 1 : 0);
    counts.push_back((int)n->tris.size());
    tris.insert(tris.end(), n->tris.begin(), n->tris.end());
    if (!n->leaf) for (int c = 0; c < 8; ++c) dump(n->ch[c], kinds, counts, tris);
}

struct Weights { std::vector<std::map<int, double>> rows; };

// resampler.cpp:142-167
int bary_weight_maps(const orc_octree* T, int n, const double* pts, std::vector<std::map<int, double>>& out) {
    out.assign(n, {});
    int err = 0;
    #pragma omp parallel for
    for (int k = 0; k < n; ++k) {
        P3 p{pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]};
        int st;
        int t = closest_triangle(T, p, &st, nullptr);
        if (t < 0) {
            #pragma omp critical
            { if (!err) err = st; }
            continue;
        }
        int idx[3]; double w[3];
        int m = bary_weights_of(T, p, t, idx, w);
        for (int j = 0; j < m; ++j) out[k][idx[j]] = w[j];
    }
    return err;
}

void vertex_areas(int nv, const double* xyz, int nt, const int* tri, double* out) {
    // Triangle::calc_area (triangle.cpp:47-50) summed per vertex in push order (= ascending
    // triangle id, mesh.cpp:112-118), divided by the number of adjacent triangles (mesh.cpp:1275)
    std::vector<double> sum(nv, 0.0);
    std::vector<int> cnt(nv, 0);
    for (int t = 0; t < nt; ++t) {
        P3 a{xyz[3 * tri[3 * t]], xyz[3 * tri[3 * t] + 1], xyz[3 * tri[3 * t] + 2]};
        P3 b{xyz[3 * tri[3 * t + 1]], xyz[3 * tri[3 * t + 1] + 1], xyz[3 * tri[3 * t + 1] + 2]};
        P3 c{xyz[3 * tri[3 * t + 2]], xyz[3 * tri[3 * t + 2] + 1], xyz[3 * tri[3 * t + 2] + 2]};
        double area = 0.5 * norm(cross(sub(c, a), sub(b, a)));
        for (int k = 0; k < 3; ++k) { sum[tri[3 * t + k]] += area; cnt[tri[3 * t + k]]++; }
    }
    for (int i = 0; i < nv; ++i) out[i] = sum[i] / cnt[i];
}

// resampler.cpp:72-140 without EXCL, single-thread order
int adaptive_maps(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                  int nv_low, const double* xyz_low, int nt_low, const int* tri_low,
                  std::vector<std::map<int, double>>& adapt, const double* area_xyz_in = nullptr, const double* excl = nullptr) {
    // area_xyz_in: coordinates the source mesh had when its Triangle objects were last (re)created. Triangle::area is
    // cached at construction and NOT refreshed by Mesh::set_coord (triangle.cpp:31,39), so a mesh that was
    // copied and then moved (DiscreteGroupModel.cpp:94-103) keeps the vertex areas of its pre-move geometry.
    orc_octree* tin = build_tree(nv_in, xyz_in, nt_in, tri_in);
    std::vector<std::map<int, double>> forward, reverse;
    int e = bary_weight_maps(tin, nv_low, xyz_low, forward);
    // exclusion mask (resampler.cpp:100, 121): a target is used iff EXCL at its closest source vertex is non-zero
    std::vector<char> active(nv_low, 1);
    if (excl && !e)
        for (int n = 0; n < nv_low; ++n) {
            const P3 p{xyz_low[3 * n], xyz_low[3 * n + 1], xyz_low[3 * n + 2]};
            int st = 0;
            const int id = closest_triangle(tin, p, &st, nullptr);
            active[n] = id >= 0 && excl[closest_vertex_of(tin, p, id)] != 0;
        }
    delete tin;
    if (e) return e;
    orc_octree* tlow = build_tree(nv_low, xyz_low, nt_low, tri_low);
    e = bary_weight_maps(tlow, nv_in, xyz_in, reverse);
    delete tlow;
    if (e) return e;

    std::vector<double> newA(nv_low), oldA(nv_in), correction(nv_in, 0.0);
    vertex_areas(nv_low, xyz_low, nt_low, tri_low, newA.data());
    vertex_areas(nv_in, area_xyz_in ? area_xyz_in : xyz_in, nt_in, tri_in, oldA.data());
    std::vector<std::map<int, double>> rr(nv_low);
    adapt.assign(nv_low, {});
    for (int o = 0; o < nv_in; ++o)
        for (auto& it : reverse[o]) rr[it.first][o] = it.second;
    for (int n = 0; n < nv_low; ++n) {
        if (!active[n]) continue;
        if (rr[n].size() <= forward[n].size()) adapt[n] = forward[n];
        else adapt[n] = rr[n];
        for (auto& it : adapt[n]) { it.second *= newA[n]; correction[it.first] += it.second; }
    }
    for (int n = 0; n < nv_low; ++n) {
        if (!active[n]) continue;
        double ws = 0.0;
        for (auto& it : adapt[n]) { it.second *= oldA[it.first] / correction[it.first]; ws += it.second; }
        if (ws != 0.0) for (auto& it : adapt[n]) it.second /= ws;
    }
    return 0;
}

} // namespace

extern "C" {

orc_octree* orc_octree_build(int nv, const double* xyz, int nt, const int* tri) { return build_tree(nv, xyz, nt, tri); }
void orc_octree_free(orc_octree* t) { delete t; }

int orc_octree_dump(const orc_octree* t, int* kinds, int* counts, int node_cap, int* tris, int tri_cap, int* n_tris) {
    std::vector<int> k, c, tr;
    dump(t->root, k, c, tr);
    if ((int)k.size() <= node_cap) {
        std::memcpy(kinds, k.data(), k.size() * sizeof(int));
        std::memcpy(counts, c.data(), c.size() * sizeof(int));
    }
    if ((int)tr.size() <= tri_cap) std::memcpy(tris, tr.data(), tr.size() * sizeof(int));
    *n_tris = (int)tr.size();
    return (int)k.size();
}

void orc_octree_query(const orc_octree* t, int n, const double* pts, int* out_tri, int* out_vertex,
                      int* status, int* path, int nthreads) {
    #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < n; ++i) {
        P3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        int st = 0, pa = 0;
        int id = closest_triangle(t, p, &st, &pa);
        out_tri[i] = id;
        if (status) status[i] = st;
        if (path) path[i] = pa;
        if (out_vertex) out_vertex[i] = id >= 0 ? closest_vertex_of(t, p, id) : -1;
    }
}

int orc_bary_weights(const orc_octree* t, int n, const double* pts, int* idx, double* w, int* n_entries, int nthreads) {
    int err = 0;
    #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < n; ++i) {
        P3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        int st;
        int id = closest_triangle(t, p, &st, nullptr);
        if (id < 0) {
            for (int j = 0; j < 3; ++j) { idx[3 * i + j] = -1; w[3 * i + j] = 0.0; }
            if (n_entries) n_entries[i] = 0;
            #pragma omp critical
            { if (!err) err = st; }
            continue;
        }
        int m = bary_weights_of(t, p, id, idx + 3 * i, w + 3 * i);
        if (n_entries) n_entries[i] = m;
    }
    return err;
}

void orc_vertex_areas(int nv, const double* xyz, int nt, const int* tri, double* out) { vertex_areas(nv, xyz, nt, tri, out); }

int orc_adaptive_weights(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                         int nv_low, const double* xyz_low, int nt_low, const int* tri_low,
                         int* rowptr, int* col, double* val, int cap) {
    std::vector<std::map<int, double>> adapt;
    if (adaptive_maps(nv_in, xyz_in, nt_in, tri_in, nv_low, xyz_low, nt_low, tri_low, adapt)) return -1;
    int pos = 0;
    for (int r = 0; r < nv_low; ++r) {
        rowptr[r] = pos;
        for (auto& it : adapt[r]) { if (pos < cap) { col[pos] = it.first; val[pos] = it.second; } ++pos; }
    }
    rowptr[nv_low] = pos;
    return pos;
}

int orc_metric_resample(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                        int nv_low, const double* xyz_low, int nt_low, const int* tri_low,
                        int D, const double* feat_in, double* feat_out, int nthreads) {
    std::vector<std::map<int, double>> adapt;
    if (adaptive_maps(nv_in, xyz_in, nt_in, tri_in, nv_low, xyz_low, nt_low, tri_low, adapt)) return 1;
    for (int d = 0; d < D; ++d) { // resampler.cpp:40-52
        #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
        for (int k = 0; k < nv_low; ++k) {
            double val = 0.0;
            for (auto& it : adapt[k]) val += feat_in[(size_t)d * nv_in + it.first] * it.second;
            feat_out[(size_t)d * nv_low + k] = val;
        }
    }
    return 0;
}

// resampler.cpp:30-70 with EXCL: masked weights, sums that skip masked source vertices, and the resampled mask (excl_out);
// rowptr / col / val (optional) = the masked weight maps as CSR
int orc_metric_resample_excl(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                             int nv_low, const double* xyz_low, int nt_low, const int* tri_low,
                             int D, const double* feat_in, const double* excl, double* feat_out, double* excl_out,
                             int* rowptr, int* col, double* val, int cap) {
    std::vector<std::map<int, double>> adapt;
    if (adaptive_maps(nv_in, xyz_in, nt_in, tri_in, nv_low, xyz_low, nt_low, tri_low, adapt, nullptr, excl)) return -1;
    for (int d = 0; d < D; ++d)
        for (int k = 0; k < nv_low; ++k) {
            double v = 0.0;
            for (auto& it : adapt[k])
                if (excl[it.first] != 0) v += feat_in[(size_t)d * nv_in + it.first] * it.second;
            feat_out[(size_t)d * nv_low + k] = v;
        }
    for (int k = 0; k < nv_low; ++k) {
        double v = 0.0;
        for (auto& it : adapt[k])
            if (excl[it.first] != 0) v += excl[it.first] * it.second;
        excl_out[k] = v;
    }
    int pos = 0;
    if (rowptr) {
        for (int r = 0; r < nv_low; ++r) {
            rowptr[r] = pos;
            for (auto& it : adapt[r]) { if (pos < cap) { col[pos] = it.first; val[pos] = it.second; } ++pos; }
        }
        rowptr[nv_low] = pos;
    }
    return pos;
}

// resampler.cpp:232-258 with EXCL
int orc_nn_resample_excl(int n, const double* low_xyz, int nv, const double* xyz, int nt, const int* tri,
                         int D, const double* feat_in, const double* excl, double* feat_out, double* excl_out) {
    orc_octree* t = build_tree(nv, xyz, nt, tri);
    std::vector<int> tr(n), vx(n), st(n);
    orc_octree_query(t, n, low_xyz, tr.data(), vx.data(), st.data(), nullptr, 1);
    delete t;
    for (int i = 0; i < n; ++i) if (st[i]) return st[i];
    for (int i = 0; i < n; ++i) {
        const bool on = excl[vx[i]] != 0;
        excl_out[i] = on ? excl[vx[i]] : 0.0;
        for (int d = 0; d < D; ++d) feat_out[(size_t)d * n + i] = on ? feat_in[(size_t)d * nv + vx[i]] : 0.0;
    }
    return 0;
}

// resampler.cpp:169-230: Gaussian smoothing of orig's data over sphLow's vertices (n vertices, feat [D][n_feat] indexed by sphLow ids like
// the reference does), optional exclusion mask
int orc_smooth_data(int nv_orig, const double* orig_xyz, int nt_orig, const int* orig_tri, int n, const double* low_xyz, double sigma,
                    int D, const double* feat, const double* excl, double* out, double* excl_out) {
    orc_octree* t = build_tree(nv_orig, orig_xyz, nt_orig, orig_tri);
    const double ang = 4 * std::asin(sigma / (2 * RAD));
    for (size_t k = 0; k < (size_t)D * n; ++k) out[k] = 0.0;
    int err = 0;
    for (int i = 0; i < n && !err; ++i) {
        if (excl) excl_out[i] = 0.0;
        const P3 ci{low_xyz[3 * i], low_xyz[3 * i + 1], low_xyz[3 * i + 2]};
        int st = 0;
        const int id = closest_triangle(t, ci, &st, nullptr);
        if (id < 0) { err = st ? st : 2; break; }
        const int cv = closest_vertex_of(t, ci, id);
        P3 ref{low_xyz[3 * cv], low_xyz[3 * cv + 1], low_xyz[3 * cv + 2]};
        normalize(ref);
        std::vector<std::pair<int, double>> nb;
        for (int m = 0; m < n; ++m) {
            P3 actual{low_xyz[3 * m], low_xyz[3 * m + 1], low_xyz[3 * m + 2]};
            normalize(actual);
            if (dot(actual, ref) >= std::cos(ang)) nb.emplace_back(m, norm(sub(ref, actual)));
        }
        if (excl && !(excl[cv] > 0)) continue;
        double SUM = 0.0, excl_sum = 0.0;
        for (const auto& q : nb) {
            const double g = 2 * RAD * std::asin(q.second / (2 * RAD));
            double w = (1 / std::sqrt(2 * M_PI * sigma * sigma)) * std::exp(-(g * g) / (2 * sigma * sigma));
            excl_sum += w;
            if (excl) w = excl[q.first] * w;
            SUM += w;
            for (int d = 0; d < D; ++d) out[(size_t)d * n + i] += feat[(size_t)d * nv_orig + q.first] * w;
        }
        if (excl_sum != 0.0 && excl) excl_out[i] = SUM / excl_sum;
        for (int d = 0; d < D; ++d)
            if (SUM != 0.0) out[(size_t)d * n + i] /= SUM;
    }
    delete t;
    return err;
}

int orc_bary_resample(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                      int n_low, const double* xyz_low, int D, const double* feat_in, double* feat_out, int nthreads) {
    orc_octree* t = build_tree(nv_in, xyz_in, nt_in, tri_in);
    std::vector<int> idx(3 * (size_t)n_low), ne(n_low);
    std::vector<double> w(3 * (size_t)n_low);
    int e = orc_bary_weights(t, n_low, xyz_low, idx.data(), w.data(), ne.data(), nthreads);
    delete t;
    if (e) return e;
    for (int d = 0; d < D; ++d)
        #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
        for (int k = 0; k < n_low; ++k) {
            double val = 0.0;
            for (int j = 0; j < ne[k]; ++j) val += feat_in[(size_t)d * nv_in + idx[3 * k + j]] * w[3 * k + j];
            feat_out[(size_t)d * n_low + k] = val;
        }
    return 0;
}

static int blend_coords(int n, const double* q_xyz, int nv, const double* mesh_xyz, int nt, const int* tri,
                        const double* payload_xyz, double* out_xyz, bool reproject, int nthreads) {
    orc_octree* t = build_tree(nv, mesh_xyz, nt, tri);
    std::vector<int> idx(3 * (size_t)n), ne(n);
    std::vector<double> w(3 * (size_t)n);
    int e = orc_bary_weights(t, n, q_xyz, idx.data(), w.data(), ne.data(), nthreads);
    delete t;
    if (e) return e;
    for (int i = 0; i < n; ++i) {
        P3 np{0, 0, 0};
        for (int j = 0; j < ne[i]; ++j) { // Point*double then += (point.cpp:198,224)
            const double* c = payload_xyz + 3 * (size_t)idx[3 * i + j];
            P3 s = mul(P3{c[0], c[1], c[2]}, w[3 * i + j]);
            np.X += s.X; np.Y += s.Y; np.Z += s.Z;
        }
        if (reproject) { normalize(np); np.X *= 100; np.Y *= 100; np.Z *= 100; } // resampler.cpp:324-325
        out_xyz[3 * i] = np.X; out_xyz[3 * i + 1] = np.Y; out_xyz[3 * i + 2] = np.Z;
    }
    return 0;
}

int orc_sphere_project_warp(int n, const double* sphere_xyz, int nv, const double* from_xyz, int nt, const int* tri,
                            const double* to_xyz, double* out_xyz, int nthreads) {
    return blend_coords(n, sphere_xyz, nv, from_xyz, nt, tri, to_xyz, out_xyz, true, nthreads);
}

int orc_surface_resample(int n, const double* low_xyz, int nv, const double* sph_xyz, int nt, const int* tri,
                         const double* anat_xyz, double* out_xyz, int nthreads) {
    return blend_coords(n, low_xyz, nv, sph_xyz, nt, tri, anat_xyz, out_xyz, false, nthreads);
}

int orc_nn_resample(int n, const double* low_xyz, int nv, const double* xyz, int nt, const int* tri,
                    int D, const double* feat_in, double* feat_out, int nthreads) {
    orc_octree* t = build_tree(nv, xyz, nt, tri);
    std::vector<int> tr(n), vx(n), st(n);
    orc_octree_query(t, n, low_xyz, tr.data(), vx.data(), st.data(), nullptr, nthreads);
    delete t;
    for (int i = 0; i < n; ++i) if (st[i]) return st[i];
    for (int d = 0; d < D; ++d)
        for (int i = 0; i < n; ++i) feat_out[(size_t)d * n + i] = feat_in[(size_t)d * nv + vx[i]];
    return 0;
}

// point.cpp:97-152, with the 3x3 algebra spelled out in the same evaluation order as
// R = I + u*sin(theta) + (1-cos(theta))*(u*u) (left-to-right sums, s = 0 + a + b + c products).
int orc_rotation_matrix(const double* ci_, const double* index_, double* R) {
    P3 ci{ci_[0], ci_[1], ci_[2]}, index{index_[0], index_[1], index_[2]};
    normalize(ci); normalize(index);
    const double c = dot(ci, index);
    const double theta = std::acos(c);
    if (theta > M_PI) return 1;
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    P3 cr = cross(ci, index);
    normalize(cr);
    if (std::fabs(1 - c) < EPS) { std::memcpy(R, I, sizeof(I)); return 0; }
    if (norm(cr) < EPS) { for (int i = 0; i < 9; ++i) R[i] = -I[i]; return 0; }
    const double u[9] = {0, -cr.Z, cr.Y, cr.Z, 0, -cr.X, -cr.Y, cr.X, 0};
    if (std::fabs(-1 - c) < EPS) {
        const double op[9] = {cr.X * cr.X, cr.X * cr.Y, cr.X * cr.Z, cr.Y * cr.X, cr.Y * cr.Y, cr.Y * cr.Z,
                              cr.Z * cr.X, cr.Z * cr.Y, cr.Z * cr.Z};
        for (int i = 0; i < 9; ++i) R[i] = 2 * op[i] - I[i];
        return 0;
    }
    const double s = std::sin(theta), omc = 1 - std::cos(theta);
    double uu[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += u[3 * i + k] * u[3 * k + j];
            uu[3 * i + j] = a;
        }
    for (int i = 0; i < 9; ++i) R[i] = (I[i] + u[i] * s) + omc * uu[i];
    return 0;
}

// similarities.cpp:129-158
double orc_corr(int n, const double* A, const double* B, const double* w) {
    double prod = 0.0, varA = 0.0, varB = 0.0, meanA = 0.0, meanB = 0.0, sum = 0.0;
    for (int i = 0; i < n; ++i) sum += w[i];
    for (int i = 0; i < n; ++i) { meanA += w[i] * A[i]; meanB += w[i] * B[i]; }
    if (sum > 0.0) { meanA /= sum; meanB /= sum; }
    for (int s = 0; s < n; ++s) {
        prod += w[s] * (A[s] - meanA) * (B[s] - meanB);
        varA += w[s] * (A[s] - meanA) * (A[s] - meanA);
        varB += w[s] * (B[s] - meanB) * (B[s] - meanB);
    }
    if (sum > 0.0) { prod /= sum; varA /= sum; varB /= sum; }
    if (varA == 0.0 || varB == 0.0) return 0.0;
    return prod / (std::sqrt(varA) * std::sqrt(varB));
}

// similarities.cpp:179-188
double orc_ssd(int n, const double* A, const double* B, const double* w) {
    double prod = 0.0;
    for (int i = 0; i < n; ++i) prod += w[i] * (A[i] - B[i]) * (A[i] - B[i]);
    return std::sqrt(prod) / n;
}

// similarities.cpp:201-253: DICE / genDICE on the samples thresholded at the order statistic floor(percentile * n) of each vector
static double g_percentile = 0.75;   // sparsesimkernel::percentile (similarities.h:68)
void orc_set_percentile(double p) { g_percentile = p; }
double orc_dice(int n, const double* A, const double* B, int general) {
    const int idx = (int)std::floor(g_percentile * n);
    if (idx >= n) return std::numeric_limits<double>::quiet_NaN();   // the reference indexes past the end here
    std::vector<double> As(A, A + n), Bs(B, B + n);
    std::sort(As.begin(), As.end());
    std::sort(Bs.begin(), Bs.end());
    int size_A = n, size_B = n, common = 0;
    for (int i = 0; i < n; ++i) {
        int keep = 1;
        if (A[i] < As[idx]) { size_A--; keep = 0; }
        if (B[i] < Bs[idx]) { size_B--; keep = 0; }
        common += keep;
    }
    if (!general) return 1.0 - ((2.0 * common) / (size_A + size_B));
    return 1.0 - (2.0 * (((common / std::pow(size_B, 2))) / ((size_A + size_B) / std::pow(size_B, 2))));
}

// similarities.h:48-58 (simmeasure 1 = SSD, 2 = correlation, 4 = DICE, 5 = genDICE)
double orc_sim_for_min(int simmeasure, int n, const double* A, const double* B, const double* w) {
    if (simmeasure == 1) return orc_ssd(n, A, B, w);
    if (simmeasure == 2) return 1 - (1 + orc_corr(n, A, B, w)) * 0.5;
    if (simmeasure == 4) return orc_dice(n, A, B, 0);
    if (simmeasure == 5) return orc_dice(n, A, B, 1);
    return std::numeric_limits<double>::quiet_NaN();
}

int orc_patch_membership(int ncp, const double* cp, int nsrc, const double* src,
                         const double* maxsep, double range, int* rowptr, int* members, int cap, int nthreads) {
    std::vector<std::vector<int>> lists(ncp);
    #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int k = 0; k < ncp; ++k) {
        P3 c{cp[3 * k], cp[3 * k + 1], cp[3 * k + 2]};
        for (int i = 0; i < nsrc; ++i) { // DiscreteCostFunction.cpp:102-107
            P3 s{src[3 * i], src[3 * i + 1], src[3 * i + 2]};
            if ((2 * RAD * std::asin(norm(sub(c, s)) / (2 * RAD))) < range * maxsep[k]) lists[k].push_back(i);
        }
    }
    int pos = 0;
    for (int k = 0; k < ncp; ++k) {
        rowptr[k] = pos;
        for (int i : lists[k]) { if (pos < cap) members[pos] = i; ++pos; }
    }
    rowptr[ncp] = pos;
    return pos;
}

int orc_unary_costs(int kind, int simmeasure, const orc_octree* T,
                    int ncp, const double* cp_xyz, const double* rot, int L, const double* labels,
                    int nsrc, const double* src_xyz, const int* prow, const int* pmem,
                    int D, const double* src_feat, const double* ref_feat,
                    int cfw_rows, const double* cfw, const double* absw,
                    double* out, int* tri_out, int nthreads) {
    const int nvt = T->nv;
    const int total = prow[ncp];
    int err = 0;
    for (int l = 0; l < L; ++l) { // DiscreteCostFunction.cpp:236-243
        #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
        for (int k = 0; k < ncp; ++k) {
            const P3 lab{labels[3 * l], labels[3 * l + 1], labels[3 * l + 2]};
            const P3 dest = matvec(rot + 9 * (size_t)k, lab);
            double R[9];
            const double ci[3] = {cp_xyz[3 * k], cp_xyz[3 * k + 1], cp_xyz[3 * k + 2]};
            const double de[3] = {dest.X, dest.Y, dest.Z};
            orc_rotation_matrix(ci, de, R);
            const int P = prow[k + 1] - prow[k];
            std::vector<double> tgt((size_t)P * D), wts(3 * (size_t)P);
            std::vector<int> ids(3 * (size_t)P);
            bool bad = false;
            for (int i = 0; i < P; ++i) { // get_target_data, cpp:353-376 / 410-442 / 652-678
                const int sv = pmem[prow[k] + i];
                const P3 tmp = matvec(R, P3{src_xyz[3 * sv], src_xyz[3 * sv + 1], src_xyz[3 * sv + 2]});
                int st;
                const int t = closest_triangle(T, tmp, &st, nullptr);
                if (tri_out) tri_out[(size_t)l * total + prow[k] + i] = t;
                if (t < 0) { bad = true; break; }
                double w[3];
                bary_interp_weights(T->tv(t, 0), T->tv(t, 1), T->tv(t, 2), tmp, w);
                for (int d = 0; d < D; ++d) { // triangle.cpp:156: Aa*va1 + Ab*va2 + Ac*va3
                    const double* rf = ref_feat + (size_t)d * nvt;
                    tgt[(size_t)i * D + d] = w[0] * rf[T->tri[3 * t]] + w[1] * rf[T->tri[3 * t + 1]] + w[2] * rf[T->tri[3 * t + 2]];
                }
            }
            if (bad) {
                #pragma omp critical
                err = 2;
                out[(size_t)l * ncp + k] = std::numeric_limits<double>::quiet_NaN();
                continue;
            }
            double cost = 0.0;
            std::vector<double> a, b, w;
            if (kind == 0) { // cpp:378-383
                a.resize(P); b.resize(P); w.resize(P);
                for (int i = 0; i < P; ++i) {
                    const int sv = pmem[prow[k] + i];
                    a[i] = src_feat[sv]; b[i] = tgt[(size_t)i * D];
                    w[i] = cfw_rows >= 1 ? cfw[sv] : 1.0;
                }
                cost = orc_sim_for_min(simmeasure, P, a.data(), b.data(), w.data());
            } else if (kind == 1) { // cpp:444-458
                a.resize(D); b.resize(D); w.resize(D);
                for (int i = 0; i < P; ++i) {
                    const int sv = pmem[prow[k] + i];
                    for (int d = 0; d < D; ++d) {
                        a[d] = src_feat[(size_t)d * nsrc + sv]; b[d] = tgt[(size_t)i * D + d];
                        w[d] = cfw_rows >= d + 1 ? cfw[(size_t)d * nsrc + sv] : 1.0;
                    }
                    cost += orc_sim_for_min(simmeasure, D, a.data(), b.data(), w.data());
                }
                if (P > 0) cost /= P;
            } else { // cpp:681-692
                a.resize(P); b.resize(P); w.resize(P);
                for (int i = 0; i < P; ++i) w[i] = cfw_rows >= 1 ? cfw[pmem[prow[k] + i]] : 1.0;
                for (int d = 0; d < D; ++d) {
                    for (int i = 0; i < P; ++i) {
                        a[i] = src_feat[(size_t)d * nsrc + pmem[prow[k] + i]];
                        b[i] = tgt[(size_t)i * D + d];
                    }
                    cost += orc_sim_for_min(simmeasure, P, a.data(), b.data(), w.data());
                }
                cost /= D;
            }
            out[(size_t)l * ncp + k] = absw[k] * cost;
        }
    }
    return err;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// Triplet costs: regulariser (strain) + HO likelihood. PARITY UNPINNED: the reference's meshreg
// library cannot be compiled here (FSL NEWMAT/armawrap `.i()`, `Determinant()` and the FSL
// BFMatrix/OptionParser stack are absent), so this is a restatement of
// DiscreteCostFunction.cpp:135-188, 468-618 and reg_tools.cpp:267-313, 551-743 with the small
// dense-matrix algebra written out (2x2 inverse by adjugate / determinant, 3x3 determinant by
// first-row cofactors, products summed left to right).
// ------------------------------------------------------------------------------------------------
namespace {

inline P3 tri_normal(const P3& v0, const P3& v1, const P3& v2) { // triangle.cpp:42-47
    P3 r = cross(sub(v2, v0), sub(v1, v0));
    normalize(r);
    return r;
}

struct Tangs { P3 e1, e2; };
Tangs calculate_tri(const P3& a) { // reg_tools.cpp:267-313
    Tangs T;
    P3 b{1.0, 0.0, 0.0};
    P3 c = cross(a, b);
    double len = c.X * c.X + c.Y * c.Y + c.Z * c.Z;
    if (len == 0.0) {
        b = P3{0.0, 1.0, 0.0};
        c = cross(a, b);
        len = c.X * c.X + c.Y * c.Y + c.Z * c.Z;
    }
    len = std::sqrt(len);
    if (len == 0.0) len = 1;
    T.e1 = P3{c.X / len, c.Y / len, c.Z / len};
    b = cross(a, c);
    len = std::sqrt(b.X * b.X + b.Y * b.Y + b.Z * b.Z);
    if (len == 0) len = 1;
    T.e2 = P3{b.X / len, b.Y / len, b.Z / len};
    return T;
}

inline double det3(const double* m) { // row-major 3x3, first-row cofactors
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// reg_tools.cpp:551-646 without the optional principal-strain output
double triangle_strain(const double* AA, const double* BB, double MU, double KAPPA, double k_exp) { // AA, BB row-major 3x3, columns 1,2 used
    const double c0 = AA[3] - AA[0], c1 = AA[4] - AA[1], c4 = AA[6] - AA[0], c5 = AA[7] - AA[1];
    const double c0c = BB[3] - BB[0], c1c = BB[4] - BB[1], c4c = BB[6] - BB[0], c5c = BB[7] - BB[1];
    // Edges = [c0 c4; c1 c5], edges = [c0c c4c; c1c c5c]; F = edges * Edges^-1
    const double det = c0 * c5 - c4 * c1;
    const double i11 = c5 / det, i12 = -c4 / det, i21 = -c1 / det, i22 = c0 / det;
    const double F11 = c0c * i11 + c4c * i21, F12 = c0c * i12 + c4c * i22;
    const double F21 = c1c * i11 + c5c * i21, F22 = c1c * i12 + c5c * i22;
    // F3D = [F 0; 0 0 1]; F3D_2 = F3D^T F3D
    const double G11 = F11 * F11 + F21 * F21 + 0.0 * 0.0, G12 = F11 * F12 + F21 * F22 + 0.0 * 0.0;
    const double G21 = F12 * F11 + F22 * F21 + 0.0 * 0.0, G22 = F12 * F12 + F22 * F22 + 0.0 * 0.0;
    const double G33 = 0.0 * 0.0 + 0.0 * 0.0 + 1.0 * 1.0;
    const double I1 = G11 + G22 + G33;
    const double g[9] = {G11, G12, 0.0, G21, G22, 0.0, 0.0, 0.0, G33};
    const double I3 = det3(g);
    const double J = std::sqrt(I3);
    const double I1st_new = (I1 - 1.0) / J;
    double R;
    if (I1st_new <= 2) R = 1.0;
    else R = 0.5 * (I1st_new + std::sqrt(I1st_new * I1st_new - 4));
    const double Rshared = std::pow(R, k_exp), Jshared = std::pow(J, k_exp);
    return 0.5 * (MU * (Rshared + 1.0 / Rshared - 2) + KAPPA * (Jshared + 1.0 / Jshared - 2));
}

// reg_tools.cpp:698-743 (Triangle overload)
double triangular_strain(const P3* O, const P3* Fv, double mu, double kappa, double k_exp) {
    const P3 NO = tri_normal(O[0], O[1], O[2]), NF = tri_normal(Fv[0], Fv[1], Fv[2]);
    const Tangs T = calculate_tri(NO), T2 = calculate_tri(NF);
    double TR[9] = {T.e1.X, T.e2.X, NO.X, T.e1.Y, T.e2.Y, NO.Y, T.e1.Z, T.e2.Z, NO.Z};     // point.cpp:77-95
    double TR2[9] = {T2.e1.X, T2.e2.X, NF.X, T2.e1.Y, T2.e2.Y, NF.Y, T2.e1.Z, T2.e2.Z, NF.Z};
    if (det3(TR) < 0) { std::swap(TR[0], TR[1]); std::swap(TR[3], TR[4]); std::swap(TR[6], TR[7]); }
    if (det3(TR) < 0) { std::swap(TR2[0], TR2[1]); std::swap(TR2[3], TR2[4]); std::swap(TR2[6], TR2[7]); } // sic: tests TRANS again (reg_tools.cpp:722)
    double A[9], B[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[3 * i + j] = O[i].X * TR[j] + O[i].Y * TR[3 + j] + O[i].Z * TR[6 + j];
            B[3 * i + j] = Fv[i].X * TR2[j] + Fv[i].Y * TR2[3 + j] + Fv[i].Z * TR2[6 + j];
        }
    return triangle_strain(A, B, mu, kappa, k_exp);
}

inline P3 bary_blend(const P3& v1, const P3& v2, const P3& v3, const P3& vref, const P3& a1, const P3& a2, const P3& a3) { // triangle.cpp:159-170
    double w[3];
    bary_interp_weights(v1, v2, v3, vref, w);
    const P3 x = mul(a1, w[0]), y = mul(a2, w[1]), z = mul(a3, w[2]);
    return P3{x.X + y.X + z.X, x.Y + y.Y + z.Y, x.Z + y.Z + z.Z};
}

} // namespace

extern "C" {

// HO*::get_source_data (DiscreteCostFunction.cpp:468-485, 541-563): sources grouped by their nearest CP-grid triangle
int orc_ho_patches(int ncp, const double* cp_xyz, int ntri, const int* cp_tri, int nsrc, const double* src_xyz,
                   int* rowptr, int* members, int cap) {
    orc_octree* T = build_tree(ncp, cp_xyz, ntri, cp_tri);
    std::vector<std::vector<int>> lists(ntri);
    int bad = 0;
    for (int i = 0; i < nsrc; ++i) {
        int st;
        const int t = closest_triangle(T, P3{src_xyz[3 * i], src_xyz[3 * i + 1], src_xyz[3 * i + 2]}, &st, nullptr);
        if (t < 0) { bad = 1; continue; }
        lists[t].push_back(i);
    }
    delete T;
    if (bad) return -1;
    int pos = 0;
    for (int k = 0; k < ntri; ++k) {
        rowptr[k] = pos;
        for (int i : lists[k]) { if (pos < cap) members[pos] = i; ++pos; }
    }
    rowptr[ntri] = pos;
    return pos;
}

// computeTripletCost (DiscreteCostFunction.cpp:135-188) for n requests (triplet, la, lb, lc).
// kind: 0..2 = non-HO classes (likelihood 0), 3 = HOUnivariate (cpp:487-531), 4 = HOMultivariate (cpp:565-618).
// triplets [T][3] node ids; patches (HO kinds) CSR over triplets. rmode 2/3 only (spherical strain). Returns 0 or 2 (a query failed).
int orc_triplet_costs(int kind, int simmeasure, const orc_octree* T, int ncp, const double* cp_xyz, const double* orig_cp_xyz,
                      const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                      int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                      int nsrc, const double* src_xyz, const int* prow, const int* pmem, int D, const double* src_feat,
                      const double* ref_feat, int cfw_rows, const double* cfw, const double* absw,
                      double lambda, double mu, double kappa, double k_exp, double rexp, double* out, int nthreads) {
    return orc_triplet_costs_anat(kind, simmeasure, T, ncp, cp_xyz, orig_cp_xyz, rot, L, labels, ntrip, triplets, n, req_triplet, req_la, req_lb, req_lc,
                                  nsrc, src_xyz, prow, pmem, D, src_feat, ref_feat, cfw_rows, cfw, absw, lambda, mu, kappa, k_exp, rexp, 3, nullptr, out,
                                  nthreads);
}

// the same with the regulariser of regoption 4/5 when rmode >= 4 (anat != NULL): mean strain energy of the anatomical faces of the
// triplet, each deformed by deform_anatomy (DiscreteCostFunction.cpp:169-181, 245-301)
int orc_triplet_costs_anat(int kind, int simmeasure, const orc_octree* T, int ncp, const double* cp_xyz, const double* orig_cp_xyz,
                           const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                           int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                           int nsrc, const double* src_xyz, const int* prow, const int* pmem, int D, const double* src_feat,
                           const double* ref_feat, int cfw_rows, const double* cfw, const double* absw,
                           double lambda, double mu, double kappa, double k_exp, double rexp, int rmode, const orc_anat* anat,
                           double* out, int nthreads) {
    int err = 0;
    const int nvt = T ? T->nv : 0;
    orc_octree* AT = nullptr;   // anattree = Octree(_TARGEThi), DiscreteCostFunction.h:160-163
    if (rmode >= 4) {
        if (!anat) return 3;
        AT = orc_octree_build(anat->n_hv, anat->thi_xyz, anat->n_ht, anat->thi_tri);
    }
    #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int r = 0; r < n; ++r) {
        const int t = req_triplet[r];
        const int ids[3] = {triplets[3 * t], triplets[3 * t + 1], triplets[3 * t + 2]};
        const int lab[3] = {req_la[r], req_lb[r], req_lc[r]};
        P3 def[3], cur[3], org[3];
        for (int k = 0; k < 3; ++k) {
            def[k] = matvec(rot + 9 * (size_t)ids[k], P3{labels[3 * lab[k]], labels[3 * lab[k] + 1], labels[3 * lab[k] + 2]});
            cur[k] = P3{cp_xyz[3 * ids[k]], cp_xyz[3 * ids[k] + 1], cp_xyz[3 * ids[k] + 2]};
            org[k] = P3{orig_cp_xyz[3 * ids[k]], orig_cp_xyz[3 * ids[k] + 1], orig_cp_xyz[3 * ids[k] + 2]};
        }
        if (dot(tri_normal(def[0], def[1], def[2]), tri_normal(cur[0], cur[1], cur[2])) < 0.0) { out[r] = 1e7 * lambda; continue; } // FOLDING
        double likelihood = 0.0;
        if (kind >= 3) {
            const int P = prow[t + 1] - prow[t];
            std::vector<double> tgt((size_t)P * D);
            bool bad = false;
            for (int i = 0; i < P && !bad; ++i) {
                const int sv = pmem[prow[t] + i];
                const P3 SP = project_point(P3{src_xyz[3 * sv], src_xyz[3 * sv + 1], src_xyz[3 * sv + 2]}, cur[0], cur[1], cur[2]);
                P3 tmp = bary_blend(cur[0], cur[1], cur[2], SP, def[0], def[1], def[2]);
                normalize(tmp);
                tmp = mul(tmp, RAD);
                int st;
                const int tt = closest_triangle(T, tmp, &st, nullptr);
                if (tt < 0) { bad = true; break; }
                double w[3];
                bary_interp_weights(T->tv(tt, 0), T->tv(tt, 1), T->tv(tt, 2), tmp, w);
                for (int d = 0; d < D; ++d) {
                    const double* rf = ref_feat + (size_t)d * nvt;
                    tgt[(size_t)i * D + d] = w[0] * rf[T->tri[3 * tt]] + w[1] * rf[T->tri[3 * tt + 1]] + w[2] * rf[T->tri[3 * tt + 2]];
                }
            }
            if (bad) {
                #pragma omp critical
                err = 2;
                out[r] = std::numeric_limits<double>::quiet_NaN();
                continue;
            }
            double cost = 0.0;
            std::vector<double> a, b, w;
            if (kind == 3) { // cpp:522-531
                a.resize(P); b.resize(P); w.resize(P);
                for (int i = 0; i < P; ++i) {
                    const int sv = pmem[prow[t] + i];
                    a[i] = src_feat[sv]; b[i] = tgt[(size_t)i * D];
                    w[i] = cfw_rows >= 1 ? cfw[sv] : 1.0;
                }
                cost = orc_sim_for_min(simmeasure, P, a.data(), b.data(), w.data());
            } else { // cpp:601-618
                a.resize(D); b.resize(D); w.resize(D);
                for (int i = 0; i < P; ++i) {
                    const int sv = pmem[prow[t] + i];
                    for (int d = 0; d < D; ++d) {
                        a[d] = src_feat[(size_t)d * nsrc + sv]; b[d] = tgt[(size_t)i * D + d];
                        w[d] = cfw_rows >= d + 1 ? cfw[(size_t)d * nsrc + sv] : 1.0;
                    }
                    cost += orc_sim_for_min(simmeasure, D, a.data(), b.data(), w.data());
                }
                if (P > 0) cost /= P;
            }
            likelihood = (absw[ids[0]] + absw[ids[1]] + absw[ids[2]]) / 3.0 * cost;
        }
        double W;
        if (rmode <= 3) {
            W = triangular_strain(org, def, mu, kappa, k_exp); // rmode 2/3, cpp:158-166
        } else {   // rmode 4/5, cpp:169-181
            const int f0 = anat->face_ptr[t], f1 = anat->face_ptr[t + 1];
            std::map<int, P3> transformed;   // `moved2` / `transformed_points`: every vertex of the neighbourhood is deformed once per request
            W = 0.0;
            for (int f = f0; f < f1; ++f) {
                const int face = anat->face_ids[f];
                P3 O[3], Fv[3];
                for (int i = 0; i < 3; ++i) {   // deform_anatomy, cpp:245-301
                    const int tindex = anat->asource_tri[3 * face + i];
                    O[i] = P3{anat->asource_xyz[3 * tindex], anat->asource_xyz[3 * tindex + 1], anat->asource_xyz[3 * tindex + 2]};
                    auto it = transformed.find(tindex);
                    if (it == transformed.end()) {
                        P3 np{0, 0, 0};
                        // `vertex[it.first]` on a std::map holding the three displaced control points: a key that is not one of the
                        // triplet's nodes default-constructs a zero Point (mesh_registration.cpp:307-327 overwrites the weights of a
                        // vertex shared by several control triangles with those of the LAST one)
                        for (int e = anat->bary_ptr[tindex]; e < anat->bary_ptr[tindex + 1]; ++e) {
                            P3 v{0, 0, 0};
                            for (int k = 0; k < 3; ++k) if (ids[k] == anat->bary_key[e]) v = def[k];
                            const P3 c = mul(v, anat->bary_w[e]);
                            np.X += c.X; np.Y += c.Y; np.Z += c.Z;
                        }
                        int st;
                        const int tt = closest_triangle(AT, np, &st, nullptr);
                        P3 res{0, 0, 0};
                        if (tt < 0) {
                            // the reference catches the exception and continues with an all-zero Triangle (ids 0, 1, 2): its weights are 0/0
                            res.X = res.Y = res.Z = std::numeric_limits<double>::quiet_NaN();
                        } else {
                            int idx[3]; double w[3];
                            const int ne = bary_weights_of(AT, np, tt, idx, w);
                            for (int j = 0; j < ne; ++j) {
                                const P3 c = mul(P3{anat->atarget_xyz[3 * idx[j]], anat->atarget_xyz[3 * idx[j] + 1], anat->atarget_xyz[3 * idx[j] + 2]}, w[j]);
                                res.X += c.X; res.Y += c.Y; res.Z += c.Z;
                            }
                        }
                        it = transformed.emplace(tindex, res).first;
                    }
                    Fv[i] = it->second;
                }
                W += triangular_strain(O, Fv, mu, kappa, k_exp);
            }
            W = W / (double)(f1 - f0);
        }
        out[r] = likelihood + lambda * std::pow(W, rexp);
    }
    if (AT) orc_octree_free(AT);
    return err;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// Groupwise (gMSM): DiscreteGroupModel::get_patch_data (DiscreteGroupModel.cpp:88-121) and
// DiscreteGroupCostFunction::computePairwiseCost (DiscreteGroupCostFunction.cpp:54-97).
// PARITY UNPINNED for the same reason as the triplet costs (meshreg library not buildable here);
// the resampling inside is the pinned adaptive-barycentric path.
// ------------------------------------------------------------------------------------------------
extern "C" {

// fields[s][l][d][p] (channel-major per (s,l), like Mesh::pvalues): subject s's data mesh rigidly "rotated" by label l
// (estimate_rotation_matrix(centre, vertex) * label per vertex; label 0 = no rotation) and metric_resampled onto the template.
int orc_group_fields(int S, int nv, const double* data_xyz /*[S][nv][3]*/, int nt, const int* tri, int D, const double* feat /*[S][D][nv]*/,
                     int L, const double* labels, const double* centre, int n_tpl, const double* tpl_xyz, int nt_tpl, const int* tpl_tri,
                     double* fields /*[S][L][D][n_tpl]*/, int nthreads) {
    int err = 0;
    #pragma omp parallel for collapse(2) num_threads(nthreads > 0 ? nthreads : 1)
    for (int s = 0; s < S; ++s)
        for (int l = 0; l < L; ++l) {
            std::vector<double> xyz(data_xyz + (size_t)s * nv * 3, data_xyz + (size_t)(s + 1) * nv * 3);
            if (l > 0)
                for (int p = 0; p < nv; ++p) {
                    double R[9];
                    orc_rotation_matrix(centre, &xyz[3 * (size_t)p], R);
                    const P3 q = matvec(R, P3{labels[3 * l], labels[3 * l + 1], labels[3 * l + 2]});
                    xyz[3 * (size_t)p] = q.X; xyz[3 * (size_t)p + 1] = q.Y; xyz[3 * (size_t)p + 2] = q.Z;
                }
            // metric_resample of the moved copy: source vertex areas are still those of the un-moved data mesh (see adaptive_maps)
            std::vector<std::map<int, double>> adapt;
            if (adaptive_maps(nv, xyz.data(), nt, tri, n_tpl, tpl_xyz, nt_tpl, tpl_tri, adapt, data_xyz + (size_t)s * nv * 3)) {
                #pragma omp critical
                err = 1;
                continue;
            }
            const double* fin = feat + (size_t)s * D * nv;
            double* fout = fields + ((size_t)s * L + l) * D * n_tpl;
            for (int d = 0; d < D; ++d)
                for (int k = 0; k < n_tpl; ++k) {
                    double val = 0.0;
                    for (auto& it : adapt[k]) val += fin[(size_t)d * nv + it.first] * it.second;
                    fout[(size_t)d * n_tpl + k] = val;
                }
        }
    return err;
}

// pair cost for n requests (pair, la, lb). pairs [P][2] global node ids (subject * ncp + vertex); rot [S*ncp][9];
// spacings [S*ncp]; patch(s,v,l) = template vertices p with 2R asin(|rot*label_l - tpl_p| / 2R) < range * spacing.
int orc_group_pair_costs(int simmeasure, int S, int ncp, int L, int D, int n_tpl, const double* tpl_xyz, const double* fields,
                         const double* rot, const double* labels, const double* spacings, double range, const int* pairs,
                         int n, const int* req_pair, const int* req_la, const int* req_lb, double* out, int nthreads) {
    return orc_group_pair_costs_masked(simmeasure, S, ncp, L, D, n_tpl, tpl_xyz, fields, rot, labels, spacings, range, pairs, n, req_pair, req_la,
                                       req_lb, nullptr, out, nthreads);
}

// with a cost mask (DiscreteGroupCostFunction.cpp:77): weight of a common template vertex = |mask[vertex]|; mask == NULL: unit weights
int orc_group_pair_costs_masked(int simmeasure, int S, int ncp, int L, int D, int n_tpl, const double* tpl_xyz, const double* fields,
                                const double* rot, const double* labels, const double* spacings, double range, const int* pairs,
                                int n, const int* req_pair, const int* req_la, const int* req_lb, const double* mask, double* out, int nthreads) {
    #pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int r = 0; r < n; ++r) {
        const int nodes[2] = {pairs[2 * req_pair[r]], pairs[2 * req_pair[r] + 1]};
        const int lab[2] = {req_la[r], req_lb[r]};
        std::vector<int> patch[2];
        for (int k = 0; k < 2; ++k) {
            const P3 rcp = matvec(rot + 9 * (size_t)nodes[k], P3{labels[3 * lab[k]], labels[3 * lab[k] + 1], labels[3 * lab[k] + 2]});
            for (int p = 0; p < n_tpl; ++p) {
                const P3 t{tpl_xyz[3 * p], tpl_xyz[3 * p + 1], tpl_xyz[3 * p + 2]};
                if ((2 * RAD * std::asin(norm(sub(rcp, t)) / (2 * RAD))) < range * spacings[nodes[k]]) patch[k].push_back(p);
            }
        }
        std::vector<int> common;
        std::set_intersection(patch[0].begin(), patch[0].end(), patch[1].begin(), patch[1].end(), std::back_inserter(common));
        const int n_c = (int)common.size();
        if (n_c == 0) { out[r] = std::numeric_limits<double>::quiet_NaN(); continue; }   // the reference dereferences an empty vector here
        const int sa = nodes[0] / ncp, sb = nodes[1] / ncp;
        std::vector<double> a(n_c), b(n_c), w(n_c, 1.0);
        if (mask) for (int i = 0; i < n_c; ++i) w[i] = std::fabs(mask[common[i]]);
        double cost = 0.0;
        for (int d = 0; d < D; ++d) {
            const double* fa = fields + (((size_t)sa * L + lab[0]) * D + d) * n_tpl;
            const double* fb = fields + (((size_t)sb * L + lab[1]) * D + d) * n_tpl;
            for (int i = 0; i < n_c; ++i) { a[i] = fa[common[i]]; b[i] = fb[common[i]]; }
            cost += orc_sim_for_min(simmeasure, n_c, a.data(), b.data(), w.data());
        }
        out[r] = cost / D;
    }
    return 0;
}

} // extern "C"

// ======================================================================================================================
// RIGID / AFFINE level: restatement of msm-newmeshreg/src/rigid_costfunction.cpp:32-236 with
// Neighbourhood::update (reg_tools.cpp:31-58), calculate_tangs (reg_tools.cpp:205-266), the full-matrix part of
// sparsesimkernel (similarities.cpp:27-128), euler_rotate (point.cpp:154-171), Mesh::local_normal / calculate_MeanVD /
// push_triangle adjacency order (mesh.cpp:112-141, 276-294). Pinned by tests/golden/rigid.npz (outputs of the compiled reference).
// ======================================================================================================================
namespace {

struct RigidMesh {
    int nv = 0, nt = 0;
    std::vector<P3> v;
    std::vector<int> tri;
    std::vector<std::vector<int>> nbr, inc;   // vertex neighbours / incident triangles in push_triangle order (mesh.cpp:112-131)
    void build(int nv_, const double* xyz, int nt_, const int* t) {
        nv = nv_; nt = nt_;
        v.resize(nv); tri.assign(t, t + 3 * (size_t)nt);
        for (int i = 0; i < nv; ++i) v[i] = P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        nbr.assign(nv, {}); inc.assign(nv, {});
        auto add = [&](int a, int b) { if (std::find(nbr[a].begin(), nbr[a].end(), b) == nbr[a].end()) nbr[a].push_back(b); };
        for (int k = 0; k < nt; ++k) {
            const int n0 = tri[3 * k], n1 = tri[3 * k + 1], n2 = tri[3 * k + 2];
            inc[n0].push_back(k); inc[n1].push_back(k); inc[n2].push_back(k);
            add(n0, n1); add(n0, n2); add(n1, n0); add(n1, n2); add(n2, n0); add(n2, n1);
        }
    }
    P3 tri_n(int t) const { return tri_normal(v[tri[3 * t]], v[tri[3 * t + 1]], v[tri[3 * t + 2]]); }
    P3 local_normal(int pt) const {   // mesh.cpp:133-141
        P3 a{0, 0, 0};
        for (int t : inc[pt]) { const P3 n = tri_n(t); a.X += n.X; a.Y += n.Y; a.Z += n.Z; }
        normalize(a);
        return a;
    }
    double mean_vd() const {          // mesh.cpp:276-294
        int k = 0;
        double kr = 0.0;
        for (int i = 0; i < nv; ++i)
            for (int j : nbr[i]) { ++k; kr += norm(sub(v[j], v[i])); }
        return kr / k;
    }
};

struct RTangs { P3 e1, e2; };
RTangs rigid_tangs(int ind, const RigidMesh& M) {   // reg_tools.cpp:205-266
    RTangs T;
    double mag;
    P3 a = M.local_normal(ind);
    if (dot(a, M.v[ind]) < 0) a = mul(a, -1);
    // `abs(a.X)` in reg_tools.cpp:213-229 is the C library's int abs(int) in that translation unit (no `using std::abs`, no <cmath>
    // overload in scope): the components of a unit normal are truncated to 0 (or +-1), so the first branch is taken unless a
    // component is exactly +-1. Restated as compiled.
    auto iabs = [](double x) { return std::abs((int)x); };
    if (iabs(a.X) >= iabs(a.Y) && iabs(a.X) >= iabs(a.Z)) {
        mag = std::sqrt(a.Z * a.Z + a.Y * a.Y);
        if (mag == 0) T.e1 = P3{0, 0, 1};
        else T.e1 = P3{0, -a.Z / mag, a.Y / mag};
    } else if (iabs(a.Y) >= iabs(a.X) && iabs(a.Y) >= iabs(a.Z)) {
        mag = std::sqrt(a.Z * a.Z + a.X * a.X);
        if (mag == 0) T.e1 = P3{0, 0, 1};
        else T.e1 = P3{-a.Z / mag, 0, a.X / mag};
    } else {
        mag = std::sqrt(a.Y * a.Y + a.X * a.X);
        if (mag == 0) T.e1 = P3{1, 0, 0};
        else T.e1 = P3{-a.Y / mag, a.X / mag, 0};
    }
    T.e2 = cross(a, T.e1);
    normalize(T.e2);
    return T;
}

struct RigidState {
    RigidMesh target, source;
    orc_octree* tree = nullptr;
    int D = 0, simmeasure = 2;
    const double* A = nullptr;   // input data  [D][nv_s]
    const double* B = nullptr;   // reference   [D][nv_t]
    std::vector<double> meanA, meanB, current_sim;
    std::vector<std::vector<int>> nbh;
    std::vector<std::map<int, double>> sim;   // sim[source column] : target row -> value (SpMat::Set / Peek)
    double min_sigma = 0.0;
    ~RigidState() { delete tree; }

    static void means(int D, int n, const double* M, std::vector<double>& out) {   // similarities.cpp:106-126
        out.assign(n, 0.0);
        if (D == 1) {
            double sum = 0.0;
            for (int i = 0; i < n; ++i) sum += M[i];
            for (int i = 0; i < n; ++i) out[i] = sum / n;
        } else
            for (int i = 0; i < n; ++i) {
                double sum = 0.0;
                for (int j = 0; j < D; ++j) sum += M[(size_t)j * n + i];
                out[i] = sum / D;
            }
    }
    double corr(int i, int j) const {   // similarities.cpp:52-85, full matrices (i: source column, j: target column; 0-based here)
        double prod = 0.0, varA = 0.0, varB = 0.0;
        const int ns = source.nv, ntg = target.nv;
        for (int r = 0; r < D; ++r) {
            const double a = A[(size_t)r * ns + i], b = B[(size_t)r * ntg + j];
            prod += (a - meanA[i]) * (b - meanB[j]);
            varA += (a - meanA[i]) * (a - meanA[i]);
            varB += (b - meanB[j]) * (b - meanB[j]);
        }
        if (varA == 0.0 || varB == 0.0) return 0.0;
        return prod / (std::sqrt(varA) * std::sqrt(varB));
    }
    double ssd(int i, int j) const {    // similarities.cpp:87-104
        double prod = 0.0;
        const int ns = source.nv, ntg = target.nv;
        for (int r = 0; r < D; ++r) {
            const double a = A[(size_t)r * ns + i], b = B[(size_t)r * ntg + j];
            prod += (a - b) * (a - b);
        }
        return std::sqrt(prod) / D;
    }
    void sim_column(int ind) {          // similarities.cpp:37-50 (a neighbour with id 0 is skipped, sic)
        for (int nb : nbh[ind])
            if (nb != 0) sim[ind][nb] = simmeasure == 1 ? -ssd(ind, nb) : corr(ind, nb);
    }
    double peek(int row, int col) const {
        auto it = sim[col].find(row);
        return it == sim[col].end() ? 0.0 : it->second;
    }

    void neighbourhoods(double ang) {   // reg_tools.cpp:31-58
        nbh.assign(source.nv, {});
        for (int index = 0; index < source.nv; ++index) {
            P3 cr = source.v[index];
            normalize(cr);
            std::vector<std::pair<double, int>> c;
            for (int n = 0; n < target.nv; ++n) {
                P3 actual = target.v[n];
                normalize(actual);
                if (dot(actual, cr) >= std::cos(ang)) c.emplace_back(norm(sub(actual, cr)), n);
            }
            std::sort(c.begin(), c.end(), [](const auto& l, const auto& r) -> bool { return l.first < r.first; });
            for (const auto& n : c) nbh[index].push_back(n.second);
        }
    }

    void initialise() {                 // rigid_costfunction.cpp:32-50
        current_sim.assign(source.nv, 0.0);
        const double MVD = source.mean_vd();
        min_sigma = MVD;
        sim.assign(source.nv, {});
        means(D, source.nv, A, meanA);
        means(D, target.nv, B, meanB);
        neighbourhoods(2 * std::asin(4 * MVD / (2 * RAD)));
        for (int i = 0; i < source.nv; ++i) sim_column(i);
    }

    bool all_neighbours(int index, std::vector<int>& N, int n, std::vector<char>& found) {   // rigid_costfunction.cpp:143-165
        bool update = false;
        for (int j : target.inc[n]) {
            const int n0 = target.tri[3 * j], n1 = target.tri[3 * j + 1], n2 = target.tri[3 * j + 2];
            if (nbh[index][0] != n0 || nbh[index][0] != n1 || nbh[index][0] != n2) update = true;
            if (!found[n0]) { N.push_back(n0); found[n0] = 1; }
            if (!found[n1]) { N.push_back(n1); found[n1] = 1; }
            if (!found[n2]) { N.push_back(n2); found[n2] = 1; }
        }
        return update;
    }

    void wls(const RTangs& tg, int index, const std::vector<int>& q) {   // rigid_costfunction.cpp:63-90
        double SUM = 0.0, JPsim = 0.0;
        P3 origin = cross(tg.e1, tg.e2);
        normalize(origin);
        origin = mul(origin, RAD);
        const P3 ys = sub(source.v[index], origin);
        const double y11 = dot(ys, tg.e1), y21 = dot(ys, tg.e2);
        for (int qp : q) {
            const P3 xs = sub(target.v[qp], origin);
            const double x11 = dot(xs, tg.e1), x21 = dot(xs, tg.e2);
            const double d1 = x11 - y11, d2 = x21 - y21;
            if ((d1 * d1 + d2 * d2) > 0) {
                const double weight = std::exp(-(d1 * d1 + d2 * d2) / (2 * min_sigma * min_sigma));
                SUM += weight;
                JPsim += peek(qp, index) * weight;
            }
        }
        if (SUM > 0) JPsim /= SUM;
        current_sim[index] = JPsim;
    }

    void evaluate(int i, const RTangs& tg) {   // rigid_costfunction.cpp:92-114
        if (nbh[i].empty()) return;
        std::vector<char> found(target.nv, 0);
        std::vector<int> q;
        int st = 0;
        const int t = closest_triangle(tree, source.v[i], &st, nullptr);
        if (t < 0) throw 1;
        bool update = false;
        if (all_neighbours(i, q, target.tri[3 * t], found)) update = true;
        if (all_neighbours(i, q, target.tri[3 * t + 1], found) || update) update = true;
        if (all_neighbours(i, q, target.tri[3 * t + 2], found) || update) update = true;
        if (update) { nbh[i] = q; sim_column(i); }
        wls(tg, i, q);
    }

    void rotate(double w1, double w2, double w3) {   // rigid_costfunction.cpp:116-128 + point.cpp:154-171
        const double R[9] = {std::cos(w2) * std::cos(w3), -std::cos(w1) * std::sin(w3) + std::sin(w1) * std::sin(w2) * std::cos(w3),
                             std::sin(w1) * std::sin(w3) + std::cos(w1) * std::sin(w2) * std::cos(w3),
                             std::cos(w2) * std::sin(w3), std::cos(w1) * std::cos(w3) + std::sin(w1) * std::sin(w2) * std::sin(w3),
                             -std::sin(w1) * std::cos(w3) + std::cos(w1) * std::sin(w2) * std::sin(w3),
                             -std::sin(w2), std::sin(w1) * std::cos(w2), std::cos(w1) * std::cos(w2)};
        for (P3& p : source.v) {      // rotation.t() * vector: row r of the product = sum over k of R(k, r) * v(k), from zero, k ascending
            const double vv[3] = {p.X, p.Y, p.Z};
            double o[3];
            for (int r = 0; r < 3; ++r) {
                double sum = 0.0;
                for (int k = 0; k < 3; ++k) sum += R[3 * k + r] * vv[k];
                o[r] = sum;
            }
            p = P3{o[0], o[1], o[2]};
        }
    }

    double cost(double dw1, double dw2, double dw3) {   // rigid_costfunction.cpp:130-141
        const std::vector<P3> keep = source.v;
        rotate(dw1, dw2, dw3);
        for (int index = 0; index < source.nv; ++index) evaluate(index, rigid_tangs(index, source));
        double SUM = 0.0;
        for (int i = 0; i < source.nv; ++i) SUM += current_sim[i];
        source.v = keep;
        return SUM;
    }

    void run(int iters, double stepsize, double spacing) {   // rigid_costfunction.cpp:167-236
        double Euler1 = 0.0, Euler2 = 0.0, Euler3 = 0.0;
        int min_iter = 0, loop = 0;
        double grad_zero = cost(Euler1, Euler2, Euler3);
        double mingrad_zero = grad_zero;
        while (spacing > 0.05) {
            double step = stepsize;
            const double per = spacing;
            for (int it = 1; it <= iters; ++it) {
                Euler1 = 0.0; Euler2 = 0.0; Euler3 = 0.0;
                P3 grad;
                grad.X = (cost(Euler1 + per, Euler2, Euler3) - grad_zero) / per;
                grad.Y = (cost(Euler1, Euler2 + per, Euler3) - grad_zero) / per;
                grad.Z = (cost(Euler1, Euler2, Euler3 + per) - grad_zero) / per;
                normalize(grad);
                Euler1 += step * grad.X; Euler2 += step * grad.Y; Euler3 += step * grad.Z;
                const std::vector<P3> tmp = source.v;
                rotate(Euler1, Euler2, Euler3);
                grad_zero = cost(Euler1, Euler2, Euler3);
                if (grad_zero > mingrad_zero) { mingrad_zero = grad_zero; min_iter = (loop * iters) + it; }
                if ((loop * iters) + it - min_iter > 0) { step *= 0.5; source.v = tmp; }
                if (step < 1e-3) break;
            }
            ++loop;
            spacing *= 0.5;
        }
    }
};

} // namespace

int orc_rigid(int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri, int nv_s, const double* src_xyz, int nt_s, const int* src_tri,
              int D, const double* src_feat, const double* ref_feat, int simmeasure, int iters, double stepsize, double gradsampling,
              double* out_xyz, double* out_cost0, int* nbh_rowptr, int* nbh_members, int cap) {
    try {
        RigidState S;
        S.target.build(nv_t, tgt_xyz, nt_t, tgt_tri);
        S.source.build(nv_s, src_xyz, nt_s, src_tri);
        S.tree = build_tree(nv_t, tgt_xyz, nt_t, tgt_tri);
        S.D = D; S.simmeasure = simmeasure; S.A = src_feat; S.B = ref_feat;
        S.initialise();
        int pos = 0;
        for (int i = 0; i < nv_s; ++i) {
            nbh_rowptr[i] = pos;
            for (int n : S.nbh[i]) { if (pos < cap) nbh_members[pos] = n; ++pos; }
        }
        nbh_rowptr[nv_s] = pos;
        if (out_cost0) *out_cost0 = S.cost(0.0, 0.0, 0.0);
        S.run(iters, stepsize, gradsampling);
        for (int i = 0; i < nv_s; ++i) { out_xyz[3 * i] = S.source.v[i].X; out_xyz[3 * i + 1] = S.source.v[i].Y; out_xyz[3 * i + 2] = S.source.v[i].Z; }
        return pos;
    } catch (...) { return -1; }
}


// variance_normalise, reg_tools.cpp:804-844: the compacted per-channel vectors, Welford's recurrence in vertex order (cpp:820-825),
// var / (size - 1) in size_t (cpp:827), then (x - mean) and, when var > 0, / sqrt(var) (cpp:829-833), written back to the kept vertices
void orc_variance_normalise(int D, int n, double* data, const double* excl) {
    for (int k = 0; k < D; ++k) {
        std::vector<double> v;
        for (int i = 0; i < n; ++i)
            if (!excl || excl[i] > 0.0) v.push_back(data[(size_t)k * n + i]);
        double mean = 0.0, var = 0.0;
        for (unsigned int j = 0; j < v.size(); j++) {
            const double delta = v[j] - mean;
            mean += delta / (j + 1);
            var += delta * (v[j] - mean);
        }
        var /= (v.size() - 1);
        for (unsigned int j = 0; j < v.size(); ++j) {
            v[j] -= mean;
            if (var > 0.0) v[j] /= std::sqrt(var);
        }
        int idx = 0;
        for (int i = 0; i < n; ++i)
            if (!excl || excl[i] > 0.0) data[(size_t)k * n + i] = v[idx++];
    }
}
