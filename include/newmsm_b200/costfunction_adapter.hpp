// costfunction_adapter.hpp — the reference-side binding of include/msmgpu.h for msm-newmeshreg's cost functions.
//
// Header-only. GpuCostFunction<Base, KIND> derives from the reference's own cost-function class
// (msm-newmeshreg/src/DiscreteCostFunction.h:205-244) and overrides exactly the virtuals that do the data-parallel work
// (DiscreteCostFunction.h:41-59: get_source_data, computeUnaryCosts, computeUnaryCost, computeTripletCost,
// computePairwiseCosts); everything else (parameters, meshes, labels, the cost tables the solvers read, ownership)
// is inherited unchanged, so NonLinearSRegDiscreteModel, FPD::FastPD and Fusion::optimize (Fusion.h:118-246)
// use it through the same pointers and the same call order (DiscreteModel.cpp:216-262).
//
//   * get_source_data()          -> msmgpu_costfn_set_cpgrid[_ho]   (patch membership, cpp:102-107, 334-351, 468-485)
//   * computeUnaryCosts()        -> msmgpu_costfn_unary_table        (cpp:236-243)
//   * computeUnaryCost(n, l)     -> entry of that table, built once per iteration (Fusion asks per call, Fusion.h:148-155)
//   * computeTripletCost(t,a,b,c)-> Fusion asks 8 combinations per triplet for one candidate label, from OpenMP workers
//                                   (Fusion.h:181-196). The first request of a (labeling, label) phase evaluates ALL triplets x 8
//                                   combinations in one msmgpu_costfn_triplet_batch launch; the others are table look-ups.
//   * computePairwiseCosts(p)    -> the rotation regulariser table (cpp:190-243) evaluated in parallel on the host (libm-bound:
//                                   acos / pow per entry) with the N*L rotation matrices computed once instead of 2*P*L^2 times,
//                                   identical values.
//
// There is no CPU fallback for the device paths: a failing msmgpu call throws MeshregException with msmgpu_last_error().
// regoption 4/5 (anatomical strain, cpp:169-181, 245-301): the anatomical meshes and maps the model received through set_anatomical /
// set_anatomical_neighbourhood are uploaded once per level (msmgpu_costfn_set_anatomical) and the same batches evaluate them.
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <omp.h>

#ifndef NEWMSM_B200_DISCRETEMODEL_HEADER
#define NEWMSM_B200_DISCRETEMODEL_HEADER "NewMeshReg/DiscreteModel.h"
#endif
#include NEWMSM_B200_DISCRETEMODEL_HEADER   // the reference's DiscreteModel / cost-function classes

#include "../msmgpu.h"

namespace newmeshreg_gpu {

using newmeshreg::MeshregException;
using newresampler::Mesh;
using newresampler::Point;

namespace detail {

inline void check(msmgpu_status st) {
    if (st != MSMGPU_OK) {
        static thread_local std::string msg;
        msg = std::string("msmgpu: ") + msmgpu_last_error();
        throw MeshregException(msg.c_str());
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

inline std::vector<double> coords_of(const Mesh& m, int n = -1) {
    if (n < 0) n = m.nvertices();
    std::vector<double> xyz(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const Point& p = m.get_coord(i);
        xyz[3 * (size_t)i] = p.X; xyz[3 * (size_t)i + 1] = p.Y; xyz[3 * (size_t)i + 2] = p.Z;
    }
    return xyz;
}

inline std::vector<int32_t> triangles_of(const Mesh& m) {
    std::vector<int32_t> tri(3 * (size_t)m.ntriangles());
    for (int t = 0; t < m.ntriangles(); ++t)
        for (int k = 0; k < 3; ++k) tri[3 * (size_t)t + k] = m.get_triangle_vertexID(t, k);
    return tri;
}

struct Timers {   // wall-clock split printed by the integration binary (seconds)
    double source = 0, unary = 0, triplet = 0, pairwise = 0;
    long unary_tables = 0, triplet_batches = 0;
    double source_done_at = 0;   // omp_get_wtime() when the last get_source_data returned (the optimiser runs next)
};
inline Timers& timers() { static Timers t; return t; }

// MSMGPU_VERIFY=1: diagnostic mode. The reference's own CPU implementation of every overridden virtual runs side by side,
// in-process, on the same object state, and the results are compared bit for bit (report on stderr).
inline bool verify() { static const bool v = std::getenv("MSMGPU_VERIFY") != nullptr; return v; }
inline long count_diff(const double* a, const double* b, size_t n, double* worst) {
    long bad = 0;
    *worst = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::memcmp(a + i, b + i, sizeof(double)) != 0) { ++bad; *worst = std::max(*worst, std::fabs(a[i] - b[i])); }
    return bad;
}

}  // namespace detail

template <class Base, msmgpu_cost_kind KIND>
class GpuCostFunction : public Base {
    static constexpr bool kHO = KIND == MSMGPU_COST_HO_UNIVARIATE || KIND == MSMGPU_COST_HO_MULTIVARIATE;

    newmeshreg::DiscreteModel* model_;   // owner of labeling[] (DiscreteModel.h:46)
    double percentile_ = 0.75;           // "percentile" parameter (DICE measures)
    msmgpu_mesh* d_target_ = nullptr;
    msmgpu_octree* d_tree_ = nullptr;
    msmgpu_costfn* d_cf_ = nullptr;

    // per-iteration host copies handed to the C ABI
    std::vector<double> labels_, rot_, orig_cp_;
    std::vector<int32_t> trip_;
    bool iter_arrays_ready_ = false, anat_ready_ = false;

    std::mutex mu_;
    std::atomic<bool> unary_ready_{false};

    // triplet caches. Readers (OpenMP workers inside Fusion::optimize) only do an acquire-load of a raw pointer; a table is
    // never modified after publication and retired tables stay alive until the next initialize() (a serial point), so a
    // reader that still holds the previous pointer keeps reading valid memory.
    struct Table {
        std::vector<int32_t> snap;    // labeling snapshot the table was computed for
        std::vector<double> val;      // [T] (current labels) or [T][8] (Fusion's combinations)
        int label = -1;
    };
    std::atomic<const Table*> cur_{nullptr}, fus_{nullptr};
    std::vector<std::unique_ptr<Table>> tables_;

    void release() {
        if (d_cf_) msmgpu_costfn_destroy(d_cf_);
        if (d_tree_) msmgpu_octree_destroy(d_tree_);
        if (d_target_) msmgpu_mesh_destroy(d_target_);
        d_cf_ = nullptr; d_tree_ = nullptr; d_target_ = nullptr;
        anat_ready_ = false;
    }

    void ensure_device() {
        if (d_cf_) return;
        msmgpu_ctx* ctx = detail::context();
        const Mesh& T = this->_TARGET;
        const std::vector<double> txyz = detail::coords_of(T);
        const std::vector<int32_t> ttri = detail::triangles_of(T);
        detail::check(msmgpu_mesh_create(ctx, T.nvertices(), txyz.data(), T.ntriangles(), ttri.data(), &d_target_));
        detail::check(msmgpu_octree_build(d_target_, &d_tree_));
        const int D = (int)this->FEAT->get_dim(), ns = this->_SOURCE.nvertices(), nt = T.nvertices();
        std::vector<double> sf((size_t)D * ns), rf((size_t)D * nt);
        for (int d = 0; d < D; ++d) {
            for (int i = 0; i < ns; ++i) sf[(size_t)d * ns + i] = this->FEAT->get_input_val(d + 1, i + 1);
            for (int i = 0; i < nt; ++i) rf[(size_t)d * nt + i] = this->FEAT->get_ref_val(d + 1, i + 1);
        }
        const std::vector<double> sxyz = detail::coords_of(this->_SOURCE);
        detail::check(msmgpu_costfn_create(d_tree_, KIND, this->_simmeasure, ns, sxyz.data(), D, sf.data(), rf.data(), &d_cf_));
        detail::check(msmgpu_costfn_set_percentile(d_cf_, percentile_));
    }

    // labels / rotations / triplets / undeformed control points of the current iteration (set_labels, setTriplets)
    void ensure_iter_arrays() {
        if (iter_arrays_ready_) return;
        const int L = (int)this->_labels.size(), N = this->_CPgrid.nvertices();
        labels_.resize(3 * (size_t)L);
        for (int l = 0; l < L; ++l) { labels_[3 * l] = this->_labels[l].X; labels_[3 * l + 1] = this->_labels[l].Y; labels_[3 * l + 2] = this->_labels[l].Z; }
        rot_.resize(9 * (size_t)N);
        for (int k = 0; k < N; ++k)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) rot_[9 * (size_t)k + 3 * r + c] = (*this->ROTATIONS)[k](r + 1, c + 1);
        if (this->_triplets && this->m_num_triplets > 0) {
            trip_.assign(this->_triplets, this->_triplets + 3 * (size_t)this->m_num_triplets);
            orig_cp_ = detail::coords_of(this->_ORIG, N);   // _ORIG.get_coord(node) (cpp:166-168): the nested icosphere's first N vertices
        }
        if (this->_rmode >= 4) ensure_anatomical();
        iter_arrays_ready_ = true;
    }

    // regoption 4/5: _aSOURCE, _TARGEThi, _aTARGET, NEARESTFACES and _ANATbaryweights (DiscreteCostFunction.h:160-169) are fixed for the
    // lifetime of this object (mesh_registration.cpp:93-98 sets them once per level, before the first iteration)
    void ensure_anatomical() {
        if (anat_ready_) return;
        const int T = this->m_num_triplets;
        if ((int)this->NEARESTFACES.size() < T || this->_aSOURCE.nvertices() == 0)
            throw MeshregException("msmgpu: regoption 4/5 without anatomical meshes (set_anatomical / set_anatomical_neighbourhood)");
        const std::vector<double> as = detail::coords_of(this->_aSOURCE), th = detail::coords_of(this->_TARGEThi), at = detail::coords_of(this->_aTARGET);
        const std::vector<int32_t> ast = detail::triangles_of(this->_aSOURCE), tht = detail::triangles_of(this->_TARGEThi);
        if (this->_aTARGET.nvertices() != this->_TARGEThi.nvertices()) throw MeshregException("msmgpu: _aTARGET and _TARGEThi differ in size");
        std::vector<int32_t> fptr(T + 1, 0), fids, bptr(this->_aSOURCE.nvertices() + 1, 0), bkey;
        std::vector<double> bw;
        for (int t = 0; t < T; ++t) {
            fids.insert(fids.end(), this->NEARESTFACES[t].begin(), this->NEARESTFACES[t].end());
            fptr[t + 1] = (int32_t)fids.size();
        }
        for (int v = 0; v < this->_aSOURCE.nvertices(); ++v) {
            if (v < (int)this->_ANATbaryweights.size())
                for (const auto& it : this->_ANATbaryweights[v]) { bkey.push_back(it.first); bw.push_back(it.second); }   // std::map: ascending keys
            bptr[v + 1] = (int32_t)bkey.size();
        }
        msmgpu_anatomical A;
        A.n_av = this->_aSOURCE.nvertices(); A.asource_xyz = as.data(); A.n_at = this->_aSOURCE.ntriangles(); A.asource_tri = ast.data();
        A.n_hv = this->_TARGEThi.nvertices(); A.thi_xyz = th.data(); A.n_ht = this->_TARGEThi.ntriangles(); A.thi_tri = tht.data();
        A.atarget_xyz = at.data();
        A.face_ptr = fptr.data(); A.face_ids = fids.data(); A.bary_ptr = bptr.data(); A.bary_key = bkey.data(); A.bary_w = bw.data();
        detail::check(msmgpu_costfn_set_anatomical(d_cf_, T, &A));
        anat_ready_ = true;
    }

    msmgpu_reg_params reg_params() const {
        msmgpu_reg_params p;
        p.lambda = this->_reglambda; p.shear_modulus = this->_mu; p.bulk_modulus = this->_kappa;
        p.k_exponent = this->_k_exp; p.exponent = this->_rexp; p.rmode = this->_rmode;
        return p;
    }

    void build_unary_table() {
        const double t0 = omp_get_wtime();
        ensure_iter_arrays();
        detail::check(msmgpu_costfn_unary_table(d_cf_, this->m_num_labels, labels_.data(), rot_.data(), this->unarycosts, nullptr));
        detail::timers().unary += omp_get_wtime() - t0;
        detail::timers().unary_tables++;
    }

    const Table* make_table(int label) {
        const double t0 = omp_get_wtime();
        ensure_iter_arrays();
        const int T = this->m_num_triplets, N = this->m_num_nodes;
        tables_.emplace_back(new Table());
        Table* tb = tables_.back().get();
        const int* lab = model_->getLabeling();
        tb->snap.assign(lab, lab + N);
        tb->label = label;
        const msmgpu_reg_params prm = reg_params();
        if (label < 0) {   // costs at the current labeling
            std::vector<int32_t> rt(T), la(T), lb(T), lc(T);
            for (int t = 0; t < T; ++t) { rt[t] = t; la[t] = lab[trip_[3 * t]]; lb[t] = lab[trip_[3 * t + 1]]; lc[t] = lab[trip_[3 * t + 2]]; }
            tb->val.resize(T);
            detail::check(msmgpu_costfn_triplet_costs(d_cf_, T, trip_.data(), this->m_num_labels, labels_.data(), rot_.data(), orig_cp_.data(), &prm,
                                                      T, rt.data(), la.data(), lb.data(), lc.data(), tb->val.data()));
        } else {           // Fusion.h:181-196: combination c = (A?4:0)|(B?2:0)|(C?1:0), bit set = candidate label
            tb->val.resize(8 * (size_t)T);
            detail::check(msmgpu_costfn_triplet_batch(d_cf_, T, trip_.data(), this->m_num_labels, labels_.data(), rot_.data(), orig_cp_.data(), &prm,
                                                      tb->snap.data(), label, tb->val.data()));
        }
        detail::timers().triplet += omp_get_wtime() - t0;
        detail::timers().triplet_batches++;
        if (detail::verify()) {
            long bad = 0, n = 0;
            double worst = 0;
            for (int t = 0; t < T; t += 5) {
                const int a = trip_[3 * t], b = trip_[3 * t + 1], c = trip_[3 * t + 2];
                for (int combo = 0; combo < (label < 0 ? 1 : 8); ++combo, ++n) {
                    const int la = (combo & 4) ? label : tb->snap[a], lb = (combo & 2) ? label : tb->snap[b], lc = (combo & 1) ? label : tb->snap[c];
                    const double r = Base::computeTripletCost(t, la, lb, lc), g = label < 0 ? tb->val[t] : tb->val[8 * (size_t)t + combo];
                    if (std::memcmp(&r, &g, sizeof(double)) != 0) { ++bad; worst = std::max(worst, std::fabs(r - g) / std::max(std::fabs(r), 1e-300)); }
                }
            }
            std::fprintf(stderr, "[msmgpu verify] triplet batch (label %d): %ld of %ld sampled costs differ (max rel %.3g)\n", label, bad, n, worst);
        }
        return tb;
    }

    static bool fresh(const Table* tb, const int* lab, int a, int b, int c, int label) {
        return tb && tb->label == label && tb->snap[a] == lab[a] && tb->snap[b] == lab[b] && tb->snap[c] == lab[c];
    }

public:
    explicit GpuCostFunction(newmeshreg::DiscreteModel* model) : model_(model) {}

    void set_parameters(newmeshreg::myparam& P) override {   // cpp:119-133; `sim` keeps the percentile private, so it is read here too
        Base::set_parameters(P);
        auto it = P.find("percentile");
        if (it != P.end()) percentile_ = std::get<double>(it->second);
    }
    ~GpuCostFunction() override { release(); }

    void initialize(int numNodes, int numLabels, int numPairs, int numTriplets) override {
        Base::initialize(numNodes, numLabels, numPairs, numTriplets);
        iter_arrays_ready_ = false;
        unary_ready_.store(false);
        cur_.store(nullptr);
        fus_.store(nullptr);
        tables_.clear();
    }

    // once per discrete iteration (DiscreteModel.cpp:254): patches of the current source / control-point grid
    void get_source_data() override {
        const double t0 = omp_get_wtime();
        ensure_device();
        const std::vector<double> sxyz = detail::coords_of(this->_SOURCE);
        detail::check(msmgpu_costfn_reset_source(d_cf_, sxyz.data()));
        this->resample_weights();   // AbsoluteWeights (cpp:303-322): newresampler::metric_resample, itself bound to the GPU by the resampler adapter
        const int N = this->_CPgrid.nvertices(), ns = this->_SOURCE.nvertices();
        const std::vector<double> cp = detail::coords_of(this->_CPgrid);
        const int cfw_rows = this->_HIGHREScfweight.Nrows();
        std::vector<double> cfw((size_t)cfw_rows * ns), absw(N);
        for (int r = 0; r < cfw_rows; ++r)
            for (int i = 0; i < ns; ++i) cfw[(size_t)r * ns + i] = this->_HIGHREScfweight(r + 1, i + 1);
        for (int k = 0; k < N; ++k) absw[k] = this->AbsoluteWeights(k + 1);
        if (kHO) {
            const std::vector<int32_t> ctri = detail::triangles_of(this->_CPgrid);
            detail::check(msmgpu_costfn_set_cpgrid_ho(d_cf_, N, cp.data(), this->_CPgrid.ntriangles(), ctri.data(), cfw_rows, cfw.data(), absw.data()));
        } else {
            std::vector<double> sep(N);
            for (int k = 0; k < N; ++k) sep[k] = this->MAXSEP(k + 1);
            detail::check(msmgpu_costfn_set_cpgrid(d_cf_, N, cp.data(), sep.data(), this->_controlptrange, cfw_rows, cfw.data(), absw.data()));
        }
        detail::timers().source += omp_get_wtime() - t0;
        detail::timers().source_done_at = omp_get_wtime();
        if (detail::verify()) {
            Base::get_source_data();   // the reference's patches (and AbsoluteWeights again) for the side-by-side evaluation
            const int rows = kHO ? this->_CPgrid.ntriangles() : N;
            std::vector<int32_t> rowptr((size_t)rows + 1), mem((size_t)64 * ns + 1024);
            detail::check(msmgpu_costfn_patches(d_cf_, rowptr.data(), nullptr));
            if ((size_t)rowptr[rows] > mem.size()) mem.resize(rowptr[rows]);
            detail::check(msmgpu_costfn_patches(d_cf_, rowptr.data(), mem.data()));
            long bad = 0;
            for (int k = 0; k < rows; ++k) {
                const std::vector<int>& ref = this->_sourceinrange[k];
                const bool same = (int)ref.size() == rowptr[k + 1] - rowptr[k] && std::equal(ref.begin(), ref.end(), mem.begin() + rowptr[k]);
                bad += !same;
            }
            std::fprintf(stderr, "[msmgpu verify] get_source_data: %d patches, %d entries, %ld patches differ\n", rows, rowptr[rows], bad);
        }
    }

    void computeUnaryCosts() override {
        if (kHO) { Base::computeUnaryCosts(); return; }   // HO kinds: computeUnaryCost is the constant 0 (DiscreteCostFunction.h:222)
        std::lock_guard<std::mutex> g(mu_);
        build_unary_table();
        unary_ready_.store(true);
        if (detail::verify()) {
            const int N = this->m_num_nodes, L = this->m_num_labels;
            std::vector<double> ref((size_t)N * L);
            for (int j = 0; j < L; ++j)
                for (int k = 0; k < N; ++k) ref[(size_t)j * N + k] = Base::computeUnaryCost(k, j);   // single thread: see the races noted in tests/
            double worst;
            const long bad = detail::count_diff(ref.data(), this->unarycosts, ref.size(), &worst);
            std::fprintf(stderr, "[msmgpu verify] unary table %d x %d: %ld entries differ (max |diff| %.3g)\n", L, N, bad, worst);
        }
    }

    double computeUnaryCost(int node, int label) override {
        if (kHO) return 0;
        if (!unary_ready_.load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> g(mu_);
            if (!unary_ready_.load()) { build_unary_table(); unary_ready_.store(true, std::memory_order_release); }
        }
        return this->unarycosts[label * this->m_num_nodes + node];
    }

    double computeTripletCost(int triplet, int labelA, int labelB, int labelC) override {
        const int* lab = model_->getLabeling();
        const int a = this->_triplets[3 * triplet], b = this->_triplets[3 * triplet + 1], c = this->_triplets[3 * triplet + 2];
        const bool da = labelA != lab[a], db = labelB != lab[b], dc = labelC != lab[c];
        if (!da && !db && !dc) {
            const Table* tb = cur_.load(std::memory_order_acquire);
            if (!fresh(tb, lab, a, b, c, -1)) {
                std::lock_guard<std::mutex> g(mu_);
                tb = cur_.load(std::memory_order_acquire);
                if (!fresh(tb, lab, a, b, c, -1)) { tb = make_table(-1); cur_.store(tb, std::memory_order_release); }
            }
            return tb->val[triplet];
        }
        const int label = da ? labelA : (db ? labelB : labelC);
        if ((da && labelA != label) || (db && labelB != label) || (dc && labelC != label)) {
            // not one of Fusion's combinations (two different non-current labels): evaluate this request on its own
            const msmgpu_reg_params prm = reg_params();
            const int32_t rt = triplet, la = labelA, lb = labelB, lc = labelC;
            double out = 0;
            std::lock_guard<std::mutex> g(mu_);
            ensure_iter_arrays();
            detail::check(msmgpu_costfn_triplet_costs(d_cf_, this->m_num_triplets, trip_.data(), this->m_num_labels, labels_.data(), rot_.data(), orig_cp_.data(),
                                                      &prm, 1, &rt, &la, &lb, &lc, &out));
            return out;
        }
        const Table* tb = fus_.load(std::memory_order_acquire);
        if (!fresh(tb, lab, a, b, c, label)) {
            std::lock_guard<std::mutex> g(mu_);
            tb = fus_.load(std::memory_order_acquire);
            if (!fresh(tb, lab, a, b, c, label)) { tb = make_table(label); fus_.store(tb, std::memory_order_release); }
        }
        return tb->val[8 * (size_t)triplet + ((da ? 4 : 0) | (db ? 2 : 0) | (dc ? 1 : 0))];
    }

    // The pairwise (regoption 1 / FastPD) table: cost(pair, la, lb) = lambda * (sqrt(2) * theta / theta_MVD)^rexp with theta the angle of
    // R1^T R2, R1 = estimate_rotation_matrix(CP[n0], ROT[n0] * label[la]), R2 likewise for (n1, lb); FOLDING if a triangle around n0
    // flips (cpp:190-233). The reference evaluates it serially, entry by entry, through the member mesh (two set_coord + restore per
    // entry) and recomputes both rotation matrices P * L^2 * 2 times although they only depend on (node, label). Here the N * L
    // matrices and moved points are computed once (the reference's own estimate_rotation_matrix and operator*), and the table is
    // filled in parallel with the same expressions in the same order (host libm: acos, pow), so every entry is the reference's value.
    void computePairwiseCosts(const int* pairs) override {
        const double t0 = omp_get_wtime();
        const int P = this->m_num_pairs, L = (int)this->_labels.size(), LL = this->m_num_labels, N = this->_CPgrid.nvertices();
        int nthreads = omp_get_max_threads();
        if (const char* e = std::getenv("MSMGPU_HOST_THREADS")) nthreads = std::max(1, std::atoi(e));
        const Mesh& G = this->_CPgrid;
        std::vector<Point> cp((size_t)N), moved((size_t)N * L);
        std::vector<double> R((size_t)N * L * 9);
        #pragma omp parallel for num_threads(nthreads)
        for (int n = 0; n < N; ++n) {
            cp[n] = G.get_coord(n);
            for (int l = 0; l < L; ++l) {
                const Point q = (*this->ROTATIONS)[n] * this->_labels[l];
                moved[(size_t)n * L + l] = q;
                const NEWMAT::Matrix M = newresampler::estimate_rotation_matrix(cp[n], q);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) R[((size_t)n * L + l) * 9 + 3 * r + c] = M(r + 1, c + 1);
            }
        }
        const int T = G.ntriangles();
        std::vector<int> tv(3 * (size_t)T);
        std::vector<Point> onormal((size_t)T);
        for (int t = 0; t < T; ++t) {
            for (int k = 0; k < 3; ++k) tv[3 * (size_t)t + k] = G.get_triangle_vertexID(t, k);
            onormal[t] = this->_oCPgrid.get_triangle(t).normal();
        }
        const double theta_MVD = 2 * asin(this->MVDmax / (2 * RAD));
        const double lambda = this->_reglambda, rexp = this->_rexp;
        #pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
        for (int i = 0; i < P; ++i) {
            const int n0 = this->_pairs[2 * i], n1 = this->_pairs[2 * i + 1];
            const std::vector<int> around(G.tIDbegin(n0), G.tIDend(n0));
            for (int j = 0; j < L; ++j)
                for (int k = 0; k < L; ++k) {
                    const double* R1 = &R[((size_t)n0 * L + j) * 9];
                    const double* R2 = &R[((size_t)n1 * L + k) * 9];
                    // Trace of R1.t() * R2 with the matrix product's own summation order (left to right from 0)
                    double trace = 0.0;
                    for (int d = 0; d < 3; ++d) {
                        double sum = 0.0;
                        for (int m = 0; m < 3; ++m) sum += R1[3 * m + d] * R2[3 * m + d];
                        trace += sum;
                    }
                    double cost = 0.0;
                    if (fabs(1 - (trace - 1) / 2) > EPSILON) {
                        const Point& p0 = moved[(size_t)n0 * L + j];
                        const Point& p1 = moved[(size_t)n1 * L + k];
                        auto at = [&](int v) -> const Point& { return v == n0 ? p0 : (v == n1 ? p1 : cp[v]); };
                        bool folded = false;
                        for (int t : around) {
                            const Point &a = at(tv[3 * (size_t)t]), &b = at(tv[3 * (size_t)t + 1]), &c = at(tv[3 * (size_t)t + 2]);
                            Point nrm = (c - a) * (b - a);          // Triangle::normal (triangle.cpp:42-47)
                            nrm.normalize();
                            if ((onormal[t] | nrm) < 0.0) { folded = true; break; }
                        }
                        if (folded) cost = FOLDING;
                        else {
                            const double theta = acos((trace - 1) / 2);
                            if (rexp == 1) cost = lambda * ((sqrt(2) * theta) / theta_MVD);
                            else cost = lambda * std::pow(((sqrt(2) * theta) / theta_MVD), rexp);
                        }
                    }
                    this->paircosts[i * LL * LL + k * LL + j] = cost;
                }
        }
        (void)pairs;
        detail::timers().pairwise += omp_get_wtime() - t0;
        if (detail::verify()) {
            long bad = 0, n = 0;
            for (int i = 0; i < P; i += 7)
                for (int j = 0; j < L; ++j)
                    for (int k = 0; k < L; ++k, ++n) {
                        const double r = Base::computePairwiseCost(i, j, k), g = this->paircosts[i * LL * LL + k * LL + j];
                        bad += std::memcmp(&r, &g, sizeof(double)) != 0;
                    }
            std::fprintf(stderr, "[msmgpu verify] pairwise table: %ld of %ld sampled entries differ\n", bad, n);
        }
    }

};

using GpuUnivariate = GpuCostFunction<newmeshreg::UnivariateNonLinearSRegDiscreteCostFunction, MSMGPU_COST_UNIVARIATE>;
using GpuMultivariate = GpuCostFunction<newmeshreg::MultivariateNonLinearSRegDiscreteCostFunction, MSMGPU_COST_MULTIVARIATE>;
using GpuPatchwise = GpuCostFunction<newmeshreg::PatchwiseMultivariateNonLinearSRegDiscreteCostFunction, MSMGPU_COST_PATCHWISE>;
using GpuHOUnivariate = GpuCostFunction<newmeshreg::HOUnivariateNonLinearSRegDiscreteCostFunction, MSMGPU_COST_HO_UNIVARIATE>;
using GpuHOMultivariate = GpuCostFunction<newmeshreg::HOMultivariateNonLinearSRegDiscreteCostFunction, MSMGPU_COST_HO_MULTIVARIATE>;

// The selection NonLinearSRegDiscreteModel::initialize_cost_function makes (DiscreteModel.cpp:44-60), returning the GPU-backed
// class of the same kind. `model` owns labeling[]; the result still needs set_parameters(P) like the original.
inline std::shared_ptr<newmeshreg::NonLinearSRegDiscreteCostFunction> make_gpu_costfunction(newmeshreg::DiscreteModel* model, bool multivariate,
                                                                                           bool patchwise, bool triclique) {
    if (multivariate) {
        if (patchwise) return std::make_shared<GpuPatchwise>(model);
        if (triclique) return std::make_shared<GpuHOMultivariate>(model);
        return std::make_shared<GpuMultivariate>(model);
    }
    if (triclique) return std::make_shared<GpuHOUnivariate>(model);
    return std::make_shared<GpuUnivariate>(model);
}

}  // namespace newmeshreg_gpu
