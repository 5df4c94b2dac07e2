// group_adapter.hpp — the reference-side binding of include/msmgpu.h for groupwise registration (gMSM).
//
// Header-only. The reference keeps the groupwise state private to DiscreteGroupModel / DiscreteGroupCostFunction
// (msm-newmeshreg/src/DiscreteGroupModel.h:30-42, DiscreteGroupCostFunction.h:57-67) and constructs both inline, so there is
// no virtual seam to derive from as for the pairwise cost functions (costfunction_adapter.hpp). The binding is therefore four
// free functions with the bodies a maintainer would put INTO the reference's own members:
//
//   DiscreteGroupModel::estimate_pairs()                    (DiscreteGroupModel.cpp:37-55)  -> GroupBinding::estimate_pairs
//   DiscreteGroupModel::get_patch_data()                    (DiscreteGroupModel.cpp:88-121) -> GroupBinding::get_patch_data
//   DiscreteGroupCostFunction::computePairwiseCost(p,a,b)   (DiscreteGroupCostFunction.cpp:54-97) -> GroupBinding::pairwise
//   DiscreteGroupCostFunction::computeTripletCost(t,a,b,c)  (DiscreteGroupCostFunction.cpp:26-52) -> GroupBinding::triplet
//
// They read the classes' private members, so a translation unit that includes this header outside the classes is compiled
// with -fno-access-control (integration/newmsm_gpu_group_hooks.cpp); inside the classes no flag is needed.
//
// Fusion::optimize (Fusion.h:118-246) asks per (pair, labelA, labelB) and per (triplet, a, b, c) from OpenMP workers. The first
// request of a (labeling, candidate label) phase evaluates ALL pairs x 4 combinations (all triplets x 8) in one batch on the
// device; every other request of the phase is a table look-up. The costs at the CURRENT labeling (combination 0 and
// evaluateTotalCostSum, DiscreteCostFunction.cpp:55-77) after a phase are read out of that phase's table, since every node either
// kept its label or took the candidate.
//
// MSMGPU_DEVICES=n (default 1): one context per device; the subjects of get_patch_data and the pair blocks of every batch are
// sharded over them (subject fields are exchanged device to device once per iteration); results do not depend on n.
//
// Cost masks (`set_masks`, DiscreteGroupModel.cpp:164): the mask mesh's values go to the device with the iteration state
// (msmgpu_group_set_mask); the pair costs weight every common template vertex by |mask| (DiscreteGroupCostFunction.cpp:77).
#pragma once

#include <thread>

#include "costfunction_adapter.hpp"

#ifndef NEWMSM_B200_GROUPMODEL_HEADER
#define NEWMSM_B200_GROUPMODEL_HEADER "NewMeshReg/DiscreteGroupModel.h"
#endif
#include NEWMSM_B200_GROUPMODEL_HEADER

namespace newmeshreg_gpu {

struct GroupTimers {
    double pairs = 0, fields = 0, pair_batches = 0, triplet_batches = 0;
    long n_pair_batches = 0, n_triplet_batches = 0, n_iterations = 0;
    long long pair_costs = 0, triplet_costs = 0;
};
inline GroupTimers& group_timers() { static GroupTimers t; return t; }

class GroupBinding {
    using Model = newmeshreg::DiscreteGroupModel;
    using CostFn = newmeshreg::DiscreteGroupCostFunction;

    struct Device {
        msmgpu_ctx* ctx = nullptr;
        msmgpu_mesh* tpl = nullptr;
        msmgpu_octree* tpl_tree = nullptr;
        double* fields = nullptr;      // [S][L][n_tpl][D] for ALL subjects
        size_t fields_bytes = 0;
        double* batch_out = nullptr;   // [pairs of this device's block][4]: result of one 4-combination batch before it goes to the host
        size_t batch_bytes = 0;
        msmgpu_group* group = nullptr;
        msmgpu_triplet_plan* plan = nullptr;   // control grids, rotations, labels and triplets of the iteration, resident for the label phases
    };
    std::vector<Device> dev_;

    // a table is never modified after publication; the previous one of its chain stays alive (a worker may still be comparing
    // against it), older ones are unreachable because the phases of Fusion::optimize are separated by OpenMP barriers
    struct Table {
        std::vector<int32_t> snap;
        std::vector<double> val;
        int label = -1;
        int width = 1;
    };
    struct Chain {
        std::atomic<const Table*> cur{nullptr};
        std::unique_ptr<Table> live, previous;
        void publish(std::unique_ptr<Table> t) {
            previous = std::move(live);
            live = std::move(t);
            cur.store(live.get(), std::memory_order_release);
        }
        void clear() { cur.store(nullptr); live.reset(); previous.reset(); }
    };
    Chain pair_cur_, pair_fus_, trip_cur_, trip_fus_;
    std::mutex mu_;

    Model* model_ = nullptr;
    CostFn* cf_ = nullptr;
    bool active_ = false;
    int S_ = 0, ncp_ = 0, L_ = 0, D_ = 0, P_ = 0, T_ = 0;
    std::vector<int32_t> pairs_, trip_;
    std::vector<double> labels_, rot_, cps_, orig_;

    static int wanted_devices() {
        int n = 1;
        if (const char* e = std::getenv("MSMGPU_DEVICES")) n = std::max(1, std::atoi(e));
        return std::min(n, std::max(1, msmgpu_device_count()));
    }

    void ensure_devices() {
        if (!dev_.empty()) return;
        const int n = wanted_devices();
        dev_.resize(n);
        dev_[0].ctx = detail::context();
        const char* e = std::getenv("MSMGPU_DEVICE");
        const int first = e ? std::atoi(e) : 0;
        for (int d = 1; d < n; ++d) detail::check(msmgpu_ctx_create(first + d, nullptr, &dev_[d].ctx));
    }

    void drop_iteration_state() {
        for (Device& d : dev_) {
            if (d.group) msmgpu_group_destroy(d.group);
            if (d.plan) msmgpu_triplet_plan_destroy(d.plan);
            d.plan = nullptr;
            if (d.tpl_tree) msmgpu_octree_destroy(d.tpl_tree);
            if (d.tpl) msmgpu_mesh_destroy(d.tpl);
            d.group = nullptr; d.tpl_tree = nullptr; d.tpl = nullptr;
        }
        pair_cur_.clear(); pair_fus_.clear(); trip_cur_.clear(); trip_fus_.clear();
        active_ = false;
    }

    static void shard(int n, int part, int parts, int& b, int& e) {   // contiguous blocks, remainder to the first ranks
        const int q = n / parts, r = n % parts;
        b = part * q + std::min(part, r);
        e = b + q + (part < r ? 1 : 0);
    }

    template <class F>
    void on_devices(F&& f) {   // f(device index) on one host thread per device; the first failure is rethrown
        const int n = (int)dev_.size();
        if (n == 1) { f(0); return; }
        std::vector<std::string> err(n);
        std::vector<std::thread> th;
        for (int d = 0; d < n; ++d)
            th.emplace_back([&, d] {
                try { f(d); } catch (const std::exception& ex) { err[d] = ex.what(); if (err[d].empty()) err[d] = "msmgpu: device worker failed"; }
            });
        for (auto& t : th) t.join();
        for (const std::string& m : err)
            if (!m.empty()) { static thread_local std::string keep; keep = m; throw MeshregException(keep.c_str()); }
    }

    msmgpu_reg_params reg_params() const {
        msmgpu_reg_params p;
        p.lambda = cf_->_reglambda; p.shear_modulus = cf_->_mu; p.bulk_modulus = cf_->_kappa;
        p.k_exponent = cf_->_k_exp; p.exponent = cf_->_rexp; p.rmode = 3;
        return p;
    }

    // ---- pair tables -----------------------------------------------------------------------------------------------------
    std::unique_ptr<Table> pair_table(int label) {   // label < 0: costs at the current labeling
        const double t0 = omp_get_wtime();
        std::unique_ptr<Table> tb(new Table());
        const int* lab = model_->getLabeling();
        tb->snap.assign(lab, lab + S_ * ncp_);
        tb->label = label;
        tb->width = label < 0 ? 1 : 4;
        tb->val.resize((size_t)tb->width * P_);
        const int n = (int)dev_.size();
        if (label < 0) {
            std::vector<int32_t> rp(P_), la(P_), lb(P_);
            for (int p = 0; p < P_; ++p) { rp[p] = p; la[p] = lab[pairs_[2 * (size_t)p]]; lb[p] = lab[pairs_[2 * (size_t)p + 1]]; }
            on_devices([&](int d) {
                int b, e;
                shard(P_, d, n, b, e);
                if (e > b)
                    detail::check(msmgpu_group_pair_costs(dev_[d].group, P_, pairs_.data(), e - b, rp.data() + b, la.data() + b, lb.data() + b, tb->val.data() + b));
            });
            group_timers().pair_costs += P_;
        } else {
            on_devices([&](int d) {   // the pair list is resident on every device (get_patch_data); only the labeling goes up, the block's costs come down
                int b, e;
                shard(P_, d, n, b, e);
                if (e <= b) return;
                detail::check(msmgpu_group_pair_batch_dev(dev_[d].group, b, e - b, tb->snap.data(), label, dev_[d].batch_out));
                detail::check(msmgpu_device_download(dev_[d].ctx, tb->val.data() + 4 * (size_t)b, dev_[d].batch_out, 4 * (size_t)(e - b) * sizeof(double)));
            });
            group_timers().pair_costs += 4LL * P_;
        }
        if (cf_->fixnan)
            for (double& v : tb->val)
                if (std::isnan(v)) v = FIX_NAN;
        group_timers().pair_batches += omp_get_wtime() - t0;
        group_timers().n_pair_batches++;
        return tb;
    }

    // costs at the current labeling out of the last phase's table: possible when every node kept its label or took the candidate
    template <int NODES>
    static std::unique_ptr<Table> carry_over(const Table* fus, const int* lab, int n_nodes, const std::vector<int32_t>& items, int n_items) {
        if (!fus) return nullptr;
        for (int k = 0; k < n_nodes; ++k)
            if (lab[k] != fus->snap[k] && lab[k] != fus->label) return nullptr;
        std::unique_ptr<Table> tb(new Table());
        tb->snap.assign(lab, lab + n_nodes);
        tb->val.resize(n_items);
        const int W = 1 << NODES;
        for (int i = 0; i < n_items; ++i) {
            int combo = 0;
            for (int k = 0; k < NODES; ++k) {
                const int node = items[(size_t)NODES * i + k];
                // a node whose old label already was the candidate reads the same cost from either slot
                combo = (combo << 1) | (lab[node] != fus->snap[node] ? 1 : 0);
            }
            tb->val[i] = fus->val[(size_t)W * i + combo];
        }
        return tb;
    }

    std::unique_ptr<Table> triplet_table(int label) {
        const double t0 = omp_get_wtime();
        std::unique_ptr<Table> tb(new Table());
        const int* lab = model_->getLabeling();
        const int n_nodes = S_ * ncp_;
        tb->snap.assign(lab, lab + n_nodes);
        tb->label = label;
        tb->width = label < 0 ? 1 : 8;
        tb->val.resize((size_t)tb->width * T_);
        const msmgpu_reg_params prm = reg_params();
        msmgpu_ctx* ctx = dev_[0].ctx;
        if (label < 0) {
            std::vector<int32_t> rt(T_), la(T_), lb(T_), lc(T_);
            for (int t = 0; t < T_; ++t) { rt[t] = t; la[t] = lab[trip_[3 * (size_t)t]]; lb[t] = lab[trip_[3 * (size_t)t + 1]]; lc[t] = lab[trip_[3 * (size_t)t + 2]]; }
            detail::check(msmgpu_group_triplet_costs(ctx, n_nodes, cps_.data(), orig_.data(), rot_.data(), L_, labels_.data(), T_, trip_.data(), &prm,
                                                     cf_->subcorr, cf_->fixnan ? 1 : 0, T_, rt.data(), la.data(), lb.data(), lc.data(), tb->val.data()));
            group_timers().triplet_costs += T_;
        } else {   // the iteration's arrays are resident on every device (get_patch_data); triplet blocks sharded like the pair blocks
            const int n = (int)dev_.size();
            on_devices([&](int d) {
                int b, e;
                shard(T_, d, n, b, e);
                if (e > b)
                    detail::check(msmgpu_triplet_plan_batch(dev_[d].plan, &prm, cf_->subcorr, cf_->fixnan ? 1 : 0, b, e - b, tb->snap.data(), label,
                                                            tb->val.data() + 8 * (size_t)b));
            });
            group_timers().triplet_costs += 8LL * T_;
        }
        group_timers().triplet_batches += omp_get_wtime() - t0;
        group_timers().n_triplet_batches++;
        return tb;
    }

    static bool fresh2(const Table* tb, const int* lab, int a, int b, int label) {
        return tb && tb->label == label && tb->snap[a] == lab[a] && tb->snap[b] == lab[b];
    }
    static bool fresh3(const Table* tb, const int* lab, int a, int b, int c, int label) {
        return tb && tb->label == label && tb->snap[a] == lab[a] && tb->snap[b] == lab[b] && tb->snap[c] == lab[c];
    }

public:
    static GroupBinding& instance() { static GroupBinding g; return g; }

    bool active_for(const CostFn* cf) const { return active_ && cf == cf_; }

    // DiscreteGroupModel.cpp:37-55: for every control point of subject A, the closest control point of every later subject B.
    // One forest over the S control grids, then one batched query per subject B with the control points of all A < B.
    bool estimate_pairs(Model& m) {
        const double t0 = omp_get_wtime();
        ensure_devices();
        msmgpu_ctx* ctx = dev_[0].ctx;
        const int S = m.m_num_subjects, ncp = m.control_grid_size;
        std::vector<msmgpu_mesh*> meshes(S, nullptr);
        std::vector<msmgpu_octree*> trees(S, nullptr);
        std::vector<std::vector<double>> cp(S);
        struct Cleanup {
            std::vector<msmgpu_mesh*>& m; std::vector<msmgpu_octree*>& t;
            ~Cleanup() { for (auto* x : t) if (x) msmgpu_octree_destroy(x); for (auto* x : m) if (x) msmgpu_mesh_destroy(x); }
        } cleanup{meshes, trees};
        const std::vector<int32_t> tri = detail::triangles_of(m.m_controlmeshes[0]);
        for (int s = 0; s < S; ++s) {
            cp[s] = detail::coords_of(m.m_controlmeshes[s]);
            detail::check(msmgpu_mesh_create(ctx, ncp, cp[s].data(), (int)tri.size() / 3, tri.data(), &meshes[s]));
        }
        detail::check(msmgpu_octree_build_batch(ctx, S, meshes.data(), trees.data()));
        std::vector<std::vector<int32_t>> closest(S);   // closest[b][a * ncp + v], a < b
        std::vector<double> q;
        for (int b = 1; b < S; ++b) {
            q.resize(3 * (size_t)b * ncp);
            for (int a = 0; a < b; ++a) std::copy(cp[a].begin(), cp[a].end(), q.begin() + 3 * (size_t)a * ncp);
            closest[b].resize((size_t)b * ncp);
            detail::check(msmgpu_nearest_triangle(trees[b], b * ncp, q.data(), nullptr, closest[b].data(), nullptr));
        }
        int pair = 0;
        for (int a = 0; a < S; ++a)
            for (int v = 0; v < ncp; ++v)
                for (int b = a + 1; b < S; ++b) {
                    m.pairs[2 * pair] = a * ncp + v;
                    m.pairs[2 * pair + 1] = b * ncp + closest[b][(size_t)a * ncp + v];
                    ++pair;
                }
        group_timers().pairs += omp_get_wtime() - t0;
        return true;
    }

    // DiscreteGroupModel.cpp:88-121 + set_patch_data: the per-(subject, label) resampled fields and the patch geometry stay on
    // the device(s); the reference's patch maps are not built.
    bool get_patch_data(Model& m) {
        drop_iteration_state();
        auto* cf = dynamic_cast<CostFn*>(m.costfct.get());
        if (!cf || m.m_num_subjects < 2) return false;
        const double t0 = omp_get_wtime();
        ensure_devices();
        model_ = &m; cf_ = cf;
        S_ = m.m_num_subjects; ncp_ = m.control_grid_size; L_ = m.m_num_labels; P_ = m.m_num_pairs; T_ = m.m_num_triplets;
        const Mesh& first = m.m_datameshes[0];
        const int nv = first.nvertices(), nt = first.ntriangles();
        const std::vector<int32_t> tri = detail::triangles_of(first);
        D_ = (int)m.FEAT->get_dim();
        std::vector<double> xyz(3 * (size_t)S_ * nv), feat((size_t)S_ * D_ * nv);
        #pragma omp parallel for
        for (int s = 0; s < S_; ++s) {
            const Mesh& M = m.m_datameshes[s];
            if (M.nvertices() != nv || M.ntriangles() != nt) continue;   // checked below
            for (int i = 0; i < nv; ++i) {
                const Point& p = M.get_coord(i);
                double* o = &xyz[3 * ((size_t)s * nv + i)];
                o[0] = p.X; o[1] = p.Y; o[2] = p.Z;
            }
            const NEWMAT::Matrix F = m.FEAT->get_data_matrix(s);   // D x nv, 1-based
            for (int d = 0; d < D_; ++d)
                for (int i = 0; i < nv; ++i) feat[((size_t)s * D_ + d) * nv + i] = F(d + 1, i + 1);
        }
        for (int s = 0; s < S_; ++s)
            if (m.m_datameshes[s].nvertices() != nv || m.m_datameshes[s].ntriangles() != nt) return false;
        labels_.resize(3 * (size_t)L_);
        for (int l = 0; l < L_; ++l) { labels_[3 * l] = m.m_labels[l].X; labels_[3 * l + 1] = m.m_labels[l].Y; labels_[3 * l + 2] = m.m_labels[l].Z; }
        const int n_nodes = S_ * ncp_;
        rot_.resize(9 * (size_t)n_nodes);
        for (int k = 0; k < n_nodes; ++k)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) rot_[9 * (size_t)k + 3 * r + c] = m.m_ROT[k](r + 1, c + 1);
        std::vector<double> spacing((size_t)n_nodes);
        for (int s = 0; s < S_; ++s)
            for (int v = 0; v < ncp_; ++v) spacing[(size_t)s * ncp_ + v] = m.spacings[s](v + 1);
        const double centre[3] = {m.centre.X, m.centre.Y, m.centre.Z};
        const Mesh& T = m.target_space;
        const std::vector<double> txyz = detail::coords_of(T);
        const std::vector<int32_t> ttri = detail::triangles_of(T);
        std::vector<double> tarea((size_t)T.ntriangles());
        for (int t = 0; t < T.ntriangles(); ++t) tarea[t] = T.get_triangle_area(t);   // the cached values this object holds (triangle.cpp:31,39)
        const int n_tpl = T.nvertices(), n = (int)dev_.size();
        const size_t per_subject = (size_t)L_ * n_tpl * D_, bytes = (size_t)S_ * per_subject * sizeof(double);
        on_devices([&](int d) {
            Device& dv = dev_[d];
            detail::check(msmgpu_mesh_create(dv.ctx, n_tpl, txyz.data(), T.ntriangles(), ttri.data(), &dv.tpl));
            detail::check(msmgpu_mesh_set_triangle_areas(dv.tpl, tarea.data()));
            detail::check(msmgpu_octree_build(dv.tpl, &dv.tpl_tree));
            if (dv.fields_bytes < bytes) {
                msmgpu_device_free(dv.ctx, dv.fields);
                dv.fields = nullptr; dv.fields_bytes = 0;
                void* p = nullptr;
                detail::check(msmgpu_device_malloc(dv.ctx, bytes, &p));
                dv.fields = static_cast<double*>(p); dv.fields_bytes = bytes;
            }
            int b, e;
            shard(S_, d, n, b, e);
            if (e > b)
                detail::check(msmgpu_group_fields(dv.ctx, e - b, nv, xyz.data() + 3 * (size_t)b * nv, nt, tri.data(), D_, feat.data() + (size_t)b * D_ * nv, L_,
                                                  labels_.data(), centre, dv.tpl, dv.tpl_tree, dv.fields + (size_t)b * per_subject));
        });
        // every device needs every subject's fields to evaluate any pair: one exchange per iteration
        for (int dst = 0; dst < n; ++dst)
            for (int src = 0; src < n; ++src) {
                if (src == dst) continue;
                int b, e;
                shard(S_, src, n, b, e);
                if (e > b)
                    detail::check(msmgpu_device_copy_peer(dev_[dst].ctx, dev_[dst].fields + (size_t)b * per_subject, dev_[src].ctx,
                                                          dev_[src].fields + (size_t)b * per_subject, (size_t)(e - b) * per_subject * sizeof(double)));
            }
        std::vector<double> mask;
        if (cf->is_masked) {   // set_masks (DiscreteGroupCostFunction.h:50): channel 0 of the mask mesh, one value per template vertex
            if (cf->_MASK.nvertices() != n_tpl) throw MeshregException("Group cost mask differs in nvertices from the template");
            mask.resize((size_t)n_tpl);
            for (int p = 0; p < n_tpl; ++p) mask[p] = cf->_MASK.get_pvalue(p);
        }
        on_devices([&](int d) {
            Device& dv = dev_[d];
            detail::check(msmgpu_group_create(dv.ctx, cf->_simmeasure, S_, ncp_, L_, D_, dv.tpl, dv.fields, rot_.data(), labels_.data(), spacing.data(), m.range,
                                              &dv.group));
            if (cf->is_masked) detail::check(msmgpu_group_set_mask(dv.group, mask.data()));
        });
        pairs_.assign(m.pairs, m.pairs + 2 * (size_t)P_);
        on_devices([&](int d) {
            Device& dv = dev_[d];
            detail::check(msmgpu_group_set_pairs(dv.group, P_, pairs_.data()));
            int b, e;
            shard(P_, d, n, b, e);
            const size_t need = 4 * (size_t)std::max(e - b, 1) * sizeof(double);
            if (dv.batch_bytes < need) {
                msmgpu_device_free(dv.ctx, dv.batch_out);
                dv.batch_out = nullptr; dv.batch_bytes = 0;
                void* p = nullptr;
                detail::check(msmgpu_device_malloc(dv.ctx, need, &p));
                dv.batch_out = static_cast<double*>(p); dv.batch_bytes = need;
            }
        });
        trip_.assign(m.triplets, m.triplets + 3 * (size_t)T_);
        cps_.resize(3 * (size_t)n_nodes); orig_.resize(3 * (size_t)n_nodes);
        for (int s = 0; s < S_; ++s)
            for (int v = 0; v < ncp_; ++v) {
                const Point &c = cf->_CONTROLMESHES[s].get_coord(v), &o = cf->_ORIG_MESHES[s].get_coord(v);
                double* pc = &cps_[3 * ((size_t)s * ncp_ + v)];
                double* po = &orig_[3 * ((size_t)s * ncp_ + v)];
                pc[0] = c.X; pc[1] = c.Y; pc[2] = c.Z;
                po[0] = o.X; po[1] = o.Y; po[2] = o.Z;
            }
        on_devices([&](int d) {
            detail::check(msmgpu_triplet_plan_create(dev_[d].ctx, n_nodes, cps_.data(), orig_.data(), rot_.data(), L_, labels_.data(), T_, trip_.data(),
                                                     &dev_[d].plan));
        });
        active_ = true;
        group_timers().fields += omp_get_wtime() - t0;
        group_timers().n_iterations++;
        return true;
    }

    double pairwise(CostFn& cf, int pair, int labelA, int labelB) {
        const int* lab = model_->getLabeling();
        const int a = pairs_[2 * (size_t)pair], b = pairs_[2 * (size_t)pair + 1];
        const bool da = labelA != lab[a], db = labelB != lab[b];
        if (!da && !db) {
            const Table* tb = pair_cur_.cur.load(std::memory_order_acquire);
            if (!fresh2(tb, lab, a, b, -1)) {
                std::lock_guard<std::mutex> g(mu_);
                tb = pair_cur_.cur.load(std::memory_order_acquire);
                if (!fresh2(tb, lab, a, b, -1)) {
                    std::unique_ptr<Table> t = carry_over<2>(pair_fus_.cur.load(), lab, S_ * ncp_, pairs_, P_);
                    if (!t) t = pair_table(-1);
                    pair_cur_.publish(std::move(t));
                    tb = pair_cur_.cur.load();
                }
            }
            return tb->val[pair];
        }
        const int label = da ? labelA : labelB;
        if (da && db && labelA != labelB) {   // not one of Fusion's combinations: evaluate this request on its own
            const int32_t rp = pair, la = labelA, lb = labelB;
            double out = 0;
            std::lock_guard<std::mutex> g(mu_);
            detail::check(msmgpu_group_pair_costs(dev_[0].group, P_, pairs_.data(), 1, &rp, &la, &lb, &out));
            if (cf.fixnan && std::isnan(out)) out = FIX_NAN;
            return out;
        }
        const Table* tb = pair_fus_.cur.load(std::memory_order_acquire);
        if (!fresh2(tb, lab, a, b, label)) {
            std::lock_guard<std::mutex> g(mu_);
            tb = pair_fus_.cur.load(std::memory_order_acquire);
            if (!fresh2(tb, lab, a, b, label)) { pair_fus_.publish(pair_table(label)); tb = pair_fus_.cur.load(); }
        }
        return tb->val[4 * (size_t)pair + ((da ? 2 : 0) | (db ? 1 : 0))];
    }

    double triplet(CostFn& cf, int triplet, int labelA, int labelB, int labelC) {
        (void)cf;
        const int* lab = model_->getLabeling();
        const int a = trip_[3 * (size_t)triplet], b = trip_[3 * (size_t)triplet + 1], c = trip_[3 * (size_t)triplet + 2];
        const bool da = labelA != lab[a], db = labelB != lab[b], dc = labelC != lab[c];
        if (!da && !db && !dc) {
            const Table* tb = trip_cur_.cur.load(std::memory_order_acquire);
            if (!fresh3(tb, lab, a, b, c, -1)) {
                std::lock_guard<std::mutex> g(mu_);
                tb = trip_cur_.cur.load(std::memory_order_acquire);
                if (!fresh3(tb, lab, a, b, c, -1)) {
                    std::unique_ptr<Table> t = carry_over<3>(trip_fus_.cur.load(), lab, S_ * ncp_, trip_, T_);
                    if (!t) t = triplet_table(-1);
                    trip_cur_.publish(std::move(t));
                    tb = trip_cur_.cur.load();
                }
            }
            return tb->val[triplet];
        }
        const int label = da ? labelA : (db ? labelB : labelC);
        if ((da && labelA != label) || (db && labelB != label) || (dc && labelC != label)) {
            const msmgpu_reg_params prm = reg_params();
            const int32_t rt = triplet, la = labelA, lb = labelB, lc = labelC;
            double out = 0;
            std::lock_guard<std::mutex> g(mu_);
            detail::check(msmgpu_group_triplet_costs(dev_[0].ctx, S_ * ncp_, cps_.data(), orig_.data(), rot_.data(), L_, labels_.data(), T_, trip_.data(), &prm,
                                                     cf_->subcorr, cf_->fixnan ? 1 : 0, 1, &rt, &la, &lb, &lc, &out));
            return out;
        }
        const Table* tb = trip_fus_.cur.load(std::memory_order_acquire);
        if (!fresh3(tb, lab, a, b, c, label)) {
            std::lock_guard<std::mutex> g(mu_);
            tb = trip_fus_.cur.load(std::memory_order_acquire);
            if (!fresh3(tb, lab, a, b, c, label)) { trip_fus_.publish(triplet_table(label)); tb = trip_fus_.cur.load(); }
        }
        return tb->val[8 * (size_t)triplet + ((da ? 4 : 0) | (db ? 2 : 0) | (dc ? 1 : 0))];
    }
};

}  // namespace newmeshreg_gpu
