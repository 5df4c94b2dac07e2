// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// extern "C" driver around the UNMODIFIED reference registration library, compiled in place from
// /root/reference/libraries/msm-newmeshreg/src/*.cpp by oracle/Makefile into
// oracle/_ref/libref_newmeshreg.so (FSL replaced by the header shim in oracle/shim).
// It instantiates the reference's own cost-function classes (DiscreteCostFunction.h:104-244,
// DiscreteGroupCostFunction.h:30-66, DiscreteGroupModel.h:30-91), feeds them plain arrays and returns
// what they compute, so that oracle/msm_oracle.cpp (the CPU restatement) and the CUDA path can be pinned
// against the reference itself: patch membership, unary cost tables, HO patches, triplet costs, groupwise
// patch data and pair costs.
// Built with -fno-access-control: the reference wires these objects together inside
// NonLinearSRegDiscreteModel/Mesh_registration; here the same protected members are set directly.
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "DiscreteGroupModel.h"
#include "DiscreteGroupCostFunction.h"
#include "DiscreteModel.h"
#include "rigid_costfunction.h"

using namespace newmeshreg;
using newresampler::Mesh;
using newresampler::Mpoint;
using newresampler::Octree;
using newresampler::Point;
using newresampler::Triangle;

namespace {

Mesh make_mesh(int nv, const double* xyz, int nt, const int* tri) {
    Mesh tmp;
    for (int i = 0; i < nv; ++i) tmp.push_point(std::make_shared<Mpoint>(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], i));
    auto& pts = tmp.get_all_points();
    for (int t = 0; t < nt; ++t) {
        Triangle tr(pts[tri[3 * t]], pts[tri[3 * t + 1]], pts[tri[3 * t + 2]], t);
        tmp.push_triangle(tr);
    }
    tmp.initialize_pvalues(1);
    return Mesh(tmp);  // copy: triangles re-created, cached areas refreshed (mesh.cpp:37-53)
}

NEWMAT::Matrix to_matrix(int rows, int cols, const double* a) {
    NEWMAT::Matrix m(rows, cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) m(r + 1, c + 1) = a[(size_t)r * cols + c];
    return m;
}

std::vector<NEWMAT::Matrix> to_rotations(int n, const double* rot) {
    std::vector<NEWMAT::Matrix> R(n);
    for (int k = 0; k < n; ++k) R[k] = to_matrix(3, 3, rot + 9 * (size_t)k);
    return R;
}

std::vector<Point> to_points(int n, const double* p) {
    std::vector<Point> v(n);
    for (int i = 0; i < n; ++i) v[i] = Point(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}

std::shared_ptr<featurespace> make_feat(int D, int nsrc, const double* src_feat, int nref, const double* ref_feat) {
    auto F = std::make_shared<featurespace>(std::string("in"), std::string("ref"));
    F->DATA.clear();
    F->DATA.push_back(std::make_shared<MISCMATHS::FullBFMatrix>(to_matrix(D, nsrc, src_feat)));
    F->DATA.push_back(std::make_shared<MISCMATHS::FullBFMatrix>(to_matrix(D, nref, ref_feat)));
    return F;
}

std::shared_ptr<NonLinearSRegDiscreteCostFunction> make_costfn(int kind) {
    switch (kind) {
        case 0: return std::make_shared<UnivariateNonLinearSRegDiscreteCostFunction>();
        case 1: return std::make_shared<MultivariateNonLinearSRegDiscreteCostFunction>();
        case 2: return std::make_shared<PatchwiseMultivariateNonLinearSRegDiscreteCostFunction>();
        case 3: return std::make_shared<HOUnivariateNonLinearSRegDiscreteCostFunction>();
        case 4: return std::make_shared<HOMultivariateNonLinearSRegDiscreteCostFunction>();
    }
    return nullptr;
}

int flatten_lists(const std::vector<std::vector<int>>& lists, int* rowptr, int* members, int cap) {
    int pos = 0;
    for (size_t k = 0; k < lists.size(); ++k) {
        if (rowptr) rowptr[k] = pos;
        for (int v : lists[k]) { if (members && pos < cap) members[pos] = v; ++pos; }
    }
    if (rowptr) rowptr[lists.size()] = pos;
    return pos;
}

double g_percentile = 0.75;   // refmr_set_percentile -> sparsesimkernel::set_percentile (similarities.h:40)

struct Setup {
    std::shared_ptr<NonLinearSRegDiscreteCostFunction> cf;
    std::shared_ptr<Octree> tree;
};

// The wiring NonLinearSRegDiscreteModel::Initialize / setupCostFunction perform (DiscreteModel.cpp:60-262).
Setup wire(int kind, int simmeasure, int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri,
           int ncp, const double* cp_xyz, int ncp_tri, const int* cp_tri, const double* rot, int L, const double* labels,
           int nsrc, const double* src_xyz, int nsrc_tri, const int* src_tri, int D, const double* src_feat, const double* ref_feat,
           int cfw_rows, const double* cfw, const double* maxsep, double range, int nthreads) {
    Setup s;
    s.cf = make_costfn(kind);
    auto& cf = *s.cf;
    Mesh target = make_mesh(nv_t, tgt_xyz, nt_t, tgt_tri);
    Mesh source = make_mesh(nsrc, src_xyz, nsrc_tri, src_tri);
    Mesh grid = make_mesh(ncp, cp_xyz, ncp_tri, cp_tri);
    cf.set_meshes(target, source, grid, 0);
    cf.set_featurespace(make_feat(D, nsrc, src_feat, nv_t, ref_feat));
    cf._simmeasure = simmeasure;
    cf.sim.set_simval(simmeasure);
    cf.sim.set_percentile(g_percentile);
    cf._controlptrange = range;
    cf._threads = nthreads;
    NEWMAT::ColumnVector sep(ncp);
    for (int k = 0; k < ncp; ++k) sep(k + 1) = maxsep ? maxsep[k] : 0.0;
    cf.set_spacings(sep, 0.0);
    s.tree = std::make_shared<Octree>(target);
    cf.set_octrees(s.tree);
    cf.set_labels(to_points(L, labels), to_rotations(ncp, rot));
    if (cfw_rows > 0 && cfw) cf.set_dataaffintyweighting(to_matrix(cfw_rows, nsrc, cfw));
    else { NEWMAT::Matrix one(1, nsrc); one = 1.0; cf.set_dataaffintyweighting(one); }
    return s;
}

}  // namespace

extern "C" {

void refmr_set_percentile(double p) { g_percentile = p; }

// get_source_data() + computeUnaryCosts() of kinds 0 (Univariate, cpp:326-383), 1 (Multivariate, 385-458),
// 2 (Patchwise, 620-692). out[L][ncp] label-major like unarycosts (cpp:242). absw_in != NULL replaces the
// AbsoluteWeights that resample_weights() (cpp:303) produced; absw_out (optional) returns the ones used.
// Returns the number of patch entries (CSR written up to cap), or -1 on a reference exception.
int refmr_unary(int kind, int simmeasure, int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri,
                int ncp, const double* cp_xyz, int ncp_tri, const int* cp_tri, const double* rot, int L, const double* labels,
                int nsrc, const double* src_xyz, int nsrc_tri, const int* src_tri, int D, const double* src_feat, const double* ref_feat,
                int cfw_rows, const double* cfw, const double* absw_in, const double* maxsep, double range,
                double* out, int* patch_rowptr, int* patch_members, int cap, double* absw_out, int nthreads) {
    try {
        Setup s = wire(kind, simmeasure, nv_t, tgt_xyz, nt_t, tgt_tri, ncp, cp_xyz, ncp_tri, cp_tri, rot, L, labels,
                       nsrc, src_xyz, nsrc_tri, src_tri, D, src_feat, ref_feat, cfw_rows, cfw, maxsep, range, nthreads);
        auto& cf = *s.cf;
        cf.initialize(ncp, L, 0, 0);
        cf.get_source_data();
        if (absw_in) for (int k = 0; k < ncp; ++k) cf.AbsoluteWeights(k + 1) = absw_in[k];
        if (absw_out) for (int k = 0; k < ncp; ++k) absw_out[k] = cf.AbsoluteWeights(k + 1);
        int total = flatten_lists(cf._sourceinrange, patch_rowptr, patch_members, cap);
        if (out) {
            cf.computeUnaryCosts();
            std::memcpy(out, cf.getUnaryCosts(), sizeof(double) * (size_t)L * ncp);
        }
        return total;
    } catch (...) { return -1; }
}

// computeTripletCost (cpp:135-188) for n requests on any kind 0..4; HO kinds add triplet_likelihood
// (cpp:487-531, 565-618). rmode 2/3 only (spherical strain, reg_tools.cpp:551-743). orig_cp_xyz supplies the
// coordinates _ORIG.get_coord(node) reads for the undeformed triangle. Returns HO patch entries (kinds 3/4; CSR
// over CP-grid triangles) or 0, -1 on a reference exception.
// the anatomical meshes and maps of regoption 4/5 (set_anatomical / set_anatomical_neighbourhood, DiscreteCostFunction.h:164-169);
// same layout as orc_anat (oracle/msm_oracle.h) and msmgpu_anatomical (include/msmgpu.h)
struct refmr_anat {
    int n_av; const double* asource_xyz; int n_at; const int* asource_tri;
    int n_hv; const double* thi_xyz; int n_ht; const int* thi_tri;
    const double* atarget_xyz;
    const int* face_ptr; const int* face_ids;
    const int* bary_ptr; const int* bary_key; const double* bary_w;
};
int refmr_triplet_anat(int kind, int simmeasure, int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri,
                  int ncp, const double* cp_xyz, int ncp_tri, const int* cp_tri, const double* orig_cp_xyz,
                  const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                  int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                  int nsrc, const double* src_xyz, int nsrc_tri, const int* src_tri, int D, const double* src_feat, const double* ref_feat,
                  int cfw_rows, const double* cfw, const double* absw_in,
                  double lambda, double mu, double kappa, double k_exp, double rexp, int rmode, const refmr_anat* anat,
                  double* out, int* patch_rowptr, int* patch_members, int cap, int nthreads);
int refmr_triplet(int kind, int simmeasure, int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri,
                  int ncp, const double* cp_xyz, int ncp_tri, const int* cp_tri, const double* orig_cp_xyz,
                  const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                  int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                  int nsrc, const double* src_xyz, int nsrc_tri, const int* src_tri, int D, const double* src_feat, const double* ref_feat,
                  int cfw_rows, const double* cfw, const double* absw_in,
                  double lambda, double mu, double kappa, double k_exp, double rexp, int rmode,
                  double* out, int* patch_rowptr, int* patch_members, int cap, int nthreads) {
    return refmr_triplet_anat(kind, simmeasure, nv_t, tgt_xyz, nt_t, tgt_tri, ncp, cp_xyz, ncp_tri, cp_tri, orig_cp_xyz, rot, L, labels, ntrip, triplets,
                              n, req_triplet, req_la, req_lb, req_lc, nsrc, src_xyz, nsrc_tri, src_tri, D, src_feat, ref_feat, cfw_rows, cfw, absw_in,
                              lambda, mu, kappa, k_exp, rexp, rmode, nullptr, out, patch_rowptr, patch_members, cap, nthreads);
}

int refmr_triplet_anat(int kind, int simmeasure, int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri,
                  int ncp, const double* cp_xyz, int ncp_tri, const int* cp_tri, const double* orig_cp_xyz,
                  const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                  int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                  int nsrc, const double* src_xyz, int nsrc_tri, const int* src_tri, int D, const double* src_feat, const double* ref_feat,
                  int cfw_rows, const double* cfw, const double* absw_in,
                  double lambda, double mu, double kappa, double k_exp, double rexp, int rmode, const refmr_anat* anat,
                  double* out, int* patch_rowptr, int* patch_members, int cap, int nthreads) {
    try {
        std::vector<double> sep(ncp, 0.0);
        Setup s = wire(kind, simmeasure, nv_t, tgt_xyz, nt_t, tgt_tri, ncp, cp_xyz, ncp_tri, cp_tri, rot, L, labels,
                       nsrc, src_xyz, nsrc_tri, src_tri, D, src_feat, ref_feat, cfw_rows, cfw, sep.data(), 1.0, nthreads);
        auto& cf = *s.cf;
        cf._reglambda = lambda; cf._mu = mu; cf._kappa = kappa; cf._k_exp = k_exp; cf._rexp = rexp; cf._rmode = rmode;
        cf._ORIG = make_mesh(ncp, orig_cp_xyz, ncp_tri, cp_tri);
        std::vector<int> trip(triplets, triplets + 3 * (size_t)ntrip);
        cf.initialize(ncp, L, 0, ntrip);
        int total = 0;
        if (kind >= 3) {
            cf.get_source_data();
            if (absw_in) for (int k = 0; k < ncp; ++k) cf.AbsoluteWeights(k + 1) = absw_in[k];
            total = flatten_lists(cf._sourceinrange, patch_rowptr, patch_members, cap);
        }
        cf.setTriplets(trip.data());
        if (anat) {   // what Mesh_registration::run_discrete_opt hands over before the optimisation (mesh_registration.cpp:93-98)
            Mesh thi = make_mesh(anat->n_hv, anat->thi_xyz, anat->n_ht, anat->thi_tri);
            cf.set_anatomical(thi, make_mesh(anat->n_hv, anat->atarget_xyz, anat->n_ht, anat->thi_tri), thi,
                              make_mesh(anat->n_av, anat->asource_xyz, anat->n_at, anat->asource_tri));
            std::vector<std::map<int, double>> bw(anat->n_av);
            for (int v = 0; v < anat->n_av; ++v)
                for (int e = anat->bary_ptr[v]; e < anat->bary_ptr[v + 1]; ++e) bw[v][anat->bary_key[e]] = anat->bary_w[e];
            std::vector<std::vector<int>> nf(ntrip);
            for (int t = 0; t < ntrip; ++t) nf[t].assign(anat->face_ids + anat->face_ptr[t], anat->face_ids + anat->face_ptr[t + 1]);
            cf.set_anatomical_neighbourhood(bw, nf);
            cf.initialize_regulariser();
        }
        #pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) out[i] = cf.computeTripletCost(req_triplet[i], req_la[i], req_lb[i], req_lc[i]);
        return total;
    } catch (...) { return -1; }
}

// NonLinearSRegDiscreteCostFunction::computePairwiseCost (cpp:190-233): the rotation-difference regulariser
// of the pairwise (regoption 1 / FastPD) model. out[n].
int refmr_pairwise_reg(int ncp, const double* cp_xyz, int ncp_tri, const int* cp_tri, const double* rot, int L, const double* labels,
                       int npairs, const int* pairs, double lambda, double rexp, double mvdmax,
                       int n, const int* req_pair, const int* req_la, const int* req_lb, double* out) {
    try {
        UnivariateNonLinearSRegDiscreteCostFunction cf;
        Mesh grid = make_mesh(ncp, cp_xyz, ncp_tri, cp_tri);
        cf._CPgrid = grid; cf._oCPgrid = grid;
        cf.set_labels(to_points(L, labels), to_rotations(ncp, rot));
        cf._reglambda = lambda; cf._rexp = rexp; cf.MVDmax = mvdmax;
        std::vector<int> pr(pairs, pairs + 2 * (size_t)npairs);
        cf.setPairs(pr.data());
        for (int i = 0; i < n; ++i) out[i] = cf.computePairwiseCost(req_pair[i], req_la[i], req_lb[i]);
        return 0;
    } catch (...) { return -1; }
}

// gMSM: DiscreteGroupModel::get_patch_data (DiscreteGroupModel.cpp:88-121) followed by
// DiscreteGroupCostFunction::computePairwiseCost (DiscreteGroupCostFunction.cpp:54-97) for n requests.
// data_xyz [S][nv][3] (one data mesh per subject, shared faces), feat [S][D][nv], rot [S*ncp][9],
// spacings [S][ncp], pairs [P][2] global node ids. fields_out (optional) [S][L][D][n_tpl] receives the resampled
// values of every template vertex that is a member of at least one patch (others stay NaN-free: untouched).
int refmr_group_pair_costs_masked(int simmeasure, int S, int nv, const double* data_xyz, int nt, const int* tri, int D, const double* feat,
                                  int L, const double* labels, const double* centre, int n_tpl, const double* tpl_xyz, int nt_tpl, const int* tpl_tri,
                                  int ncp, const double* rot, const double* spacings, double range, int P, const int* pairs,
                                  int n, const int* req_pair, const int* req_la, const int* req_lb, const double* mask, double* out,
                                  double* fields_out, int nthreads);
int refmr_group_pair_costs(int simmeasure, int S, int nv, const double* data_xyz, int nt, const int* tri, int D, const double* feat,
                           int L, const double* labels, const double* centre, int n_tpl, const double* tpl_xyz, int nt_tpl, const int* tpl_tri,
                           int ncp, const double* rot, const double* spacings, double range, int P, const int* pairs,
                           int n, const int* req_pair, const int* req_la, const int* req_lb, double* out, double* fields_out, int nthreads) {
    return refmr_group_pair_costs_masked(simmeasure, S, nv, data_xyz, nt, tri, D, feat, L, labels, centre, n_tpl, tpl_xyz, nt_tpl, tpl_tri, ncp, rot,
                                         spacings, range, P, pairs, n, req_pair, req_la, req_lb, nullptr, out, fields_out, nthreads);
}

// the same with a cost mask installed like DiscreteGroupModel::Initialize does (`costfct->set_masks(mask)`, DiscreteGroupModel.cpp:164):
// mask [n_tpl] = channel 0 of a mask mesh on the template
int refmr_group_pair_costs_masked(int simmeasure, int S, int nv, const double* data_xyz, int nt, const int* tri, int D, const double* feat,
                                  int L, const double* labels, const double* centre, int n_tpl, const double* tpl_xyz, int nt_tpl, const int* tpl_tri,
                                  int ncp, const double* rot, const double* spacings, double range, int P, const int* pairs,
                                  int n, const int* req_pair, const int* req_la, const int* req_lb, const double* mask, double* out,
                                  double* fields_out, int nthreads) {
    try {
        DiscreteGroupModel model;
        model._nthreads = nthreads;
        model.costfct = std::make_shared<DiscreteGroupCostFunction>();
        model.costfct->_simmeasure = simmeasure;
        model.costfct->sim.set_simval(simmeasure);
        model.m_num_subjects = S;
        model.control_grid_size = ncp;
        model.m_num_labels = L;
        model.range = range;
        model.centre = Point(centre[0], centre[1], centre[2]);
        model.m_labels = to_points(L, labels);
        model.m_ROT = to_rotations(S * ncp, rot);
        model.target_space = make_mesh(n_tpl, tpl_xyz, nt_tpl, tpl_tri);
        model.m_datameshes.clear();
        auto F = std::make_shared<featurespace>(std::string("in"), std::string("ref"));
        F->DATA.clear();
        for (int s = 0; s < S; ++s) {
            model.m_datameshes.push_back(make_mesh(nv, data_xyz + (size_t)s * nv * 3, nt, tri));
            F->DATA.push_back(std::make_shared<MISCMATHS::FullBFMatrix>(to_matrix(D, nv, feat + (size_t)s * D * nv)));
            NEWMAT::ColumnVector sp(ncp);
            for (int k = 0; k < ncp; ++k) sp(k + 1) = spacings[(size_t)s * ncp + k];
            model.spacings.push_back(sp);
        }
        model.FEAT = F;
        model.get_patch_data();
        auto cf = std::dynamic_pointer_cast<DiscreteGroupCostFunction>(model.costfct);
        cf->VERTICES_PER_SUBJ = ncp;
        cf->m_num_labels = L;
        std::vector<int> pr(pairs, pairs + 2 * (size_t)P);
        cf->setPairs(pr.data());
        if (mask) {
            Mesh mm = model.target_space;
            mm.initialize_pvalues(1);
            for (int p = 0; p < n_tpl; ++p) mm.set_pvalue(p, mask[p]);
            model.costfct->set_masks(mm);
        }
        if (fields_out)
            for (int s = 0; s < S; ++s)
                for (int k = 0; k < ncp; ++k)
                    for (int l = 0; l < L; ++l)
                        for (const auto& e : cf->patch_data[(size_t)s * ncp * L + (size_t)k * L + l])
                            for (int d = 0; d < D; ++d)
                                fields_out[(((size_t)s * L + l) * D + d) * n_tpl + e.first] = e.second[d];
        #pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
        for (int i = 0; i < n; ++i) out[i] = cf->computePairwiseCost(req_pair[i], req_la[i], req_lb[i]);
        return 0;
    } catch (...) { return -1; }
}

// DiscreteGroupCostFunction::computeTripletCost (DiscreteGroupCostFunction.cpp:26-52). cp_xyz/orig_xyz [S][ncp][3].
int refmr_group_triplet_costs(int S, int ncp, const double* cp_xyz, const double* orig_xyz, int ncp_tri, const int* cp_tri,
                              const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                              double lambda, double mu, double kappa, double k_exp, double rexp,
                              int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc, double* out) {
    try {
        DiscreteGroupCostFunction cf;
        std::vector<Mesh> orig;
        for (int s = 0; s < S; ++s) orig.push_back(make_mesh(ncp, orig_xyz + (size_t)s * ncp * 3, ncp_tri, cp_tri));
        Mesh grid0 = make_mesh(ncp, cp_xyz, ncp_tri, cp_tri);
        cf.set_meshes(orig, grid0, S);
        for (int s = 0; s < S; ++s) cf.reset_CPgrid(make_mesh(ncp, cp_xyz + (size_t)s * ncp * 3, ncp_tri, cp_tri), s);
        cf.set_labels(to_points(L, labels), to_rotations(S * ncp, rot));
        cf._reglambda = lambda; cf._mu = mu; cf._kappa = kappa; cf._k_exp = k_exp; cf._rexp = rexp;
        std::vector<int> trip(triplets, triplets + 3 * (size_t)ntrip);
        cf.setTriplets(trip.data());
        for (int i = 0; i < n; ++i) out[i] = cf.computeTripletCost(req_triplet[i], req_la[i], req_lb[i], req_lc[i]);
        return 0;
    } catch (...) { return -1; }
}

// label_sampling_grid (DiscreteModel.cpp:124-190): the label sets of a sampling grid of resolution sgres around
// its first 6-neighbour vertex, for a control grid with the given MaxVD. Returns the number of vertex labels;
// *n_bary the number of barycentre labels; arrays [<=cap][3].
int refmr_label_sets(int sgres, double maxvd, double* samples, double* barycentres, int cap, int* n_bary, double* centre) {
    try {
        NonLinearSRegDiscreteModel m;
        m.m_SGres = sgres;
        m.m_maxs_dist = m._labeldist * maxvd;
        m.Initialize_sampling_grid();
        for (size_t i = 0; i < m.m_samples.size() && (int)i < cap; ++i) {
            samples[3 * i] = m.m_samples[i].X; samples[3 * i + 1] = m.m_samples[i].Y; samples[3 * i + 2] = m.m_samples[i].Z;
        }
        for (size_t i = 0; i < m.m_barycentres.size() && (int)i < cap; ++i) {
            barycentres[3 * i] = m.m_barycentres[i].X; barycentres[3 * i + 1] = m.m_barycentres[i].Y; barycentres[3 * i + 2] = m.m_barycentres[i].Z;
        }
        if (n_bary) *n_bary = (int)m.m_barycentres.size();
        if (centre) { centre[0] = m.centre.X; centre[1] = m.centre.Y; centre[2] = m.centre.Z; }
        return (int)m.m_samples.size();
    } catch (...) { return -1; }
}

// The RIGID / AFFINE level (rigid_costfunction.cpp:32-236, mesh_registration.cpp:68-73, 117-121): initialise() (neighbourhoods within
// 4 mean vertex distances, reg_tools.cpp:31-58, + the sparse similarity columns), the cost at zero rotation, and run(). Outputs: the
// rotated source coordinates [nv_s][3], the neighbour lists after initialise() (CSR, nearest first), the initial cost.
// not accelerated yet; these outputs pin the CPU restatement (msm_oracle.cpp: orc_rigid, tests/golden/rigid.npz).
int refmr_rigid(int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri, int nv_s, const double* src_xyz, int nt_s, const int* src_tri,
                int D, const double* src_feat, const double* ref_feat, int simmeasure, int iters, double stepsize, double gradsampling, int nthreads,
                double* out_xyz, double* out_cost0, int* nbh_rowptr, int* nbh_members, int cap) {
    try {
        Mesh target = make_mesh(nv_t, tgt_xyz, nt_t, tgt_tri);
        Mesh source = make_mesh(nv_s, src_xyz, nt_s, src_tri);
        auto F = make_feat(D, nv_s, src_feat, nv_t, ref_feat);
        Rigid_cost_function rc(target, source, F);
        myparam P;
        P.insert(parameterPair("iters", iters));
        P.insert(parameterPair("simmeasure", simmeasure));
        P.insert(parameterPair("verbosity", false));
        P.insert(parameterPair("stepsize", stepsize));
        P.insert(parameterPair("gradsampling", gradsampling));
        P.insert(parameterPair("numthreads", nthreads));
        rc.set_parameters(P);
        rc.initialise();
        const int n = flatten_lists(rc.nbh->neighbours, nbh_rowptr, nbh_members, cap);
        if (out_cost0) *out_cost0 = rc.rigid_cost_mesh(0.0, 0.0, 0.0);
        Mesh out = rc.run();
        for (int i = 0; i < nv_s; ++i) {
            const Point& p = out.get_coord(i);
            out_xyz[3 * i] = p.X; out_xyz[3 * i + 1] = p.Y; out_xyz[3 * i + 2] = p.Z;
        }
        return n;
    } catch (...) { return -1; }
}

}  // extern "C"

// newmeshreg::variance_normalise (reg_tools.cpp:804-844) on data [D][n] in place; excl NULL or [n] = the first channel of the EXCL mesh
extern "C" int refmr_variance_normalise(int D, int n, double* data, const double* excl, int nthreads) {
    try {
        NEWMAT::Matrix M(D, n);
        for (int d = 0; d < D; ++d)
            for (int i = 0; i < n; ++i) M(d + 1, i + 1) = data[(size_t)d * n + i];
        std::shared_ptr<MISCMATHS::BFMatrix> B = std::make_shared<MISCMATHS::FullBFMatrix>(M);
        std::shared_ptr<Mesh> E;
        if (excl) {
            E = std::make_shared<Mesh>();
            for (int i = 0; i < n; ++i) E->push_point(std::make_shared<newresampler::Mpoint>(0.0, 0.0, 0.0, i));
            E->initialize_pvalues(1);
            for (int i = 0; i < n; ++i) E->set_pvalue(i, excl[i]);
        }
        newmeshreg::variance_normalise(B, E, nthreads);
        for (int d = 0; d < D; ++d)
            for (int i = 0; i < n; ++i) data[(size_t)d * n + i] = B->Peek(d + 1, i + 1);
        return 0;
    } catch (...) { return -1; }
}
