/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library. The product (newmsm_b200/) never links or calls it.
 *
 * Parity status: the reference ships NO tests or golden vectors, so this
 * oracle is pinned against the reference itself: oracle/_ref/libref_newresampler.so (the
 * unmodified reference sources compiled by oracle/Makefile) in tests/test_oracle_vs_ref.py,
 * and against fixtures generated from that build (tests/golden/, tests/golden/make_golden.py).
 *
 * All arrays are plain C: xyz = [n][3] doubles (AoS), tri = [nt][3] int32, features are
 * channel-major [D][V] doubles like the reference's Mesh::pvalues (mesh.h:44).
 */
#ifndef MSM_ORACLE_H
#define MSM_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_octree orc_octree;

/* octree.cpp:31-141 (sequential insertion with the split heuristic) */
orc_octree* orc_octree_build(int nv, const double* xyz, int nt, const int* tri);
void orc_octree_free(orc_octree* t);
/* pre-order dump, children in [i][j][k] nesting: kinds (1 leaf / 0 internal), per-node triangle
 * counts, concatenated triangle ids. Returns node count; *n_tris = total ids. */
int orc_octree_dump(const orc_octree* t, int* kinds, int* counts, int node_cap, int* tris, int tri_cap, int* n_tris);

/* octree.cpp:156-214 / 216-233. status: 0 ok, 1 outside root cube, 2 nothing found.
 * path[i] (optional): 0 = found in leaf, 1 = fallback 1, 2 = fallback 2. */
void orc_octree_query(const orc_octree* t, int n, const double* pts, int* out_tri, int* out_vertex,
                      int* status, int* path, int nthreads);

/* resampler.cpp:142-167 + triangle.cpp:124-143: weights keyed by ascending vertex id.
 * idx/w are [n][3]; n_entries[n] counts distinct ids. Returns 0, or the first non-zero status. */
int orc_bary_weights(const orc_octree* t, int n, const double* pts, int* idx, double* w, int* n_entries, int nthreads);

/* mesh.cpp:1275: mean cached area of adjacent triangles (areas computed from current coords,
 * i.e. the "freshly copied mesh" convention). */
void orc_vertex_areas(int nv, const double* xyz, int nt, const int* tri, double* out);

/* resampler.cpp:72-140 (no exclusion mask), single-thread summation order. CSR out.
 * Returns nnz, or -1 on a failed query. rowptr[n_low+1]; col/val written up to cap. */
/* exclusion masks (resampler.cpp:30-140, 169-258 with EXCL) */
int orc_metric_resample_excl(int nv_in, const double* xyz_in, int nt_in, const int* tri_in, int nv_low, const double* xyz_low, int nt_low,
                             const int* tri_low, int D, const double* feat_in, const double* excl, double* feat_out, double* excl_out,
                             int* rowptr, int* col, double* val, int cap);
int orc_nn_resample_excl(int n, const double* low_xyz, int nv, const double* xyz, int nt, const int* tri, int D, const double* feat_in,
                         const double* excl, double* feat_out, double* excl_out);
int orc_smooth_data(int nv_orig, const double* orig_xyz, int nt_orig, const int* orig_tri, int n, const double* low_xyz, double sigma, int D,
                    const double* feat, const double* excl, double* out, double* excl_out);
int orc_adaptive_weights(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                         int nv_low, const double* xyz_low, int nt_low, const int* tri_low,
                         int* rowptr, int* col, double* val, int cap);

/* resampler.cpp:30-70 / 304: out[d][k] = sum over row k (ascending col) of in[d][col]*val */
int orc_metric_resample(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                        int nv_low, const double* xyz_low, int nt_low, const int* tri_low,
                        int D, const double* feat_in, double* feat_out, int nthreads);

/* plain barycentric resample: Octree(in) + get_barycentric_weights + loop of resampler.cpp:40-52 */
int orc_bary_resample(int nv_in, const double* xyz_in, int nt_in, const int* tri_in,
                      int n_low, const double* xyz_low, int D, const double* feat_in, double* feat_out, int nthreads);

/* resampler.cpp:311-328: new = normalize(sum w * to[idx]) * 100 */
int orc_sphere_project_warp(int n, const double* sphere_xyz, int nv, const double* from_xyz, int nt, const int* tri,
                            const double* to_xyz, double* out_xyz, int nthreads);
/* resampler.cpp:284-302 (surface_resample) — same blend without the re-projection */
int orc_surface_resample(int n, const double* low_xyz, int nv, const double* sph_xyz, int nt, const int* tri,
                         const double* anat_xyz, double* out_xyz, int nthreads);
/* resampler.cpp:232-258 (no exclusion) */
int orc_nn_resample(int n, const double* low_xyz, int nv, const double* xyz, int nt, const int* tri,
                    int D, const double* feat_in, double* feat_out, int nthreads);

/* point.cpp:97-152 -> row-major 3x3. Returns 1 if the reference would throw. */
int orc_rotation_matrix(const double* ci, const double* index, double* R);

/* similarities.cpp:129-158 (weighted corr), 179-188 (weighted SSD); similarities.h:48-58 */
double orc_corr(int n, const double* A, const double* B, const double* w);
double orc_ssd(int n, const double* A, const double* B, const double* w);
double orc_sim_for_min(int simmeasure, int n, const double* A, const double* B, const double* w);
/* similarities.cpp:201-253 DICE (general = 0) / genDICE (1); threshold percentile set process-wide like sparsesimkernel::set_percentile */
void orc_set_percentile(double p);
double orc_dice(int n, const double* A, const double* B, int general);

/* DiscreteCostFunction.cpp:102-107 + 334-351: patch membership lists (CSR over CPs, ascending
 * source id). Returns total entries, writes up to cap. */
int orc_patch_membership(int ncp, const double* cp_xyz, int nsrc, const double* src_xyz,
                         const double* maxsep, double range, int* rowptr, int* members, int cap, int nthreads);

/* Unary cost table, label-major out[l*ncp + k] (DiscreteCostFunction.cpp:236-243).
 * kind: 0 univariate (cpp:353-383), 1 multivariate (cpp:410-458), 2 patchwise (cpp:652-692).
 * rot: [ncp][9] row-major ROTATIONS; labels [L][3]; src_feat [D][nsrc]; ref_feat [D][nv_t];
 * cfw (HIGHREScfweight) [cfw_rows][nsrc] or NULL (=> 1.0); absw [ncp] AbsoluteWeights.
 * tri_out (optional) [L][total patch entries] nearest-triangle ids for parity of indices. */
int orc_unary_costs(int kind, int simmeasure, const orc_octree* target_tree,
                    int ncp, const double* cp_xyz, const double* rot, int L, const double* labels,
                    int nsrc, const double* src_xyz, const int* patch_rowptr, const int* patch_members,
                    int D, const double* src_feat, const double* ref_feat,
                    int cfw_rows, const double* cfw, const double* absw,
                    double* out, int* tri_out, int nthreads);

/* HO*::get_source_data (DiscreteCostFunction.cpp:468-485, 541-563): CSR over CP-grid triangles. Returns entries or -1. */
int orc_ho_patches(int ncp, const double* cp_xyz, int ntri, const int* cp_tri, int nsrc, const double* src_xyz,
                   int* rowptr, int* members, int cap);

/* variance_normalise (msm-newmeshreg/src/reg_tools.cpp:804-844) in place on data [D][n]; excl NULL or [n] (kept where > 0). */
void orc_variance_normalise(int D, int n, double* data, const double* excl);

/* computeTripletCost (DiscreteCostFunction.cpp:135-188) for n requests; PARITY UNPINNED (see msm_oracle.cpp).
 * kind 0..2: likelihood 0; 3: HOUnivariate; 4: HOMultivariate. rmode 2/3 (spherical strain) only. */
/* regoption 4/5 (anatomical strain, DiscreteCostFunction.cpp:169-181, 245-301): the meshes and maps of set_anatomical /
 * set_anatomical_neighbourhood (DiscreteCostFunction.h:160-168). Same layout as msmgpu_anatomical (include/msmgpu.h). */
typedef struct {
    int n_av; const double* asource_xyz; int n_at; const int* asource_tri;     /* _aSOURCE */
    int n_hv; const double* thi_xyz; int n_ht; const int* thi_tri;             /* _TARGEThi (the octree `anattree` is built over it) */
    const double* atarget_xyz;                                                 /* _aTARGET coordinates [n_hv][3] */
    const int* face_ptr; const int* face_ids;                                  /* NEARESTFACES as CSR over the triplets */
    const int* bary_ptr; const int* bary_key; const double* bary_w;            /* _ANATbaryweights as CSR over _aSOURCE vertices, ascending key */
} orc_anat;
int orc_triplet_costs_anat(int kind, int simmeasure, const orc_octree* T, int ncp, const double* cp_xyz, const double* orig_cp_xyz,
                           const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                           int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                           int nsrc, const double* src_xyz, const int* prow, const int* pmem, int D, const double* src_feat,
                           const double* ref_feat, int cfw_rows, const double* cfw, const double* absw,
                           double lambda, double mu, double kappa, double k_exp, double rexp, int rmode, const orc_anat* anat,
                           double* out, int nthreads);
int orc_triplet_costs(int kind, int simmeasure, const orc_octree* T, int ncp, const double* cp_xyz, const double* orig_cp_xyz,
                      const double* rot, int L, const double* labels, int ntrip, const int* triplets,
                      int n, const int* req_triplet, const int* req_la, const int* req_lb, const int* req_lc,
                      int nsrc, const double* src_xyz, const int* prow, const int* pmem, int D, const double* src_feat,
                      const double* ref_feat, int cfw_rows, const double* cfw, const double* absw,
                      double lambda, double mu, double kappa, double k_exp, double rexp, double* out, int nthreads);

/* gMSM, PARITY UNPINNED: DiscreteGroupModel::get_patch_data (DiscreteGroupModel.cpp:88-121) as resampled fields
 * [S][L][D][n_tpl], and DiscreteGroupCostFunction::computePairwiseCost (DiscreteGroupCostFunction.cpp:54-97). */
int orc_group_fields(int S, int nv, const double* data_xyz, int nt, const int* tri, int D, const double* feat, int L, const double* labels,
                     const double* centre, int n_tpl, const double* tpl_xyz, int nt_tpl, const int* tpl_tri, double* fields, int nthreads);
int orc_group_pair_costs(int simmeasure, int S, int ncp, int L, int D, int n_tpl, const double* tpl_xyz, const double* fields,
                         const double* rot, const double* labels, const double* spacings, double range, const int* pairs,
                         int n, const int* req_pair, const int* req_la, const int* req_lb, double* out, int nthreads);
int orc_group_pair_costs_masked(int simmeasure, int S, int ncp, int L, int D, int n_tpl, const double* tpl_xyz, const double* fields,
                         const double* rot, const double* labels, const double* spacings, double range, const int* pairs,
                         int n, const int* req_pair, const int* req_la, const int* req_lb, const double* mask, double* out, int nthreads);

/* RIGID / AFFINE level (rigid_costfunction.cpp:32-236): initialise + cost at zero rotation + run; same outputs as
 * oracle/ref_meshreg_driver.cpp: refmr_rigid. src_feat [D][nv_s], ref_feat [D][nv_t]. Returns the number of neighbour entries. */
int orc_rigid(int nv_t, const double* tgt_xyz, int nt_t, const int* tgt_tri, int nv_s, const double* src_xyz, int nt_s, const int* src_tri,
              int D, const double* src_feat, const double* ref_feat, int simmeasure, int iters, double stepsize, double gradsampling,
              double* out_xyz, double* out_cost0, int* nbh_rowptr, int* nbh_members, int cap);

#ifdef __cplusplus
}
#endif
#endif
