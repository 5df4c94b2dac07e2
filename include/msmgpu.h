/* msmgpu.h — C ABI of the B200 (sm_100a) implementation of newMSM's data-parallel hot path.
 *
 * The reference (rbesenczi/newMSM) has no FFI: its hot path is C++ classes called in-process.
 * Each entry point below replaces one reference interface; the citation after "replaces:"
 * is relative to /root/reference/libraries/. A reference-side adapter keeps
 * the C++ signatures (newresampler::Octree / Resampler / free functions,
 * newmeshreg::DiscreteCostFunction) and forwards to these calls.
 *
 * Conventions
 *  - plain pointers and sizes only; `_dev` suffix = pointers are DEVICE pointers and the call
 *    is asynchronous on the context stream; otherwise pointers are HOST pointers and the call
 *    copies in/out and synchronises before returning.
 *  - coordinates: xyz = [n][3] double (AoS, the reference's Point), triangles = [nt][3] int32,
 *    0-based vertex ids. Sphere radius 100 (point.h:32).
 *  - features at the host boundary are CHANNEL-major [D][V] like Mesh::pvalues
 *    (msm-newresampler/src/mesh.h:44); on the device they are VERTEX-major rows [V][D].
 *  - every function returns msmgpu_status; msmgpu_last_error() gives the message
 *    (the reference's MeshException texts where one exists). There is no CPU fallback:
 *    without a CUDA device every compute entry point returns MSMGPU_ERR_CUDA.
 */
#ifndef MSMGPU_H
#define MSMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MSMGPU_OK = 0,
    MSMGPU_ERR_CUDA = 1,          /* CUDA runtime error / no device */
    MSMGPU_ERR_INVALID = 2,       /* bad argument */
    MSMGPU_ERR_OUT_OF_BOX = 3,    /* "Point is not in the bounding box of the mesh" (octree.cpp:158) */
    MSMGPU_ERR_NO_TRIANGLE = 4,   /* "Error in octree. ..." (octree.cpp:211) */
    MSMGPU_ERR_CAPACITY = 5       /* internal buffer bound exceeded */
} msmgpu_status;

typedef struct msmgpu_ctx msmgpu_ctx;         /* device + stream + scratch */
typedef struct msmgpu_mesh msmgpu_mesh;       /* device-resident vertices/faces (+ per-face tables) */
typedef struct msmgpu_octree msmgpu_octree;   /* flattened octree of one mesh */
typedef struct msmgpu_weights msmgpu_weights; /* device CSR resampling matrix */
typedef struct msmgpu_costfn msmgpu_costfn;   /* device state of one DiscreteCostFunction */
typedef struct msmgpu_fwd msmgpu_fwd;         /* barycentric weight maps of a batch of subjects, kept on the device */

const char* msmgpu_last_error(void);
const char* msmgpu_version(void);
int msmgpu_device_count(void);
/* number of kernels this library has launched since it was loaded (bench.py reports it as gpu_launches) */
unsigned long long msmgpu_launch_count(void);
/* debugging aid: text of a pending CUDA runtime error left by an unchecked call (clears it); "" if none */
const char* msmgpu_debug_take_cuda_error(void);

/* Tuning knob with no effect on results: how many lanes cooperate on one nearest-triangle query
 * (1, 2, 4, 8, 16 or 32; default 1 or $MSMGPU_QUERY_GROUP). */
msmgpu_status msmgpu_set_query_group(int lanes);
int msmgpu_get_query_group(void);

/* Other tuning knobs with no effect on results (launch variants that profiles/ compares): "gather" (1 = bulk-copy row gather for
 * rows >= 128 bytes, 0 = register-path kernels), "gather_variant", "gather_variant_bary", "resample_variant", "weights_minb",
 * "query_order", "build_top" (-1 auto / 0 off / k: depth of the one-pass construction of the first octree levels), "build_top_order",
 * "build_top_agg", "build_fused_levels", "lazy_records" (1: view-batch meshes defer their 128-byte records), "reverse_across" (1: reverse queries of a batch
 * with the subject on the lanes), "tables_minb", "across_minb", "apply_minb", "batch_chunk" ... Each starts from its MSMGPU_<NAME>
 * environment variable. */
msmgpu_status msmgpu_set_tuning(const char* name, int value);

/* stream == NULL -> a private non-blocking stream; otherwise a cudaStream_t owned by the caller.
 * A context (and the handles created from it) serves ONE host thread at a time; use one context per thread (or per device) otherwise. */
msmgpu_status msmgpu_ctx_create(int device, void* stream, msmgpu_ctx** out);
void msmgpu_ctx_destroy(msmgpu_ctx* ctx);
msmgpu_status msmgpu_ctx_sync(msmgpu_ctx* ctx);
void* msmgpu_ctx_stream(msmgpu_ctx* ctx);
/* Device buffers for hosts that have no allocator of their own (the C++ adapters; Python callers pass torch storage instead):
 * plain cudaMalloc / cudaFree on the context's device, and copies ordered on the context's stream (the call returns when done).
 * msmgpu_device_copy_peer copies between buffers of two contexts (devices), e.g. the per-iteration field shards of gMSM. */
msmgpu_status msmgpu_device_malloc(msmgpu_ctx* ctx, size_t bytes, void** out);
void msmgpu_device_free(msmgpu_ctx* ctx, void* ptr);
msmgpu_status msmgpu_device_download(msmgpu_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
msmgpu_status msmgpu_device_copy_peer(msmgpu_ctx* dst_ctx, void* dst, msmgpu_ctx* src_ctx, const void* src, size_t bytes);
/* Page-locked host memory (cudaHostAlloc) for callers that flatten their own containers anyway (the reference keeps Mesh::pvalues as
 * vector<vector<double>>, mesh.h:44): a host-pointer entry point handed such a buffer copies by DMA at link speed instead of through the
 * driver's staging copy of pageable memory. Any host pointer stays valid input; this is an optimisation, not a requirement. */
msmgpu_status msmgpu_host_alloc(msmgpu_ctx* ctx, size_t bytes, void** out);
void msmgpu_host_free(msmgpu_ctx* ctx, void* ptr);

/* replaces: newresampler::Mesh as geometry carrier (msm-newresampler/src/mesh.h:37-58) */
msmgpu_status msmgpu_mesh_create(msmgpu_ctx* ctx, int nv, const double* xyz, int nt, const int32_t* tri, msmgpu_mesh** out);
msmgpu_status msmgpu_mesh_create_dev(msmgpu_ctx* ctx, int nv, const double* d_xyz, int nt, const int32_t* d_tri, msmgpu_mesh** out);
/* Device-resident pipelines: n meshes that share one topology and VIEW the caller's device buffers (d_xyz[i] [nv][3], d_tri [nt][3];
 * no copies; the buffers must stay valid and unchanged for the lifetime of the meshes and of the trees / weights built from them).
 * The per-triangle tables of the batch live in one allocation. msmgpu_mesh_set_coords is refused on such a mesh. */
msmgpu_status msmgpu_mesh_create_view_batch(msmgpu_ctx* ctx, int n, int nv, const double* const* d_xyz, int nt, const int32_t* d_tri, msmgpu_mesh** out);
msmgpu_status msmgpu_mesh_set_coords(msmgpu_mesh* m, const double* xyz);   /* Mesh::set_coord for all vertices */
void msmgpu_mesh_destroy(msmgpu_mesh* m);
msmgpu_status msmgpu_mesh_shape(msmgpu_mesh* m, int* nv, int* nt);
/* replaces: Mesh::pvalues / set_pvalues (mesh.h:44, mesh.cpp:206) as a device-resident FP32 payload: channel-major host floats
 * [D][nv] uploaded ONCE and reused by msmgpu_mesh_bary_resample_f32 / msmgpu_mesh_metric_resample_f32 */
msmgpu_status msmgpu_mesh_set_features_f32(msmgpu_mesh* m, int D, const float* feat_cm);
/* Triangle::area is cached when a Triangle is constructed (triangle.cpp:31,39) and NOT refreshed by Mesh::set_coord; only a Mesh
 * copy/assignment recomputes it (mesh.cpp:37-53). compute_vertex_area (mesh.cpp:1275) reads the cached value, so a reference
 * Mesh that was copied from A and then moved still has A's vertex areas (e.g. DiscreteGroupModel.cpp:94-103). Mirror of that
 * state: vertex areas of `m` (msmgpu_mesh_vertex_areas, the adaptive weights) are taken from `area_mesh`'s geometry.
 * area_mesh = NULL (or m) restores "freshly copied" semantics. Same context, same nv/nt; area_mesh must outlive its use. */
msmgpu_status msmgpu_mesh_set_area_source(msmgpu_mesh* m, msmgpu_mesh* area_mesh);
/* Same purpose with explicit values: areas[nt] = Mesh::get_triangle_area(t) (mesh.cpp:793) of the reference object being mirrored,
 * whatever its history of copies and set_coord calls. NULL returns to areas computed from the current coordinates. Takes precedence
 * over the mesh's own geometry (an area source's explicit areas are honoured too). */
msmgpu_status msmgpu_mesh_set_triangle_areas(msmgpu_mesh* m, const double* areas);

/* replaces: compute_vertex_area (msm-newresampler/src/mesh.cpp:1275) for all vertices */
msmgpu_status msmgpu_mesh_vertex_areas(msmgpu_mesh* m, double* out);

/* replaces: Octree::Octree / initialize_tree / add_triangle (msm-newresampler/src/octree.cpp:31-141).
 * Built on the device, level by level, with the same leaf sets and leaf order as sequential insertion. */
msmgpu_status msmgpu_octree_build(msmgpu_mesh* m, msmgpu_octree** out);
/* same for a batch of meshes in one pass (a forest: every kernel launch covers all meshes) */
msmgpu_status msmgpu_octree_build_batch(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* meshes, msmgpu_octree** out);
void msmgpu_octree_destroy(msmgpu_octree* t);
/* sizes: n_nodes, n_leaf_refs (sum of leaf triangle counts), depth */
msmgpu_status msmgpu_octree_stats(msmgpu_octree* t, int* n_nodes, int* n_leaf_refs, int* depth);
/* pre-order dump (children in the reference's [i][j][k] order): kinds[n_nodes] (1 leaf / 0 internal),
 * counts[n_nodes], tris[n_leaf_refs] — for parity tests of the topology */
msmgpu_status msmgpu_octree_dump(msmgpu_octree* t, int32_t* kinds, int32_t* counts, int32_t* tris);

/* replaces: Octree::get_closest_triangle (octree.cpp:156-214) and get_closest_vertex_ID (216-233) for n points.
 * out_tri / out_vertex may be NULL. status[i]: 0 ok, MSMGPU_ERR_OUT_OF_BOX, MSMGPU_ERR_NO_TRIANGLE; status may be
 * NULL, then the first failing point makes the call fail like the reference's throw. */
msmgpu_status msmgpu_nearest_triangle(msmgpu_octree* t, int n, const double* pts, int32_t* out_tri, int32_t* out_vertex, int32_t* status);
msmgpu_status msmgpu_nearest_triangle_dev(msmgpu_octree* t, int n, const double* d_pts, int32_t* d_tri, int32_t* d_vertex, int32_t* d_status);

/* replaces: Resampler::get_barycentric_weights (msm-newresampler/src/resampler.cpp:142-167): idx/w are [n][3], entries in
 * ascending vertex id (std::map order); n_entries[n] (<3 when ids repeat), may be NULL. */
msmgpu_status msmgpu_bary_weights(msmgpu_octree* t, int n, const double* pts, int32_t* idx, double* w, int32_t* n_entries);
msmgpu_status msmgpu_bary_weights_dev(msmgpu_octree* t, int n, const double* d_pts, int32_t* d_idx, double* d_w, int32_t* d_n_entries, int32_t* d_status);

/* replaces: Resampler::get_adaptive_barycentric_weights (resampler.cpp:72-140, no exclusion mask).
 * Result: CSR over target vertices, columns ascending. */
msmgpu_status msmgpu_adaptive_weights(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, msmgpu_weights** out);
/* optional pre-built trees (NULL = build): lets a batch share the target tree */
msmgpu_status msmgpu_adaptive_weights_ex(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, msmgpu_weights** out);
/* batch of n subjects onto one target mesh: one set of kernel launches for all of them. in_trees may be NULL or hold
 * NULL entries (built here, as one forest), low_tree may be NULL. out[n] share one device store. */
msmgpu_status msmgpu_adaptive_weights_batch(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* in_meshes, msmgpu_octree* const* in_trees,
                                            msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, msmgpu_weights** out);
msmgpu_status msmgpu_weights_shape(msmgpu_weights* w, int* n_rows, int* n_cols, int64_t* nnz);
msmgpu_status msmgpu_weights_export(msmgpu_weights* w, int32_t* rowptr, int32_t* col, double* val);
void msmgpu_weights_destroy(msmgpu_weights* w);

/* replaces: the interpolation loop of Resampler::barycentric_data_interpolation (resampler.cpp:40-52).
 * d_in: [n_cols][D] float rows, d_out: [n_rows][D] float rows (vertex-major). */
msmgpu_status msmgpu_weights_apply_f32_dev(msmgpu_weights* w, int D, const float* d_in, float* d_out);
/* one launch for the n matrices of one msmgpu_adaptive_weights_batch call: d_in[i] / d_out[i] row pointers per subject */
msmgpu_status msmgpu_weights_apply_batch_f32_dev(msmgpu_ctx* ctx, int n, msmgpu_weights* const* ws, int D, const float* const* d_in, float* const* d_out);

/* FP64 payload (bit-exact with the reference's double features; used by the groupwise fields) */
msmgpu_status msmgpu_weights_apply_batch_f64_dev(msmgpu_ctx* ctx, int n, msmgpu_weights* const* ws, int D, const double* const* d_in, double* const* d_out);

/* replaces: metric_resample(in, low) (resampler.cpp:304): adaptive-barycentric resampling of D channels.
 * Host buffers, channel-major: feat_in [D][nv_in] double, feat_out [D][nv_low] double. */
msmgpu_status msmgpu_metric_resample(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, int D, const double* feat_in, double* feat_out);
/* f32 payload variant (what the reference reads from / writes to GIFTI, mesh.cpp:625): channel-major host floats */
msmgpu_status msmgpu_metric_resample_f32(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree,
                                         int D, const float* feat_in, float* feat_out);

/* Plain barycentric resample = Octree(in) + get_barycentric_weights + loop of resampler.cpp:40-52, fused in one
 * kernel (query -> weights -> 3-row gather). Device rows: d_feat_in [nv_in][D] float, d_feat_out [n][D] float. */
msmgpu_status msmgpu_bary_resample_f32_dev(msmgpu_octree* t, int n, const double* d_pts, int D, const float* d_feat_in, float* d_feat_out, int32_t* d_status);
/* batched over subjects (one launch): trees[s], pts shared, d_feat_in[s], d_feat_out[s] */
msmgpu_status msmgpu_bary_resample_batch_f32_dev(msmgpu_ctx* ctx, int n_subjects, msmgpu_octree* const* trees, int n, const double* d_pts,
                                                 int D, const float* const* d_feat_in, float* const* d_feat_out, int32_t* d_status);
/* host buffers, FP32 payload, channel-major like Mesh::pvalues: feat_in [D][nv] float -> feat_out [D][n] float */
/* A batch job that runs BOTH resampling methods on the same (subjects, targets) computes get_barycentric_weights(targets, subject)
 * twice in the reference: once for the barycentric resample (resampler.cpp:142) and once as the `forward` half of
 * get_adaptive_barycentric_weights (resampler.cpp:74-75). The _keep variant of the fused resample stores those maps (3 ids, 3 weights,
 * entry count per target) and msmgpu_adaptive_weights_batch_fwd consumes them instead of querying again: same values, one query pass.
 * The store must have been filled for exactly these trees and for targets = the vertices of low_mesh (shape and trees are checked). */
msmgpu_status msmgpu_fwd_create(msmgpu_ctx* ctx, int n_subjects, int n, msmgpu_fwd** out);
void msmgpu_fwd_destroy(msmgpu_fwd* f);
msmgpu_status msmgpu_bary_resample_batch_f32_dev_keep(msmgpu_ctx* ctx, int n_subjects, msmgpu_octree* const* trees, int n, const double* d_pts,
                                                      int D, const float* const* d_feat_in, float* const* d_feat_out, int32_t* d_status,
                                                      msmgpu_fwd* keep);
msmgpu_status msmgpu_adaptive_weights_batch_fwd(msmgpu_ctx* ctx, int n, msmgpu_mesh* const* in_meshes, msmgpu_octree* const* in_trees,
                                                msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, const msmgpu_fwd* fwd, msmgpu_weights** out);
/* replaces: the interpolation loop of barycentric_data_interpolation (resampler.cpp:40-52) over weight maps that are already known:
 * applies the kept forward maps of a batch to (other) feature rows of the same subjects — the gather kernel alone.
 * Rows must be 16-byte aligned multiples of 16 bytes between 128 B and 2 KB. */
msmgpu_status msmgpu_fwd_apply_batch_f32_dev(msmgpu_ctx* ctx, const msmgpu_fwd* fwd, int D, const float* const* d_feat_in, float* const* d_feat_out);
msmgpu_status msmgpu_bary_resample_f32(msmgpu_octree* t, int n, const double* pts, int D, const float* feat_in, float* feat_out);
/* the same two resamplers on the mesh's resident features: only the result crosses PCIe. feat_out channel-major [D][n] floats */
msmgpu_status msmgpu_mesh_bary_resample_f32(msmgpu_octree* t, int n, const double* pts, float* feat_out);
msmgpu_status msmgpu_mesh_metric_resample_f32(msmgpu_mesh* in_mesh, msmgpu_octree* in_tree, msmgpu_mesh* low_mesh, msmgpu_octree* low_tree, float* feat_out);
/* replaces: the per-subject loop of a batch resampling job on HOST buffers (BASELINE configs[1]: S subjects with their own coordinates
 * over one topology, resampled onto one target sphere with the barycentric method — Octree + get_barycentric_weights +
 * resampler.cpp:40-52 — and / or metric_resample, resampler.cpp:304). One call, pipelined inside the library: chunk k+1 is uploaded
 * while chunk k runs through the BATCHED kernels and chunk k-1 is downloaded, so the job runs at the rate of the host link.
 *   xyz[s]      host [nv][3] doubles            tri        host [nt][3], shared by the subjects
 *   feat_cm[s]  host [D][nv] floats (Mesh::pvalues layout)
 *   out_bary_cm[s] / out_adaptive_cm[s]  host [D][n_low] floats; either array may be NULL to skip that method
 *   chunk       subjects per pipeline stage (0: default 4, knob "batch_chunk")
 * Page-locked host buffers (msmgpu_host_alloc or the caller's own) make the copies asynchronous; pageable ones work, serialised.
 * Results are those of the per-subject calls (msmgpu_mesh_bary_resample_f32 / msmgpu_mesh_metric_resample_f32) bit for bit. */
msmgpu_status msmgpu_resample_batch_host_f32(msmgpu_ctx* ctx, int n_subjects, int nv, const double* const* xyz, int nt, const int32_t* tri,
                                             int n_low, const double* low_xyz, int n_low_tri, const int32_t* low_tri, int D,
                                             const float* const* feat_cm, float* const* out_bary_cm, float* const* out_adaptive_cm, int chunk);
/* host-buffer convenience (channel-major double in/out), used by the parity tests */
msmgpu_status msmgpu_bary_resample(msmgpu_mesh* in_mesh, int n, const double* pts, int D, const double* feat_in, double* feat_out);

/* replaces: sphere_project_warp (resampler.cpp:311-328): out = normalize(sum w*to[idx])*100 for n points located in `from` */
msmgpu_status msmgpu_sphere_project_warp(msmgpu_mesh* from_mesh, const double* to_xyz, int n, const double* sphere_xyz, double* out_xyz);
/* replaces: surface_resample (resampler.cpp:284-302) / project_anatomical_mesh (260-282): out = sum w*anat[idx] */
msmgpu_status msmgpu_surface_resample(msmgpu_mesh* sph_mesh, const double* anat_xyz, int n, const double* low_xyz, double* out_xyz);
/* replaces: nearest_neighbour_interpolation (resampler.cpp:232-258, no exclusion): channel-major host doubles */
msmgpu_status msmgpu_nn_resample(msmgpu_mesh* in_mesh, int n, const double* low_xyz, int D, const double* feat_in, double* feat_out);

/* replaces: estimate_rotation_matrix (msm-newresampler/src/point.cpp:97-152) for n (ci,index) pairs -> [n][9] row-major */
/* replaces: the O(V^2) neighbourhood search of newresampler::smooth_data (resampler.cpp:186-200), the IEEE-only part of the Gaussian
 * smoothing. For target i: ref = unit(low_xyz[closest[i]]) (closest = get_closest_vertex_ID of the target in `orig`); the list holds,
 * in ascending n, every vertex n with unit(low_xyz[n]) . ref >= cos_ang, and chords[] = |ref - unit(low_xyz[n])|. CSR output:
 * rowptr[n+1] is always written; members / chords only when non-NULL and cap >= rowptr[n] (call once with NULL to size them).
 * The Gaussian weights (asin, exp: resampler.cpp:204-205) are the caller's, on the host libm. */
msmgpu_status msmgpu_smooth_neighbourhoods(msmgpu_ctx* ctx, int n, const double* low_xyz, const int32_t* closest, double cos_ang,
                                           int32_t* rowptr, int64_t cap, int32_t* members, double* chords);

/* replaces: newresampler::smooth_data (resampler.cpp:169-230) as one call, exclusion mask included (201-225): neighbourhood scan on the
 * device, Gaussian weights and the reference's sequential sums on the host libm. feat_cm [D][n_feat] = orig's pvalues (indexed by the
 * vertex ids of low_xyz, as the reference does), excl NULL or [n_excl] = EXCL's values; out_cm [D][n]; excl_out [n] (with a mask). */
msmgpu_status msmgpu_smooth_data(msmgpu_ctx* ctx, int n, const double* low_xyz, const int32_t* closest, double sigma, int D, int n_feat,
                                 const double* feat_cm, int n_excl, const double* excl, double* out_cm, double* excl_out);

/* replaces: newmeshreg::variance_normalise (msm-newmeshreg/src/reg_tools.cpp:804-844), the last stage of featurespace::initialise
 * (featurespace.cpp:79-82): every channel of data_cm [D][n] becomes (x - mean) / sqrt(var) over the vertices with excl > 0 (all when
 * excl is NULL), mean / var from the reference's sequential Welford recurrence in vertex order; excluded vertices keep their values. */
msmgpu_status msmgpu_variance_normalise(msmgpu_ctx* ctx, int D, int n, double* data_cm, const double* excl);

/* ---- exclusion masks (`--excl` / cut thresholds, featurespace.cpp:61-70): EXCL = a Mesh whose first channel is 0 where data is ignored ---- */
/* replaces: get_adaptive_barycentric_weights(in, low, nthreads, EXCL) (resampler.cpp:72-140): targets whose closest source vertex is masked
 * out (cpp:100) get empty rows and do not enter the correction sums. excl [nv_in] host doubles. */
msmgpu_status msmgpu_adaptive_weights_excl(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, const double* excl, msmgpu_weights** out);
/* replaces: barycentric_data_interpolation / metric_resample with EXCL (resampler.cpp:30-70, 304): masked weights, sums that skip masked
 * source vertices (cpp:46-47), and the mask resampled with the same weights (cpp:55-67; excl_out [nv_low] replaces *EXCL) */
msmgpu_status msmgpu_metric_resample_excl(msmgpu_mesh* in_mesh, msmgpu_mesh* low_mesh, int D, const double* feat_in, const double* excl_in,
                                          double* feat_out, double* excl_out);
/* replaces: nearest_neighbour_interpolation with EXCL (resampler.cpp:232-258) */
msmgpu_status msmgpu_nn_resample_excl(msmgpu_mesh* in_mesh, int n, const double* low_xyz, int D, const double* feat_in, const double* excl_in,
                                      double* feat_out, double* excl_out);

msmgpu_status msmgpu_rotation_matrices(msmgpu_ctx* ctx, int n, const double* ci, const double* index, double* R);

/* ---- discrete-optimisation cost evaluation (msm-newmeshreg/src/DiscreteCostFunction.{h,cpp}) ---- */
typedef enum {
    MSMGPU_COST_UNIVARIATE = 0,   /* UnivariateNonLinearSRegDiscreteCostFunction  (cpp:326-383) */
    MSMGPU_COST_MULTIVARIATE = 1, /* MultivariateNonLinearSRegDiscreteCostFunction (cpp:385-458) */
    MSMGPU_COST_PATCHWISE = 2,    /* PatchwiseMultivariate...                      (cpp:620-692) */
    MSMGPU_COST_HO_UNIVARIATE = 3,   /* HOUnivariateNonLinearSRegDiscreteCostFunction   (cpp:460-531): triplet likelihood */
    MSMGPU_COST_HO_MULTIVARIATE = 4  /* HOMultivariateNonLinearSRegDiscreteCostFunction (cpp:533-618) */
} msmgpu_cost_kind;

/* regulariser parameters of NonLinearSRegDiscreteCostFunction::set_parameters (cpp:119-133) */
typedef struct {
    double lambda;          /* "lambda"        _reglambda */
    double shear_modulus;   /* "shearmodulus"  _mu     (default 0.4) */
    double bulk_modulus;    /* "bulkmodulus"   _kappa  (default 1.6) */
    double k_exponent;      /* "kexponent"     _k_exp  (default 2) */
    double exponent;        /* "exponent"      _rexp   (default 2) */
    int rmode;              /* "regularisermode": 2 or 3 = spherical strain; 4 or 5 = anatomical strain (msmgpu_costfn_set_anatomical first) */
} msmgpu_reg_params;

/* replaces: set_meshes + set_featurespace + set_octrees (DiscreteCostFunction.h:173-189).
 * target: TARGET mesh + its octree; source_xyz [nsrc][3]; features channel-major doubles:
 * src_feat [D][nsrc] (FEAT input), ref_feat [D][nv_target] (FEAT reference). simmeasure 1 = SSD, 2 = correlation. */
msmgpu_status msmgpu_costfn_create(msmgpu_octree* target_tree, msmgpu_cost_kind kind, int simmeasure,
                                   int nsrc, const double* source_xyz, int D, const double* src_feat, const double* ref_feat,
                                   msmgpu_costfn** out);
void msmgpu_costfn_destroy(msmgpu_costfn* c);
/* replaces: reset_source (DiscreteCostFunction.h:190) */
/* sparsesimkernel::set_percentile (similarities.h:40, "--percentile"): threshold of the DICE measures (simmeasure 4 / 5), default 0.75 */
msmgpu_status msmgpu_costfn_set_percentile(msmgpu_costfn* c, double percentile);
msmgpu_status msmgpu_costfn_reset_source(msmgpu_costfn* c, const double* source_xyz);
/* replaces: reset_CPgrid + set_spacings + set_dataaffintyweighting + get_source_data (cpp:334-351: patch membership by
 * within_controlpt_range, cpp:102-107) + resample_weights (cpp:303-323) inputs.
 * cp_xyz [ncp][3], maxsep [ncp], range = _controlptrange, cfw [cfw_rows][nsrc] or NULL (weights 1), absw [ncp] AbsoluteWeights. */
msmgpu_status msmgpu_costfn_set_cpgrid(msmgpu_costfn* c, int ncp, const double* cp_xyz, const double* maxsep, double range,
                                       int cfw_rows, const double* cfw, const double* absw);
/* patch lists as computed on the device: rowptr[ncp+1], members[rowptr[ncp]] (ascending source id); members may be NULL */
msmgpu_status msmgpu_costfn_patches(msmgpu_costfn* c, int32_t* rowptr, int32_t* members);
/* replaces: set_labels (h:180) + computeUnaryCosts (cpp:236-243): labels [L][3], rotations [ncp][9] row-major (HOST arrays:
 * the per-(cp,label) matrices estimate_rotation_matrix(CP_k, ROT_k*label_l) are built with the host libm).
 * out [L][ncp] doubles (label-major like unarycosts[l*N+k]); tri_out (optional) [L][n_patch_entries] nearest-triangle ids.
 * A failed query makes that entry NaN and the call return MSMGPU_ERR_NO_TRIANGLE (the reference throws, octree.cpp:211). */
msmgpu_status msmgpu_costfn_unary_table(msmgpu_costfn* c, int L, const double* labels, const double* rotations,
                                        double* out, int32_t* tri_out);
/* device-resident result: d_out [L][ncp] (and d_tri_out) stay on the device for a device-side consumer; labels / rotations are host arrays */
msmgpu_status msmgpu_costfn_unary_table_dev(msmgpu_costfn* c, int L, const double* labels, const double* rotations, double* d_out, int32_t* d_tri_out);

/* replaces: HO*::get_source_data (cpp:468-485, 541-563): patches = source vertices grouped by their nearest CP-GRID triangle
 * (cp_tri [ntri][3]); msmgpu_costfn_patches then returns rowptr[ntri+1] / members. HO kinds only. */
msmgpu_status msmgpu_costfn_set_cpgrid_ho(msmgpu_costfn* c, int ncp, const double* cp_xyz, int ntri, const int32_t* cp_tri,
                                          int cfw_rows, const double* cfw, const double* absw);
/* replaces: computeTripletCost (cpp:135-188) = HO likelihood (cpp:487-531, 565-618; 0 for the non-HO kinds) + lambda * strain^exponent
 * (reg_tools.cpp:551-743) for n requests (triplet, la, lb, lc). triplets [ntrip][3] node ids (ascending, DiscreteModel.cpp:303),
 * rotations [ncp][9], labels [L][3], orig_cp_xyz [ncp][3] = _ORIG. Folded triangles cost FOLDING * lambda = 1e7 * lambda. */
msmgpu_status msmgpu_costfn_triplet_costs(msmgpu_costfn* c, int ntrip, const int32_t* triplets, int L, const double* labels, const double* rotations,
                                          const double* orig_cp_xyz, const msmgpu_reg_params* prm, int n, const int32_t* req_triplet,
                                          const int32_t* req_la, const int32_t* req_lb, const int32_t* req_lc, double* out);
/* the 8 combinations Fusion::optimize asks per triplet for one candidate label (Fusion.h:181-196): out [ntrip][8],
 * out[t][b] = cost(t, b&4 ? label : labeling[A], b&2 ? label : labeling[B], b&1 ? label : labeling[C]) */
msmgpu_status msmgpu_costfn_triplet_batch(msmgpu_costfn* c, int ntrip, const int32_t* triplets, int L, const double* labels, const double* rotations,
                                          const double* orig_cp_xyz, const msmgpu_reg_params* prm, const int32_t* labeling, int label, double* out);

/* replaces: set_anatomical + set_anatomical_neighbourhood + initialize_regulariser (DiscreteCostFunction.h:160-169) for regoption 4 / 5:
 * the triplet regulariser becomes the MEAN strain energy of the anatomical faces that belong to the control triangle, each deformed by
 * deform_anatomy (DiscreteCostFunction.cpp:169-181, 245-301): a face vertex moves with its barycentric weights in the displaced control
 * triangle, is located on _TARGEThi (octree `anattree`) and takes the barycentric blend of the _aTARGET coordinates there.
 *   asource:  _aSOURCE (n_av vertices, n_at faces)      thi: _TARGEThi (n_hv vertices, n_ht faces), atarget_xyz: _aTARGET [n_hv][3]
 *   face_ptr / face_ids: NEARESTFACES as CSR over the ntrip control triangles (ids into asource_tri, the reference's order)
 *   bary_ptr / bary_key / bary_w: _ANATbaryweights as CSR over the _aSOURCE vertices, keys (control-point ids) ascending like std::map
 * A vertex whose location query fails gets NaN coordinates and the cost is NaN: the reference catches the exception and carries on with
 * a zero triangle whose weights are 0/0 (cpp:268-274). */
typedef struct {
    int n_av; const double* asource_xyz; int n_at; const int32_t* asource_tri;
    int n_hv; const double* thi_xyz; int n_ht; const int32_t* thi_tri;
    const double* atarget_xyz;
    const int32_t* face_ptr; const int32_t* face_ids;
    const int32_t* bary_ptr; const int32_t* bary_key; const double* bary_w;
} msmgpu_anatomical;
msmgpu_status msmgpu_costfn_set_anatomical(msmgpu_costfn* c, int ntrip, const msmgpu_anatomical* a);

/* ---- AFFINE / RIGID level (msm-newmeshreg/src/rigid_costfunction.cpp) ---- */
typedef struct msmgpu_rigid msmgpu_rigid;
/* replaces: Mesh::calculate_MeanVD (msm-newresampler/src/mesh.cpp:276-294) for a mesh built by push_triangle in triangle order. Host code. */
msmgpu_status msmgpu_mean_vertex_distance(int nv, const double* xyz, int nt, const int32_t* tri, double* out);
/* replaces: Rigid_cost_function::initialise (rigid_costfunction.cpp:32-50): TARGET octree, similarity means (similarities.cpp:106-126),
 * Neighbourhood::update (reg_tools.cpp:31-58; only "does vertex i have a neighbour" outlives the first evaluation). Features are
 * channel-major doubles: src_feat [D][nv_s] = FEAT input data, ref_feat [D][nv_t] = FEAT reference data; simmeasure 1 = SSD, 2 = correlation.
 * mean_vertex_distance = SOURCE.calculate_MeanVD() (cpp:35: MVD = min_sigma). */
msmgpu_status msmgpu_rigid_create(msmgpu_ctx* ctx, int nv_t, const double* tgt_xyz, int nt_t, const int32_t* tgt_tri, int nv_s, const double* src_xyz,
                                  int nt_s, const int32_t* src_tri, int D, const double* src_feat, const double* ref_feat, int simmeasure,
                                  double mean_vertex_distance, msmgpu_rigid** out);
void msmgpu_rigid_destroy(msmgpu_rigid* r);
/* replaces: Rigid_cost_function::rigid_cost_mesh (rigid_costfunction.cpp:130-141) = rotate_in_mesh + Evaluate_SIMGradient for every source
 * vertex (calculate_tangs reg_tools.cpp:205-266, closest TARGET triangle, get_all_neighbours, calculate_sim_column_nbh, WLS_simgradient) + the
 * sum of current_sim. src_xyz [nv_s][3] = the CURRENT source coordinates (NULL: those of the previous call); the source is not modified
 * (the reference restores it, cpp:139). exp() of the weights and the sequential sums run on the host libm inside this call. */
msmgpu_status msmgpu_rigid_cost(msmgpu_rigid* r, const double* src_xyz, double dw1, double dw2, double dw3, double* cost);

/* ---- pow() of the host C library on the device (csrc/hostpow.cuh) ----
 * The three std::pow calls per triplet cost (reg_tools.cpp:596-597, DiscreteCostFunction.cpp:187, DiscreteGroupCostFunction.cpp:51) are
 * evaluated on the device with glibc's own algorithm and the tables of the libm mapped into the process, after a host-side self-test
 * against std::pow. Returns 1 when that path is active, 0 when the costs are finished with the host libm (MSMGPU_DEVICE_POW=0 forces 0). */
int msmgpu_device_pow_enabled(void);
/* test hook: out[i] = pow(x[i], y[i]) evaluated on the device (compared with the host's std::pow by the tests) */
msmgpu_status msmgpu_debug_device_pow(msmgpu_ctx* ctx, int n, const double* x, const double* y, double* out);

/* ---- groupwise registration (gMSM): msm-newmeshreg/src/DiscreteGroupModel.cpp, DiscreteGroupCostFunction.cpp ---- */
typedef struct msmgpu_group msmgpu_group;

/* replaces: the per-(subject,label) part of DiscreteGroupModel::get_patch_data (DiscreteGroupModel.cpp:92-106): every data mesh
 * is "rotated" by every label (estimate_rotation_matrix(centre, vertex) * label, label 0 = identity) and metric_resampled onto the
 * template. Host in: data_xyz [n][nv][3], tri [nt][3], feat_cm [n][D][nv] doubles. Device out: d_fields [n][L][n_tpl][D] doubles.
 * A rank calls it for ITS shard of subjects; the shards are all-gathered by the host side (one collective per iteration). */
msmgpu_status msmgpu_group_fields(msmgpu_ctx* ctx, int n_subjects, int nv, const double* data_xyz, int nt, const int32_t* tri, int D,
                                  const double* feat_cm, int L, const double* labels, const double* centre, msmgpu_mesh* tpl,
                                  msmgpu_octree* tpl_tree, double* d_fields);
/* per-iteration state for the pair costs: rotated control points ROT[node]*label (rotations [S*ncp][9] = m_ROT,
 * DiscreteGroupModel.cpp:77-86), patch radii range*spacings[S*ncp] (cpp:111), and d_fields [S][L][n_tpl][D] for ALL S subjects */
msmgpu_status msmgpu_group_create(msmgpu_ctx* ctx, int simmeasure, int S, int ncp, int L, int D, msmgpu_mesh* tpl, const double* d_fields,
                                  const double* rotations, const double* labels, const double* spacings, double range, msmgpu_group** out);
void msmgpu_group_destroy(msmgpu_group* g);
/* replaces: DiscreteGroupCostFunction::set_masks (DiscreteGroupCostFunction.h:50; DiscreteGroupModel.cpp:164): with a cost mask the
 * weight of a common template vertex p in the weighted similarity is std::abs(_MASK.get_pvalue(p)) instead of 1
 * (DiscreteGroupCostFunction.cpp:77). mask: [n_tpl] host doubles (channel 0 of the mask mesh); NULL removes the mask. */
msmgpu_status msmgpu_group_set_mask(msmgpu_group* g, const double* mask);
/* replaces: DiscreteGroupCostFunction::computePairwiseCost (DiscreteGroupCostFunction.cpp:54-97) for n requests (pair, la, lb);
 * pairs [P][2] global node ids (subject*ncp + vertex, DiscreteGroupModel.cpp:37-55). Any sub-array of pairs may be passed (sharding). */
msmgpu_status msmgpu_group_pair_costs(msmgpu_group* g, int P, const int32_t* pairs, int n, const int32_t* req_pair, const int32_t* req_la,
                                      const int32_t* req_lb, double* out);
/* the 4 combinations Fusion::optimize asks per pair for one candidate label (Fusion.h:164-174): out [P][4] =
 * (cur,cur), (cur,label), (label,cur), (label,label) */
msmgpu_status msmgpu_group_pair_batch(msmgpu_group* g, int P, const int32_t* pairs, const int32_t* labeling, int label, double* out);
/* the same with the pair list resident on the device (set once per iteration: estimate_pairs runs once per setupCostFunction) and
 * the result left on the device: d_out [n_pairs][4] for the pairs [first_pair, first_pair + n_pairs) — any block (sharding).
 * Asynchronous on the context's stream; labeling is a host array. */
msmgpu_status msmgpu_group_set_pairs(msmgpu_group* g, int P, const int32_t* pairs);
msmgpu_status msmgpu_group_pair_batch_dev(msmgpu_group* g, int first_pair, int n_pairs, const int32_t* labeling, int label, double* d_out);

/* replaces: DiscreteGroupCostFunction::computeTripletCost (msm-newmeshreg/src/DiscreteGroupCostFunction.cpp:26-52): the strain energy
 * of a control-grid triangle of one subject under three candidate labels, `subcorr * lambda * W^rexp` with subcorr = 0.1 * S
 * (DiscreteGroupCostFunction.h:45); a folded triangle costs FOLDING = 1e7, a NaN energy FIX_NAN = 1e7 when fixnan != 0.
 * cp_xyz / orig_xyz / rotations are the per-subject control grids concatenated ([S*ncp]), triplets hold global node ids
 * (DiscreteGroupModel.cpp:57-74). _costs: request list; _batch: Fusion's 8 combinations per triplet (Fusion.h:181-196), out[T][8]. */
msmgpu_status msmgpu_group_triplet_costs(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                         const double* labels, int ntrip, const int32_t* triplets, const msmgpu_reg_params* prm, double subcorr, int fixnan,
                                         int n, const int32_t* req_triplet, const int32_t* req_la, const int32_t* req_lb, const int32_t* req_lc, double* out);
msmgpu_status msmgpu_group_triplet_batch(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                         const double* labels, int ntrip, const int32_t* triplets, const msmgpu_reg_params* prm, double subcorr, int fixnan,
                                         const int32_t* labeling, int label, double* out);
/* the same with the per-iteration arrays (control grids, rotations, labels, triplets) resident on the device: a label phase of
 * Fusion::optimize then only sends the labeling. _batch evaluates the triplets [first_triplet, first_triplet + n_triplets)
 * (any block: sharding), out [n_triplets][8]. */
typedef struct msmgpu_triplet_plan msmgpu_triplet_plan;
msmgpu_status msmgpu_triplet_plan_create(msmgpu_ctx* ctx, int n_nodes, const double* cp_xyz, const double* orig_xyz, const double* rotations, int L,
                                         const double* labels, int ntrip, const int32_t* triplets, msmgpu_triplet_plan** out);
void msmgpu_triplet_plan_destroy(msmgpu_triplet_plan* p);
msmgpu_status msmgpu_triplet_plan_batch(msmgpu_triplet_plan* p, const msmgpu_reg_params* prm, double subcorr, int fixnan, int first_triplet,
                                        int n_triplets, const int32_t* labeling, int label, double* out);
/* the same with the costs left on the device (d_out [n_triplets][8] device doubles, the call returns when they are complete): a sharded host gathers the blocks device to device and copies once */
msmgpu_status msmgpu_triplet_plan_batch_dev(msmgpu_triplet_plan* p, const msmgpu_reg_params* prm, double subcorr, int fixnan, int first_triplet,
                                            int n_triplets, const int32_t* labeling, int label, double* d_out);

#ifdef __cplusplus
}
#endif
#endif /* MSMGPU_H */
