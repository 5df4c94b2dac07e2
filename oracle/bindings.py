"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle (oracle/libmsm_oracle.so,
our restatement) and for the compiled reference (oracle/_ref/libref_newresampler.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package newmsm_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libmsm_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_newresampler.so")
REFMR_SO = os.path.join(_HERE, "_ref", "libref_newmeshreg.so")
NEWMSM_REF = os.path.join(_HERE, "_ref", "newmsm_ref")
REFERENCE_ROOT = "/root/reference"

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double


def build(ref: bool | None = None) -> None:
    """Compile the oracle restatement, and the reference build when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref is None:
        ref = os.path.isdir(REFERENCE_ROOT)
    if ref:
        subprocess.run(["make", "-s", "-j8", "-C", _HERE, "ref", "refmr"], check=True)
        if os.path.exists(os.path.join(os.path.dirname(_HERE), "newmsm_b200", "lib", "libmsmgpu.so")):
            # in-process drop-in check (C++ adapter) and the reference program linked with the GPU library (integration/_build/)
            subprocess.run(["make", "-s", "-j8", "-C", _HERE, "adapter_check", "newmsm_gpu"], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_refmr() -> bool:
    return os.path.exists(REFMR_SO)


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# --------------------------------------------------------------------------------------
# our restatement
# --------------------------------------------------------------------------------------
class Anat(C.Structure):
    """orc_anat / refmr_anat / msmgpu_anatomical: the anatomical meshes and maps of regoption 4/5 (DiscreteCostFunction.h:164-169)."""
    _fields_ = [("n_av", _i), ("asource_xyz", _vp), ("n_at", _i), ("asource_tri", _vp),
                ("n_hv", _i), ("thi_xyz", _vp), ("n_ht", _i), ("thi_tri", _vp), ("atarget_xyz", _vp),
                ("face_ptr", _vp), ("face_ids", _vp), ("bary_ptr", _vp), ("bary_key", _vp), ("bary_w", _vp)]


def make_anat(a):
    """dict with asource_xyz, asource_tri, thi_xyz, thi_tri, atarget_xyz, face_ptr, face_ids, bary_ptr, bary_key, bary_w -> (Anat, keep-alive dict)"""
    keep = dict(asource_xyz=_f64(a["asource_xyz"]), asource_tri=_i32(a["asource_tri"]), thi_xyz=_f64(a["thi_xyz"]), thi_tri=_i32(a["thi_tri"]),
                atarget_xyz=_f64(a["atarget_xyz"]), face_ptr=_i32(a["face_ptr"]), face_ids=_i32(a["face_ids"]), bary_ptr=_i32(a["bary_ptr"]),
                bary_key=_i32(a["bary_key"]), bary_w=_f64(a["bary_w"]))
    A = Anat(len(keep["asource_xyz"]), _p(keep["asource_xyz"]), len(keep["asource_tri"]), _p(keep["asource_tri"]),
             len(keep["thi_xyz"]), _p(keep["thi_xyz"]), len(keep["thi_tri"]), _p(keep["thi_tri"]), _p(keep["atarget_xyz"]),
             _p(keep["face_ptr"]), _p(keep["face_ids"]), _p(keep["bary_ptr"]), _p(keep["bary_key"]), _p(keep["bary_w"]))
    return A, keep


class Oracle:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build(ref=False)
            L = C.CDLL(ORACLE_SO)
            L.orc_octree_build.restype = _vp
            L.orc_octree_build.argtypes = [_i, _vp, _i, _vp]
            L.orc_octree_free.argtypes = [_vp]
            L.orc_octree_dump.argtypes = [_vp, _vp, _vp, _i, _vp, _i, _vp]
            L.orc_octree_query.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i]
            L.orc_bary_weights.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _i]
            L.orc_vertex_areas.argtypes = [_i, _vp, _i, _vp, _vp]
            L.orc_adaptive_weights.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i]
            L.orc_metric_resample.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i]
            L.orc_bary_resample.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i]
            L.orc_sphere_project_warp.argtypes = [_i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i]
            L.orc_surface_resample.argtypes = [_i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i]
            L.orc_nn_resample.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i]
            L.orc_rotation_matrix.argtypes = [_vp, _vp, _vp]
            for f in (L.orc_corr, L.orc_ssd):
                f.restype = _d
                f.argtypes = [_i, _vp, _vp, _vp]
            L.orc_sim_for_min.restype = _d
            L.orc_sim_for_min.argtypes = [_i, _i, _vp, _vp, _vp]
            L.orc_patch_membership.argtypes = [_i, _vp, _i, _vp, _vp, _d, _vp, _vp, _i, _i]
            L.orc_unary_costs.argtypes = [_i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp,
                                          _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i]
            L.orc_group_fields.argtypes = [_i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i]
            L.orc_group_pair_costs.argtypes = [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _vp, _i, _vp, _vp, _vp, _vp, _i]
            L.orc_group_pair_costs_masked.argtypes = [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i]
            L.orc_ho_patches.argtypes = [_i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i]
            L.orc_rigid.restype = _i
            L.orc_rigid.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _i]
            L.orc_triplet_costs.argtypes = [_i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp,
                                            _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _d, _d, _d, _d, _d, _vp, _i]
            L.orc_triplet_costs_anat.argtypes = [_i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp,
                                                 _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _d, _d, _d, _d, _d, _i, _vp, _vp, _i]
            cls._lib = L
        return cls._lib


class OracleOctree:
    """octree.cpp restated; see oracle/msm_oracle.h."""

    def __init__(self, xyz, tri):
        self.xyz, self.tri = _f64(xyz), _i32(tri)
        self.L = Oracle.lib()
        self.h = self.L.orc_octree_build(len(self.xyz), _p(self.xyz), len(self.tri), _p(self.tri))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_octree_free(self.h)
            self.h = None

    def dump(self):
        nt = C.c_int()
        n = self.L.orc_octree_dump(self.h, None, None, 0, None, 0, C.byref(nt))
        kinds, counts, tris = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(nt.value, np.int32)
        self.L.orc_octree_dump(self.h, _p(kinds), _p(counts), n, _p(tris), nt.value, C.byref(nt))
        return kinds, counts, tris

    def query(self, pts, nthreads=8):
        pts = _f64(pts)
        n = len(pts)
        tri, vtx = np.zeros(n, np.int32), np.zeros(n, np.int32)
        st, path = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.orc_octree_query(self.h, n, _p(pts), _p(tri), _p(vtx), _p(st), _p(path), nthreads)
        return tri, vtx, st, path

    def bary_weights(self, pts, nthreads=8):
        pts = _f64(pts)
        n = len(pts)
        idx, w, ne = np.zeros((n, 3), np.int32), np.zeros((n, 3)), np.zeros(n, np.int32)
        err = self.L.orc_bary_weights(self.h, n, _p(pts), _p(idx), _p(w), _p(ne), nthreads)
        return idx, w, ne, err


def oracle_vertex_areas(xyz, tri):
    xyz, tri = _f64(xyz), _i32(tri)
    out = np.zeros(len(xyz))
    Oracle.lib().orc_vertex_areas(len(xyz), _p(xyz), len(tri), _p(tri), _p(out))
    return out


def _csr_call(fn, nrows, *args):
    rowptr = np.zeros(nrows + 1, np.int32)
    nnz = fn(*args, _p(rowptr), None, None, 0)
    if nnz < 0:
        raise RuntimeError("query failed")
    col, val = np.zeros(nnz, np.int32), np.zeros(nnz)
    fn(*args, _p(rowptr), _p(col), _p(val), nnz)
    return rowptr, col, val


def oracle_adaptive_weights(xyz_in, tri_in, xyz_low, tri_low):
    a, b, c, d = _f64(xyz_in), _i32(tri_in), _f64(xyz_low), _i32(tri_low)
    return _csr_call(Oracle.lib().orc_adaptive_weights, len(c), len(a), _p(a), len(b), _p(b), len(c), _p(c), len(d), _p(d))


def oracle_metric_resample(xyz_in, tri_in, xyz_low, tri_low, feat, nthreads=8):
    a, b, c, d, f = _f64(xyz_in), _i32(tri_in), _f64(xyz_low), _i32(tri_low), _f64(feat)
    out = np.zeros((f.shape[0], len(c)))
    e = Oracle.lib().orc_metric_resample(len(a), _p(a), len(b), _p(b), len(c), _p(c), len(d), _p(d),
                                         f.shape[0], _p(f), _p(out), nthreads)
    if e:
        raise RuntimeError(f"oracle metric_resample failed ({e})")
    return out


def oracle_bary_resample(xyz_in, tri_in, xyz_low, feat, nthreads=8):
    a, b, c, f = _f64(xyz_in), _i32(tri_in), _f64(xyz_low), _f64(feat)
    out = np.zeros((f.shape[0], len(c)))
    e = Oracle.lib().orc_bary_resample(len(a), _p(a), len(b), _p(b), len(c), _p(c), f.shape[0], _p(f), _p(out), nthreads)
    if e:
        raise RuntimeError(f"oracle bary_resample failed ({e})")
    return out


def oracle_metric_resample_excl(xyz_in, tri_in, xyz_low, tri_low, feat, excl):
    """metric_resample with an exclusion mask -> (out [D][n_low], resampled mask, (rowptr, col, val) of the masked weights)."""
    a, b, c, d = _f64(xyz_in), _i32(tri_in), _f64(xyz_low), _i32(tri_low)
    f, e = _f64(np.atleast_2d(feat)), _f64(excl)
    out, eo = np.zeros((f.shape[0], len(c))), np.zeros(len(c))
    rowptr = np.zeros(len(c) + 1, np.int32)
    cap = 64 * max(len(a), len(c))
    col, val = np.zeros(cap, np.int32), np.zeros(cap)
    L = Oracle.lib()
    L.orc_metric_resample_excl.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]
    n = L.orc_metric_resample_excl(len(a), _p(a), len(b), _p(b), len(c), _p(c), len(d), _p(d), f.shape[0], _p(f), _p(e), _p(out), _p(eo),
                                   _p(rowptr), _p(col), _p(val), cap)
    if n < 0 or n > cap:
        raise RuntimeError("oracle metric_resample (EXCL) failed")
    return out, eo, (rowptr, col[:n].copy(), val[:n].copy())


def oracle_nn_resample_excl(low, xyz, tri, feat, excl):
    a, b, c, f, e = _f64(low), _f64(xyz), _i32(tri), _f64(np.atleast_2d(feat)), _f64(excl)
    out, eo = np.zeros((f.shape[0], len(a))), np.zeros(len(a))
    L = Oracle.lib()
    L.orc_nn_resample_excl.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]
    if L.orc_nn_resample_excl(len(a), _p(a), len(b), _p(b), len(c), _p(c), f.shape[0], _p(f), _p(e), _p(out), _p(eo)):
        raise RuntimeError("oracle nn resample (EXCL) failed")
    return out, eo


def oracle_smooth_data(xyz, tri, low, sigma, feat, excl=None):
    a, b, c, f = _f64(xyz), _i32(tri), _f64(low), _f64(np.atleast_2d(feat))
    e = _f64(excl) if excl is not None else None
    out, eo = np.zeros((f.shape[0], len(c))), np.zeros(len(c))
    L = Oracle.lib()
    L.orc_smooth_data.argtypes = [_i, _vp, _i, _vp, _i, _vp, _d, _i, _vp, _vp, _vp, _vp]
    if L.orc_smooth_data(len(a), _p(a), len(b), _p(b), len(c), _p(c), float(sigma), f.shape[0], _p(f), _p(e) if e is not None else None, _p(out), _p(eo)):
        raise RuntimeError("oracle smooth_data failed")
    return out, (eo if e is not None else None)


def oracle_sphere_project_warp(sphere, from_xyz, tri, to_xyz, nthreads=8):
    s, a, t, b = _f64(sphere), _f64(from_xyz), _i32(tri), _f64(to_xyz)
    out = np.zeros_like(s)
    e = Oracle.lib().orc_sphere_project_warp(len(s), _p(s), len(a), _p(a), len(t), _p(t), _p(b), _p(out), nthreads)
    if e:
        raise RuntimeError("oracle sphere_project_warp failed")
    return out


def oracle_surface_resample(low, sph, tri, anat, nthreads=8):
    s, a, t, b = _f64(low), _f64(sph), _i32(tri), _f64(anat)
    out = np.zeros_like(s)
    e = Oracle.lib().orc_surface_resample(len(s), _p(s), len(a), _p(a), len(t), _p(t), _p(b), _p(out), nthreads)
    if e:
        raise RuntimeError("oracle surface_resample failed")
    return out


def oracle_nn_resample(low, xyz, tri, feat, nthreads=8):
    s, a, t, f = _f64(low), _f64(xyz), _i32(tri), _f64(feat)
    out = np.zeros((f.shape[0], len(s)))
    e = Oracle.lib().orc_nn_resample(len(s), _p(s), len(a), _p(a), len(t), _p(t), f.shape[0], _p(f), _p(out), nthreads)
    if e:
        raise RuntimeError("oracle nn_resample failed")
    return out


def oracle_rotation_matrix(ci, index):
    a, b, R = _f64(ci), _f64(index), np.zeros(9)
    Oracle.lib().orc_rotation_matrix(_p(a), _p(b), _p(R))
    return R.reshape(3, 3)


def oracle_sim(simmeasure, A, B, w):
    A, B, w = _f64(A), _f64(B), _f64(w)
    return Oracle.lib().orc_sim_for_min(simmeasure, len(A), _p(A), _p(B), _p(w))


def oracle_set_percentile(p):
    Oracle.lib().orc_set_percentile(C.c_double(p))


def oracle_patch_membership(cp_xyz, src_xyz, maxsep, rng, nthreads=8):
    cp, s, ms = _f64(cp_xyz), _f64(src_xyz), _f64(maxsep)
    rowptr = np.zeros(len(cp) + 1, np.int32)
    L = Oracle.lib()
    n = L.orc_patch_membership(len(cp), _p(cp), len(s), _p(s), _p(ms), rng, _p(rowptr), None, 0, nthreads)
    mem = np.zeros(n, np.int32)
    L.orc_patch_membership(len(cp), _p(cp), len(s), _p(s), _p(ms), rng, _p(rowptr), _p(mem), n, nthreads)
    return rowptr, mem


def oracle_unary_costs(kind, simmeasure, tree: OracleOctree, cp_xyz, rot, labels, src_xyz, prow, pmem,
                       src_feat, ref_feat, cfw, absw, want_tri=False, nthreads=8):
    cp, rot, labels, src = _f64(cp_xyz), _f64(rot), _f64(labels), _f64(src_xyz)
    prow, pmem = _i32(prow), _i32(pmem)
    sf, rf, absw = _f64(np.atleast_2d(src_feat)), _f64(np.atleast_2d(ref_feat)), _f64(absw)
    D = sf.shape[0]
    cfw_rows = 0 if cfw is None else np.atleast_2d(cfw).shape[0]
    cfw_a = None if cfw is None else _f64(np.atleast_2d(cfw))
    Lb = len(labels)
    out = np.zeros((Lb, len(cp)))
    tri_out = np.zeros((Lb, int(prow[-1])), np.int32) if want_tri else None
    e = Oracle.lib().orc_unary_costs(kind, simmeasure, tree.h, len(cp), _p(cp), _p(rot), Lb, _p(labels),
                                     len(src), _p(src), _p(prow), _p(pmem), D, _p(sf), _p(rf),
                                     cfw_rows, _p(cfw_a), _p(absw), _p(out), _p(tri_out), nthreads)
    if e:
        raise RuntimeError("oracle unary costs: a query failed")
    return (out, tri_out) if want_tri else out


def oracle_ho_patches(cp_xyz, cp_tri, src_xyz):
    cp, tri, s = _f64(cp_xyz), _i32(cp_tri), _f64(src_xyz)
    rowptr, mem = np.zeros(len(tri) + 1, np.int32), np.zeros(len(s), np.int32)
    n = Oracle.lib().orc_ho_patches(len(cp), _p(cp), len(tri), _p(tri), len(s), _p(s), _p(rowptr), _p(mem), len(s))
    if n < 0:
        raise RuntimeError("oracle HO patches: a query failed")
    return rowptr, mem[:n].copy()


def oracle_triplet_costs(kind, simmeasure, tree, cp_xyz, orig_cp_xyz, rot, labels, triplets, req_t, req_la, req_lb, req_lc,
                         src_xyz, prow, pmem, src_feat, ref_feat, cfw, absw, lambda_, mu=0.4, kappa=1.6, k_exp=2.0, rexp=2.0, nthreads=8, rmode=3, anat=None):
    cp, org, rot, labels = _f64(cp_xyz), _f64(orig_cp_xyz), _f64(rot), _f64(labels)
    trip = _i32(triplets)
    rt, la, lb, lc = _i32(req_t), _i32(req_la), _i32(req_lb), _i32(req_lc)
    src = _f64(src_xyz)
    prow = _i32(prow) if prow is not None else np.zeros(len(trip) + 1, np.int32)
    pmem = _i32(pmem) if pmem is not None else np.zeros(1, np.int32)
    sf, rf, absw = _f64(np.atleast_2d(src_feat)), _f64(np.atleast_2d(ref_feat)), _f64(absw)
    cfw_rows = 0 if cfw is None else np.atleast_2d(cfw).shape[0]
    cfw_a = None if cfw is None else _f64(np.atleast_2d(cfw))
    out = np.zeros(len(rt))
    A, keep = make_anat(anat) if anat is not None else (None, None)
    e = Oracle.lib().orc_triplet_costs_anat(kind, simmeasure, tree.h if tree is not None else None, len(cp), _p(cp), _p(org), _p(rot), len(labels), _p(labels),
                                            len(trip), _p(trip), len(rt), _p(rt), _p(la), _p(lb), _p(lc), len(src), _p(src), _p(prow), _p(pmem),
                                            sf.shape[0], _p(sf), _p(rf), cfw_rows, _p(cfw_a), _p(absw), lambda_, mu, kappa, k_exp, rexp, int(rmode),
                                            C.byref(A) if A is not None else None, _p(out), nthreads)
    if e:
        raise RuntimeError("oracle triplet costs: a query failed")
    return out


def oracle_group_fields(data_xyz, tri, feat, labels, centre, tpl_xyz, tpl_tri, nthreads=8):
    """-> [S][L][D][n_tpl] (channel-major per (subject,label))"""
    xyz, tri, feat, labels, centre = _f64(data_xyz), _i32(tri), _f64(feat), _f64(labels), _f64(centre)
    tx, tt = _f64(tpl_xyz), _i32(tpl_tri)
    S, nv, D, L = xyz.shape[0], xyz.shape[1], feat.shape[1], len(labels)
    out = np.zeros((S, L, D, len(tx)))
    e = Oracle.lib().orc_group_fields(S, nv, _p(xyz), len(tri), _p(tri), D, _p(feat), L, _p(labels), _p(centre), len(tx), _p(tx), len(tt), _p(tt),
                                      _p(out), nthreads)
    if e:
        raise RuntimeError("oracle group fields: a query failed")
    return out


def oracle_group_pair_costs(simmeasure, ncp, tpl_xyz, fields, rot, labels, spacings, range_, pairs, req_pair, req_la, req_lb, nthreads=8, mask=None):
    tx, fields, rot, labels, sp = _f64(tpl_xyz), _f64(fields), _f64(rot), _f64(labels), _f64(spacings).reshape(-1)
    pairs, rp, la, lb = _i32(pairs), _i32(req_pair), _i32(req_la), _i32(req_lb)
    S, L, D = fields.shape[0], fields.shape[1], fields.shape[2]
    out = np.zeros(len(rp))
    mask = None if mask is None else _f64(mask).reshape(-1)      # cost mask on the template (DiscreteGroupCostFunction.cpp:77)
    Oracle.lib().orc_group_pair_costs_masked(simmeasure, S, ncp, L, D, len(tx), _p(tx), _p(fields), _p(rot), _p(labels), _p(sp), float(range_), _p(pairs),
                                             len(rp), _p(rp), _p(la), _p(lb), _p(mask), _p(out), nthreads)
    return out


def oracle_group_triplet_costs(cps, orig_cps, rot, labels, triplets, req_t, req_la, req_lb, req_lc, lambda_, mu=0.4, kappa=1.6, k_exp=2.0, rexp=2.0):
    """DiscreteGroupCostFunction::computeTripletCost (cpp:26-52) from the pairwise-model restatement: no likelihood, lambda' = subcorr * lambda
    (the reference multiplies left to right: (subcorr * lambda) * W^rexp), FOLDING instead of FOLDING * lambda."""
    cp, org = _f64(cps), _f64(orig_cps)
    S = cp.shape[0]
    lam = (0.1 * S) * lambda_
    n_nodes = cp.shape[0] * cp.shape[1]
    dummy = np.zeros((1, n_nodes))
    out = oracle_triplet_costs(0, 2, None, cp.reshape(-1, 3), org.reshape(-1, 3), rot, labels, triplets, req_t, req_la, req_lb, req_lc,
                               cp.reshape(-1, 3), None, None, dummy, dummy, None, np.ones(n_nodes), lam, mu, kappa, k_exp, rexp)
    out[out == 1e7 * lam] = 1e7
    return out


def oracle_rigid(tgt_xyz, tgt_tri, src_xyz, src_tri, src_feat, ref_feat, simmeasure=2, iters=4, stepsize=0.01, gradsampling=0.5):
    """Restatement of the RIGID / AFFINE level (rigid_costfunction.cpp:32-236); same outputs as refmr_rigid."""
    tx, tt, sx, st = _f64(tgt_xyz), _i32(tgt_tri), _f64(src_xyz), _i32(src_tri)
    sf, rf = _f64(np.atleast_2d(src_feat)), _f64(np.atleast_2d(ref_feat))
    out = np.zeros((len(sx), 3))
    cost0 = C.c_double(0.0)
    rowptr = np.zeros(len(sx) + 1, np.int32)
    cap = 512 * len(sx)
    mem = np.zeros(cap, np.int32)
    n = Oracle.lib().orc_rigid(len(tx), _p(tx), len(tt), _p(tt), len(sx), _p(sx), len(st), _p(st), sf.shape[0], _p(sf), _p(rf), int(simmeasure),
                               int(iters), float(stepsize), float(gradsampling), _p(out), C.cast(C.byref(cost0), C.c_void_p), _p(rowptr), _p(mem), cap)
    if n < 0 or n > cap:
        raise RuntimeError("oracle rigid level failed")
    return out, cost0.value, rowptr, mem[:n].copy()


# --------------------------------------------------------------------------------------
# compiled reference (oracle/_ref)
# --------------------------------------------------------------------------------------
class Ref:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not have_ref():
                raise RuntimeError("oracle/_ref/libref_newresampler.so not built (needs /root/reference; run `make -C oracle ref`)")
            L = C.CDLL(REF_SO)
            L.ref_mesh_new.restype = _vp
            L.ref_mesh_new.argtypes = [_i, _vp, _i, _vp]
            L.ref_mesh_icosa.restype = _vp
            L.ref_mesh_icosa.argtypes = [_i, _d]
            L.ref_mesh_free.argtypes = [_vp]
            L.ref_mesh_nvertices.argtypes = [_vp]
            L.ref_mesh_ntriangles.argtypes = [_vp]
            L.ref_mesh_export.argtypes = [_vp, _vp, _vp]
            L.ref_mesh_set_pvalues.argtypes = [_vp, _i, _vp]
            L.ref_vertex_areas.argtypes = [_vp, _vp]
            L.ref_octree_new.restype = _vp
            L.ref_octree_new.argtypes = [_vp]
            L.ref_octree_free.argtypes = [_vp]
            L.ref_octree_query.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _i]
            L.ref_octree_dump.argtypes = [_vp, _vp, _vp, _i, _vp, _i, _vp]
            L.ref_bary_weights.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i]
            L.ref_adaptive_weights.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _i]
            L.ref_metric_resample.restype = _d
            L.ref_metric_resample.argtypes = [_vp, _vp, _i, _vp]
            L.ref_bary_resample.restype = _d
            L.ref_bary_resample.argtypes = [_vp, _vp, _i, _vp]
            L.ref_sphere_project_warp.argtypes = [_vp, _vp, _vp, _i, _vp]
            L.ref_surface_resample.argtypes = [_vp, _vp, _vp, _i, _vp]
            L.ref_nn_resample.argtypes = [_vp, _vp, _i, _vp]
            L.ref_rotation_matrix.argtypes = [_vp, _vp, _vp]
            L.ref_metric_resample_excl.argtypes = [_vp, _vp, _i, _vp, _vp, _vp]
            L.ref_adaptive_weights_excl.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i]
            L.ref_smooth_data.argtypes = [_vp, _vp, _d, _i, _vp, _vp, _vp]
            L.ref_nn_resample_excl.argtypes = [_vp, _vp, _i, _vp, _vp, _vp]
            cls._lib = L
        return cls._lib


class RefMesh:
    def __init__(self, xyz=None, tri=None, icosa=None, radius=100.0, feat=None):
        self.L = Ref.lib()
        if icosa is not None:
            self.h = self.L.ref_mesh_icosa(int(icosa), float(radius))
        else:
            xyz, tri = _f64(xyz), _i32(tri)
            self.h = self.L.ref_mesh_new(len(xyz), _p(xyz), len(tri), _p(tri))
        self.nv = self.L.ref_mesh_nvertices(self.h)
        self.nt = self.L.ref_mesh_ntriangles(self.h)
        if feat is not None:
            self.set_features(feat)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_mesh_free(self.h)
            self.h = None

    def export(self):
        xyz, tri = np.zeros((self.nv, 3)), np.zeros((self.nt, 3), np.int32)
        self.L.ref_mesh_export(self.h, _p(xyz), _p(tri))
        return xyz, tri

    def set_features(self, feat):
        f = _f64(np.atleast_2d(feat))
        assert f.shape[1] == self.nv
        self.D = f.shape[0]
        self.L.ref_mesh_set_pvalues(self.h, f.shape[0], _p(f))

    def vertex_areas(self):
        out = np.zeros(self.nv)
        self.L.ref_vertex_areas(self.h, _p(out))
        return out


class RefOctree:
    def __init__(self, mesh: RefMesh):
        self.mesh = mesh
        self.L = Ref.lib()
        self.h = self.L.ref_octree_new(mesh.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_octree_free(self.h)
            self.h = None

    def dump(self):
        nt = C.c_int()
        n = self.L.ref_octree_dump(self.h, None, None, 0, None, 0, C.byref(nt))
        kinds, counts, tris = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(nt.value, np.int32)
        self.L.ref_octree_dump(self.h, _p(kinds), _p(counts), n, _p(tris), nt.value, C.byref(nt))
        return kinds, counts, tris

    def query(self, pts, nthreads=8):
        pts = _f64(pts)
        n = len(pts)
        tri, vtx, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.ref_octree_query(self.h, n, _p(pts), _p(tri), _p(vtx), _p(st), nthreads)
        return tri, vtx, st

    def bary_weights(self, low: RefMesh, nthreads=8):
        n = low.nv
        idx, w, ne = np.zeros((n, 3), np.int32), np.zeros((n, 3)), np.zeros(n, np.int32)
        err = self.L.ref_bary_weights(low.h, self.mesh.h, self.h, _p(idx), _p(w), _p(ne), nthreads)
        return idx, w, ne, err


def ref_adaptive_weights(m_in: RefMesh, m_low: RefMesh, nthreads=1):
    L = Ref.lib()
    rowptr = np.zeros(m_low.nv + 1, np.int32)
    cap = 64 * max(m_low.nv, m_in.nv)
    col, val = np.zeros(cap, np.int32), np.zeros(cap)
    nnz = L.ref_adaptive_weights(m_in.h, m_low.h, nthreads, _p(rowptr), _p(col), _p(val), cap)
    if nnz < 0 or nnz > cap:
        raise RuntimeError("reference adaptive weights failed")
    return rowptr, col[:nnz].copy(), val[:nnz].copy()


def ref_metric_resample(m_in: RefMesh, m_low: RefMesh, nthreads=1, want_out=True):
    out = np.zeros((m_in.D, m_low.nv)) if want_out else None
    secs = Ref.lib().ref_metric_resample(m_in.h, m_low.h, nthreads, _p(out))
    if secs < 0:
        raise RuntimeError("reference metric_resample failed")
    return out, secs


def ref_bary_resample(m_in: RefMesh, m_low: RefMesh, nthreads=1, want_out=True):
    out = np.zeros((m_in.D, m_low.nv)) if want_out else None
    secs = Ref.lib().ref_bary_resample(m_in.h, m_low.h, nthreads, _p(out))
    if secs < 0:
        raise RuntimeError("reference bary_resample failed")
    return out, secs


def ref_sphere_project_warp(sphere: RefMesh, m_from: RefMesh, m_to: RefMesh, nthreads=1):
    out = np.zeros((sphere.nv, 3))
    if Ref.lib().ref_sphere_project_warp(sphere.h, m_from.h, m_to.h, nthreads, _p(out)):
        raise RuntimeError("reference sphere_project_warp failed")
    return out


def ref_surface_resample(anat: RefMesh, sph: RefMesh, low: RefMesh, nthreads=1):
    out = np.zeros((low.nv, 3))
    if Ref.lib().ref_surface_resample(anat.h, sph.h, low.h, nthreads, _p(out)):
        raise RuntimeError("reference surface_resample failed")
    return out


def ref_nn_resample(m_in: RefMesh, m_low: RefMesh, nthreads=1):
    out = np.zeros((m_in.D, m_low.nv))
    if Ref.lib().ref_nn_resample(m_in.h, m_low.h, nthreads, _p(out)):
        raise RuntimeError("reference nn resample failed")
    return out


def ref_metric_resample_excl(m_in: RefMesh, m_low: RefMesh, excl, nthreads=1):
    """metric_resample with an exclusion mask (resampler.cpp:30-70): (out [D][n_low], resampled mask [n_low])."""
    e = _f64(excl)
    out, eo = np.zeros((m_in.D, m_low.nv)), np.zeros(m_low.nv)
    if Ref.lib().ref_metric_resample_excl(m_in.h, m_low.h, nthreads, _p(e), _p(out), _p(eo)):
        raise RuntimeError("reference metric_resample (EXCL) failed")
    return out, eo


def ref_adaptive_weights_excl(m_in: RefMesh, m_low: RefMesh, excl, nthreads=1):
    L = Ref.lib()
    e = _f64(excl)
    rowptr = np.zeros(m_low.nv + 1, np.int32)
    cap = 64 * max(m_low.nv, m_in.nv)
    col, val = np.zeros(cap, np.int32), np.zeros(cap)
    nnz = L.ref_adaptive_weights_excl(m_in.h, m_low.h, nthreads, _p(e), _p(rowptr), _p(col), _p(val), cap)
    if nnz < 0 or nnz > cap:
        raise RuntimeError("reference adaptive weights (EXCL) failed")
    return rowptr, col[:nnz].copy(), val[:nnz].copy()


def ref_smooth_data(m_orig: RefMesh, m_low: RefMesh, sigma, excl=None, nthreads=1):
    """smooth_data (resampler.cpp:169-230), optional exclusion mask: (out [D][n_low], new mask or None)."""
    out, eo = np.zeros((m_orig.D, m_low.nv)), np.zeros(m_low.nv)
    e = _f64(excl) if excl is not None else None
    if Ref.lib().ref_smooth_data(m_orig.h, m_low.h, float(sigma), nthreads, _p(e) if e is not None else None, _p(out), _p(eo)):
        raise RuntimeError("reference smooth_data failed")
    return out, (eo if e is not None else None)


def ref_nn_resample_excl(m_in: RefMesh, m_low: RefMesh, excl, nthreads=1):
    e = _f64(excl)
    out, eo = np.zeros((m_in.D, m_low.nv)), np.zeros(m_low.nv)
    if Ref.lib().ref_nn_resample_excl(m_in.h, m_low.h, nthreads, _p(e), _p(out), _p(eo)):
        raise RuntimeError("reference nn resample (EXCL) failed")
    return out, eo


def ref_rotation_matrix(ci, index):
    a, b, R = _f64(ci), _f64(index), np.zeros(9)
    Ref.lib().ref_rotation_matrix(_p(a), _p(b), _p(R))
    return R.reshape(3, 3)


# --------------------------------------------------------------------------------------
# compiled reference, registration library (oracle/_ref/libref_newmeshreg.so, oracle/ref_meshreg_driver.cpp)
# --------------------------------------------------------------------------------------
class RefMR:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not have_refmr():
                raise RuntimeError("oracle/_ref/libref_newmeshreg.so not built (needs /root/reference; run `make -C oracle refmr`)")
            L = C.CDLL(REFMR_SO)
            for fn in ("refmr_unary", "refmr_triplet", "refmr_pairwise_reg", "refmr_group_pair_costs", "refmr_group_triplet_costs", "refmr_label_sets"):
                getattr(L, fn).restype = _i
            L.refmr_unary.argtypes = [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp,
                                      _i, _vp, _vp, _vp, _d, _vp, _vp, _vp, _i, _vp, _i]
            L.refmr_triplet.argtypes = [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp,
                                        _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp,
                                        _d, _d, _d, _d, _d, _i, _vp, _vp, _vp, _i, _i]
            if hasattr(L, "refmr_triplet_anat"):
                L.refmr_triplet_anat.restype = _i
                L.refmr_triplet_anat.argtypes = [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp,
                                                 _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp,
                                                 _d, _d, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _i, _i]
            L.refmr_pairwise_reg.argtypes = [_i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _d, _d, _d, _i, _vp, _vp, _vp, _vp]
            L.refmr_group_pair_costs.argtypes = [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _d, _i, _vp,
                                                 _i, _vp, _vp, _vp, _vp, _vp, _i]
            if hasattr(L, "refmr_group_pair_costs_masked"):
                L.refmr_group_pair_costs_masked.restype = _i
                L.refmr_group_pair_costs_masked.argtypes = [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _d, _i, _vp,
                                                            _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]
            L.refmr_group_triplet_costs.argtypes = [_i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _d, _d, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp]
            L.refmr_label_sets.argtypes = [_i, _d, _vp, _vp, _i, _vp, _vp]
            if hasattr(L, "refmr_rigid"):
                L.refmr_rigid.restype = _i
                L.refmr_rigid.argtypes = [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _i]
            cls._lib = L
        return cls._lib


def refmr_variance_normalise(data, excl=None, nthreads=1):
    """The reference's newmeshreg::variance_normalise (reg_tools.cpp:804-844) on [D][n] -> normalised copy."""
    L = RefMR.lib()
    L.refmr_variance_normalise.restype = _i
    L.refmr_variance_normalise.argtypes = [_i, _i, _vp, _vp, _i]
    d = _f64(np.atleast_2d(data)).copy()
    e = None if excl is None else _f64(excl)
    if L.refmr_variance_normalise(d.shape[0], d.shape[1], _p(d), _p(e), nthreads) != 0:
        raise RuntimeError("reference variance_normalise failed")
    return d


def oracle_variance_normalise(data, excl=None):
    L = Oracle.lib()
    L.orc_variance_normalise.restype = None
    L.orc_variance_normalise.argtypes = [_i, _i, _vp, _vp]
    d = _f64(np.atleast_2d(data)).copy()
    e = None if excl is None else _f64(excl)
    L.orc_variance_normalise(d.shape[0], d.shape[1], _p(d), _p(e))
    return d


def refmr_set_percentile(p):
    RefMR.lib().refmr_set_percentile(C.c_double(p))


def refmr_unary(kind, simmeasure, tgt_xyz, tgt_tri, cp_xyz, cp_tri, rot, labels, src_xyz, src_tri, src_feat, ref_feat, cfw, absw, maxsep, range_,
                want_costs=True, nthreads=1):
    """The reference's get_source_data() + computeUnaryCosts(). -> (costs [L][ncp] | None, rowptr, members, AbsoluteWeights used)"""
    tx, tt, cp, ct, rot, labels = _f64(tgt_xyz), _i32(tgt_tri), _f64(cp_xyz), _i32(cp_tri), _f64(rot), _f64(labels)
    sx, st = _f64(src_xyz), _i32(src_tri)
    sf, rf = _f64(np.atleast_2d(src_feat)), _f64(np.atleast_2d(ref_feat))
    cfw_rows = 0 if cfw is None else np.atleast_2d(cfw).shape[0]
    cfw_a = None if cfw is None else _f64(np.atleast_2d(cfw))
    absw_a = None if absw is None else _f64(absw)
    ms = _f64(maxsep)
    out = np.zeros((len(labels), len(cp))) if want_costs else None
    rowptr = np.zeros(len(cp) + 1, np.int32)
    cap = 64 * len(sx)
    mem = np.zeros(cap, np.int32)
    absw_out = np.zeros(len(cp))
    n = RefMR.lib().refmr_unary(kind, simmeasure, len(tx), _p(tx), len(tt), _p(tt), len(cp), _p(cp), len(ct), _p(ct), _p(rot), len(labels), _p(labels),
                                len(sx), _p(sx), len(st), _p(st), sf.shape[0], _p(sf), _p(rf), cfw_rows, _p(cfw_a), _p(absw_a), _p(ms), float(range_),
                                _p(out), _p(rowptr), _p(mem), cap, _p(absw_out), nthreads)
    if n < 0 or n > cap:
        raise RuntimeError("reference unary costs failed")
    return out, rowptr, mem[:n].copy(), absw_out


def refmr_rigid(tgt_xyz, tgt_tri, src_xyz, src_tri, src_feat, ref_feat, simmeasure=2, iters=4, stepsize=0.01, gradsampling=0.5, nthreads=1):
    """The reference's RIGID / AFFINE level (Rigid_cost_function::initialise + run, rigid_costfunction.cpp:32-236).
    -> (rotated source coords [nv_s][3], cost at zero rotation, neighbour rowptr, neighbour members)"""
    tx, tt, sx, st = _f64(tgt_xyz), _i32(tgt_tri), _f64(src_xyz), _i32(src_tri)
    sf, rf = _f64(np.atleast_2d(src_feat)), _f64(np.atleast_2d(ref_feat))
    out = np.zeros((len(sx), 3))
    cost0 = C.c_double(0.0)
    rowptr = np.zeros(len(sx) + 1, np.int32)
    cap = 512 * len(sx)
    mem = np.zeros(cap, np.int32)
    n = RefMR.lib().refmr_rigid(len(tx), _p(tx), len(tt), _p(tt), len(sx), _p(sx), len(st), _p(st), sf.shape[0], _p(sf), _p(rf), int(simmeasure),
                                int(iters), float(stepsize), float(gradsampling), int(nthreads), _p(out), C.cast(C.byref(cost0), C.c_void_p), _p(rowptr),
                                _p(mem), cap)
    if n < 0 or n > cap:
        raise RuntimeError("reference rigid level failed")
    return out, cost0.value, rowptr, mem[:n].copy()


def refmr_triplet(kind, simmeasure, tgt_xyz, tgt_tri, cp_xyz, cp_tri, orig_cp_xyz, rot, labels, triplets, req_t, req_la, req_lb, req_lc,
                  src_xyz, src_tri, src_feat, ref_feat, cfw, absw, lambda_, mu=0.4, kappa=1.6, k_exp=2.0, rexp=2.0, rmode=3, nthreads=8, anat=None):
    """The reference's computeTripletCost for a request list. -> (costs [n], HO patch rowptr, members). anat: the inputs of regoption 4/5."""
    tx, tt, cp, ct, org, rot, labels = _f64(tgt_xyz), _i32(tgt_tri), _f64(cp_xyz), _i32(cp_tri), _f64(orig_cp_xyz), _f64(rot), _f64(labels)
    trip, rt, la, lb, lc = _i32(triplets), _i32(req_t), _i32(req_la), _i32(req_lb), _i32(req_lc)
    sx, st = _f64(src_xyz), _i32(src_tri)
    sf, rf = _f64(np.atleast_2d(src_feat)), _f64(np.atleast_2d(ref_feat))
    cfw_rows = 0 if cfw is None else np.atleast_2d(cfw).shape[0]
    cfw_a = None if cfw is None else _f64(np.atleast_2d(cfw))
    absw_a = None if absw is None else _f64(absw)
    out = np.zeros(len(rt))
    rowptr, mem = np.zeros(len(ct) + 1, np.int32), np.zeros(len(sx), np.int32)
    A, keep = make_anat(anat) if anat is not None else (None, None)
    n = RefMR.lib().refmr_triplet_anat(kind, simmeasure, len(tx), _p(tx), len(tt), _p(tt), len(cp), _p(cp), len(ct), _p(ct), _p(org), _p(rot), len(labels),
                                       _p(labels), len(trip), _p(trip), len(rt), _p(rt), _p(la), _p(lb), _p(lc), len(sx), _p(sx), len(st), _p(st),
                                       sf.shape[0], _p(sf), _p(rf), cfw_rows, _p(cfw_a), _p(absw_a), lambda_, mu, kappa, k_exp, rexp, rmode,
                                       C.byref(A) if A is not None else None, _p(out), _p(rowptr), _p(mem), len(sx), nthreads)
    if n < 0:
        raise RuntimeError("reference triplet costs failed")
    return out, rowptr, mem[:n].copy()


def refmr_pairwise_reg(cp_xyz, cp_tri, rot, labels, pairs, lambda_, rexp, mvdmax, req_pair, req_la, req_lb):
    cp, ct, rot, labels, pairs = _f64(cp_xyz), _i32(cp_tri), _f64(rot), _f64(labels), _i32(pairs)
    rp, la, lb = _i32(req_pair), _i32(req_la), _i32(req_lb)
    out = np.zeros(len(rp))
    if RefMR.lib().refmr_pairwise_reg(len(cp), _p(cp), len(ct), _p(ct), _p(rot), len(labels), _p(labels), len(pairs), _p(pairs), lambda_, rexp, mvdmax,
                                      len(rp), _p(rp), _p(la), _p(lb), _p(out)):
        raise RuntimeError("reference pairwise regulariser failed")
    return out


def refmr_group_pair_costs(simmeasure, data_xyz, tri, feat, labels, centre, tpl_xyz, tpl_tri, ncp, rot, spacings, range_, pairs, req_pair, req_la, req_lb,
                           want_fields=False, nthreads=8, mask=None):
    xyz, tri, feat, labels, centre = _f64(data_xyz), _i32(tri), _f64(feat), _f64(labels), _f64(centre)
    tx, tt, rot, sp, pairs = _f64(tpl_xyz), _i32(tpl_tri), _f64(rot), _f64(spacings).reshape(-1), _i32(pairs)
    rp, la, lb = _i32(req_pair), _i32(req_la), _i32(req_lb)
    S, nv, D, L = xyz.shape[0], xyz.shape[1], feat.shape[1], len(labels)
    out = np.zeros(len(rp))
    fields = np.full((S, L, D, len(tx)), np.nan) if want_fields else None
    mask = None if mask is None else _f64(mask).reshape(-1)
    if RefMR.lib().refmr_group_pair_costs_masked(simmeasure, S, nv, _p(xyz), len(tri), _p(tri), D, _p(feat), L, _p(labels), _p(centre), len(tx), _p(tx),
                                                 len(tt), _p(tt), ncp, _p(rot), _p(sp), float(range_), len(pairs), _p(pairs), len(rp), _p(rp), _p(la),
                                                 _p(lb), _p(mask), _p(out), _p(fields), nthreads):
        raise RuntimeError("reference group pair costs failed")
    return (out, fields) if want_fields else out


def refmr_group_triplet_costs(cp_xyz, orig_xyz, cp_tri, rot, labels, triplets, lambda_, mu, kappa, k_exp, rexp, req_t, req_la, req_lb, req_lc):
    cp, org, ct, rot, labels, trip = _f64(cp_xyz), _f64(orig_xyz), _i32(cp_tri), _f64(rot), _f64(labels), _i32(triplets)
    rt, la, lb, lc = _i32(req_t), _i32(req_la), _i32(req_lb), _i32(req_lc)
    out = np.zeros(len(rt))
    if RefMR.lib().refmr_group_triplet_costs(cp.shape[0], cp.shape[1], _p(cp), _p(org), len(ct), _p(ct), _p(rot), len(labels), _p(labels), len(trip), _p(trip),
                                             lambda_, mu, kappa, k_exp, rexp, len(rt), _p(rt), _p(la), _p(lb), _p(lc), _p(out)):
        raise RuntimeError("reference group triplet costs failed")
    return out


def refmr_label_sets(sgres, maxvd, cap=64):
    """label_sampling_grid: -> (vertex labels [n][3], barycentre labels [m][3], centre [3])"""
    s, b, c = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(3)
    nb = C.c_int(0)
    n = RefMR.lib().refmr_label_sets(sgres, float(maxvd), _p(s), _p(b), cap, C.byref(nb), _p(c))
    if n < 0:
        raise RuntimeError("reference label sets failed")
    return s[:n].copy(), b[:nb.value].copy(), c
