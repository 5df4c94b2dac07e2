"""Host-side mirror of the reference's resampler interface (msm-newresampler/src/resampler.h:38-53,
octree.h:39-59, mesh.h:37-58) on top of the C ABI. Names, argument meaning and error behaviour
follow the reference; the `nthreads` arguments are accepted and ignored (the work runs on the GPU).

Everything here calls libmsmgpu.so; nothing computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, f32, f64, i32, ptr

MeshException = capi.MsmGpuError   # meshException.h:31


class Context:
    """One device + stream (msmgpu_ctx). `stream` = a cudaStream_t integer (e.g. torch's) or None."""

    _default = {}

    def __init__(self, device: int = 0, stream: int | None = None):
        self.L = capi.lib()
        self.h = C.c_void_p()
        check(self.L.msmgpu_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.device = device

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def sync(self):
        check(self.L.msmgpu_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.msmgpu_ctx_stream(self.h) or 0)

    def close(self):
        if getattr(self, "h", None):
            self.L.msmgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mesh:
    """newresampler::Mesh as far as the hot path needs it: coordinates, triangles and channel-major
    `pvalues` [D][V] (mesh.h:44)."""

    def __init__(self, xyz, tri, pvalues=None, ctx: Context | None = None):
        self.ctx = ctx or Context.default()
        self.L = self.ctx.L
        self.xyz, self.tri = f64(xyz), i32(tri)
        self.pvalues = None if pvalues is None else f64(np.atleast_2d(pvalues))
        self.h = C.c_void_p()
        check(self.L.msmgpu_mesh_create(self.ctx.h, len(self.xyz), ptr(self.xyz), len(self.tri), ptr(self.tri), C.byref(self.h)))

    @classmethod
    def from_device(cls, ctx: Context, nv: int, d_xyz, nt: int, d_tri) -> "Mesh":
        m = cls.__new__(cls)
        m.ctx, m.L, m.xyz, m.tri, m.pvalues = ctx, ctx.L, None, None, None
        m.h = C.c_void_p()
        check(m.L.msmgpu_mesh_create_dev(ctx.h, nv, ptr(d_xyz), nt, ptr(d_tri), C.byref(m.h)))
        return m

    @classmethod
    def views_from_device(cls, ctx: Context, nv: int, d_xyz_list, nt: int, d_tri) -> "list[Mesh]":
        """msmgpu_mesh_create_view_batch: meshes that view the caller's device coordinate buffers (no copies, one table allocation)."""
        n = len(d_xyz_list)
        xs = (C.c_void_p * n)(*[x.data_ptr() if hasattr(x, "data_ptr") else int(x) for x in d_xyz_list])
        hs = (C.c_void_p * n)()
        check(ctx.L.msmgpu_mesh_create_view_batch(ctx.h, n, nv, xs, nt, ptr(d_tri), hs))
        out = []
        for i in range(n):
            m = cls.__new__(cls)
            m.ctx, m.L, m.xyz, m.tri, m.pvalues = ctx, ctx.L, None, None, None
            m._keep = (d_xyz_list[i], d_tri)     # the viewed buffers must outlive the mesh
            m.h = C.c_void_p(hs[i])
            out.append(m)
        return out

    def nvertices(self) -> int:
        nv = C.c_int()
        check(self.L.msmgpu_mesh_shape(self.h, C.byref(nv), None))
        return nv.value

    def ntriangles(self) -> int:
        nt = C.c_int()
        check(self.L.msmgpu_mesh_shape(self.h, None, C.byref(nt)))
        return nt.value

    def set_coords(self, xyz):
        self.xyz = f64(xyz)
        check(self.L.msmgpu_mesh_set_coords(self.h, ptr(self.xyz)))

    def vertex_areas(self):
        """compute_vertex_area for every vertex (mesh.cpp:1275)."""
        out = np.zeros(self.nvertices())
        check(self.L.msmgpu_mesh_vertex_areas(self.h, ptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.msmgpu_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Octree:
    """newresampler::Octree (octree.h:39-59), queries batched over points."""

    def __init__(self, mesh: Mesh, _handle=None):
        self.mesh = mesh
        self.L = mesh.L
        if _handle is not None:
            self.h = _handle
        else:
            self.h = C.c_void_p()
            check(self.L.msmgpu_octree_build(mesh.h, C.byref(self.h)))

    @classmethod
    def build_batch(cls, meshes) -> list["Octree"]:
        """One forest for many meshes: every build kernel launch covers all of them."""
        ctx = meshes[0].ctx
        hs = (C.c_void_p * len(meshes))(*[m.h.value for m in meshes])
        out = (C.c_void_p * len(meshes))()
        check(ctx.L.msmgpu_octree_build_batch(ctx.h, len(meshes), hs, out))
        return [cls(m, C.c_void_p(out[i])) for i, m in enumerate(meshes)]

    def stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.L.msmgpu_octree_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def dump(self):
        n_nodes, n_refs, _ = self.stats()
        kinds, counts, tris = np.zeros(n_nodes, np.int32), np.zeros(n_nodes, np.int32), np.zeros(n_refs, np.int32)
        check(self.L.msmgpu_octree_dump(self.h, ptr(kinds), ptr(counts), ptr(tris)))
        return kinds, counts, tris

    def get_closest_triangle(self, pts, want_status=False):
        """Triangle ids for [n,3] points. Like the reference it raises for a point outside the root cube
        (octree.cpp:158) or without any candidate (211) unless want_status."""
        pts = f64(pts).reshape(-1, 3)
        n = len(pts)
        tri, st = np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(self.L.msmgpu_nearest_triangle(self.h, n, ptr(pts), ptr(tri), None, ptr(st) if want_status else None))
        return (tri, st) if want_status else tri

    def get_closest_vertex_ID(self, pts):
        pts = f64(pts).reshape(-1, 3)
        vtx = np.zeros(len(pts), np.int32)
        check(self.L.msmgpu_nearest_triangle(self.h, len(pts), ptr(pts), None, ptr(vtx), None))
        return vtx

    def query(self, pts):
        pts = f64(pts).reshape(-1, 3)
        n = len(pts)
        tri, vtx, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(self.L.msmgpu_nearest_triangle(self.h, n, ptr(pts), ptr(tri), ptr(vtx), ptr(st)))
        return tri, vtx, st

    def close(self):
        if getattr(self, "h", None):
            self.L.msmgpu_octree_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Weights:
    """Device CSR resampling matrix; `rows()` gives the reference's vector<map<int,double>> view."""

    def __init__(self, L, handle, owner=None):
        self.L, self.h = L, handle
        self._owner = owner     # the Context (or a mesh of it): destroying the matrix needs the context's stream alive

    def shape(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int64()
        check(self.L.msmgpu_weights_shape(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def csr(self):
        n_rows, _, nnz = self.shape()
        rowptr, col, val = np.zeros(n_rows + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        check(self.L.msmgpu_weights_export(self.h, ptr(rowptr), ptr(col), ptr(val)))
        return rowptr, col, val

    def rows(self):
        rowptr, col, val = self.csr()
        return [dict(zip(col[a:b].tolist(), val[a:b].tolist())) for a, b in zip(rowptr[:-1], rowptr[1:])]

    def apply_f32_dev(self, D, d_in, d_out):
        check(self.L.msmgpu_weights_apply_f32_dev(self.h, D, ptr(d_in), ptr(d_out)))

    def close(self):
        if getattr(self, "h", None):
            self.L.msmgpu_weights_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Resampler:
    """newresampler::Resampler (resampler.h:38-53)."""

    def get_barycentric_weights(self, low: Mesh, orig: Mesh, oct: Octree, nthreads: int = 1):
        """resampler.cpp:142-167 -> (idx [n,3] ascending ids, w [n,3], n_entries [n])."""
        n = low.nvertices()
        idx, w, ne = np.zeros((n, 3), np.int32), np.zeros((n, 3)), np.zeros(n, np.int32)
        check(oct.L.msmgpu_bary_weights(oct.h, n, ptr(low.xyz), ptr(idx), ptr(w), ptr(ne)))
        return idx, w, ne

    def get_adaptive_barycentric_weights(self, in_mesh: Mesh, sphLow: Mesh, EXCL=None, nthreads: int = 1) -> Weights:
        """resampler.cpp:72-140 (EXCL masks are not on the accelerated path)."""
        if EXCL is not None:
            raise NotImplementedError("exclusion masks are outside the accelerated path (SURVEY §8)")
        h = C.c_void_p()
        check(in_mesh.L.msmgpu_adaptive_weights(in_mesh.h, sphLow.h, C.byref(h)))
        return Weights(in_mesh.L, h, owner=in_mesh.ctx)

    def get_adaptive_barycentric_weights_batch(self, in_meshes, sphLow: Mesh, in_trees=None, low_tree: Octree | None = None):
        """The same for a batch of subjects resampled onto one target: one set of launches (msmgpu_adaptive_weights_batch)."""
        n = len(in_meshes)
        ctx = sphLow.ctx
        mh = (C.c_void_p * n)(*[m.h.value for m in in_meshes])
        th = (C.c_void_p * n)(*[(t.h.value if t is not None else None) for t in in_trees]) if in_trees else None
        out = (C.c_void_p * n)()
        check(ctx.L.msmgpu_adaptive_weights_batch(ctx.h, n, mh, th, sphLow.h, low_tree.h if low_tree else None, out))
        return [Weights(ctx.L, C.c_void_p(out[i]), owner=ctx) for i in range(n)]

    def barycentric_data_interpolation(self, metric_in: Mesh, sphLow: Mesh, nthreads: int = 1, EXCL=None):
        """resampler.cpp:30-70: adaptive-barycentric resampling of metric_in.pvalues -> [D, n_low]."""
        if EXCL is not None:
            raise NotImplementedError("exclusion masks are outside the accelerated path (SURVEY §8)")
        feat = metric_in.pvalues
        out = np.zeros((feat.shape[0], sphLow.nvertices()))
        check(metric_in.L.msmgpu_metric_resample(metric_in.h, sphLow.h, feat.shape[0], ptr(feat), ptr(out)))
        return out


def metric_resample(in_mesh: Mesh, target: Mesh, nthreads: int = 1, EXCL=None):
    """resampler.cpp:304-309. EXCL (optional): the exclusion mask's values [nv_in]; then returns (resampled data, resampled mask),
    the mask being what the reference writes back into *EXCL (resampler.cpp:55-67)."""
    if EXCL is None:
        return Resampler().barycentric_data_interpolation(in_mesh, target, nthreads, None)
    feat, excl = in_mesh.pvalues, f64(EXCL)
    out, eo = np.zeros((feat.shape[0], target.nvertices())), np.zeros(target.nvertices())
    check(in_mesh.L.msmgpu_metric_resample_excl(in_mesh.h, target.h, feat.shape[0], ptr(feat), ptr(excl), ptr(out), ptr(eo)))
    return out, eo


def adaptive_weights_excl(in_mesh: Mesh, target: Mesh, EXCL) -> "Weights":
    """get_adaptive_barycentric_weights with an exclusion mask (resampler.cpp:72-140)."""
    h = C.c_void_p()
    check(in_mesh.L.msmgpu_adaptive_weights_excl(in_mesh.h, target.h, ptr(f64(EXCL)), C.byref(h)))
    return Weights(in_mesh.L, h, owner=in_mesh.ctx)


def smooth_data(orig: Mesh, sphLow: Mesh, sigma: float, nthreads: int = 1, EXCL=None):
    """resampler.cpp:169-230 -> smoothed data [D, n_low] (and the new mask when EXCL is given)."""
    n = sphLow.nvertices()
    closest = Octree(orig).get_closest_vertex_ID(sphLow.xyz).astype(np.int32)
    feat = orig.pvalues
    out = np.zeros((feat.shape[0], n))
    excl = f64(EXCL) if EXCL is not None else None
    eo = np.zeros(n) if excl is not None else None
    check(orig.L.msmgpu_smooth_data(orig.ctx.h, n, ptr(sphLow.xyz), ptr(closest), float(sigma), feat.shape[0], feat.shape[1], ptr(feat),
                                    len(excl) if excl is not None else 0, ptr(excl), ptr(out), ptr(eo)))
    return out if excl is None else (out, eo)


def resample_batch_host(ctx: "Context", xyz_list, tri, low_xyz, low_tri, feat_list, out_bary=None, out_adaptive=None, chunk: int = 0):
    """Batch job on HOST buffers, pipelined inside the library (msmgpu_resample_batch_host_f32): subject s has coordinates xyz_list[s]
    [nv][3] f64 over the shared topology `tri`, channel-major FP32 features feat_list[s] [D][nv]; the results land in out_bary[s] /
    out_adaptive[s] [D][n_low] f32 (lists of numpy arrays or CPU torch tensors, ideally page-locked; either list may be None)."""
    S = len(xyz_list)

    def addr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

    def arr(lst):
        return None if lst is None else (C.c_void_p * S)(*[addr(a) for a in lst])
    t32, l32, low = i32(tri), i32(low_tri), f64(low_xyz)
    nv = len(xyz_list[0]) if not hasattr(xyz_list[0], "data_ptr") else xyz_list[0].shape[0]
    f0 = feat_list[0]
    D = f0.shape[0]
    check(capi.lib().msmgpu_resample_batch_host_f32(ctx.h, S, nv, arr(xyz_list), len(t32), ptr(t32), len(low), ptr(low), len(l32), ptr(l32), D,
                                                    arr(feat_list), arr(out_bary), arr(out_adaptive), int(chunk)))


def variance_normalise(ctx: "Context", DATA, EXCL=None):
    """newmeshreg::variance_normalise (reg_tools.cpp:804-844): [D, n] -> normalised copy; EXCL = mask values [n] or None."""
    data = f64(np.atleast_2d(DATA)).copy()
    excl = f64(EXCL) if EXCL is not None else None
    check(capi.lib().msmgpu_variance_normalise(ctx.h, data.shape[0], data.shape[1], ptr(data), ptr(excl)))
    return data


def metric_resample_f32(in_mesh: Mesh, target: Mesh, feat_f32, in_tree: Octree | None = None, target_tree: Octree | None = None):
    """FP32 payload variant of metric_resample (GIFTI stores floats, mesh.cpp:625): [D,V] -> [D,n]."""
    feat = f32(feat_f32)
    out = np.zeros((feat.shape[0], target.nvertices()), np.float32)
    check(in_mesh.L.msmgpu_metric_resample_f32(in_mesh.h, in_tree.h if in_tree else None, target.h, target_tree.h if target_tree else None,
                                               feat.shape[0], ptr(feat), ptr(out)))
    return out


def barycentric_resample(in_mesh: Mesh, low_xyz, feat=None):
    """Octree(in) + get_barycentric_weights + the loop of resampler.cpp:40-52 in one fused kernel."""
    feat = in_mesh.pvalues if feat is None else f64(np.atleast_2d(feat))
    low = f64(low_xyz)
    out = np.zeros((feat.shape[0], len(low)))
    check(in_mesh.L.msmgpu_bary_resample(in_mesh.h, len(low), ptr(low), feat.shape[0], ptr(feat), ptr(out)))
    return out


def sphere_project_warp(sphere_xyz, mesh_from: Mesh, to_xyz, nthreads: int = 1):
    """resampler.cpp:311-328; returns the warped coordinates instead of mutating `sphere`."""
    s, to = f64(sphere_xyz), f64(to_xyz)
    out = np.zeros_like(s)
    check(mesh_from.L.msmgpu_sphere_project_warp(mesh_from.h, ptr(to), len(s), ptr(s), ptr(out)))
    return out


def surface_resample(anat_xyz, sph: Mesh, low_xyz, nthreads: int = 1):
    """resampler.cpp:284-302 (project_anatomical_mesh, 260-282, is the same blend)."""
    a, low = f64(anat_xyz), f64(low_xyz)
    out = np.zeros_like(low)
    check(sph.L.msmgpu_surface_resample(sph.h, ptr(a), len(low), ptr(low), ptr(out)))
    return out


project_anatomical_mesh = surface_resample


def nearest_neighbour_interpolation(in_mesh: Mesh, low_xyz, feat=None, nthreads: int = 1, EXCL=None):
    """resampler.cpp:232-258. With EXCL (the mask's values [nv_in]) returns (data, new mask)."""
    feat = in_mesh.pvalues if feat is None else f64(np.atleast_2d(feat))
    low = f64(low_xyz)
    out = np.zeros((feat.shape[0], len(low)))
    if EXCL is None:
        check(in_mesh.L.msmgpu_nn_resample(in_mesh.h, len(low), ptr(low), feat.shape[0], ptr(feat), ptr(out)))
        return out
    eo = np.zeros(len(low))
    check(in_mesh.L.msmgpu_nn_resample_excl(in_mesh.h, len(low), ptr(low), feat.shape[0], ptr(feat), ptr(f64(EXCL)), ptr(out), ptr(eo)))
    return out, eo


def estimate_rotation_matrix(ci, index):
    """point.cpp:97-152 for [n,3] pairs -> [n,3,3]."""
    a, b = f64(ci).reshape(-1, 3), f64(index).reshape(-1, 3)
    R = np.zeros((len(a), 9))
    check(capi.lib().msmgpu_rotation_matrices(None, len(a), ptr(a), ptr(b), ptr(R)))
    return R.reshape(-1, 3, 3)
