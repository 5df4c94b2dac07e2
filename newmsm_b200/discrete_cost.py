"""Host-side mirror of the reference's unary cost classes (msm-newmeshreg/src/DiscreteCostFunction.h:
82-244) on top of the C ABI. The class and method names are the reference's; the per-call virtual
`computeUnaryCost(node, label)` becomes a lookup into the table `computeUnaryCosts()` filled on the
GPU (label-major `unarycosts[label * N + node]`, DiscreteCostFunction.cpp:242).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, f64, ptr
from .resampler import Mesh, Octree

MeshregException = capi.MsmGpuError   # reg_tools.h:38

UNIVARIATE, MULTIVARIATE, PATCHWISE, HO_UNIVARIATE, HO_MULTIVARIATE = 0, 1, 2, 3, 4
SSD, CORRELATION = 1, 2   # similarities.h:48-58 (_simmeasure)


class NonLinearSRegDiscreteCostFunction:
    """Common state: TARGET mesh + octree, SOURCE mesh, FEAT (input/reference features), CP grid."""

    KIND = UNIVARIATE

    def __init__(self, simmeasure: int = CORRELATION, percentile: float = 0.75):
        self.simmeasure = simmeasure      # 1 SSD, 2 correlation, 4 DICE, 5 genDICE (similarities.h:48-58)
        self.percentile = percentile      # sparsesimkernel::set_percentile (DICE measures)
        self.h = None
        self.unarycosts = None

    # set_meshes + set_featurespace + set_octrees (DiscreteCostFunction.h:173-189)
    def set_meshes(self, target: Mesh, source_xyz, src_feat, ref_feat, target_tree: Octree | None = None):
        self.target = target
        self.tree = target_tree or Octree(target)
        self.L = target.L
        s, a, b = f64(source_xyz), f64(np.atleast_2d(src_feat)), f64(np.atleast_2d(ref_feat))
        assert a.shape[1] == len(s) and b.shape[1] == target.nvertices() and a.shape[0] == b.shape[0]
        self.D, self.nsrc = a.shape[0], len(s)
        self.h = C.c_void_p()
        check(self.L.msmgpu_costfn_create(self.tree.h, self.KIND, self.simmeasure, len(s), ptr(s), self.D, ptr(a), ptr(b), C.byref(self.h)))
        check(self.L.msmgpu_costfn_set_percentile(self.h, float(self.percentile)))

    def reset_source(self, source_xyz):
        s = f64(source_xyz)
        check(self.L.msmgpu_costfn_reset_source(self.h, ptr(s)))

    # reset_CPgrid + set_spacings + get_source_data (cpp:334-351)
    def reset_CPgrid(self, cp_xyz, maxsep, controlptrange: float, HIGHREScfweight=None, AbsoluteWeights=None):
        cp, ms = f64(cp_xyz), f64(maxsep)
        self.ncp = len(cp)
        absw = np.ones(self.ncp) if AbsoluteWeights is None else f64(AbsoluteWeights)
        cfw = None if HIGHREScfweight is None else f64(np.atleast_2d(HIGHREScfweight))
        check(self.L.msmgpu_costfn_set_cpgrid(self.h, self.ncp, ptr(cp), ptr(ms), float(controlptrange),
                                              0 if cfw is None else cfw.shape[0], ptr(cfw), ptr(absw)))

    def get_source_data(self):
        """patch lists (CSR rowptr, members) as `_sourceinrange` (cpp:334-351)."""
        rowptr = np.zeros(self.ncp + 1, np.int32)
        check(self.L.msmgpu_costfn_patches(self.h, ptr(rowptr), None))
        mem = np.zeros(int(rowptr[-1]), np.int32)
        check(self.L.msmgpu_costfn_patches(self.h, ptr(rowptr), ptr(mem)))
        return rowptr, mem

    # set_labels + computeUnaryCosts (h:180, cpp:236-243)
    def computeUnaryCosts(self, labels, ROTATIONS, want_triangles: bool = False):
        lab, rot = f64(labels).reshape(-1, 3), f64(ROTATIONS).reshape(-1, 9)
        assert len(rot) == self.ncp
        self.nlabels = len(lab)
        out = np.zeros((self.nlabels, self.ncp))
        tri = None
        if want_triangles:
            rowptr = np.zeros(self.ncp + 1, np.int32)
            check(self.L.msmgpu_costfn_patches(self.h, ptr(rowptr), None))
            tri = np.zeros((self.nlabels, int(rowptr[-1])), np.int32)
        check(self.L.msmgpu_costfn_unary_table(self.h, self.nlabels, ptr(lab), ptr(rot), ptr(out), ptr(tri)))
        self.unarycosts = out
        return (out, tri) if want_triangles else out

    def computeUnaryCost(self, node: int, label: int) -> float:
        return float(self.unarycosts[label, node])

    # set_parameters (cpp:119-133): the regulariser part
    def set_parameters(self, lambda_: float, shearmodulus: float = 0.4, bulkmodulus: float = 1.6, kexponent: float = 2.0,
                       exponent: float = 2.0, regularisermode: int = 3):
        self.reg = capi.RegParams(lambda_, shearmodulus, bulkmodulus, kexponent, exponent, regularisermode)

    # setTriplets (h:57) + the state computeTripletCost reads: ROTATIONS, _labels, _ORIG
    def setTriplets(self, triplets, labels, ROTATIONS, ORIG_xyz):
        self.triplets = capi.i32(triplets).reshape(-1, 3)
        self._labels = f64(labels).reshape(-1, 3)
        self._rot = f64(ROTATIONS).reshape(-1, 9)
        self._orig = f64(ORIG_xyz).reshape(-1, 3)

    # set_anatomical + set_anatomical_neighbourhood (h:160-169): what regoption 4/5 reads (cpp:169-181, 245-301)
    def set_anatomical(self, asource_xyz, asource_tri, thi_xyz, thi_tri, atarget_xyz, face_ptr, face_ids, bary_ptr, bary_key, bary_w):
        k = dict(asource_xyz=f64(asource_xyz), asource_tri=capi.i32(asource_tri), thi_xyz=f64(thi_xyz), thi_tri=capi.i32(thi_tri),
                 atarget_xyz=f64(atarget_xyz), face_ptr=capi.i32(face_ptr), face_ids=capi.i32(face_ids), bary_ptr=capi.i32(bary_ptr),
                 bary_key=capi.i32(bary_key), bary_w=f64(bary_w))
        A = capi.Anatomical(len(k["asource_xyz"].reshape(-1, 3)), k["asource_xyz"].ctypes.data, len(k["asource_tri"].reshape(-1, 3)), k["asource_tri"].ctypes.data,
                            len(k["thi_xyz"].reshape(-1, 3)), k["thi_xyz"].ctypes.data, len(k["thi_tri"].reshape(-1, 3)), k["thi_tri"].ctypes.data,
                            k["atarget_xyz"].ctypes.data, k["face_ptr"].ctypes.data, k["face_ids"].ctypes.data, k["bary_ptr"].ctypes.data,
                            k["bary_key"].ctypes.data, k["bary_w"].ctypes.data)
        check(self.L.msmgpu_costfn_set_anatomical(self.h, len(k["face_ptr"]) - 1, C.byref(A)))

    def computeTripletCostList(self, triplet, la, lb, lc):
        """computeTripletCost (cpp:135-188) for arrays of requests."""
        t, a, b, c_ = (capi.i32(x) for x in (triplet, la, lb, lc))
        out = np.zeros(len(t))
        check(self.L.msmgpu_costfn_triplet_costs(self.h, len(self.triplets), ptr(self.triplets), len(self._labels), ptr(self._labels), ptr(self._rot),
                                                 ptr(self._orig), C.byref(self.reg), len(t), ptr(t), ptr(a), ptr(b), ptr(c_), ptr(out)))
        return out

    def computeTripletCost(self, triplet: int, labelA: int, labelB: int, labelC: int) -> float:
        return float(self.computeTripletCostList([triplet], [labelA], [labelB], [labelC])[0])

    def computeTripletCostsForLabel(self, labeling, label: int):
        """The 8 combinations per triplet Fusion::optimize evaluates for one candidate label (Fusion.h:181-196) -> [T, 8]."""
        lab = capi.i32(labeling)
        out = np.zeros((len(self.triplets), 8))
        check(self.L.msmgpu_costfn_triplet_batch(self.h, len(self.triplets), ptr(self.triplets), len(self._labels), ptr(self._labels), ptr(self._rot),
                                                 ptr(self._orig), C.byref(self.reg), ptr(lab), int(label), ptr(out)))
        return out

    def getUnaryCosts(self):
        return self.unarycosts.reshape(-1)

    def close(self):
        if getattr(self, "h", None):
            self.L.msmgpu_costfn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _HOMixin:
    """HO (triclique) classes: patches hang on CP-grid triangles (cpp:468-485, 541-563)."""

    def reset_CPgrid(self, cp_xyz, cp_tri, HIGHREScfweight=None, AbsoluteWeights=None):   # noqa: N803
        cp, tri = f64(cp_xyz), capi.i32(cp_tri)
        self.ncp = len(cp)
        self.n_cp_tri = len(tri)
        absw = np.ones(self.ncp) if AbsoluteWeights is None else f64(AbsoluteWeights)
        cfw = None if HIGHREScfweight is None else f64(np.atleast_2d(HIGHREScfweight))
        check(self.L.msmgpu_costfn_set_cpgrid_ho(self.h, self.ncp, ptr(cp), len(tri), ptr(tri), 0 if cfw is None else cfw.shape[0], ptr(cfw), ptr(absw)))

    def get_source_data(self):
        rowptr = np.zeros(self.n_cp_tri + 1, np.int32)
        check(self.L.msmgpu_costfn_patches(self.h, ptr(rowptr), None))
        mem = np.zeros(int(rowptr[-1]), np.int32)
        check(self.L.msmgpu_costfn_patches(self.h, ptr(rowptr), ptr(mem)))
        return rowptr, mem


class HOUnivariateNonLinearSRegDiscreteCostFunction(_HOMixin, NonLinearSRegDiscreteCostFunction):
    KIND = HO_UNIVARIATE


class HOMultivariateNonLinearSRegDiscreteCostFunction(_HOMixin, NonLinearSRegDiscreteCostFunction):
    KIND = HO_MULTIVARIATE


class UnivariateNonLinearSRegDiscreteCostFunction(NonLinearSRegDiscreteCostFunction):
    KIND = UNIVARIATE


class MultivariateNonLinearSRegDiscreteCostFunction(NonLinearSRegDiscreteCostFunction):
    KIND = MULTIVARIATE


class PatchwiseMultivariateNonLinearSRegDiscreteCostFunction(NonLinearSRegDiscreteCostFunction):
    KIND = PATCHWISE
