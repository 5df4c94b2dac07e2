"""Host-side mirror of the AFFINE / RIGID level (msm-newmeshreg/src/rigid_costfunction.{h,cpp}) over the C ABI.

`Rigid_cost_function::initialise` and `rigid_cost_mesh` run through msmgpu_rigid_create / msmgpu_rigid_cost (csrc/rigid.cu: device
kernels for everything built from + - * / sqrt, host libm for exp / cos / sin inside the library); the gradient-ascent driver `run`
(rigid_costfunction.cpp:167-236) and `rotate_in_mesh` (116-128) are host logic, restated here like the reference-side adapter keeps
them (integration/newmsm_gpu_rigid_hooks.cpp leaves the reference's own `run` in place).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi
from .capi import check, f64, i32, ptr
from .resampler import Context


def euler_rotate(xyz, w1: float, w2: float, w3: float):
    """point.cpp:154-171 for every row of xyz: rotation.t() * v, each output = ((0 + R[0][r] v0) + R[1][r] v1) + R[2][r] v2.
    cos / sin from the C library (math), like the reference."""
    c1, s1, c2, s2, c3, s3 = math.cos(w1), math.sin(w1), math.cos(w2), math.sin(w2), math.cos(w3), math.sin(w3)
    R = [[c2 * c3, -c1 * s3 + s1 * s2 * c3, s1 * s3 + c1 * s2 * c3],
         [c2 * s3, c1 * c3 + s1 * s2 * s3, -s1 * c3 + c1 * s2 * s3],
         [-s2, s1 * c2, c1 * c2]]
    v = f64(xyz)
    out = np.empty_like(v)
    for r in range(3):
        acc = np.zeros(len(v))
        for k in range(3):
            acc = acc + R[k][r] * v[:, k]
        out[:, r] = acc
    return out


class Rigid_cost_function:
    """newmeshreg::Rigid_cost_function (rigid_costfunction.h)."""

    def __init__(self, target_xyz, target_tri, source_xyz, source_tri, input_data, reference_data, ctx: Context | None = None):
        self.ctx = ctx or Context.default()
        self.L = capi.lib()
        self.tx, self.tt = f64(target_xyz), i32(target_tri)
        self.SOURCE, self.st = f64(source_xyz).copy(), i32(source_tri)
        self.A, self.B = f64(np.atleast_2d(input_data)), f64(np.atleast_2d(reference_data))
        self.iters, self.simmeasure, self.stepsize, self.spacing = 20, 2, 0.01, 0.5
        self.h = None

    def set_parameters(self, iters=None, simmeasure=None, stepsize=None, gradsampling=None):
        if iters is not None: self.iters = int(iters)
        if simmeasure is not None: self.simmeasure = int(simmeasure)
        if stepsize is not None: self.stepsize = float(stepsize)
        if gradsampling is not None: self.spacing = float(gradsampling)

    def initialise(self):
        """rigid_costfunction.cpp:32-50."""
        mvd = C.c_double(0.0)
        check(self.L.msmgpu_mean_vertex_distance(len(self.SOURCE), ptr(self.SOURCE), len(self.st), ptr(self.st), C.cast(C.byref(mvd), C.c_void_p)))
        self.MVD = mvd.value
        self.close()
        h = C.c_void_p()
        check(self.L.msmgpu_rigid_create(self.ctx.h, len(self.tx), ptr(self.tx), len(self.tt), ptr(self.tt), len(self.SOURCE), ptr(self.SOURCE),
                                         len(self.st), ptr(self.st), self.A.shape[0], ptr(self.A), ptr(self.B), self.simmeasure, self.MVD, C.byref(h)))
        self.h = h

    def rigid_cost_mesh(self, dw1: float, dw2: float, dw3: float) -> float:
        """rigid_costfunction.cpp:130-141 (SOURCE is left as it was)."""
        cost = C.c_double(0.0)
        check(self.L.msmgpu_rigid_cost(self.h, ptr(self.SOURCE), float(dw1), float(dw2), float(dw3), C.cast(C.byref(cost), C.c_void_p)))
        return cost.value

    def rotate_in_mesh(self, a1: float, a2: float, a3: float):
        self.SOURCE = euler_rotate(self.SOURCE, a1, a2, a3)

    def run(self):
        """rigid_costfunction.cpp:167-236: finite-difference gradient ascent over three Euler angles; returns the rotated SOURCE."""
        e1 = e2 = e3 = 0.0
        min_iter = loop = 0
        grad_zero = self.rigid_cost_mesh(e1, e2, e3)
        mingrad_zero = grad_zero
        spacing = self.spacing
        while spacing > 0.05:
            step, per = self.stepsize, spacing
            for it in range(1, self.iters + 1):
                e1 = e2 = e3 = 0.0
                g = np.array([(self.rigid_cost_mesh(e1 + per, e2, e3) - grad_zero) / per,
                              (self.rigid_cost_mesh(e1, e2 + per, e3) - grad_zero) / per,
                              (self.rigid_cost_mesh(e1, e2, e3 + per) - grad_zero) / per])
                n = math.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2])      # Point::normalize (point.cpp:26-34)
                if n > 1e-8:
                    g = np.array([g[0] / n, g[1] / n, g[2] / n])
                e1 += step * g[0]; e2 += step * g[1]; e3 += step * g[2]
                tmp = self.SOURCE.copy()
                self.rotate_in_mesh(e1, e2, e3)
                grad_zero = self.rigid_cost_mesh(e1, e2, e3)
                if grad_zero > mingrad_zero:
                    mingrad_zero = grad_zero
                    min_iter = loop * self.iters + it
                if loop * self.iters + it - min_iter > 0:
                    step *= 0.5
                    self.SOURCE = tmp
                if step < 1e-3:
                    break
            loop += 1
            spacing *= 0.5
        return self.SOURCE

    def close(self):
        if getattr(self, "h", None):
            self.L.msmgpu_rigid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
