"""Where the first call's time goes (process start -> first result): library load, context creation, first mesh / octree / resample.
Usage (GPU box): python tools/init_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
from newmsm_b200 import capi, synth  # noqa: E402
L = capi.lib()
t1 = time.perf_counter()
print(f"dlopen libmsmgpu.so            {1e3 * (t1 - t0):9.1f} ms")
C = capi.C
ctx = C.c_void_p()
capi.check(L.msmgpu_ctx_create(0, None, C.byref(ctx)))
t2 = time.perf_counter()
print(f"msmgpu_ctx_create              {1e3 * (t2 - t1):9.1f} ms")
xyz, tri = synth.icosphere(6)
low, ltri = synth.icosphere(4)
low = synth.rotate_sphere(low)
feat = np.ascontiguousarray(synth.smooth_fields(xyz, 4))
tri32, ltri32 = capi.i32(tri), capi.i32(ltri)
for rep in range(3):
    ta = time.perf_counter()
    m, ml = C.c_void_p(), C.c_void_p()
    capi.check(L.msmgpu_mesh_create(ctx, len(xyz), capi.ptr(xyz), len(tri32), capi.ptr(tri32), C.byref(m)))
    capi.check(L.msmgpu_mesh_create(ctx, len(low), capi.ptr(low), len(ltri32), capi.ptr(ltri32), C.byref(ml)))
    tb = time.perf_counter()
    t = C.c_void_p()
    capi.check(L.msmgpu_octree_build(m, C.byref(t)))
    tc = time.perf_counter()
    out = np.zeros((4, len(low)))
    capi.check(L.msmgpu_metric_resample(m, ml, 4, capi.ptr(feat), capi.ptr(out)))
    td = time.perf_counter()
    L.msmgpu_octree_destroy(t); L.msmgpu_mesh_destroy(m); L.msmgpu_mesh_destroy(ml)
    print(f"rep {rep}: 2x mesh_create {1e3 * (tb - ta):8.1f} ms | octree_build {1e3 * (tc - tb):8.1f} ms | metric_resample {1e3 * (td - tc):8.1f} ms")
print(f"process start -> end           {1e3 * (time.perf_counter() - t0):9.1f} ms")
