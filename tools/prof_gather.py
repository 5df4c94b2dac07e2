"""Small driver for ncu captures of the gather kernels: S subjects of the BASELINE configs[1] shape, one launch of every kernel form.
Usage: [ncu ...] python tools/prof_gather.py [subjects] [channels]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newmsm_b200 import capi, resampler as R, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = R.Context(0, stream=stream.cuda_stream)
L = capi.lib()
xyz0, tri = synth.icosphere(7)
low_xyz, low_tri = synth.geodesic_sphere(57)
n_low, nv, nt = len(low_xyz), len(xyz0), len(tri)
d_low = torch.from_numpy(low_xyz).to(dev)
d_low_tri = torch.from_numpy(low_tri).to(dev)
d_tri = torch.from_numpy(tri).to(dev)
d_xyz = [torch.from_numpy(synth.jitter_sphere(xyz0, tri, frac=0.3, seed=1234 + s)).to(dev) for s in range(S)]
feat = [torch.randn(nv, D, device=dev) for _ in range(S)]
out = [torch.empty(n_low, D, device=dev) for _ in range(S)]
fp = (C.c_void_p * S)(*[t.data_ptr() for t in feat])
op = (C.c_void_p * S)(*[t.data_ptr() for t in out])
tune = lambda k, v: capi.check(L.msmgpu_set_tuning(k.encode(), int(v)))
with torch.cuda.stream(stream):
    low = R.Mesh.from_device(ctx, n_low, d_low, len(low_tri), d_low_tri)
    meshes = R.Mesh.views_from_device(ctx, nv, d_xyz, nt, d_tri)
    trees = R.Octree.build_batch(meshes + [low])
    tp = (C.c_void_p * S)(*[t.h.value for t in trees[:S]])
    mp = (C.c_void_p * S)(*[m.h.value for m in meshes])
    fwd = C.c_void_p()
    capi.check(L.msmgpu_fwd_create(ctx.h, S, n_low, C.byref(fwd)))
    for mode in (0, 1, 2):          # register-path fused kernel, queries + bulk gather, fused bulk
        tune("gather", mode)
        tune("bary_chunks", 1)
        for _ in range(2):
            capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None, fwd))
    w_ptrs = (C.c_void_p * S)()
    capi.check(L.msmgpu_adaptive_weights_batch_fwd(ctx.h, S, mp, tp, low.h, trees[-1].h, fwd, w_ptrs))
    for mode in (0, 1):
        tune("gather_csr", mode)
        for _ in range(2):
            capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, w_ptrs, D, fp, op))
    stream.synchronize()
print("done")
