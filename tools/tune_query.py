"""Times the fused barycentric resample kernel and the stand-alone query kernel for every lane-group
width G (MSMGPU_QUERY_GROUP). Results are identical for all G (tests/test_gpu_parity.py); this is
only a tuning aid. Usage: python tools/tune_query.py [subjects] [channels]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newmsm_b200 import capi, resampler as R, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = R.Context(0, stream=stream.cuda_stream)
L = capi.lib()
xyz0, tri = synth.icosphere(7)
low_xyz, low_tri = synth.geodesic_sphere(57)
n_low, nv = len(low_xyz), len(xyz0)
d_low = torch.from_numpy(low_xyz).to(dev)
meshes = [R.Mesh(synth.jitter_sphere(xyz0, tri, seed=1234 + s), tri, ctx=ctx) for s in range(S)]
low = R.Mesh(low_xyz, low_tri, ctx=ctx)
trees = R.Octree.build_batch(meshes + [low])
feat = [torch.randn(nv, D, device=dev) for _ in range(S)]
out = [torch.empty(n_low, D, device=dev) for _ in range(S)]
tp = (C.c_void_p * S)(*[t.h.value for t in trees[:S]])
fp = (C.c_void_p * S)(*[t.data_ptr() for t in feat])
op = (C.c_void_p * S)(*[t.data_ptr() for t in out])
d_src = torch.from_numpy(meshes[0].xyz).to(dev)
d_tri_out = torch.empty(nv, dtype=torch.int32, device=dev)
torch.cuda.synchronize()


def timed(fn, reps=5):
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        for i in range(reps):
            e[i].record(stream)
            fn()
        e[reps].record(stream)
        stream.synchronize()
    return float(np.median([e[i].elapsed_time(e[i + 1]) for i in range(reps)]))


for g in (1, 2, 4, 8, 16, 32):
    capi.check(L.msmgpu_set_query_group(g))
    t_fused = timed(lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None)))
    t_fwd = timed(lambda: capi.check(L.msmgpu_nearest_triangle_dev(trees[0].h, n_low, capi.ptr(d_low), capi.ptr(d_tri_out), None, None)))
    t_rev = timed(lambda: capi.check(L.msmgpu_nearest_triangle_dev(trees[-1].h, nv, capi.ptr(d_src), capi.ptr(d_tri_out), None, None)))
    print(f"G={g:2d}  fused batch ({S} subj, D={D}): {t_fused:8.3f} ms = {S * n_low / t_fused / 1e3:8.1f} M targets/s | "
          f"query {n_low} pts in ico7 mesh: {t_fwd * 1e3:7.1f} us ({n_low / t_fwd / 1e3:7.1f} Mq/s) | "
          f"query {nv} pts in 32k mesh: {t_rev * 1e3:7.1f} us ({nv / t_rev / 1e3:7.1f} Mq/s)", flush=True)
