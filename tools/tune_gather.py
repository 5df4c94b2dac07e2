"""Times the bulk-copy row gather (csrc/gather.cu) against the register-path kernels it replaces, for every launch variant, on the
BASELINE configs[1] shape (S subjects, ico7 -> 32 492 vertices, D FP32 channels), and checks that the outputs are bit-identical.
Usage: python tools/tune_gather.py [subjects] [channels]      (tuning aid; results land in profiles/)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newmsm_b200 import capi, resampler as R, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
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
out2 = [torch.empty(n_low, D, device=dev) for _ in range(S)]
fp = (C.c_void_p * S)(*[t.data_ptr() for t in feat])
op = (C.c_void_p * S)(*[t.data_ptr() for t in out])
op2 = (C.c_void_p * S)(*[t.data_ptr() for t in out2])
PEAK = 6543.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def tune(name, v):
    capi.check(L.msmgpu_set_tuning(name.encode(), int(v)))


def timed(fn, reps=7):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        for i in range(reps):
            e[i].record(stream)
            fn()
        e[reps].record(stream)
        stream.synchronize()
    return float(np.median([e[i].elapsed_time(e[i + 1]) for i in range(reps)]))


with torch.cuda.stream(stream):
    low = R.Mesh.from_device(ctx, n_low, d_low, len(low_tri), d_low_tri)
    meshes = R.Mesh.views_from_device(ctx, nv, d_xyz, nt, d_tri)
    trees = R.Octree.build_batch(meshes + [low])
    tp = (C.c_void_p * S)(*[t.h.value for t in trees[:S]])
    mp = (C.c_void_p * S)(*[m.h.value for m in meshes])
    fwd = C.c_void_p()
    capi.check(L.msmgpu_fwd_create(ctx.h, S, n_low, C.byref(fwd)))

    # ---- barycentric: fused register path vs query kernel + bulk gather ----
    bytes_bary = S * (24 * n_low + 24 * nv + 12 * nt + 4 * D * min(3 * n_low, nv) + 4 * D * n_low)
    bytes_gather = S * (4 * D * min(3 * n_low, nv) + 4 * D * n_low + 36 * n_low)
    tune("gather", 0)
    t_fused = timed(lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None, fwd)))
    stream.synchronize()
    ref_b = [o.clone() for o in out]
    print(f"fused k_bary_resample_f32 (register path): {t_fused:.3f} ms  frac {bytes_bary / t_fused / 1e6 / PEAK:.3f}", flush=True)
    res = {"S": S, "D": D, "fused_ms": t_fused, "bary": {}, "csr": {}}
    tune("gather", 2)
    res["fused_bulk"] = {}
    for v in range(7):
        tune("fused_bulk_variant", v)
        t = timed(lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None, fwd)))
        stream.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(ref_b, out))
        res["fused_bulk"][v] = {"ms": t, "identical": bool(same)}
        print(f"fused bulk variant {v}: {t:.3f} ms  frac of B_bary {bytes_bary / t / 1e6 / PEAK:.3f} | identical to register path: {same}", flush=True)
    tune("gather", 1)
    tune("bary_chunks", 1)
    for v in range(6):
        tune("gather_variant_bary", v)
        t_split = timed(lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None, fwd)))
        stream.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(ref_b, out))
        t_g = timed(lambda: capi.check(L.msmgpu_fwd_apply_batch_f32_dev(ctx.h, fwd, D, fp, op2)))
        stream.synchronize()
        same2 = all(torch.equal(a, b) for a, b in zip(ref_b, out2))
        res["bary"][v] = {"split_ms": t_split, "gather_ms": t_g, "identical": bool(same and same2)}
        print(f"bary variant {v}: queries + gather {t_split:.3f} ms (frac of B_bary {bytes_bary / t_split / 1e6 / PEAK:.3f}) | gather alone {t_g:.3f} ms "
              f"= {bytes_gather / t_g / 1e6:.0f} GB/s, frac {bytes_gather / t_g / 1e6 / PEAK:.3f} | identical to fused: {same and same2}", flush=True)

    # ---- overlap of queries and gather over subject chunks (two streams) ----
    tune("gather_variant_bary", 5)
    res["chunks"] = {}
    for chunks in (1, 4):
        for cap in (0, 2):
            tune("bary_chunks", chunks)
            tune("gather_ctas_per_sm", cap)
            t = timed(lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None, fwd)))
            stream.synchronize()
            same = all(torch.equal(a, b) for a, b in zip(ref_b, out))
            res["chunks"][f"{chunks}x{cap}"] = {"ms": t, "identical": bool(same)}
            print(f"bary chunks {chunks:2d}, gather CTAs/SM cap {cap}: {t:.3f} ms  frac of B_bary {bytes_bary / t / 1e6 / PEAK:.3f} | identical: {same}", flush=True)
    tune("bary_chunks", 4)
    tune("gather_ctas_per_sm", 0)

    # ---- Morton-ordered queries: split barycentric (forward queries) and adaptive weights (reverse queries) ----
    def adaptive_once():
        wp = (C.c_void_p * S)()
        capi.check(L.msmgpu_adaptive_weights_batch_fwd(ctx.h, S, mp, tp, low.h, trees[-1].h, fwd, wp))
        for i in range(S):
            L.msmgpu_weights_destroy(C.c_void_p(wp[i]))
    tune("gather_variant_bary", 0)
    for order in (0, 1, 2):
        tune("query_order", order)
        t_split = timed(lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, capi.ptr(d_low), D, fp, op, None, fwd)))
        t_adapt = timed(adaptive_once, reps=5)
        res[f"query_order_{order}"] = {"bary_split_ms": t_split, "adaptive_weights_ms": t_adapt}
        print(f"query_order={order}: bary queries + gather {t_split:.3f} ms | adaptive weights (reverse queries + rows) {t_adapt:.3f} ms", flush=True)

    # ---- adaptive: CSR apply ----
    w_ptrs = (C.c_void_p * S)()
    capi.check(L.msmgpu_adaptive_weights_batch_fwd(ctx.h, S, mp, tp, low.h, trees[-1].h, fwd, w_ptrs))
    Ws = [R.Weights(L, C.c_void_p(w_ptrs[i])) for i in range(S)]
    nnz = sum(w.shape()[2] for w in Ws)
    bytes_apply = S * (4 * D * nv + 4 * D * n_low + 4 * (n_low + 1)) + 12 * nnz
    tune("gather_csr", 0)
    t_old = timed(lambda: capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, w_ptrs, D, fp, op)))
    stream.synchronize()
    ref_a = [o.clone() for o in out]
    res["csr_old_ms"] = t_old
    print(f"k_csr_apply_f32x4 (register path): {t_old:.3f} ms  frac {bytes_apply / t_old / 1e6 / PEAK:.3f}  (nnz {nnz})", flush=True)
    tune("gather_csr", 1)
    for rows_per_tile, v in [(r, v) for r in (32, 8) for v in range(7)]:
        tune("gather_rows", rows_per_tile)
        tune("gather_variant", v)
        t_new = timed(lambda: capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, w_ptrs, D, fp, op)))
        stream.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(ref_a, out))
        res["csr"][f"{v}_rows{rows_per_tile}"] = {"ms": t_new, "identical": bool(same)}
        print(f"csr variant {v}, {rows_per_tile} rows per tile: {t_new:.3f} ms = {bytes_apply / t_new / 1e6:.0f} GB/s, frac {bytes_apply / t_new / 1e6 / PEAK:.3f} | identical: {same}", flush=True)
    print(json.dumps(res))
