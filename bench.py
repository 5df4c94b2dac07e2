#!/usr/bin/env python
"""bench.py — BASELINE.json configs[1]: standalone batch resampling of a 164k-vertex native sphere
(jittered ico7, 163 842 V / 327 680 T, one mesh per subject) onto a 32k sphere (32 492 V geodesic
sphere), 100-channel FP32 feature matrix, barycentric AND adaptive-barycentric resampling.

A step = one batch of S subjects through BOTH methods, octree builds included (the reference's
metric_resample builds its trees per call, resampler.cpp:74-78). Metric: resampled verts/s, i.e.
2 * S * 32 492 output vertices (x100 channels each) per step time.

  value : inputs (meshes, targets, features) already resident in HBM; CUDA events, max over ranks
  e2e   : the same work through the host-buffer C ABI a reference-side loop calls: ONE msmgpu_resample_batch_host_f32 per step
          (chunks of subjects pipelined over copy-in / compute / copy-out streams inside the library), pinned host buffers,
          H2D and D2H copies inside the timed region; the per-subject calls on 8 worker streams beside it ("per_subject_calls")
  secondary legs of the same line (rank 0; skippable): "parity" (subject 0 at full size against the compiled reference),
          "e2e_adapter" (the C++ adapter on reference Mesh objects), "gmsm" (BASELINE configs[4], all ranks, NCCL all-gather),
          "unary_costs" (BASELINE metric ii: unary cost tables, costs/s), "newmsm_wall_time" (BASELINE metric iii on a bounded case:
          the reference's own program with the library bound in vs the same program on the host cores)
  --impl reference : the reference's own CPU implementation (oracle/_ref, compiled from the
          unmodified sources) on the host cores, one subject per step

Launch: python bench.py --gpus N --steps K --warmup W      (N>1 via torch.distributed.run)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TARGET_FREQ = 57     # 10 f^2 + 2 = 32 492 vertices
NATIVE_LEVEL = 7       # 163 842 vertices
METRIC = "resampled verts/s"
UNIT = "verts/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--subjects", type=int, default=int(os.environ.get("BENCH_SUBJECTS", 64)), help="subjects per GPU per step")
    ap.add_argument("--channels", type=int, default=100)
    ap.add_argument("--native-level", type=int, default=NATIVE_LEVEL)
    ap.add_argument("--target-freq", type=int, default=N_TARGET_FREQ)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size check of subject 0 against the reference's outputs")
    ap.add_argument("--e2e-workers", type=int, default=8)
    ap.add_argument("--e2e-chunk", type=int, default=0, help="subjects per pipeline stage of the batch host call (0: the library's default)")
    ap.add_argument("--no-adapter-e2e", action="store_true", help="skip the C++ adapter leg (Mesh in / Mesh out, pageable FP64)")
    ap.add_argument("--no-newmsm", action="store_true", help="skip the secondary `newmsm` wall-time leg (BASELINE metric iii, ~45 s)")
    ap.add_argument("--no-unary", action="store_true", help="skip the secondary unary-cost-table leg (BASELINE metric ii)")
    ap.add_argument("--no-gmsm", action="store_true", help="skip the secondary groupwise (gMSM, BASELINE configs[4]) leg")
    ap.add_argument("--gmsm-subjects", type=int, default=int(os.environ.get("BENCH_GMSM_SUBJECTS", 64)))
    ap.add_argument("--gmsm-data-level", type=int, default=6)
    ap.add_argument("--gmsm-cp-level", type=int, default=4)
    ap.add_argument("--value-workers", type=int, default=int(os.environ.get("BENCH_VALUE_WORKERS", 1)),
                    help="host threads / CUDA streams the device-resident step is split over (subjects are independent): the "
                         "latency-bound octree builds of one group overlap the bandwidth-bound resampling of another")
    return ap.parse_args()


def workload_config(a, nv, nt, n_low):
    """Identical in both arms (the driver compares the two dicts). The metric is a RATE (resampled verts/s): the GPU arm resamples
    a batch of subjects per step, the reference arm one subject per step; the ratio of the two lines compares rates."""
    return {
        "workload": f"batch resampling ico{a.native_level} native sphere ({nv} V, jittered per subject) -> {n_low}-vertex sphere, "
                    f"{a.channels} channels, barycentric + adaptive-barycentric, octree builds included",
        "channels": a.channels, "native_vertices": nv, "native_triangles": nt, "target_vertices": n_low,
        "methods": ["barycentric", "adaptive_barycentric"],
        "subjects_per_step": {"gpu_arm_per_gpu": a.subjects, "reference_arm": 1},
        "payload": {"gpu_arm": "FP32 feature rows, FP64 geometry / weights / accumulation, one rounding to FP32 on output",
                    "reference_arm": "FP64 (Mesh::pvalues), the reference's own types"},
        "l2_policy": f"GPU arm: inputs larger than L2 ({a.subjects * nv * a.channels * 4 / 1e9:.2f} GB of feature rows streamed per step), no flush needed; "
                     "reference arm: host memory",
        "rate_note": "metric is a rate; value = resampled target vertices (x channels each) per second, both methods counted",
    }


# --------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------
def native_mesh(level, subject):
    from newmsm_b200 import synth
    xyz, tri = synth.icosphere(level)
    return synth.jitter_sphere(xyz, tri, frac=0.3, seed=1234 + subject), tri


def target_sphere(freq):
    from newmsm_b200 import synth
    return synth.geodesic_sphere(freq)


def smooth_fields_torch(xyz_t, D, seed, device):
    """f_d(x) = sum_m a cos(k.x/100 + phi): the SURVEY §8d smooth random fields, generated on the GPU -> [V, D] f32."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    M = 16
    a = torch.randn(D, M, generator=g) / torch.arange(1, M + 1)
    k = torch.randn(D, M, 3, generator=g)
    k = k / k.norm(dim=-1, keepdim=True) * (1 + 11 * torch.rand(D, M, 1, generator=g))
    phi = 2 * np.pi * torch.rand(D, M, generator=g)
    a, k, phi = a.to(device), k.to(device), phi.to(device)
    x = (xyz_t / 100.0).to(torch.float32)
    out = torch.zeros(x.shape[0], D, device=device)
    for m in range(M):
        out += a[:, m] * torch.cos(x @ k[:, m, :].T + phi[:, m])
    return out.contiguous()


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        if os.environ.get("BENCH_NO_CLOCKS"):
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """Only samples taken after this call count (the sampler itself is started before the warm-up, because
        nvidia-smi's start-up can stall CUDA calls for tens of ms)."""
        self.f.flush()
        try:
            self.skip = len(open(self.f.name).read().splitlines())
        except Exception:
            self.skip = 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines()[getattr(self, "skip", 0):] if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm (CPU)
# --------------------------------------------------------------------------------------------
def reference_subject_seconds(B, xyz, tri, low_xyz, low_tri, feat_cm, threads):
    """One subject through the reference's own code (oracle/_ref): metric_resample (resampler.cpp:304) and the plain
    barycentric path (Octree + get_barycentric_weights + resampler.cpp:40-52). Returns (t_bary, t_adaptive)."""
    mi, ml = B.RefMesh(xyz, tri, feat=feat_cm), B.RefMesh(low_xyz, low_tri)
    _, tb = B.ref_bary_resample(mi, ml, nthreads=threads, want_out=False)
    _, ta = B.ref_metric_resample(mi, ml, nthreads=threads, want_out=False)
    return tb, ta


def port_subject_seconds(B, xyz, tri, low_xyz, low_tri, feat_cm, threads):
    t0 = time.perf_counter()
    B.oracle_bary_resample(xyz, tri, low_xyz, feat_cm, nthreads=threads)
    t1 = time.perf_counter()
    B.oracle_metric_resample(xyz, tri, low_xyz, low_tri, feat_cm, nthreads=threads)
    return t1 - t0, time.perf_counter() - t1


def cpu_arm(a, steps, warmup):
    """Times the reference on the host cores: each step = one subject of the workload (bounded sample)."""
    from newmsm_b200 import synth
    from oracle import bindings as B
    threads = os.cpu_count() or 1
    kind = "reference" if B.have_ref() else "port"
    if kind == "port":
        B.build(ref=False)
    xyz, tri = native_mesh(a.native_level, 0)
    low_xyz, low_tri = target_sphere(a.target_freq)
    rng = np.random.default_rng(100)
    # feature VALUES do not change the reference's control flow or timing; cheap smooth-ish fields
    feat = (np.cos(xyz @ rng.normal(size=(3, a.channels)) / 40.0)).T.copy()
    fn = reference_subject_seconds if kind == "reference" else port_subject_seconds
    times = []
    for i in range(warmup + steps):
        tb, ta = fn(B, xyz, tri, low_xyz, low_tri, feat, threads)
        if i >= warmup:
            times.append((tb, ta))
    tb = float(np.mean([t[0] for t in times])); ta = float(np.mean([t[1] for t in times]))
    n_low = len(low_xyz)
    return {"value": 2 * n_low / (tb + ta), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"1 subject per step ({len(xyz)} -> {n_low} V, {a.channels} channels, both methods), {steps} timed steps; "
                      f"barycentric {tb:.3f} s, adaptive (metric_resample) {ta:.3f} s per subject",
            "seconds_per_subject": tb + ta}, len(xyz), len(tri), n_low


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = a.steps, a.warmup
    base, nv, nt, n_low = cpu_arm(a, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * base["seconds_per_subject"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, nv, nt, n_low),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def reference_outputs(xyz, tri, low_xyz, low_tri, feat_cm):
    """The reference's own outputs for ONE subject (oracle/_ref = the unmodified sources compiled here; the CPU restatement when
    that library is absent): barycentric resample at full thread count (deterministic), adaptive weights + metric_resample on ONE
    thread (the reference's adaptive weights race at > 1 thread). Used only as the checker of the parity block."""
    from oracle import bindings as B
    threads = os.cpu_count() or 1
    if B.have_ref():
        mi, ml = B.RefMesh(xyz, tri, feat=feat_cm), B.RefMesh(low_xyz, low_tri)
        ref_b, _ = B.ref_bary_resample(mi, ml, nthreads=threads)
        ref_a, _ = B.ref_metric_resample(mi, ml, nthreads=1)
        rp, col, val = B.ref_adaptive_weights(mi, ml, nthreads=1)
        tri_ids = B.RefOctree(mi).query(low_xyz)[0]
        return "reference (oracle/_ref)", ref_b, ref_a, (rp, col, val), tri_ids
    B.build(ref=False)
    ref_b = B.oracle_bary_resample(xyz, tri, low_xyz, feat_cm, nthreads=threads)
    ref_a = B.oracle_metric_resample(xyz, tri, low_xyz, low_tri, feat_cm, nthreads=threads)
    rp, col, val = B.oracle_adaptive_weights(xyz, tri, low_xyz, low_tri)
    tri_ids = B.OracleOctree(xyz, tri).query(low_xyz)[0]
    return "port (oracle/msm_oracle.cpp)", ref_b, ref_a, (rp, col, val), tri_ids


def parity_block(R, capi, L, xyz, tri, low_xyz, low_tri, d_feat0, d_out_b0, d_out_a0):
    """GPU outputs of the timed subject 0 against the reference's outputs for the same inputs, at the full BASELINE configs[1] size,
    outside the timed region: nearest-triangle ids, adaptive CSR and FP64 metric_resample bit-exact; FP32-payload outputs within 1e-5."""
    t0 = time.perf_counter()
    feat_cm = d_feat0.T.contiguous().double().cpu().numpy()          # [D][nv]: the FP32 payload, exactly representable
    kind, ref_b, ref_a, (rp, col, val), ref_tri = reference_outputs(xyz, tri, low_xyz, low_tri, feat_cm)
    gpu_b = d_out_b0.T.contiguous().cpu().numpy().astype(np.float64)
    gpu_a = d_out_a0.T.contiguous().cpu().numpy().astype(np.float64)
    rel = lambda got, ref: float(np.abs(got - ref).max() / np.abs(ref).max())
    m = R.Mesh(xyz, tri, feat_cm)
    low = R.Mesh(low_xyz, low_tri)
    tree = R.Octree(m)
    got_tri = tree.get_closest_triangle(low_xyz)
    W = R.Resampler().get_adaptive_barycentric_weights(m, low)
    g_rp, g_col, g_val = W.csr()
    got64 = R.metric_resample(m, low)
    out = {"checked": True, "subject": 0, "against": kind, "reference_threads": {"barycentric": os.cpu_count() or 1, "adaptive": 1},
           "nearest_triangle_ids_bit_exact": bool(np.array_equal(got_tri, ref_tri)),
           "adaptive_csr_bit_exact": bool(np.array_equal(g_rp, rp) and np.array_equal(g_col, col) and np.array_equal(g_val, val)),
           "adaptive_f64_output_bit_exact": bool(np.array_equal(got64, ref_a)),
           "barycentric_f32_max_rel": rel(gpu_b, ref_b), "adaptive_f32_max_rel": rel(gpu_a, ref_a),
           "barycentric_f32_equals_rounded_reference": bool(np.array_equal(gpu_b, ref_b.astype(np.float32).astype(np.float64))),
           "adaptive_f32_equals_rounded_reference": bool(np.array_equal(gpu_a, ref_a.astype(np.float32).astype(np.float64))),
           "tolerance_f32": 1e-5, "nnz": int(len(col)), "seconds": None}
    out["ok"] = bool(out["nearest_triangle_ids_bit_exact"] and out["adaptive_csr_bit_exact"] and out["adaptive_f64_output_bit_exact"]
                     and out["barycentric_f32_max_rel"] <= 1e-5 and out["adaptive_f32_max_rel"] <= 1e-5)
    out["seconds"] = round(time.perf_counter() - t0, 1)
    return out


# --------------------------------------------------------------------------------------------
# our arm (GPU)
# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa(torch, local):
    """Pin this rank's host threads (and, by first touch, its page-locked staging buffers) to the CPUs NVML names as closest to its GPU:
    with several ranks per node the e2e path moves ~4.8 GB of host memory per rank and step, and a rank whose buffers sit on the other
    socket pays the inter-socket link on every copy. Best effort: a cpuset that excludes those CPUs leaves the affinity as it was."""
    info = {"cpus_before": len(os.sched_getaffinity(0))}
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        allowed = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        use = ideal & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
        info.update({"gpu_pci": bus, "ideal_cpus": len(ideal), "cpus_after": len(os.sched_getaffinity(0)), "bound": bool(use) and use != allowed})
    except Exception as ex:   # no NVML, no permission: measure unbound
        info.update({"bound": False, "note": f"{type(ex).__name__}: {ex}"})
    return info


def run_ours(a):
    if "WORLD_SIZE" in os.environ and "BENCH_KEEP_OMP" not in os.environ:
        # torchrun pins OMP_NUM_THREADS to 1; the library's host-side finishes (libm pow / acos of the cost paths) are OpenMP loops
        os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // int(os.environ["WORLD_SIZE"])))
    import torch
    import torch.distributed as dist
    from newmsm_b200 import build, capi, resampler as R

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_binding = bind_to_gpu_numa(torch, local) if (world > 1 and "BENCH_NO_BIND" not in os.environ) else {"bound": False, "note": "single rank"}
    build.build_library()
    L = capi.lib()
    assert L.msmgpu_device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"

    S, D = a.subjects, a.channels
    stream = torch.cuda.Stream(device=dev)
    ctx = R.Context(local, stream=stream.cuda_stream)

    # ---- synthetic inputs, resident in HBM before the timed region -------------------------------
    low_xyz, low_tri = target_sphere(a.target_freq)
    n_low = len(low_xyz)
    d_low_xyz = torch.from_numpy(low_xyz).to(dev)
    d_low_tri = torch.from_numpy(low_tri).to(dev)
    xyz0, tri = native_mesh(a.native_level, 0)
    nv, nt = len(xyz0), len(tri)
    d_tri = torch.from_numpy(tri).to(dev)
    host_xyz = []
    d_xyz, d_feat = [], []
    for s in range(S):
        x = xyz0 if s == 0 else native_mesh(a.native_level, rank * S + s)[0]
        host_xyz.append(x)
        xt = torch.from_numpy(x).to(dev)
        d_xyz.append(xt)
        d_feat.append(smooth_fields_torch(xt, D, 100 + rank * S + s, dev))       # [nv, D] f32 rows
    d_out_b = [torch.empty(n_low, D, device=dev) for _ in range(S)]
    d_out_a = [torch.empty(n_low, D, device=dev) for _ in range(S)]
    feat_ptrs = (capi.C.c_void_p * S)(*[t.data_ptr() for t in d_feat])
    outb_ptrs = (capi.C.c_void_p * S)(*[t.data_ptr() for t in d_out_b])
    outa_ptrs = (capi.C.c_void_p * S)(*[t.data_ptr() for t in d_out_a])
    torch.cuda.synchronize()

    stage_ms = {"mesh_tables+octree_forest": [], "bary_fused_batch": [], "adaptive_weights": [], "adaptive_apply": []}

    NW = max(1, min(a.value_workers, S))
    groups = [list(range(w, S, NW)) for w in range(NW)]
    wstreams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(NW - 1)]
    wctx = [ctx] + [R.Context(local, stream=wstreams[w].cuda_stream) for w in range(1, NW)]

    def group_step(w, record=None):
        """The whole path for the subjects of group w on its own stream: mesh tables + octree forest, fused barycentric
        resample, adaptive weights, adaptive apply."""
        g, st, cx = groups[w], wstreams[w], wctx[w]
        n = len(g)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record is not None else None
        torch.cuda.set_device(local)
        with torch.cuda.stream(st):
            if ev: ev[0].record(st)
            low = R.Mesh.from_device(cx, n_low, d_low_xyz, len(low_tri), d_low_tri)
            meshes = R.Mesh.views_from_device(cx, nv, [d_xyz[s_] for s_ in g], nt, d_tri)   # the subjects' coordinates are already in HBM: viewed, not copied
            trees = R.Octree.build_batch(meshes + [low])
            low_tree = trees[-1]
            if ev: ev[1].record(st)
            tree_ptrs = (capi.C.c_void_p * n)(*[t.h.value for t in trees[:n]])
            fp = (capi.C.c_void_p * n)(*[d_feat[s_].data_ptr() for s_ in g])
            ob = (capi.C.c_void_p * n)(*[d_out_b[s_].data_ptr() for s_ in g])
            oa = (capi.C.c_void_p * n)(*[d_out_a[s_].data_ptr() for s_ in g])
            # both methods need get_barycentric_weights(targets, subject): the fused resample keeps its weight maps and the adaptive
            # weights consume them (msmgpu.h: msmgpu_fwd) instead of querying the same points in the same trees again
            fwd = capi.C.c_void_p()
            capi.check(L.msmgpu_fwd_create(cx.h, n, n_low, capi.C.byref(fwd)))
            capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(cx.h, n, tree_ptrs, n_low, capi.ptr(d_low_xyz), D, fp, ob, d_status_g[w].data_ptr(), fwd))
            if ev: ev[2].record(st)
            mesh_ptrs = (capi.C.c_void_p * n)(*[m.h.value for m in meshes])
            w_ptrs = (capi.C.c_void_p * n)()
            capi.check(L.msmgpu_adaptive_weights_batch_fwd(cx.h, n, mesh_ptrs, tree_ptrs, low.h, low_tree.h, fwd, w_ptrs))
            L.msmgpu_fwd_destroy(fwd)
            ws = [R.Weights(L, capi.C.c_void_p(w_ptrs[i])) for i in range(n)]
            if ev: ev[3].record(st)
            capi.check(L.msmgpu_weights_apply_batch_f32_dev(cx.h, n, w_ptrs, D, fp, oa))
            if ev: ev[4].record(st)
            for w_ in ws: w_.close()
            for t in trees: t.close()
            for m in meshes: m.close()
            low.close()
        st.synchronize()   # a step's results are complete when it returns (and the stream-ordered pool reuses its blocks)
        if ev:
            for name, i in zip(stage_ms, range(4)):
                record[name].append(ev[i].elapsed_time(ev[i + 1]))

    d_status_g = [torch.zeros(len(g), n_low, dtype=torch.int32, device=dev) for g in groups]
    step_errors = []

    def device_step(record=None):
        if NW == 1:
            group_step(0, record)
            return
        def run(w):
            try:
                group_step(w, record if w == 0 else None)   # the stage split is reported for group 0 (they overlap anyway)
            except Exception as ex:
                step_errors.append(ex)
        th = [threading.Thread(target=run, args=(w,)) for w in range(NW)]
        for t in th: t.start()
        for t in th: t.join()
        if step_errors:
            raise step_errors[0]

    def sync_all():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    clocks = ClockSampler(local)
    dbg = []
    for _ in range(max(a.warmup, 3)):
        t0 = time.perf_counter()
        device_step()
        stream.synchronize()
        dbg.append(round(1e3 * (time.perf_counter() - t0), 1))
    sync_all()
    if os.environ.get("BENCH_DEBUG"):
        print(f"[debug] warm-up steps, host ms: {dbg}", file=sys.stderr, flush=True)
    if os.environ.get("BENCH_DEBUG"):
        for mode in ("sync each step", "no sync"):
            ts = []
            for _ in range(6):
                t0 = time.perf_counter()
                device_step()
                if mode.startswith("sync"):
                    stream.synchronize()
                ts.append(1e3 * (time.perf_counter() - t0))
            stream.synchronize()
            print(f"[debug] {mode}: host ms per step {[round(t, 1) for t in ts]}", file=sys.stderr, flush=True)
    assert all(int(t.abs().max().item()) == 0 for t in d_status_g), "a nearest-triangle query failed"
    launches0 = L.msmgpu_launch_count()
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(a.steps):
        device_step()          # every group's stream is synchronised before it returns
    e1.record(stream)          # e0 / e1 sit on an otherwise idle stream: their difference is the device-clock span of the K steps
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = int(L.msmgpu_launch_count() - launches0)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    verts_per_step = 2 * S * n_low * world
    value = verts_per_step / (ms_step * 1e-3)

    # ---- stage breakdown + the dominant kernel alone (roofline) ----------------------------------
    for _ in range(2):
        device_step(stage_ms)
    breakdown = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    tune = lambda name, v: capi.check(L.msmgpu_set_tuning(name.encode(), int(v)))

    def timed_launches(fn, reps=7, warm=3):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        for i in range(reps + warm):      # warm launches first; inputs (4.2 GB of feature rows) exceed the 126 MB L2
            if i >= warm: evs[i - warm].record(stream)
            fn()
        evs[reps].record(stream)
        stream.synchronize()
        return float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]))

    with torch.cuda.stream(stream):
        low = R.Mesh.from_device(ctx, n_low, d_low_xyz, len(low_tri), d_low_tri)
        meshes = [R.Mesh.from_device(ctx, nv, d_xyz[s], nt, d_tri) for s in range(S)]
        trees = R.Octree.build_batch(meshes + [low])
        tree_ptrs = (capi.C.c_void_p * S)(*[t.h.value for t in trees[:S]])
        fwd = capi.C.c_void_p()
        capi.check(L.msmgpu_fwd_create(ctx.h, S, n_low, capi.C.byref(fwd)))
        bary_call = lambda: capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tree_ptrs, n_low, capi.ptr(d_low_xyz), D, feat_ptrs, outb_ptrs, None, fwd))
        bary_ms = timed_launches(bary_call)                                   # the path of the step: queries kernel + bulk-copy gather
        gather_ms = timed_launches(lambda: capi.check(L.msmgpu_fwd_apply_batch_f32_dev(ctx.h, fwd, D, feat_ptrs, outb_ptrs)))   # the gather kernel alone
        mode0 = int(os.environ.get("MSMGPU_GATHER", "1"))
        tune("gather", 0)
        fused_ms = timed_launches(bary_call)                                  # the fused register-path kernel (A/B reference)
        tune("gather", mode0)
        # adaptive apply alone (one launch for the batch)
        mesh_ptrs = (capi.C.c_void_p * S)(*[m.h.value for m in meshes])
        w_ptrs = (capi.C.c_void_p * S)()
        capi.check(L.msmgpu_adaptive_weights_batch_fwd(ctx.h, S, mesh_ptrs, tree_ptrs, low.h, trees[-1].h, fwd, w_ptrs))
        Ws = [R.Weights(L, capi.C.c_void_p(w_ptrs[s])) for s in range(S)]
        nnz = sum(w.shape()[2] for w in Ws)
        apply_ms = timed_launches(lambda: capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, w_ptrs, D, feat_ptrs, outa_ptrs)), warm=2)
        L.msmgpu_fwd_destroy(fwd)
        for W in Ws: W.close()
        for t_ in trees: t_.close()
        for m in meshes: m.close()
        low.close()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # B_bary = 24 N_t + 24 V_s + 12 T_s + 4 D min(3 N_t, V_s) + 4 D N_t  per subject; the gather kernel's share of it is the
    # compulsory feature rows + the output rows, plus the 40-byte weight map it reads per target (ids, weights, count)
    bytes_bary = 24 * n_low + 24 * nv + 12 * nt + 4 * D * min(3 * n_low, nv) + 4 * D * n_low
    bytes_gather = 4 * D * min(3 * n_low, nv) + 4 * D * n_low + 36 * n_low
    bytes_apply = S * (4 * D * nv + 4 * D * n_low + 4 * (n_low + 1)) + 12 * nnz
    bytes_adaptive = 4 * D * nv + 4 * D * n_low + 8 * (nv + n_low) + 24 * nv + 12 * nt      # SURVEY §8d "adaptive": every source row + areas + mesh
    gbs = lambda b, ms: b / (ms * 1e-3) / 1e9
    # DRAM traffic of the gather kernel per launch from the committed `ncu --set full` capture (bytes per subject x subjects in this launch)
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_gather_traffic.json")))
        traffic = float(tr["dram_bytes_per_subject"]) * S
        traffic_src = "profiles/r2_gather_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of a 16-subject launch, scaled to %d subjects)" % S
    except Exception:
        pass
    roofline = {"kernel": "k_gather_rows_bulk<BARY> (cp.async.bulk + mbarrier row gather of the barycentric resample: one launch for the batch)",
                "bound": "hbm", "achieved": gbs(S * bytes_gather, gather_ms), "peak": peak, "unit": "GB/s", "frac": gbs(S * bytes_gather, gather_ms) / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": gather_ms,
                "algorithmic_bytes_per_launch": S * bytes_gather,
                "algorithmic_bytes_note": "per subject: 4 D min(3 N_t, V_s) feature rows + 4 D N_t output + 36 N_t weight maps (SURVEY §8d feature / output terms)",
                "barycentric_total": {"kernels": "k_bary_weights_batch (queries + weight maps) + k_gather_rows_bulk", "launch_ms": bary_ms,
                                      "algorithmic_bytes_per_launch": S * bytes_bary, "achieved": gbs(S * bytes_bary, bary_ms),
                                      "frac": gbs(S * bytes_bary, bary_ms) / peak},
                "barycentric_fused_register_path": {"kernel": "k_bary_resample_f32 (round-1 kernel, MSMGPU_GATHER=0)", "launch_ms": fused_ms,
                                                    "achieved": gbs(S * bytes_bary, fused_ms), "frac": gbs(S * bytes_bary, fused_ms) / peak},
                "adaptive_apply": {"kernel": "k_csr_apply_f32x4", "launch_ms": apply_ms, "algorithmic_bytes_per_launch": bytes_apply,
                                   "achieved": gbs(bytes_apply, apply_ms), "frac": gbs(bytes_apply, apply_ms) / peak,
                                   "bound_note": "XU pipe (FP32->FP64 conversions) 47 % busy, issue 51 %: profiles/r2_gather_summary.md"},
                "whole_step": {"algorithmic_bytes_per_step": S * (bytes_bary + bytes_adaptive), "ms_per_step": ms_step,
                               "achieved": gbs(S * (bytes_bary + bytes_adaptive), ms_step), "frac": gbs(S * (bytes_bary + bytes_adaptive), ms_step) / peak,
                               "note": "both methods' SURVEY §8d bytes over the whole step (octree builds, queries, weight construction, both gathers)"}}

    # ---- end to end through the host-buffer C ABI ------------------------------------------------
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, torch, R, capi, L, local, S, D, host_xyz, tri, low_xyz, low_tri, d_feat, n_low, nv, nt, world, dist, d_out_b, d_out_a)

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            cpu_base = cpu_arm(a, 2, 1)[0]
            cpu_base = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:   # the checker is optional for the measurement itself
            cpu_base = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {ex}"}

    gmsm = None
    if not a.no_gmsm:
        try:
            gmsm = run_gmsm(a, torch, dist, rank, world, local)
        except Exception as ex:
            gmsm = {"error": f"{type(ex).__name__}: {ex}"}
    parity = None
    if rank == 0 and not a.no_parity:
        try:
            parity = parity_block(R, capi, L, host_xyz[0], tri, low_xyz, low_tri, d_feat[0], d_out_b[0], d_out_a[0])
        except Exception as ex:     # the checker is test infrastructure: its absence must not hide the measurement
            parity = {"checked": False, "error": f"{type(ex).__name__}: {ex}"}

    adapter = None
    if rank == 0 and world == 1 and not a.no_adapter_e2e:
        try:
            adapter = run_adapter_e2e(a, host_xyz[0], tri, low_xyz, low_tri)
        except Exception as ex:
            adapter = {"error": f"{type(ex).__name__}: {ex}"}

    # BASELINE.json's second quantity, "unary costs/s": one unary cost table N_cp x L (control grid ico4 on data grid ico6, 19 labels)
    # through msmgpu_costfn_unary_table, univariate and multivariate; the oracle port on the host cores beside it, compared bit for bit
    unary = None
    if rank == 0 and world == 1 and not a.no_unary:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import bench_unary
            unary = {"cases": bench_unary.run_cases(4, 6, 10, cpu=not a.no_cpu_baseline),
                     "api": "msmgpu_costfn_unary_table: host rotation matrices + upload + k_unary_table + table download, median of 10"}
        except Exception as ex:
            unary = {"error": f"{type(ex).__name__}: {ex}"}

    newmsm = None
    if rank == 0 and world == 1 and not a.no_newmsm:
        try:
            newmsm = run_newmsm_leg()
        except Exception as ex:
            newmsm = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        cfg = workload_config(a, nv, nt, n_low)
        detail = {"query_group_lanes": int(L.msmgpu_get_query_group()), "breakdown_ms_per_step": breakdown, "value_streams": NW,
                  "breakdown_note": f"stage times of one of the {NW} concurrent subject groups ({len(groups[0])} subjects)",
                  "host_binding_rank0": host_binding}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clk, "gpu_launches": launches, "roofline": roofline, "detail": detail}
        if parity is not None:
            line["parity"] = parity
        if gmsm is not None:
            line["gmsm"] = gmsm
        if e2e is not None:
            line["e2e"] = e2e
        if adapter is not None:
            line["e2e_adapter"] = adapter
        if unary is not None:
            line["unary_costs"] = unary
        if newmsm is not None:
            line["newmsm_wall_time"] = newmsm
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_gmsm(a, torch, dist, rank, world, local):
    """Secondary leg: ONE discrete iteration of groupwise registration at BASELINE configs[4] scale (S synthetic ico6 subjects, ico6
    template, ico4 control grids, 19 labels; docs/guide.md:390-407) on the ranks of this run:
      get_patch_data   subjects block-sharded, every rank resamples its subjects for all labels (msmgpu_group_fields)
      ONE ncclAllGather of the fields [S][L][N_t][D] f64 (torch.distributed all_gather_into_tensor on the device buffers)
      pair costs       pair blocks sharded, 18 label phases of Fusion's 4 combinations, blocks gathered on the device for the host solver
      triplet costs    triplet blocks sharded, 18 label phases of 8 combinations (strain; pow() finish on the host libm)
    Times are max over ranks. Rank 0 re-computes a sample unsharded and compares bit for bit."""
    from newmsm_b200 import group_cost as GC, resampler as R, synth
    S, D = a.gmsm_subjects, 1
    dev = torch.device("cuda", local)
    ctx = R.Context(local)
    cp0, cp_tri = synth.icosphere(a.gmsm_cp_level)
    d0, dtri = synth.icosphere(a.gmsm_data_level)
    tpl, tpl_tri = d0.copy(), dtri
    data = np.stack([synth.smooth_warp(d0, max_disp=2.0, seed=40 + s_) for s_ in range(S)])
    cps = np.stack([synth.smooth_warp(cp0, max_disp=1.0, seed=60 + s_) for s_ in range(S)])
    feat = np.stack([synth.smooth_fields(data[s_], D, seed0=100, noise=0.1, noise_seed=7 + s_) for s_ in range(S)])
    centre = np.array([0.0, 0.0, 100.0])
    spacing = 2 * 100 * np.arcsin(np.linalg.norm(cp0[cp_tri[:, 0]] - cp0[cp_tri[:, 1]], axis=1).max() / 200)
    labels = [centre]
    for ring, n in ((0.25, 6), (0.5, 12)):      # 19 labels like the barycentre + vertex sets of the sampling grid (DiscreteModel.cpp:124-190)
        for k in range(n):
            q = centre + ring * spacing * np.array([np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n), 0.0])
            labels.append(q / np.linalg.norm(q) * 100)
    labels = np.array(labels)
    Lb = len(labels)
    ncp = len(cp0)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def tmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    M = GC.DiscreteGroupModel(R.Mesh(tpl, tpl_tri, ctx=ctx), simmeasure=2, dist=dist if world > 1 else None)
    spac = M.get_spacings(cps, cp_tri)
    rot = M.get_rotations(centre, cps)
    pairs = M.estimate_pairs(cps, cp_tri)
    trip = np.concatenate([np.sort(cp_tri + s_ * ncp, axis=1) for s_ in range(S)]).astype(np.int32)
    orig = np.stack([cp0] * S)
    labeling = np.zeros(S * ncp, np.int32)
    launches0 = capi_launches()
    out = {}
    for it in range(2):                          # one warm iteration, one timed
        sync(); t0 = time.perf_counter()
        M.get_patch_data(data, dtri, feat, labels, centre, rot, spac, 1.0)          # fields + all-gather + group state
        sync(); t1 = time.perf_counter()
        sample = []
        for l in range(1, Lb):       # the host solver lives on rank 0: every rank computes its block and gathers, rank 0 downloads the table
            c = M.computePairwiseCostsForLabel(pairs, labeling, l, copy=False, to_host=(rank == 0))
            if c is not None: sample.append(c[::997].ravel().copy())
        sync(); t2 = time.perf_counter()
        tsum = 0.0
        M.reset_triplet_state(cps, orig, rot, labels, trip)      # once per iteration, like reset_CPgrid / estimate_triplets
        for l in range(1, Lb):
            c = M.computeTripletCostsForLabel(None, None, None, None, None, labeling, l, 0.2, copy=False, to_host=(rank == 0))
            if c is not None: tsum += float(c[::997].sum())
        sync(); t3 = time.perf_counter()
        out = {"fields_s": tmax(t1 - t0), "pair_sweep_s": tmax(t2 - t1), "triplet_sweep_s": tmax(t3 - t2), "iteration_s": tmax(t3 - t0)}
    # the collective alone: all-gather of the field shards, CUDA events on this rank's stream, max over ranks
    ag_bytes = int(M.fields.numel() * 8)
    ag_ms = None
    if world > 1:
        b, e = GC.shard_range(S, rank, world)
        local_blk = M.fields[b:e].contiguous()
        full = torch.empty_like(M.fields)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        uneven = len(set(GC.shard_counts(S, world))) > 1
        for i in range(4):
            sync()
            e0.record()
            if uneven: M.coll.all_gather_blocks(local_blk, S)
            else: dist.all_gather_into_tensor(full, local_blk)
            e1.record()
            torch.cuda.synchronize(dev)
            ag_ms = tmax(e0.elapsed_time(e1))
    # bitwise check on rank 0: the sharded results against an unsharded computation of the same batch on this rank alone
    same = None
    if world > 1:
        lab_chk = 7
        got_pairs = M.computePairwiseCostsForLabel(pairs, labeling, lab_chk).copy()
        got_trip = M.computeTripletCostsForLabel(None, None, None, None, None, labeling, lab_chk, 0.2)   # the resident plan, sharded
        # every rank holds the gathered fields; only rank 0 re-computes
        if rank == 0:
            M1 = GC.DiscreteGroupModel(R.Mesh(tpl, tpl_tri, ctx=ctx), simmeasure=2, dist=None)
            other = [s_ for s_ in (S - 1, S // 2) if not (GC.shard_range(S, 0, world)[0] <= s_ < GC.shard_range(S, 0, world)[1])][:1] or [S - 1]
            M1.get_patch_data(data[other], dtri, feat[other], labels, centre, rot[other[0] * ncp:(other[0] + 1) * ncp], spac[other], 1.0)
            same_fields = bool(torch.equal(M1.fields[0], M.fields[other[0]]))
            M1.close()
            M.coll = GC.Collective(None)         # this rank alone, on the gathered fields
            M._pairs_key = None
            ref_pairs = M.computePairwiseCostsForLabel(pairs, labeling, lab_chk)
            ref_trip = M.computeTripletCostsForLabel(cps, orig, rot, labels, trip, labeling, lab_chk, 0.2)
            eq = lambda x, y: bool(np.array_equal(np.nan_to_num(x, nan=-1.0), np.nan_to_num(y, nan=-1.0)))
            same = {"fields_of_a_remote_subject": same_fields, "pair_batch": eq(got_pairs, ref_pairs), "triplet_batch": eq(got_trip, ref_trip)}
        dist.barrier()
    P, T = len(pairs), len(trip)
    res = {"workload": f"gMSM, one discrete iteration: {S} subjects ico{a.gmsm_data_level} ({len(d0)} V), template ico{a.gmsm_data_level}, control grids ico{a.gmsm_cp_level} "
                       f"({ncp} nodes per subject), {Lb} labels, D = {D}, simmeasure 2",
           "n_gpus": world, "subjects": S, "labels": Lb, "pairs": P, "triplets": T, "resamples_per_iteration": S * Lb,
           **out,
           "pair_costs_per_s": P * 4 * (Lb - 1) / out["pair_sweep_s"], "triplet_costs_per_s": T * 8 * (Lb - 1) / out["triplet_sweep_s"],
           "allgather": {"collective": "ncclAllGather (torch.distributed.all_gather_into_tensor on the device field shards)" if world > 1 else "none (1 rank)",
                         "bytes_total": ag_bytes, "ms": ag_ms,
                         "bus_GBs": (ag_bytes * (world - 1) / world / (ag_ms * 1e-3) / 1e9) if ag_ms else None},
           "sharded_equals_unsharded_bitwise": same,
           "limiter_note": "fields: host libm rotation matrices + per-rank build; pairs: device-bound (k_group_pair_costs_thread); triplets: strain + the three "
                           "pow() per cost on the device (glibc's algorithm with the host library's tables, csrc/hostpow.cuh; device_pow_enabled below), "
                           "bound by the D2H copy of the [T][8] table per label phase; the cost tables of a label phase are gathered on every rank's device (NCCL) "
                           "and downloaded by rank 0 only, where the host solver runs",
           "device_pow_enabled": bool(capi_device_pow()),
           "gpu_launches": int(capi_launches() - launches0), "timing": "wall clock between device synchronisations + barriers, max over ranks; second of two iterations"}
    M.close()      # the context is released with the last object that holds it (meshes / trees keep a reference)
    return res


def run_adapter_e2e(a, xyz, tri, low_xyz, low_tri):
    """The reference-facing C++ adapter on the reference's own types: newresampler_gpu::metric_resample(Mesh, Mesh) with pageable FP64
    Mesh::pvalues in and a reference Mesh out (include/newmsm_b200/resampler_adapter.hpp), one subject of the bench workload per call,
    timed by integration/_build/adapter_bench (built in the container: it contains the compiled reference Mesh class) next to
    newresampler::metric_resample of the compiled reference on the same objects; outputs compared bit for bit."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "integration", "_build", "adapter_bench")
    if not os.path.exists(exe):
        return {"unavailable": "integration/_build/adapter_bench not built (needs /root/reference at build time)"}

    def write_asc(path, v, t):
        with open(path, "w") as f:
            f.write("#!ascii synthetic\n%d %d\n" % (len(v), len(t)))
            np.savetxt(f, np.column_stack([v, np.zeros(len(v))]), fmt="%.17g %.17g %.17g %d")
            np.savetxt(f, np.column_stack([t, np.zeros(len(t), np.int64)]), fmt="%d")
    with tempfile.TemporaryDirectory(prefix="adapter_e2e_") as d:
        write_asc(os.path.join(d, "in.asc"), xyz, tri)
        write_asc(os.path.join(d, "low.asc"), low_xyz, low_tri)
        r = subprocess.run([exe, "--in", os.path.join(d, "in.asc"), "--low", os.path.join(d, "low.asc"), "--D", str(a.channels), "--reps", "4",
                            "--ref-threads", str(os.cpu_count() or 1)], capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
    if r.returncode != 0:
        return {"error": (r.stdout + r.stderr)[-500:]}
    res = json.loads(r.stdout.strip().splitlines()[-1])
    n_low = len(low_xyz)
    res["verts_per_s"] = n_low / (res["adapter_ms"] * 1e-3)
    res["reference_verts_per_s"] = n_low / (res["reference_ms"] * 1e-3)
    res["speedup_vs_reference"] = res["reference_ms"] / res["adapter_ms"]
    res["note"] = ("adaptive-barycentric method only (metric_resample, resampler.cpp:304), one subject per call; every millisecond between the "
                   "caller's Mesh and the returned Mesh is inside adapter_ms (flattening Mesh::pvalues, H2D/D2H of pageable FP64, the Mesh copy "
                   "the reference also makes, resampler.cpp:37-38)")
    return res


def capi_device_pow():
    from newmsm_b200 import capi
    return int(capi.lib().msmgpu_device_pow_enabled())


def capi_launches():
    from newmsm_b200 import capi
    return int(capi.lib().msmgpu_launch_count())


def run_e2e(a, torch, R, capi, L, local, S, D, host_xyz, tri, low_xyz, low_tri, d_feat, n_low, nv, nt, world, dist, d_out_b, d_out_a):
    """Per subject, through the calls a reference-side adapter makes: Mesh upload (msmgpu_mesh_create), Octree
    (msmgpu_octree_build), barycentric resample and metric_resample on HOST channel-major FP32 buffers."""
    C = capi.C
    workers = max(1, min(a.e2e_workers, S))
    # pinned host buffers: features [D][nv] per subject (the reference's pvalues layout), outputs [D][n_low]
    h_feat = [torch.empty(D, nv, dtype=torch.float32).pin_memory() for _ in range(S)]
    for s in range(S):
        h_feat[s].copy_(d_feat[s].T)
    h_out_b = [torch.empty(D, n_low, dtype=torch.float32).pin_memory() for _ in range(S)]
    h_out_a = [torch.empty(D, n_low, dtype=torch.float32).pin_memory() for _ in range(S)]
    h_xyz = [torch.from_numpy(x).pin_memory() for x in host_xyz]
    h_tri = torch.from_numpy(tri).pin_memory()
    h_low = torch.from_numpy(low_xyz).pin_memory()
    h_low_tri = torch.from_numpy(low_tri).pin_memory()
    torch.cuda.synchronize()
    ctxs = [R.Context(local) for _ in range(workers)]
    errors = []

    def work(w):
        try:
            torch.cuda.set_device(local)
            ctx = ctxs[w]
            low = C.c_void_p(); low_tree = C.c_void_p()
            capi.check(L.msmgpu_mesh_create(ctx.h, n_low, capi.ptr(h_low), len(low_tri), capi.ptr(h_low_tri), C.byref(low)))
            capi.check(L.msmgpu_octree_build(low, C.byref(low_tree)))
            for s in range(w, S, workers):
                m = C.c_void_p(); t = C.c_void_p()
                capi.check(L.msmgpu_mesh_create(ctx.h, nv, capi.ptr(h_xyz[s]), nt, capi.ptr(h_tri), C.byref(m)))
                capi.check(L.msmgpu_mesh_set_features_f32(m, D, capi.ptr(h_feat[s])))      # Mesh::pvalues, uploaded once
                capi.check(L.msmgpu_octree_build(m, C.byref(t)))
                capi.check(L.msmgpu_mesh_bary_resample_f32(t, n_low, capi.ptr(h_low), capi.ptr(h_out_b[s])))
                capi.check(L.msmgpu_mesh_metric_resample_f32(m, t, low, low_tree, capi.ptr(h_out_a[s])))
                L.msmgpu_octree_destroy(t)
                L.msmgpu_mesh_destroy(m)
            L.msmgpu_octree_destroy(low_tree)
            L.msmgpu_mesh_destroy(low)
        except Exception as ex:   # surfaced after the join
            errors.append(ex)

    def step():
        th = [threading.Thread(target=work, args=(w,)) for w in range(workers)]
        for t in th: t.start()
        for t in th: t.join()
        if errors:
            raise errors[0]

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    # (1) the batch entry point: ONE call per step, chunks of subjects pipelined inside the library (csrc/batch.cu)
    bctx = R.Context(local)

    def batch_step():
        R.resample_batch_host(bctx, h_xyz, h_tri, h_low, h_low_tri, h_feat, h_out_b, h_out_a, chunk=a.e2e_chunk)
    dt_batch = timed(batch_step)
    tb = torch.tensor([dt_batch], device=torch.device("cuda", local), dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    dt_batch = float(tb.item())
    same_batch = True
    for s_ in sorted({0, S - 1}):
        same_batch = same_batch and bool(torch.equal(h_out_b[s_], d_out_b[s_].T.cpu())) and bool(torch.equal(h_out_a[s_], d_out_a[s_].T.cpu()))
    for t_ in h_out_b + h_out_a:
        t_.zero_()
    bctx.close()
    # (2) the per-subject calls a reference-side adapter makes, on worker threads
    dt = timed(step)
    tt = torch.tensor([dt], device=torch.device("cuda", local), dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    # the host-path outputs equal the device-path outputs of the same subjects bit for bit (same kernels, same inputs)
    same = True
    for s_ in sorted({0, S - 1}):
        same = same and bool(torch.equal(h_out_b[s_], d_out_b[s_].T.cpu())) and bool(torch.equal(h_out_a[s_], d_out_a[s_].T.cpu()))
    h2d = S * (D * nv * 4 + nv * 24 + nt * 12 + n_low * 24) + workers * (n_low * 24 + len(low_tri) * 12)
    d2h = S * 2 * D * n_low * 4
    for c in ctxs: c.close()
    h2d_batch = S * (D * nv * 4 + nv * 24) + nt * 12 + n_low * 24 + len(low_tri) * 12
    per_subject = {"value": 2 * S * n_low * world * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": 1e3 * dt / a.steps, "workers": workers, "outputs_equal_device_path": same,
                   "api": "per subject: msmgpu_mesh_create + msmgpu_mesh_set_features_f32 + msmgpu_octree_build + msmgpu_mesh_bary_resample_f32 + "
                          "msmgpu_mesh_metric_resample_f32, pinned host buffers, worker threads"}
    return {"value": 2 * S * n_low * world * a.steps / dt_batch, "unit": UNIT, "h2d_bytes_per_step": int(h2d_batch), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": 1e3 * dt_batch / a.steps, "outputs_equal_device_path": same_batch, "chunk_subjects": a.e2e_chunk or 4,
            "api": "one call per step: msmgpu_resample_batch_host_f32 (host FP32 features / FP64 coordinates in, host FP32 outputs of both methods "
                   "out; chunks of subjects pipelined over copy-in / compute / copy-out streams inside the library), pinned host buffers",
            "h2d_GBs": h2d_batch * a.steps / dt_batch / 1e9,
            "per_subject_calls": per_subject}


def run_newmsm_leg():
    """BASELINE.json's third quantity, "`newmsm` wall-time vs CPU cores", on a bounded case: the reference's own program — CLI, config
    parsing, drivers, FastPD solver, unmodified — with libmsmgpu.so bound in at link time (integration/_build/newmsm_gpu) against the
    same program on the host cores (oracle/_ref/newmsm_ref_trace, all threads), on `basic_configs/config_standard_MSMpair` semantics
    (three DISCRETE levels, 19 iterations) at ico5. Parity: the labeling and every traced mesh of every iteration against the
    single-thread reference trace recorded by tests/newmsm_e2e.py --cpu-trace-out (the multi-threaded reference races).
    Full-size cases (ico6, cfg1 / cfg3 / cfg4, gMSM) take minutes per arm: profiles/*newmsm_e2e*, DESIGN §6b."""
    trace = os.path.join(ROOT, "tests", "golden", "newmsm_cfg1_MSMpair_ico5_single_thread_trace.txt")
    for need in (trace, os.path.join(ROOT, "integration", "_build", "newmsm_gpu"), os.path.join(ROOT, "oracle", "_ref", "newmsm_ref_trace")):
        if not os.path.exists(need):
            return {"unavailable": "missing " + os.path.relpath(need, ROOT)}
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res.json")
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "newmsm_e2e.py"), "--level", "5", "--config", "MSMpair", "--D", "1",
                        "--threads", str(threads), "--cpu-trace-in", trace, "--out", out],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
        r = json.load(open(out))
    return {"workload": "newmsm, config_standard_MSMpair semantics (3 DISCRETE levels, FastPD, D = 1), ico5 input / reference meshes, synthetic warp",
            "gpu_wall_s": r["gpu_wall_s"], "cpu_wall_s": r["cpu_wall_s"], "cpu_threads": r["cpu_threads"], "speedup": r["speedup"],
            "discrete_iterations": r["discrete_iterations"], "labels_bit_exact": r.get("labels_bit_exact"),
            "all_meshes_bit_exact": r.get("all_meshes_bit_exact"), "parity_against": "single-thread reference trace " + r.get("cpu_parity_trace", "?"),
            "gpu_split": r["gpu_split"][-1] if r.get("gpu_split") else None,
            "note": "process wall clock of both programs incl. file I/O and CUDA start-up; the GPU run's remainder is the reference's host solver"}


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
