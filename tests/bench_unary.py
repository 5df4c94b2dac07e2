"""Secondary benchmark (BASELINE.json metric "unary costs/s", configs[0]/[2] shapes): one unary cost table
N_cp x L for control grid ico4 on data grid ico6 (2 562 x 19 = 48 678 costs, ~3.2 M nearest-triangle queries),
univariate (D = 1) and multivariate (D = 40) correlation. Prints one JSON line per case:
  value        costs/s through msmgpu_costfn_unary_table (host rotation matrices + upload + kernel + table download)
  kernel_ms    k_unary_table alone (CUDA events around msmgpu_costfn_unary_table_dev minus nothing: includes the R upload)
  cpu_baseline the CPU restatement of the reference loop (oracle port, all host threads) on the same inputs
TEST / MEASUREMENT INFRASTRUCTURE (its cpu_baseline leg runs the oracle port).
Usage: python tests/bench_unary.py [--cp 4] [--data 6] [--reps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newmsm_b200 import capi, discrete_cost as DC, resampler as R, synth  # noqa: E402


def label_grid(spacing):
    """19 labels: centre + 6 + 12 points on two rings around the pole (the size of the reference's SG = CP + 2 label set)."""
    centre = np.array([0.0, 0.0, 100.0])
    pts = [centre]
    for ring, n in ((0.5, 6), (1.0, 12)):
        for k in range(n):
            a = 2 * np.pi * k / n
            p = centre + ring * spacing * np.array([np.cos(a), np.sin(a), 0.0])
            pts.append(p / np.linalg.norm(p) * 100)
    return np.array(pts)


def run_cases(cp_level=4, data_level=6, reps=10, cpu=True):
    """One JSON-able dict per case (univariate D = 1, multivariate D = 40). cpu: also time the oracle port on the same table and
    compare bit for bit (the only place this module touches oracle/)."""
    cp, cp_tri = synth.icosphere(cp_level)
    xyz, tri = synth.icosphere(data_level)
    src = synth.smooth_warp(xyz, max_disp=4.0, seed=2024)
    e = np.zeros(len(cp))
    for i, j in ((0, 1), (1, 2), (0, 2)):
        d = 2 * 100 * np.arcsin(np.linalg.norm(cp[cp_tri[:, i]] - cp[cp_tri[:, j]], axis=1) / 200)
        np.maximum.at(e, cp_tri[:, i], d)
        np.maximum.at(e, cp_tri[:, j], d)
    labels = label_grid(0.5 * e.mean())
    centre = np.array([0.0, 0.0, 100.0])
    rot = R.estimate_rotation_matrix(np.tile(centre, (len(cp), 1)), cp).reshape(-1, 9)
    target = R.Mesh(xyz, tri)
    tree = R.Octree(target)
    out = []
    for name, cls, kind, D in (("univariate corr, D=1", DC.UnivariateNonLinearSRegDiscreteCostFunction, 0, 1),
                               ("multivariate corr, D=40", DC.MultivariateNonLinearSRegDiscreteCostFunction, 1, 40)):
        ref_feat = synth.smooth_fields(xyz, D)
        src_feat = synth.smooth_fields(src, D, noise=0.05)
        cf = cls(simmeasure=DC.CORRELATION)
        cf.set_meshes(target, src, src_feat, ref_feat, tree)
        t0 = time.perf_counter()
        cf.reset_CPgrid(cp, e, 1.0)
        t_patch = time.perf_counter() - t0
        prow, pmem = cf.get_source_data()
        cf.computeUnaryCosts(labels, rot)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            costs = cf.computeUnaryCosts(labels, rot)
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        n_costs = costs.size
        line = {"metric": "unary costs/s", "case": name, "cp_grid": f"ico{cp_level}", "data_grid": f"ico{data_level}", "labels": len(labels),
                "costs": int(n_costs), "patch_points": int(prow[-1]), "queries_per_table": int(prow[-1]) * len(labels),
                "value": n_costs / t, "unit": "costs/s", "ms_per_table": 1e3 * t, "resampled_verts_per_s": int(prow[-1]) * len(labels) / t,
                "patch_membership_ms": 1e3 * t_patch, "query_group_lanes": int(capi.lib().msmgpu_get_query_group())}
        if cpu:
            from oracle import bindings as O
            ot = O.OracleOctree(xyz, tri)
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            t0 = time.perf_counter()
            ref = O.oracle_unary_costs(kind, 2, ot, cp, rot, labels, src, prow, pmem, src_feat, ref_feat, None, np.ones(len(cp)), nthreads=threads)
            tc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_costs / tc, "unit": "costs/s", "cores": threads, "kind": "port", "sample": "the same full table, one pass"}
            line["bit_exact_vs_cpu"] = bool(np.array_equal(ref, costs))
        out.append(line)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cp", type=int, default=4)
    ap.add_argument("--data", type=int, default=6)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    for line in run_cases(a.cp, a.data, a.reps, not a.no_cpu):
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
