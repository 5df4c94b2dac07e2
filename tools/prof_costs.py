"""Cost kernels under the profiler: one unary table (k_unary_table), one HO triplet batch (k_triplet_costs) and one
strain-only triplet batch (k_strain_costs) at the sizes of the last level of the shipped configs (control grid ico4, data grid ico6).
The groupwise kernels are profiled through bench.py's gMSM leg. Prints the wall-clock of each call; run under
  ncu --set full --clock-control none --import-source on -k regex:'k_unary_table|k_triplet_costs|k_strain_costs' -c 6 python tools/prof_costs.py
Usage: python tools/prof_costs.py [--cp 4] [--data 6] [--D 40]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from newmsm_b200 import discrete_cost as DC, resampler as R, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cp", type=int, default=4)
    ap.add_argument("--data", type=int, default=6)
    ap.add_argument("--D", type=int, default=40)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    cp, cp_tri = synth.icosphere(a.cp)
    xyz, tri = synth.icosphere(a.data)
    src = synth.smooth_warp(xyz, max_disp=4.0, seed=2024)
    e = np.zeros(len(cp))
    for i, j in ((0, 1), (1, 2), (0, 2)):
        d = 2 * 100 * np.arcsin(np.linalg.norm(cp[cp_tri[:, i]] - cp[cp_tri[:, j]], axis=1) / 200)
        np.maximum.at(e, cp_tri[:, i], d)
        np.maximum.at(e, cp_tri[:, j], d)
    centre = np.array([0.0, 0.0, 100.0])
    labels = [centre]
    for ring, n in ((0.5, 6), (1.0, 12)):
        for k in range(n):
            ang = 2 * np.pi * k / n
            p = centre + ring * 0.5 * e.mean() * np.array([np.cos(ang), np.sin(ang), 0.0])
            labels.append(p / np.linalg.norm(p) * 100)
    labels = np.array(labels)
    rot = R.estimate_rotation_matrix(np.tile(centre, (len(cp), 1)), cp).reshape(-1, 9)
    target = R.Mesh(xyz, tri)
    tree = R.Octree(target)
    triplets = np.sort(cp_tri, axis=1).astype(np.int32)
    labeling = np.random.default_rng(3).integers(0, len(labels), len(cp)).astype(np.int32)
    for name, cls, D, ho in (("unary table, multivariate corr", DC.MultivariateNonLinearSRegDiscreteCostFunction, a.D, False),
                             ("unary table, univariate corr", DC.UnivariateNonLinearSRegDiscreteCostFunction, 1, False),
                             ("HO multivariate triplet batch", DC.HOMultivariateNonLinearSRegDiscreteCostFunction, a.D, True)):
        ref_feat = synth.smooth_fields(xyz, D)
        src_feat = synth.smooth_fields(src, D, noise=0.05)
        cf = cls(simmeasure=DC.CORRELATION)
        cf.set_meshes(target, src, src_feat, ref_feat, tree)
        if ho:
            cf.reset_CPgrid(cp, cp_tri)
            cf.set_parameters(0.01)
            cf.setTriplets(triplets, labels, rot, cp)
            fn = lambda: cf.computeTripletCostsForLabel(labeling, 2)   # noqa: E731
            n = 8 * len(triplets)
        else:
            cf.reset_CPgrid(cp, e, 1.0)
            fn = lambda: cf.computeUnaryCosts(labels, rot)            # noqa: E731
            n = len(labels) * len(cp)
        fn()
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        print(f"{name}, D={D}: {n} costs in {1e3 * min(ts):.3f} ms = {n / min(ts):.3e} costs/s", flush=True)
        if ho:   # the strain-only batch of the same grid (univariate / multivariate kinds: likelihood 0)
            cf2 = DC.UnivariateNonLinearSRegDiscreteCostFunction(simmeasure=DC.CORRELATION)
            cf2.set_meshes(target, src, src_feat[:1], ref_feat[:1], tree)
            cf2.reset_CPgrid(cp, e, 1.0)
            cf2.set_parameters(0.01)
            cf2.setTriplets(triplets, labels, rot, cp)
            cf2.computeTripletCostsForLabel(labeling, 2)
            t0 = time.perf_counter()
            cf2.computeTripletCostsForLabel(labeling, 2)
            dt = time.perf_counter() - t0
            print(f"strain-only triplet batch: {n} costs in {1e3 * dt:.3f} ms = {n / dt:.3e} costs/s", flush=True)


if __name__ == "__main__":
    main()
