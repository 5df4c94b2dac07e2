"""CPU baseline sample for the gMSM rows (TEST / MEASUREMENT INFRASTRUCTURE: runs the compiled reference under oracle/_ref).
Times the reference's own DiscreteGroupModel::get_patch_data (DiscreteGroupModel.cpp:88-121) and
DiscreteGroupCostFunction::computePairwiseCost (DiscreteGroupCostFunction.cpp:54-97) on a bounded sample of BASELINE configs[4]
(2 of the 64 subjects, ico6 data / template, ico4 control grid, 19 labels), all host threads.

    python tests/group_reference_sample.py [subjects] > gpurun_out/group_cpu_sample.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from newmsm_b200 import synth  # noqa: E402
from oracle import bindings as B  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    threads = os.cpu_count() or 1
    cp0, cp_tri = synth.icosphere(4)
    d0, dtri = synth.icosphere(6)
    tpl, tpl_tri = synth.icosphere(6)
    data = np.stack([synth.smooth_warp(d0, max_disp=2.0, seed=40 + s) for s in range(S)])
    cps = np.stack([synth.smooth_warp(cp0, max_disp=1.0, seed=60 + s) for s in range(S)])
    feat = np.stack([synth.smooth_fields(data[s], 1, seed0=100, noise=0.1, noise_seed=7 + s) for s in range(S)])
    centre = np.array([0.0, 0.0, 100.0])
    spacing = 2 * 100 * np.arcsin(np.linalg.norm(cp0[cp_tri[:, 0]] - cp0[cp_tri[:, 1]], axis=1).max() / 200)
    labels = [centre]
    for ring, n in ((0.25, 6), (0.5, 12)):
        for k in range(n):
            p = centre + ring * spacing * np.array([np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n), 0.0])
            labels.append(p / np.linalg.norm(p) * 100)
    labels = np.array(labels)
    L, ncp = len(labels), len(cp0)
    rot = np.array([B.oracle_rotation_matrix(centre, c) for c in cps.reshape(-1, 3)]).reshape(-1, 9)
    spac = np.full((S, ncp), spacing)
    pairs = np.array([[v, ncp + v] for v in range(ncp)], np.int32)
    rng = np.random.default_rng(1)

    def run(n):
        rp, la, lb = rng.integers(0, len(pairs), n), rng.integers(0, L, n), rng.integers(0, L, n)
        t0 = time.perf_counter()
        B.refmr_group_pair_costs(2, data, dtri, feat, labels, centre, tpl, tpl_tri, ncp, rot, spac, 1.0, pairs, rp, la, lb, nthreads=threads)
        return time.perf_counter() - t0

    t_small = run(16)
    n_big = 200000
    t_big = run(n_big)
    # get_patch_data parallelises over SUBJECTS only (DiscreteGroupModel.cpp:92): the sample keeps min(S, threads) threads busy
    busy = min(S, threads)
    per_resample = t_small * busy / (S * L)                       # thread-seconds per (subject, label)
    pair_rate = n_big / max(t_big - t_small, 1e-9)
    print(json.dumps({"what": "reference gMSM, CPU", "threads": threads, "subjects_in_sample": S, "labels": L, "get_patch_data_wall_s": t_small,
                      "get_patch_data_thread_s_per_subject_label": per_resample,
                      "get_patch_data_s_for_64_subjects_on_all_threads": per_resample * 64 * L / min(64, threads),
                      "pair_costs_per_s": pair_rate, "sample": f"{S} subjects x {L} labels, ico6 -> ico6 template, ico4 control grid; {n_big} pair costs"}))


if __name__ == "__main__":
    main()
