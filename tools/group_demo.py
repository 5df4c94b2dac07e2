"""gMSM on N GPUs: subjects sharded for the per-(subject,label) resampling, ONE all-gather of the fields per
iteration (NCCL), pair-cost blocks sharded, results gathered for the host solver. Rank 0 also runs the unsharded
computation and checks that the sharded result is bit-identical (reduction order never depends on the shard layout).

  python tools/group_demo.py                                            # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/group_demo.py
"""
import os as _os
if "WORLD_SIZE" in _os.environ:   # torchrun pins OMP_NUM_THREADS to 1; the library's host-side finishes (libm pow / acos) are OpenMP loops
    _os.environ["OMP_NUM_THREADS"] = str(max(1, (_os.cpu_count() or 1) // int(_os.environ["WORLD_SIZE"])))
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newmsm_b200 import group_cost as GC, resampler as R, synth  # noqa: E402


def main():
    S = int(os.environ.get("GROUP_SUBJECTS", 16))
    data_level = int(os.environ.get("GROUP_DATA_LEVEL", 5))
    cp_level = int(os.environ.get("GROUP_CP_LEVEL", 3))
    D = int(os.environ.get("GROUP_CHANNELS", 4))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = R.Context(local)
    cp0, cp_tri = synth.icosphere(cp_level)
    d0, dtri = synth.icosphere(data_level)
    tpl, tpl_tri = synth.icosphere(data_level)
    data = np.stack([synth.smooth_warp(d0, max_disp=2.0, seed=40 + s) for s in range(S)])
    cps = np.stack([synth.smooth_warp(cp0, max_disp=1.0, seed=60 + s) for s in range(S)])
    feat = np.stack([synth.smooth_fields(data[s], D, seed0=100, noise=0.1, noise_seed=7 + s) for s in range(S)])
    centre = np.array([0.0, 0.0, 100.0])
    spacing = 2 * 100 * np.arcsin(np.linalg.norm(cp0[cp_tri[:, 0]] - cp0[cp_tri[:, 1]], axis=1).max() / 200)
    labels = [centre]
    for ring, n in ((0.25, 6), (0.5, 12)):
        for k in range(n):
            p = centre + ring * spacing * np.array([np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n), 0.0])
            labels.append(p / np.linalg.norm(p) * 100)
    labels = np.array(labels)
    L = len(labels)

    def run(d):
        M = GC.DiscreteGroupModel(R.Mesh(tpl, tpl_tri, ctx=ctx), simmeasure=2, dist=d)
        spac = M.get_spacings(cps, cp_tri)
        rot = M.get_rotations(centre, cps)
        pairs = M.estimate_pairs(cps, cp_tri)
        torch.cuda.synchronize()
        if d is not None:
            d.barrier()
        t0 = time.perf_counter()
        M.get_patch_data(data, dtri, feat, labels, centre, rot, spac, 1.0)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        labeling = np.zeros(S * len(cp0), np.int32)
        costs = []
        for l in range(1, L):      # the solver consumes a batch before it asks for the next: a view of the reused host buffer, reduced at once
            c = M.computePairwiseCostsForLabel(pairs, labeling, l, copy=False)
            costs.append(c[::997].ravel().copy())      # a strided sample is kept; two whole batches are compared below, untimed
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        costs.append(M.computePairwiseCostsForLabel(pairs, labeling, 1).ravel())        # untimed: two whole batches for the bitwise comparison
        costs.append(M.computePairwiseCostsForLabel(pairs, labeling, L - 1).ravel())
        # strain triplets of every subject's control grid (DiscreteGroupCostFunction.cpp:26-52), Fusion's 8 combinations per label
        ncp = len(cp0)
        trip = np.concatenate([np.sort(cp_tri + s_ * ncp, axis=1) for s_ in range(S)]).astype(np.int32)
        orig = np.stack([cp0] * S)
        tcosts = [M.computeTripletCostsForLabel(cps, orig, rot, labels, trip, labeling, l, 0.2) for l in range(1, L)]
        t3 = time.perf_counter()
        run.triplets = (len(trip), t3 - t2, float(np.stack(tcosts).sum()))
        return np.concatenate(costs), len(pairs), t1 - t0, t2 - t1

    costs, P, t_fields, t_pairs = run(dist)
    if rank == 0:
        line = {"demo": "gMSM fields + pair costs", "n_gpus": world, "subjects": S, "labels": L, "channels": D, "data_grid": f"ico{data_level}",
                "cp_grid": f"ico{cp_level}", "pairs": P, "resamples_per_iteration": S * L, "fields_s": t_fields,
                "pair_costs_per_s": P * 4 * (L - 1) / t_pairs, "pair_batches_s": t_pairs,
                "triplets": run.triplets[0], "triplet_costs_per_s": run.triplets[0] * 8 * (L - 1) / run.triplets[1], "triplet_batches_s": run.triplets[1]}
        if world > 1:
            ref, _, tf1, tp1 = run(None)     # unsharded, on this rank alone
            same = np.array_equal(np.nan_to_num(costs, nan=-1.0), np.nan_to_num(ref, nan=-1.0))
            line.update({"sharded_equals_unsharded_bitwise": bool(same), "unsharded_fields_s": tf1, "unsharded_pair_batches_s": tp1})
            assert same, "sharded pair costs differ from the unsharded ones"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
