"""Host-side mirror of the groupwise (gMSM) model / cost function (msm-newmeshreg/src/DiscreteGroupModel.{h,cpp},
DiscreteGroupCostFunction.{h,cpp}) on top of the C ABI, with the multi-GPU sharding of SURVEY §8e:

* subjects are block-sharded over the ranks for the per-(subject,label) resampling (`get_patch_data`);
* ONE collective per iteration: all-gather of the resampled fields [S][L][N_t][D] (NCCL over NVLink on GPUs);
* pair-cost blocks are independent: pairs are block-sharded, every rank evaluates its block on the gathered
  fields and the [P_local][4] blocks are gathered for the host solver.

Single process (world size 1) needs no torch.distributed at all.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, f64, i32, ptr
from .resampler import Context, Mesh, Octree, estimate_rotation_matrix

RAD = 100.0


# ---------------------------------------------------------------------------------------------
# sharding helpers (pure host logic; tested with gloo, world size 2, on CPU)
# ---------------------------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of n items for `rank`: the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_counts(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


class Collective:
    """all-gather of variable-sized leading-dimension blocks. `dist` = torch.distributed (initialised) or None."""

    def __init__(self, dist=None):
        self.dist = dist if (dist is not None and dist.is_available() and dist.is_initialized()) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1

    def all_gather_blocks(self, local, n_total: int):
        """local: torch tensor [n_local, ...] holding this rank's shard_range block -> tensor [n_total, ...] on every rank."""
        if self.world == 1:
            return local
        import torch
        counts = shard_counts(n_total, self.world)
        assert local.shape[0] == counts[self.rank]
        tail = tuple(local.shape[1:])
        if len(set(counts)) == 1:
            out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
            self.dist.all_gather_into_tensor(out, local.contiguous())
            return out
        # uneven shards: collectives want equal sizes, so pad every block to the largest and trim after the gather
        m = max(counts)
        padded = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in counts]
        self.dist.all_gather(parts, padded)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


# ---------------------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------------------
class DiscreteGroupModel:
    """The parts of DiscreteGroupModel + DiscreteGroupCostFunction that the GPU path replaces."""

    def __init__(self, template: Mesh, simmeasure: int = 2, dist=None):
        self.template = template
        self.ctx: Context = template.ctx
        self.L_ = template.L
        self.template_tree = Octree(template)
        self.simmeasure = simmeasure
        self.coll = Collective(dist)
        self.g = None
        self.fields = None
        self.mask = None

    # DiscreteGroupModel::set_masks / DiscreteGroupCostFunction::set_masks (DiscreteGroupModel.h:51, DiscreteGroupCostFunction.h:50)
    def set_masks(self, mask):
        """mask: [n_tpl] values of the mask mesh's first channel; the pair costs weight a common vertex by |mask| (cpp:77)."""
        self.mask = None if mask is None else f64(mask).reshape(-1)
        if self.g is not None:
            check(self.L_.msmgpu_group_set_mask(self.g, ptr(self.mask) if self.mask is not None else None))

    # DiscreteGroupModel::get_spacings (cpp:123-143): largest geodesic distance to a mesh neighbour, per control point
    @staticmethod
    def get_spacings(cp_xyz, cp_tri):
        cp, tri = f64(cp_xyz), np.asarray(cp_tri)
        S, n = cp.shape[0], cp.shape[1]
        out = np.zeros((S, n))
        for s in range(S):
            for a, b in ((0, 1), (1, 2), (0, 2)):
                d = 2 * RAD * np.arcsin(np.sqrt(((cp[s, tri[:, a]] - cp[s, tri[:, b]]) ** 2) @ np.ones(3)) / (2 * RAD))
                np.maximum.at(out[s], tri[:, a], d)
                np.maximum.at(out[s], tri[:, b], d)
        return out

    # DiscreteGroupModel::get_rotations (cpp:77-86)
    @staticmethod
    def get_rotations(centre, cp_xyz):
        cp = f64(cp_xyz).reshape(-1, 3)
        return estimate_rotation_matrix(np.tile(f64(centre), (len(cp), 1)), cp).reshape(-1, 9)

    # DiscreteGroupModel::estimate_pairs (cpp:37-55)
    def estimate_pairs(self, cp_xyz, cp_tri):
        cp = f64(cp_xyz)
        S, n = cp.shape[0], cp.shape[1]
        trees = [Octree(Mesh(cp[s], cp_tri, ctx=self.ctx)) for s in range(S)]
        blocks = []
        for a in range(S - 1):        # order of the reference's loops: subject a, vertex v, partner subject b > a
            partners = np.stack([trees[b].get_closest_vertex_ID(cp[a]) + b * n for b in range(a + 1, S)], axis=1)     # [n][S-a-1]
            first = np.repeat(a * n + np.arange(n), S - a - 1)
            blocks.append(np.stack([first, partners.reshape(-1)], axis=1))
        if not blocks:
            return np.zeros((0, 2), np.int32)
        return np.ascontiguousarray(np.concatenate(blocks).astype(np.int32))

    # DiscreteGroupModel::get_patch_data (cpp:88-121) + DiscreteGroupCostFunction::set_patch_data
    def get_patch_data(self, data_xyz, data_tri, feat, labels, centre, rotations, spacings, range_):
        """data_xyz [S][nv][3], feat [S][D][nv], labels [L][3], rotations [S*ncp][9], spacings [S][ncp]."""
        import torch
        xyz, tri, feat = f64(data_xyz), i32(data_tri), f64(feat)
        labels, centre = f64(labels), f64(centre)
        S, nv = xyz.shape[0], xyz.shape[1]
        D, L = feat.shape[1], len(labels)
        n_tpl = self.template.nvertices()
        b, e = shard_range(S, self.coll.rank, self.coll.world)
        dev = torch.device("cuda", self.ctx.device)
        local = torch.empty((e - b, L, n_tpl, D), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        if e > b:
            check(self.L_.msmgpu_group_fields(self.ctx.h, e - b, nv, ptr(np.ascontiguousarray(xyz[b:e])), len(tri), ptr(tri), D,
                                              ptr(np.ascontiguousarray(feat[b:e])), L, ptr(labels), ptr(centre), self.template.h,
                                              self.template_tree.h, ptr(local)))
        self.ctx.sync()
        self.fields = self.coll.all_gather_blocks(local, S)       # the one collective of an iteration
        torch.cuda.synchronize(dev)
        self.S, self.L, self.D = S, L, D
        self.ncp = len(f64(spacings).reshape(S, -1)[0])
        if self.g is not None:
            self.L_.msmgpu_group_destroy(self.g)
        self.g = C.c_void_p()
        rot, sp = f64(rotations).reshape(-1, 9), f64(spacings).reshape(-1)
        check(self.L_.msmgpu_group_create(self.ctx.h, self.simmeasure, S, self.ncp, L, D, self.template.h, ptr(self.fields), ptr(rot),
                                          ptr(labels), ptr(sp), float(range_), C.byref(self.g)))
        if self.mask is not None:
            check(self.L_.msmgpu_group_set_mask(self.g, ptr(self.mask)))
        return self.fields

    # DiscreteGroupCostFunction::computePairwiseCost (cpp:54-97)
    def computePairwiseCostList(self, pairs, pair, la, lb):
        pairs = i32(pairs).reshape(-1, 2)
        p, a, b = i32(pair), i32(la), i32(lb)
        out = np.zeros(len(p))
        check(self.L_.msmgpu_group_pair_costs(self.g, len(pairs), ptr(pairs), len(p), ptr(p), ptr(a), ptr(b), ptr(out)))
        return out

    def computePairwiseCost(self, pairs, pair: int, labelA: int, labelB: int) -> float:
        return float(self.computePairwiseCostList(pairs, [pair], [labelA], [labelB])[0])

    def setPairs(self, pairs):
        """The pair list of the iteration (estimate_pairs), kept on the device for the label phases."""
        pairs = i32(pairs).reshape(-1, 2)
        check(self.L_.msmgpu_group_set_pairs(self.g, len(pairs), ptr(pairs)))
        self._pairs_key = (pairs.ctypes.data, len(pairs))
        self._pairs_ref = pairs
        self.P = len(pairs)

    def computePairwiseCostsForLabel(self, pairs, labeling, label: int, copy: bool = True, to_host: bool = True):
        """The 4 combinations per pair of Fusion::optimize (Fusion.h:164-174) -> [P, 4]; pairs block-sharded over the ranks.
        The pair list stays on the device between label phases, the block results are gathered on the device (NCCL) and come to the
        host once, into a reused pinned buffer (copy=False returns a view of it, valid until the next call). to_host=False: this rank
        takes part in the computation and the gather but does not download the table (the host solver runs on one rank) -> None."""
        import torch
        pairs = i32(pairs).reshape(-1, 2)
        if getattr(self, "_pairs_key", None) != (pairs.ctypes.data, len(pairs)) or getattr(self, "_pairs_group", None) is not self.g:
            self.setPairs(pairs)
            self._pairs_group = self.g
        lab = i32(labeling)
        P = self.P
        dev = torch.device("cuda", self.ctx.device)
        b, e = shard_range(P, self.coll.rank, self.coll.world)
        local = torch.empty((e - b, 4), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        if e > b:
            check(self.L_.msmgpu_group_pair_batch_dev(self.g, b, e - b, ptr(lab), int(label), ptr(local)))
        self.ctx.sync()
        full = local if self.coll.world == 1 else self.coll.all_gather_blocks(local, P)
        if not to_host:
            torch.cuda.synchronize(dev)
            return None
        if getattr(self, "_host_out", None) is None or self._host_out.shape[0] != P:
            self._host_out = torch.empty((P, 4), dtype=torch.float64, pin_memory=True)
        self._host_out.copy_(full)
        torch.cuda.synchronize(dev)
        return self._host_out.numpy().copy() if copy else self._host_out.numpy()

    # DiscreteGroupCostFunction::computeTripletCost (cpp:26-52): strain of a control-grid triangle of one subject, scaled by subcorr = 0.1 * S
    def computeTripletCostList(self, cps, orig_cps, rotations, labels, triplets, triplet, la, lb, lc, lambda_, shearmodulus=0.4, bulkmodulus=1.6,
                               kexponent=2.0, exponent=2.0, fixnan=False):
        """cps / orig_cps [S][ncp][3] current and undeformed control grids; triplets [T][3] GLOBAL node ids (estimate_triplets, cpp:57-74)."""
        cp, org = f64(cps).reshape(-1, 3), f64(orig_cps).reshape(-1, 3)
        S = f64(cps).shape[0]
        rot, labels, trip = f64(rotations).reshape(-1, 9), f64(labels), i32(triplets).reshape(-1, 3)
        t, a, b, c = i32(triplet), i32(la), i32(lb), i32(lc)
        reg = capi.RegParams(lambda_, shearmodulus, bulkmodulus, kexponent, exponent, 3)
        out = np.zeros(len(t))
        check(self.L_.msmgpu_group_triplet_costs(self.ctx.h, len(cp), ptr(cp), ptr(org), ptr(rot), len(labels), ptr(labels), len(trip), ptr(trip),
                                                 C.byref(reg), 0.1 * S, int(fixnan), len(t), ptr(t), ptr(a), ptr(b), ptr(c), ptr(out)))
        return out

    def reset_triplet_state(self, cps, orig_cps, rotations, labels, triplets):
        """The per-iteration arrays of the triplet term (control grids after reset_CPgrid, m_ROT, the label set, estimate_triplets) go to the
        device once; computeTripletCostsForLabel(None, ...) then only sends the labeling of each label phase."""
        cp, org = f64(cps).reshape(-1, 3), f64(orig_cps).reshape(-1, 3)
        rot, labels, trip = f64(rotations).reshape(-1, 9), f64(labels), i32(triplets).reshape(-1, 3)
        if getattr(self, "_plan", None) is not None:
            self.L_.msmgpu_triplet_plan_destroy(self._plan)
        self._plan = C.c_void_p()
        check(self.L_.msmgpu_triplet_plan_create(self.ctx.h, len(cp), ptr(cp), ptr(org), ptr(rot), len(labels), ptr(labels), len(trip), ptr(trip),
                                                 C.byref(self._plan)))
        self._plan_S, self._plan_T = f64(cps).shape[0], len(trip)

    def computeTripletCostsForLabel(self, cps, orig_cps, rotations, labels, triplets, labeling, label, lambda_, shearmodulus=0.4, bulkmodulus=1.6,
                                    kexponent=2.0, exponent=2.0, fixnan=False, copy=True, to_host=True):
        """The 8 combinations per triplet of Fusion::optimize (Fusion.h:181-196) -> [T, 8]. cps=None: the arrays set by reset_triplet_state
        (copy=False then returns a view of the reused pinned buffer, valid until the next call)."""
        reg = capi.RegParams(lambda_, shearmodulus, bulkmodulus, kexponent, exponent, 3)
        lab = i32(labeling)
        if cps is None:
            # resident plan: the block's costs stay on the device, the blocks are gathered device to device (NCCL) and come to the host
            # once, into a reused pinned buffer -- the same data path as the pair batches above
            import torch
            S, T = self._plan_S, self._plan_T
            dev = torch.device("cuda", self.ctx.device)
            b, e = shard_range(T, self.coll.rank, self.coll.world)       # triplets are per subject: block-sharded like the pairs
            local = torch.empty((e - b, 8), dtype=torch.float64, device=dev)
            torch.cuda.synchronize(dev)
            if e > b:
                check(self.L_.msmgpu_triplet_plan_batch_dev(self._plan, C.byref(reg), 0.1 * S, int(fixnan), b, e - b, ptr(lab), int(label), ptr(local)))
            full = local if self.coll.world == 1 else self.coll.all_gather_blocks(local, T)
            if not to_host:      # the solver's rank downloads, the others only compute and gather
                torch.cuda.synchronize(dev)
                return None
            if getattr(self, "_host_trip", None) is None or self._host_trip.shape[0] != T:
                self._host_trip = torch.empty((T, 8), dtype=torch.float64, pin_memory=True)
            self._host_trip.copy_(full)
            torch.cuda.synchronize(dev)
            return self._host_trip.numpy().copy() if copy else self._host_trip.numpy()
        else:
            cp, org = f64(cps).reshape(-1, 3), f64(orig_cps).reshape(-1, 3)
            S = f64(cps).shape[0]
            rot, labels, trip = f64(rotations).reshape(-1, 9), f64(labels), i32(triplets).reshape(-1, 3)
            T = len(trip)
            b, e = shard_range(T, self.coll.rank, self.coll.world)
            out = np.zeros((e - b, 8))
            if e > b:
                blk = np.ascontiguousarray(trip[b:e])
                check(self.L_.msmgpu_group_triplet_batch(self.ctx.h, len(cp), ptr(cp), ptr(org), ptr(rot), len(labels), ptr(labels), e - b, ptr(blk),
                                                         C.byref(reg), 0.1 * S, int(fixnan), ptr(lab), int(label), ptr(out)))
        if self.coll.world == 1:
            return out
        import torch
        dev = torch.device("cuda", self.ctx.device)
        return self.coll.all_gather_blocks(torch.from_numpy(out).to(dev), T).cpu().numpy()

    def close(self):
        if self.g is not None:
            self.L_.msmgpu_group_destroy(self.g)
            self.g = None
        if getattr(self, "_plan", None) is not None:
            self.L_.msmgpu_triplet_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
