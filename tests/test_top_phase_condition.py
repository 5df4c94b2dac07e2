"""CPU check of the claim the octree top phase rests on (newmsm_b200/csrc/octree_build.cu, note above k_top_count), against the
oracle's tree AND, when it is built, the compiled reference's own tree (oracle/_ref):

  (1) a node's final content is the ascending list of the triangles whose AABB touches its closed cube;
  (2) per triangle the number of child cubes it touches is >= its split_size (octree.cpp:69-102), so
      sum_c count(child c) >= total_size for every node;
  (3) hence  count >= 50  and  sum_c count(child c) < 3 count  is SUFFICIENT for the reference to have split the node — every
      node that satisfies it must be internal in the reference's tree — and on sphere meshes it decides the shallow levels.

Everything is recomputed here in numpy from the geometry alone (closed-interval tests of node.cpp:79-120 in FP64); the trees are
only read back through their pre-order dumps. Test infrastructure: imports oracle/."""
import numpy as np
import pytest

from newmsm_b200 import synth

BOUNDS = 101.0          # octree.h:37
MAX_TRIANGLES = 50      # node.h:33


def tri_boxes(xyz, tri):
    c = xyz[tri]                       # [T, 3 corners, 3]
    return c.min(axis=1), c.max(axis=1)


def walk(kinds, counts, leaf_tris, tlo, thi):
    """Pre-order walk of a dump (node, then its 8 children, child c = 4 x + 2 y + z): yields per node
    (depth, is_leaf, touching triangle ids (ascending), children's touching counts, total_size, stored leaf list or None)."""
    pos = {"node": 0, "tri": 0}
    out = []

    def rec(lo, half, cand, depth):
        i = pos["node"]
        pos["node"] += 1
        leaf = kinds[i] == 1
        stored = None
        if leaf:
            stored = leaf_tris[pos["tri"]:pos["tri"] + counts[i]]
            pos["tri"] += counts[i]
        b0, b1, b2 = lo, lo + half, lo + (half + half)
        lo_c, hi_c = tlo[cand], thi[cand]
        # node.cpp:79-89 containing_oct on both corners (strict '<' against the midpoint)
        same = (lo_c < b1) == (hi_c < b1)
        split_size = 8 >> same.sum(axis=1)
        # node.cpp:112-120 can_contain of the lower / upper half per axis (closed intervals)
        in0 = ~((hi_c < b0) | (lo_c > b1))
        in1 = ~((hi_c < b1) | (lo_c > b2))
        child_sets = []
        for c in range(8):
            m = np.ones(len(cand), bool)
            for a, bit in enumerate((4, 2, 1)):
                m &= in1[:, a] if c & bit else in0[:, a]
            child_sets.append(cand[m])
        child_cnt = np.array([len(s) for s in child_sets])
        touches = (in0[:, 0].astype(int) + in1[:, 0]) * (in0[:, 1].astype(int) + in1[:, 1]) * (in0[:, 2].astype(int) + in1[:, 2])   # children touched
        out.append({"depth": depth, "leaf": leaf, "list": cand, "child_cnt": child_cnt, "total_size": int(split_size.sum()),
                    "stored": stored, "touches": touches, "split_size": split_size})
        if not leaf:
            for c in range(8):
                clo = lo + np.array([half if c & 4 else 0.0, half if c & 2 else 0.0, half if c & 1 else 0.0])
                rec(clo, half / 2, child_sets[c], depth + 1)

    rec(np.array([-BOUNDS] * 3), BOUNDS, np.arange(len(tlo)), 0)
    assert pos["node"] == len(kinds) and pos["tri"] == len(leaf_tris)
    return out


def check_tree(dump, xyz, tri):
    kinds, counts, leaf_tris = dump
    tlo, thi = tri_boxes(xyz, tri)
    nodes = walk(kinds, counts, leaf_tris, tlo, thi)
    decided = undecided = 0
    for nd in nodes:
        if nd["leaf"]:                                   # (1): the stored list is the ascending list of touching triangles
            assert np.array_equal(nd["stored"], nd["list"])
        assert np.all(nd["touches"] >= nd["split_size"])     # (2), per triangle ...
        assert nd["child_cnt"].sum() == nd["touches"].sum()  # ... and the children's counts are exactly their sum
        n = len(nd["list"])
        if n >= MAX_TRIANGLES:
            if nd["child_cnt"].sum() < 3 * n:            # (3): sufficient => the reference split this node
                assert not nd["leaf"], f"depth {nd['depth']}: {n} triangles, children {nd['child_cnt'].sum()} < {3 * n}, but a leaf"
                decided += 1
            elif not nd["leaf"]:
                undecided += 1                           # split by the reference on a PREFIX of its list: left to the exact level passes
    return decided, undecided, nodes


@pytest.mark.parametrize("case", ["ico3", "ico4", "ico5", "jittered ico5", "squeezed ico4"])
def test_sufficient_split_condition_matches_the_oracle_tree(oracle_built, case):
    lvl = int(case[-1])
    xyz, tri = synth.icosphere(lvl)
    if case.startswith("jittered"):
        xyz = synth.jitter_sphere(xyz, tri, frac=0.3, seed=5)
    if case.startswith("squeezed"):                       # triangles crowded towards the poles: long lists in a few cells
        w = xyz * np.array([0.35, 0.35, 1.0])
        xyz = w / np.linalg.norm(w, axis=1, keepdims=True) * 100.0
    decided, undecided, nodes = check_tree(oracle_built.OracleOctree(xyz, tri).dump(), xyz, tri)
    internal = sum(1 for nd in nodes if not nd["leaf"])
    assert decided + undecided == internal
    if lvl >= 4 and not case.startswith("squeezed"):
        # on a sphere mesh the sufficient condition decides every split of the levels the top phase covers
        d0 = 3 if lvl == 4 else 4
        assert all(not (not nd["leaf"] and nd["depth"] < d0 and nd["child_cnt"].sum() >= 3 * len(nd["list"])) for nd in nodes)


def test_sufficient_split_condition_matches_the_reference_tree(ref_built):
    xyz, tri = synth.icosphere(4)
    xyz = synth.jitter_sphere(xyz, tri, frac=0.3, seed=9)
    m = ref_built.RefMesh(xyz, tri)
    decided, undecided, nodes = check_tree(ref_built.RefOctree(m).dump(), xyz, tri)
    assert decided > 0
