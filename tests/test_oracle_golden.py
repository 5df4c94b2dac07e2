"""CPU oracle against the committed golden fixtures (outputs of the compiled reference,
tests/golden/make_golden.py). Runs without /root/reference. Bit-exact."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_octree_golden(oracle_built):
    g = load("octree_ico3.npz")
    t = oracle_built.OracleOctree(g["xyz"], g["tri"])
    kinds, counts, leaf_tris = t.dump()
    assert np.array_equal(kinds, g["kinds"]) and np.array_equal(counts, g["counts"])
    assert np.array_equal(leaf_tris, g["leaf_tris"])
    tri, vtx, st, _ = t.query(g["q"])
    assert np.array_equal(tri, g["tri_id"]) and np.array_equal(st, g["status"])
    ok = st == 0
    assert np.array_equal(vtx[ok], g["vertex_id"][ok])
    idx, w, n, err = t.bary_weights(g["q"][ok])
    assert err == 0
    assert np.array_equal(idx, g["w_idx"]) and np.array_equal(w, g["w_val"]) and np.array_equal(n, g["w_n"])
    assert np.array_equal(oracle_built.oracle_vertex_areas(g["xyz"], g["tri"]), g["vertex_area"])


@pytest.mark.parametrize("name", ["down", "up"])
def test_resample_golden(oracle_built, name):
    g = load(f"resample_{name}.npz")
    rowptr, col, val = oracle_built.oracle_adaptive_weights(g["xyz_in"], g["tri_in"], g["xyz_low"], g["tri_low"])
    assert np.array_equal(rowptr, g["rowptr"]) and np.array_equal(col, g["col"]) and np.array_equal(val, g["val"])
    out = oracle_built.oracle_metric_resample(g["xyz_in"], g["tri_in"], g["xyz_low"], g["tri_low"], g["feat"])
    assert np.array_equal(out, g["metric_out"])
    out = oracle_built.oracle_bary_resample(g["xyz_in"], g["tri_in"], g["xyz_low"], g["feat"])
    assert np.array_equal(out, g["bary_out"])
    # rows of the adaptive matrix are normalised (resampler.cpp:130-135)
    sums = np.add.reduceat(val, rowptr[:-1])
    assert np.allclose(sums, 1.0, atol=1e-12)


def test_blend_golden(oracle_built):
    g = load("blend.npz")
    O = oracle_built
    assert np.array_equal(O.oracle_sphere_project_warp(g["sph"], g["xf"], g["tf"], g["xto"]), g["warp"])
    assert np.array_equal(O.oracle_surface_resample(g["sph"], g["xf"], g["tf"], g["anat"]), g["surf"])
    assert np.array_equal(O.oracle_nn_resample(g["sph"], g["xf"], g["tf"], g["feat"]), g["nn"])


def test_rotation_golden(oracle_built):
    g = load("rotation.npz")
    for ci, ix, R in zip(g["ci"], g["index"], g["R"]):
        assert np.array_equal(oracle_built.oracle_rotation_matrix(ci, ix), R)
    assert np.array_equal(g["R"][0], np.eye(3)) and np.array_equal(g["R"][2], np.eye(3))


def test_similarity_kats(oracle_built):
    # analytic known answers: corr(A,A)=1 -> cost 0; corr(A,-A)=-1 -> cost 1;
    # zero variance -> r=0 -> cost 0.5 (similarities.cpp:157); SSD = sqrt(sum w d^2)/n (cpp:186)
    rng = np.random.default_rng(0)
    A = rng.normal(size=65); w = rng.uniform(0.5, 1.5, size=65)
    assert abs(oracle_built.oracle_sim(2, A, A, w)) < 1e-15
    assert abs(oracle_built.oracle_sim(2, A, -A, w) - 1.0) < 1e-15
    assert oracle_built.oracle_sim(2, A, np.ones(65), w) == 0.5
    B = rng.normal(size=65)
    assert np.isclose(oracle_built.oracle_sim(1, A, B, w), np.sqrt((w * (A - B) ** 2).sum()) / 65, rtol=1e-14)
    r = np.cov(A, B, aweights=w, bias=True)
    r = r[0, 1] / np.sqrt(r[0, 0] * r[1, 1])
    assert np.isclose(oracle_built.oracle_sim(2, A, B, w), 1 - (1 + r) / 2, rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# cost functions: outputs of the reference's own classes (tests/golden/make_golden_costs.py)
# ---------------------------------------------------------------------------------------------
def test_unary_costs_golden(oracle_built):
    from cost_cases import GOLDEN_CP, GOLDEN_DATA, cost_setup, golden_digest
    g = load("costs.npz")
    for kind, D in ((0, 1), (1, 4), (2, 4)):
        s = cost_setup(oracle_built, GOLDEN_CP, GOLDEN_DATA, D)
        assert np.array_equal(golden_digest(s), g[f"unary_k{kind}_digest"]), "seeded inputs drifted: regenerate the fixture"
        prow, pmem = oracle_built.oracle_patch_membership(s["cp"], s["src"], s["maxsep"], 1.0)
        assert np.array_equal(prow, g[f"unary_k{kind}_prow"]) and np.array_equal(pmem, g[f"unary_k{kind}_pmem"])
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D if kind == 1 else 1, len(s["src"])))
        ot = oracle_built.OracleOctree(s["xyz"], s["tri"])
        for sim in (1, 2):
            got = oracle_built.oracle_unary_costs(kind, sim, ot, s["cp"], s["rot"], s["labels"], s["src"], prow, pmem,
                                                  s["src_feat"], s["ref_feat"], cfw, s["absw"])
            assert np.array_equal(got, g[f"unary_k{kind}_s{sim}"])
        for sim, pct in ((4, 0.75), (5, 0.6)):
            oracle_built.oracle_set_percentile(pct)
            got = oracle_built.oracle_unary_costs(kind, sim, ot, s["cp"], s["rot"], s["labels"], s["src"], prow, pmem,
                                                  s["src_feat"], s["ref_feat"], cfw, s["absw"])
            oracle_built.oracle_set_percentile(0.75)
            assert np.array_equal(got, g[f"unary_k{kind}_s{sim}"])


def test_triplet_costs_golden(oracle_built):
    from cost_cases import GOLDEN_CP, GOLDEN_DATA, golden_digest, triplet_setup
    g = load("costs.npz")
    for kind, D in ((0, 1), (3, 1), (4, 3)):
        s = triplet_setup(oracle_built, GOLDEN_CP, GOLDEN_DATA, D)
        assert np.array_equal(golden_digest(s), g[f"triplet_k{kind}_digest"]), "seeded inputs drifted: regenerate the fixture"
        rt, la, lb, lc = s["req"]
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
        prow = pmem = ot = None
        if kind >= 3:
            prow, pmem = oracle_built.oracle_ho_patches(s["cp_now"], s["cp_tri"], s["src"])
            assert np.array_equal(prow, g[f"triplet_k{kind}_prow"]) and np.array_equal(pmem, g[f"triplet_k{kind}_pmem"])
            ot = oracle_built.OracleOctree(s["xyz"], s["tri"])
        got = oracle_built.oracle_triplet_costs(kind, 2, ot, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                                s["src"], prow, pmem, s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05)
        assert np.array_equal(got, g[f"triplet_k{kind}"])


def test_group_costs_golden(oracle_built):
    from cost_cases import golden_digest, golden_group_glue, group_setup, group_triplet_case
    g = load("costs.npz")
    c = group_setup(S=2, cp_level=1, data_level=3, tpl_level=3, D=2)
    orig, trip, rot_t, (rt, ta, tb, tc) = group_triplet_case(oracle_built, c)
    assert np.array_equal(oracle_built.oracle_group_triplet_costs(c["cps"], orig, rot_t, c["labels"], trip, rt, ta, tb, tc, 0.05), g["group_triplet"])
    assert np.array_equal(golden_digest(c), g["group_digest"]), "seeded inputs drifted: regenerate the fixture"
    rot, spacings, pairs, (rp, la, lb) = golden_group_glue(oracle_built, c)
    fields = oracle_built.oracle_group_fields(c["data"], c["dtri"], c["feat"], c["labels"], c["centre"], c["tpl"], c["tpl_tri"])
    seen = ~np.isnan(g["group_fields"])
    assert np.array_equal(fields[seen], g["group_fields"][seen])
    for sim in (1, 2):
        got = oracle_built.oracle_group_pair_costs(sim, c["cps"].shape[1], c["tpl"], fields, rot, c["labels"], spacings, 1.0, pairs, rp, la, lb)
        ok = ~np.isnan(got)          # empty intersections: undefined behaviour in the reference, not compared
        assert ok.mean() > 0.9 and np.array_equal(got[ok], g[f"group_pair_s{sim}"][ok])
        from cost_cases import group_mask
        got_m = oracle_built.oracle_group_pair_costs(sim, c["cps"].shape[1], c["tpl"], fields, rot, c["labels"], spacings, 1.0, pairs, rp, la, lb,
                                                     mask=group_mask(c))
        assert np.array_equal(got_m[ok], g[f"group_pair_masked_s{sim}"][ok])


def test_rigid_level_golden(oracle_built):
    """RIGID / AFFINE level: the CPU restatement (oracle/msm_oracle.cpp: orc_rigid) reproduces the compiled reference's
    neighbour lists, cost at zero rotation and final rotated source bit for bit (tests/golden/rigid.npz)."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_rigid", os.path.join(here, "golden", "make_golden_rigid.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(here, "golden", "rigid.npz"))
    for name, D, sim in (("d2_corr", 2, 2), ("d1_corr", 1, 2), ("d3_ssd", 3, 1)):
        xyz, tri, src, mov, ref = mod.rigid_case(3, D)
        moved, cost0, rowptr, members = oracle_built.oracle_rigid(xyz, tri, src, tri, mov, ref, simmeasure=sim, iters=4)
        assert np.array_equal(rowptr, g[f"{name}_rowptr"]) and np.array_equal(members, g[f"{name}_members"])
        assert cost0 == float(g[f"{name}_cost0"])
        assert np.array_equal(moved, g[f"{name}_xyz"])


def test_exclusion_masks_golden(oracle_built):
    """Exclusion masks (resampler.cpp:30-140, 169-258 with EXCL): the CPU restatement reproduces the compiled reference's outputs
    (tests/golden/excl.npz, generator make_golden_excl.py) bit for bit: masked adaptive weights, metric_resample + resampled mask,
    nearest-neighbour interpolation, Gaussian smoothing with and without a mask."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_excl", os.path.join(here, "golden", "make_golden_excl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(here, "golden", "excl.npz"))
    for name, (sl, ll) in {"down": (4, 3), "up": (3, 4), "same": (3, 3)}.items():
        xyz, tri, low, low_tri, feat, excl = mod.excl_case(sl, ll)
        out, eo, (rp, col, val) = oracle_built.oracle_metric_resample_excl(xyz, tri, low, low_tri, feat, excl)
        assert np.array_equal(rp, g[f"{name}_rowptr"]) and np.array_equal(col, g[f"{name}_col"]) and np.array_equal(val, g[f"{name}_val"])
        assert np.array_equal(out, g[f"{name}_metric"]) and np.array_equal(eo, g[f"{name}_metric_excl"])
        n, en = oracle_built.oracle_nn_resample_excl(low, xyz, tri, feat, excl)
        assert np.array_equal(n, g[f"{name}_nn"]) and np.array_equal(en, g[f"{name}_nn_excl"])
    xyz, tri, _, _, feat, excl = mod.excl_case(4, 3)
    for sigma in (4.0, 9.0):
        s0, _ = oracle_built.oracle_smooth_data(xyz, tri, xyz, sigma, feat)
        s1, e1 = oracle_built.oracle_smooth_data(xyz, tri, xyz, sigma, feat, excl)
        assert np.array_equal(s0, g[f"smooth{int(sigma)}"])
        assert np.array_equal(s1, g[f"smooth{int(sigma)}_masked"]) and np.array_equal(e1, g[f"smooth{int(sigma)}_excl"])


def test_anatomical_strain_golden(oracle_built):
    """regoption 5 triplet costs of the reference's own class (tests/golden/make_golden_anat.py) vs the restatement."""
    from cost_cases import GOLDEN_CP, GOLDEN_DATA, anat_case, golden_digest, triplet_setup
    O = oracle_built
    g = load("anat.npz")
    for kind, D, depth in ((0, 1, 1), (0, 1, 2), (3, 1, 2)):
        s = triplet_setup(O, GOLDEN_CP, GOLDEN_DATA, D)
        a = anat_case(O, s, depth)
        assert np.array_equal(np.concatenate([golden_digest(s), golden_digest(a)]), g[f"k{kind}_d{depth}_digest"]), "seeded inputs drifted: regenerate"
        rt, la, lb, lc = s["req"]
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
        prow = pmem = ot = None
        if kind >= 3:
            prow, pmem = O.oracle_ho_patches(s["cp_now"], s["cp_tri"], s["src"])
            ot = O.OracleOctree(s["xyz"], s["tri"])
        got = O.oracle_triplet_costs(kind, 2, ot, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                     s["src"], prow, pmem, s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05, rmode=5, anat=a)
        assert np.array_equal(got, g[f"k{kind}_d{depth}"])
