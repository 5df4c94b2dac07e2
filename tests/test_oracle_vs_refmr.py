"""Pins the cost-function part of the CPU oracle (oracle/msm_oracle.cpp) against the UNMODIFIED reference
registration library compiled into oracle/_ref/libref_newmeshreg.so (the reference's own
DiscreteCostFunction / DiscreteGroupModel classes driven by oracle/ref_meshreg_driver.cpp).
The reference ships no tests, so this is what makes the cost rows (a9-a16) "pinned".
Bit-exact unless a tolerance is written next to the assert. Skipped when oracle/_ref is absent."""
import numpy as np
import pytest

from cost_cases import cost_setup, group_mask, group_setup, group_triplet_case, triplet_setup


@pytest.fixture(scope="module")
def O(ref_built):
    if not ref_built.have_refmr():
        pytest.skip("oracle/_ref/libref_newmeshreg.so not built (needs /root/reference)")
    return ref_built


def rel_close(a, b, tol):
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("range_", [1.0, 0.5, 1.7])
def test_patch_membership(O, range_):
    s = cost_setup(O, 3, 5, 1)
    cp_tri = __import__("newmsm_b200.synth", fromlist=["x"]).icosphere(3)[1]
    _, r1, m1, _ = O.refmr_unary(0, 2, s["xyz"], s["tri"], s["cp"], cp_tri, s["rot"], s["labels"], s["src"], s["tri"],
                                 s["src_feat"], s["ref_feat"], None, None, s["maxsep"], range_, want_costs=False)
    r0, m0 = O.oracle_patch_membership(s["cp"], s["src"], s["maxsep"], range_)
    assert np.array_equal(r0, r1) and np.array_equal(m0, m1)


@pytest.mark.parametrize("kind,D", [(0, 1), (1, 6), (2, 6)])
@pytest.mark.parametrize("sim", [1, 2])
def test_unary_costs(O, kind, D, sim):
    from newmsm_b200 import synth
    s = cost_setup(O, 3, 5, D)
    cp_tri = synth.icosphere(3)[1]
    rng = np.random.default_rng(5)
    cfw = rng.uniform(0.2, 1.0, size=(D if kind == 1 else 1, len(s["src"])))
    ot = O.OracleOctree(s["xyz"], s["tri"])
    for weights in (None, cfw):
        # nthreads = 1: Multivariate::get_target_data writes per-SOURCE-vertex buffers shared by overlapping patches
        # (DiscreteCostFunction.cpp:418-440), a data race under the reference's own omp loop (cpp:240)
        ref, prow, pmem, absw_ref = O.refmr_unary(kind, sim, s["xyz"], s["tri"], s["cp"], cp_tri, s["rot"], s["labels"], s["src"], s["tri"],
                                                  s["src_feat"], s["ref_feat"], weights, s["absw"], s["maxsep"], 1.0, nthreads=1)
        got = O.oracle_unary_costs(kind, sim, ot, s["cp"], s["rot"], s["labels"], s["src"], prow, pmem,
                                   s["src_feat"], s["ref_feat"], weights, s["absw"])
        assert np.array_equal(got, ref)
        assert np.ptp(ref) > 0


def test_resample_weights_is_metric_resample_of_row_max(O):
    """resample_weights (DiscreteCostFunction.cpp:303-322): AbsoluteWeights = adaptive resample of max_d cfweight onto the CP grid."""
    from newmsm_b200 import synth
    s = cost_setup(O, 3, 5, 3)
    cp_tri = synth.icosphere(3)[1]
    cfw = np.random.default_rng(2).uniform(0.2, 1.0, size=(3, len(s["src"])))
    _, _, _, absw = O.refmr_unary(1, 2, s["xyz"], s["tri"], s["cp"], cp_tri, s["rot"], s["labels"], s["src"], s["tri"],
                                  s["src_feat"], s["ref_feat"], cfw, None, s["maxsep"], 1.0, want_costs=False)
    mine = O.oracle_metric_resample(s["src"], s["tri"], s["cp"], cp_tri, cfw.max(axis=0)[None, :])
    assert np.array_equal(mine.reshape(-1), absw)


@pytest.mark.parametrize("kexp,rexp", [(2.0, 2.0), (2.0, 1.0), (1.5, 1.3)])
def test_triplet_strain(O, kexp, rexp):
    s = triplet_setup(O, 3, 5, 1)
    rt, la, lb, lc = s["req"]
    ref, _, _ = O.refmr_triplet(0, 2, s["xyz"], s["tri"], s["cp_now"], s["cp_tri"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                s["src"], s["tri"], s["src_feat"], s["ref_feat"], None, None, 0.1, 0.4, 1.6, kexp, rexp, 3)
    got = O.oracle_triplet_costs(0, 2, None, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                 s["src"], None, None, s["src_feat"], s["ref_feat"], None, np.ones(len(s["cp"])), 0.1, 0.4, 1.6, kexp, rexp)
    assert np.all(np.isfinite(ref))
    # the 2x2 inverse / determinant run in the FSL stand-in (oracle/shim/armawrap/newmat.h), not in FSL's NEWMAT/armadillo,
    # which the reference does not vendor or pin: operation order there is ours. Everything around it is the reference's.
    assert np.array_equal(got, ref)
    assert (ref == 1e7 * 0.1).sum() >= 0


@pytest.mark.parametrize("kind,D", [(3, 1), (4, 5)])
@pytest.mark.parametrize("sim", [1, 2])
def test_ho_triplet_likelihood(O, kind, D, sim):
    s = triplet_setup(O, 3, 5, D)
    rt, la, lb, lc = s["req"]
    cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
    # nthreads = 1: HO get_target_data writes per-TRIPLET buffers (cpp:487-513), racy when a triplet is requested twice concurrently
    ref, prow, pmem = O.refmr_triplet(kind, sim, s["xyz"], s["tri"], s["cp_now"], s["cp_tri"], s["orig"], s["rot_now"], s["labels"], s["triplets"],
                                      rt, la, lb, lc, s["src"], s["tri"], s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05, nthreads=1)
    r0, m0 = O.oracle_ho_patches(s["cp_now"], s["cp_tri"], s["src"])
    assert np.array_equal(prow, r0) and np.array_equal(pmem, m0)
    ot = O.OracleOctree(s["xyz"], s["tri"])
    got = O.oracle_triplet_costs(kind, sim, ot, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                 s["src"], prow, pmem, s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind,D,depth,kexp,rexp", [(0, 1, 1, 2.0, 2.0), (0, 1, 2, 1.5, 1.3), (3, 1, 2, 2.0, 2.0)])
def test_triplet_anatomical_strain(O, kind, D, depth, kexp, rexp):
    """regoption 5 (DiscreteCostFunction.cpp:169-181, 245-301): the restatement vs the reference's own computeTripletCost with the anatomical
    meshes and maps handed over the way Mesh_registration does (mesh_registration.cpp:93-98)."""
    from cost_cases import anat_case
    s = triplet_setup(O, 3, 5, D)
    a = anat_case(O, s, depth)
    rt, la, lb, lc = s["req"]
    cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
    ref, prow, pmem = O.refmr_triplet(kind, 2, s["xyz"], s["tri"], s["cp_now"], s["cp_tri"], s["orig"], s["rot_now"], s["labels"], s["triplets"],
                                      rt, la, lb, lc, s["src"], s["tri"], s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.1, 0.4, 1.6, kexp, rexp, 5,
                                      nthreads=1, anat=a)
    ot = O.OracleOctree(s["xyz"], s["tri"]) if kind >= 3 else None
    got = O.oracle_triplet_costs(kind, 2, ot, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                 s["src"], prow if kind >= 3 else None, pmem if kind >= 3 else None, s["src_feat"], s["ref_feat"], cfw, s["absw"],
                                 0.1, 0.4, 1.6, kexp, rexp, rmode=5, anat=a)
    assert np.all(np.isfinite(ref)) and (ref == 1e7 * 0.1).sum() > 0
    assert np.array_equal(got, ref)
    # the regulariser differs from the spherical one (the case exercises the anatomical branch)
    sph = O.oracle_triplet_costs(kind, 2, ot, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                 s["src"], prow if kind >= 3 else None, pmem if kind >= 3 else None, s["src_feat"], s["ref_feat"], cfw, s["absw"],
                                 0.1, 0.4, 1.6, kexp, rexp)
    assert (sph != got).mean() > 0.9


@pytest.mark.parametrize("sim", [2, 1])
def test_group_patch_data_and_pair_costs(O, sim):
    g = group_setup()
    S, ncp = g["cps"].shape[0], g["cps"].shape[1]
    # host glue of DiscreteGroupModel restated with numpy + the oracle's rotation matrices
    rot = np.array([O.oracle_rotation_matrix(g["centre"], c) for c in g["cps"].reshape(-1, 3)]).reshape(-1, 9)
    spacings = np.zeros((S, ncp))
    for s_ in range(S):
        for a, b in ((0, 1), (1, 2), (0, 2)):
            d = np.sqrt(((g["cps"][s_][g["cp_tri"][:, a]] - g["cps"][s_][g["cp_tri"][:, b]]) ** 2).sum(axis=1))
            geo = 2 * 100.0 * np.arcsin(d / 200.0)
            np.maximum.at(spacings[s_], g["cp_tri"][:, a], geo)
            np.maximum.at(spacings[s_], g["cp_tri"][:, b], geo)
    # estimate_pairs (DiscreteGroupModel.cpp:37-55): partner = nearest control point of subject B
    near = lambda a, v, b: int(np.argmin(((g["cps"][b] - g["cps"][a][v]) ** 2).sum(axis=1)))
    pairs = np.array([[a * ncp + v, b * ncp + near(a, v, b)] for a in range(S) for v in range(ncp) for b in range(a + 1, S)], np.int32)
    rng = np.random.default_rng(9)
    n, L = 1500, len(g["labels"])
    rp, la, lb = rng.integers(0, len(pairs), n), rng.integers(0, L, n), rng.integers(0, L, n)
    ref, ref_fields = O.refmr_group_pair_costs(sim, g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"], ncp, rot, spacings, 1.0,
                                               pairs, rp, la, lb, want_fields=True)
    fields = O.oracle_group_fields(g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"])
    seen = ~np.isnan(ref_fields)
    assert seen.mean() > 0.5
    assert np.array_equal(fields[seen], ref_fields[seen])
    got = O.oracle_group_pair_costs(sim, ncp, g["tpl"], fields, rot, g["labels"], spacings, 1.0, pairs, rp, la, lb)
    # an empty patch intersection makes the reference read patch_data_A[0] of an empty vector (DiscreteGroupCostFunction.cpp:79,
    # undefined behaviour); the oracle returns NaN there and those requests are not compared
    ok = ~np.isnan(got)
    assert ok.mean() > 0.9
    assert np.array_equal(got[ok], ref[ok])
    # with a cost mask (set_masks, DiscreteGroupModel.cpp:164): common vertices weighted by |mask| (DiscreteGroupCostFunction.cpp:77)
    mask = group_mask(g)
    ref_m = O.refmr_group_pair_costs(sim, g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"], ncp, rot, spacings, 1.0,
                                     pairs, rp, la, lb, mask=mask)
    got_m = O.oracle_group_pair_costs(sim, ncp, g["tpl"], fields, rot, g["labels"], spacings, 1.0, pairs, rp, la, lb, mask=mask)
    assert np.array_equal(np.isnan(got_m), np.isnan(got))
    assert np.array_equal(got_m[ok], ref_m[ok])
    assert (got_m[ok] != got[ok]).mean() > 0.5           # the mask matters


@pytest.mark.parametrize("kexp,rexp", [(2.0, 2.0), (1.5, 1.3)])
def test_group_triplet_costs(O, kexp, rexp):
    g = group_setup()
    orig, trip, rot, (rt, la, lb, lc) = group_triplet_case(O, g)
    ref = O.refmr_group_triplet_costs(g["cps"], orig, g["cp_tri"], rot, g["labels"], trip, 0.05, 0.4, 1.6, kexp, rexp, rt, la, lb, lc)
    got = O.oracle_group_triplet_costs(g["cps"], orig, rot, g["labels"], trip, rt, la, lb, lc, 0.05, 0.4, 1.6, kexp, rexp)
    assert np.all(np.isfinite(ref)) and np.ptp(ref) > 0
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind,D", [(0, 1), (1, 9), (2, 4)])
@pytest.mark.parametrize("sim,pct", [(4, 0.75), (4, 0.4), (5, 0.75)])
def test_unary_costs_dice(O, kind, D, sim, pct):
    """simmeasure 4 / 5: DICE / genDICE at a percentile threshold (similarities.cpp:201-253), the reference's "experimental" measures."""
    from newmsm_b200 import synth
    s = cost_setup(O, 2, 4, D)
    cp_tri = synth.icosphere(2)[1]
    ot = O.OracleOctree(s["xyz"], s["tri"])
    O.refmr_set_percentile(pct)
    O.oracle_set_percentile(pct)
    try:
        ref, prow, pmem, _ = O.refmr_unary(kind, sim, s["xyz"], s["tri"], s["cp"], cp_tri, s["rot"], s["labels"], s["src"], s["tri"],
                                           s["src_feat"], s["ref_feat"], None, s["absw"], s["maxsep"], 1.0, nthreads=1)
        got = O.oracle_unary_costs(kind, sim, ot, s["cp"], s["rot"], s["labels"], s["src"], prow, pmem, s["src_feat"], s["ref_feat"], None, s["absw"])
    finally:
        O.refmr_set_percentile(0.75)
        O.oracle_set_percentile(0.75)
    assert np.array_equal(got, ref)
    assert np.ptp(ref) > 0


def test_rigid_level_reference_reproduces_golden(O):
    """SURVEY §8 f4 groundwork: the reference's RIGID / AFFINE level (rigid_costfunction.cpp:32-236), run through the compiled reference
    here, reproduces tests/golden/rigid.npz bit for bit and does not depend on the thread count (its OpenMP loops write disjoint entries).
    The CPU restatement is pinned by the same vectors (test_oracle_golden.py::test_rigid_level_golden); no CUDA path exists yet."""
    import os
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_rigid", os.path.join(sys_path, "make_golden_rigid.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(sys_path, "rigid.npz"))
    for name, D, sim in (("d2_corr", 2, 2), ("d3_ssd", 3, 1)):
        xyz, tri, src, mov, ref = mod.rigid_case(3, D)
        for threads in (1, 4):
            moved, cost0, rowptr, members = O.refmr_rigid(xyz, tri, src, tri, mov, ref, simmeasure=sim, iters=4, nthreads=threads)
            assert np.array_equal(moved, g[f"{name}_xyz"]) and cost0 == float(g[f"{name}_cost0"])
            assert np.array_equal(rowptr, g[f"{name}_rowptr"]) and np.array_equal(members, g[f"{name}_members"])
        assert np.abs(np.linalg.norm(moved, axis=1) - 100).max() < 1e-9          # a rotation: the sphere radius is kept
        assert rowptr[-1] > 30 * len(xyz)                                       # neighbourhoods of 4 mean vertex distances


@pytest.mark.parametrize("sim,D,seed", [(2, 3, 5), (1, 2, 9)])
def test_rigid_level_oracle_matches_reference(O, sim, D, seed):
    """The restated RIGID / AFFINE level against the reference's own Rigid_cost_function on a jittered ico4 source that differs from
    the target (the golden cases use one icosphere for both): everything bit-exact."""
    from newmsm_b200 import synth
    xyz, tri = synth.icosphere(4)
    src = synth.rotate_sphere(synth.jitter_sphere(xyz, tri, frac=0.2, seed=seed), -0.02, 0.05, 0.01)
    ref = synth.smooth_fields(xyz, D, seed0=40)
    mov = synth.smooth_fields(src, D, seed0=41)
    a = O.oracle_rigid(xyz, tri, src, tri, mov, ref, simmeasure=sim, iters=2)
    b = O.refmr_rigid(xyz, tri, src, tri, mov, ref, simmeasure=sim, iters=2, nthreads=4)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])     # neighbourhoods (reg_tools.cpp:31-58, incl. the tie order of std::sort)
    assert a[1] == b[1]                                                  # cost at zero rotation
    assert np.array_equal(a[0], b[0])                                    # rotated source after run()


def variance_cases():
    rng = np.random.default_rng(77)
    n = 2562
    data = rng.normal(3.0, 2.5, size=(5, n)) * np.array([1.0, 1e-6, 1e6, 1.0, 0.0])[:, None]   # incl. a constant (zero-variance) channel
    data[3] = np.round(data[3])                                                                   # many ties
    excl = (rng.uniform(size=n) > 0.3).astype(np.float64) * rng.uniform(0.5, 2.0, size=n)       # 0 = excluded, positive = kept
    excl[rng.integers(0, n, 50)] = -1.0                                                          # negative values are excluded too (> 0.0 test)
    return data, excl


@pytest.mark.parametrize("masked", [False, True])
def test_variance_normalise(O, masked):
    """newmeshreg::variance_normalise (reg_tools.cpp:804-844, last stage of featurespace::initialise): restatement vs the reference."""
    data, excl = variance_cases()
    e = excl if masked else None
    ref = O.refmr_variance_normalise(data, e, nthreads=4)
    got = O.oracle_variance_normalise(data, e)
    assert np.array_equal(got, ref, equal_nan=True)
    keep = (excl > 0) if masked else np.ones(data.shape[1], bool)
    assert np.allclose(ref[0][keep].mean(), 0, atol=1e-12) and np.allclose(ref[0][keep].std(ddof=1), 1)
    assert np.array_equal(ref[:, ~keep], data[:, ~keep])                 # excluded vertices keep their values
    one = O.refmr_variance_normalise(data[:, :1]), O.oracle_variance_normalise(data[:, :1])   # a single value: 0 / 0 variance, value - mean = 0
    assert np.array_equal(one[0], one[1], equal_nan=True)
