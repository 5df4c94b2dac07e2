"""GPU parity tests: the CUDA path, called through the C ABI (newmsm_b200.capi / resampler /
discrete_cost), against the CPU oracle on the same seeded inputs and against the committed golden
fixtures (outputs of the compiled reference). Indices bit-exact; FP64 outputs bit-exact; FP32
payloads within 1e-5 relative (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from newmsm_b200 import capi, synth
from cost_cases import cost_setup, triplet_setup, group_setup

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ST = {0: 0, 1: capi.ERR_OUT_OF_BOX, 2: capi.ERR_NO_TRIANGLE}   # oracle status -> msmgpu_status


def load(name):
    return np.load(os.path.join(G, name))


@pytest.fixture(scope="module")
def R():
    from newmsm_b200 import build, resampler
    build.build_library()
    assert capi.lib().msmgpu_device_count() > 0, "no CUDA device: the GPU tests cannot run on a fallback"
    return resampler


@pytest.fixture(autouse=True)
def no_pending_cuda_error():
    yield
    import gc
    gc.collect()
    pending = capi.lib().msmgpu_debug_take_cuda_error()
    assert not pending, f"a call left a CUDA error behind: {pending.decode()}"


@pytest.fixture(scope="module")
def meshes():
    out = {k: synth.icosphere(k) for k in (2, 3, 4, 5, 6)}
    x5, t5 = out[5]
    out["j5"] = (synth.jitter_sphere(x5, t5, frac=0.3, seed=11), t5)
    return out


def query_points(xyz, seed=0, n_rand=20000):
    rng = np.random.default_rng(seed)
    rot = synth.rotate_sphere(xyz)
    rnd = rng.normal(size=(n_rand, 3))
    rnd = rnd / np.linalg.norm(rnd, axis=1, keepdims=True) * 100.0
    off = rnd[: n_rand // 4] * rng.uniform(0.6, 1.009, size=(n_rand // 4, 1))       # off the sphere, inside the cube
    box = rng.uniform(-101, 101, size=(2000, 3))                                     # anywhere in the root cube
    out_of_box = np.array([[101.5, 0, 0], [0, -102, 0], [50, 50, 101.0001]])
    return np.concatenate([xyz, rot, rnd, off, box, out_of_box])


def adversarial_points(xyz, tri):
    mids = (xyz[tri[:, 0]] + xyz[tri[:, 1]]) / 2      # on edges (not re-projected: inside the sphere)
    cent = xyz[tri].mean(axis=1)
    on_mid_planes = np.array([[0.0, 0, 100], [0, 100, 0], [100, 0, 0], [0, 0, -100], [50.5, 50.5, 50.5], [-50.5, 25.25, 0]])
    return np.concatenate([mids, cent, mids / np.linalg.norm(mids, axis=1, keepdims=True) * 100, on_mid_planes])


@pytest.mark.parametrize("key", [2, 3, 4, 5, 6, "j5"])
def test_octree_topology_matches_oracle(R, oracle_built, meshes, key):
    xyz, tri = meshes[key]
    ot = oracle_built.OracleOctree(xyz, tri)
    k0, c0, t0 = ot.dump()
    m = R.Mesh(xyz, tri)
    k1, c1, t1 = R.Octree(m).dump()
    assert np.array_equal(k0, k1) and np.array_equal(c0, c1) and np.array_equal(t0, t1)


def test_octree_forest_batch(R, oracle_built, meshes):
    keys = [4, "j5", 3, 5]
    ms = [R.Mesh(*meshes[k]) for k in keys]
    trees = R.Octree.build_batch(ms)
    for k, t in zip(keys, trees):
        ref = oracle_built.OracleOctree(*meshes[k]).dump()
        got = t.dump()
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
        q = synth.rotate_sphere(meshes[k][0])
        assert np.array_equal(t.get_closest_triangle(q), oracle_built.OracleOctree(*meshes[k]).query(q)[0])


def test_octree_golden(R):
    g = load("octree_ico3.npz")
    t = R.Octree(R.Mesh(g["xyz"], g["tri"]))
    kinds, counts, leaf_tris = t.dump()
    assert np.array_equal(kinds, g["kinds"]) and np.array_equal(counts, g["counts"]) and np.array_equal(leaf_tris, g["leaf_tris"])
    tri, vtx, st = t.query(g["q"])
    assert np.array_equal(tri, g["tri_id"])
    assert np.array_equal(st, np.vectorize(ST.get)(g["status"]))
    ok = st == 0
    assert np.array_equal(vtx[ok], g["vertex_id"][ok])
    idx = np.zeros((ok.sum(), 3), np.int32); w = np.zeros((ok.sum(), 3)); ne = np.zeros(ok.sum(), np.int32)
    q = np.ascontiguousarray(g["q"][ok])
    capi.check(t.L.msmgpu_bary_weights(t.h, len(q), capi.ptr(q), capi.ptr(idx), capi.ptr(w), capi.ptr(ne)))
    assert np.array_equal(idx, g["w_idx"]) and np.array_equal(w, g["w_val"]) and np.array_equal(ne, g["w_n"])
    assert np.array_equal(t.mesh.vertex_areas(), g["vertex_area"])


@pytest.mark.parametrize("group", [1, 2, 4, 8, 16, 32])
@pytest.mark.parametrize("key", [3, 5, "j5"])
def test_nearest_triangle_bit_exact(R, oracle_built, meshes, key, group):
    xyz, tri = meshes[key]
    capi.check(capi.lib().msmgpu_set_query_group(group))
    try:
        q = np.concatenate([query_points(xyz, seed=group), adversarial_points(xyz, tri)])
        t0, v0, s0, _ = oracle_built.OracleOctree(xyz, tri).query(q)
        t1, v1, s1 = R.Octree(R.Mesh(xyz, tri)).query(q)
        assert np.array_equal(s1, np.vectorize(ST.get)(s0))
        assert np.array_equal(t0, t1)
        ok = s0 == 0
        assert np.array_equal(v0[ok], v1[ok])
        assert (s0 == 1).sum() == 3
    finally:
        capi.check(capi.lib().msmgpu_set_query_group(1))


@pytest.mark.parametrize("lo,hi", [(4, 5), (3, 5), (5, 5), (4, 6)])
def test_nearest_triangle_nested_icospheres(R, oracle_built, lo, hi):
    """Vertices of a finer icosphere located in a coarser one (what project_CPgrid does between resolution levels,
    mesh_registration.cpp:141-154): most queries sit exactly on an edge or a vertex of the mesh, i.e. on the ties of the search."""
    xyz, tri = synth.icosphere(lo)
    q = synth.icosphere(hi)[0]
    ot, ov, os_, _ = oracle_built.OracleOctree(xyz, tri).query(q)
    gt, gv, gs = R.Octree(R.Mesh(xyz, tri)).query(q)
    assert np.array_equal(gs, np.vectorize(ST.get)(os_)) and np.array_equal(gt, ot)
    ok = os_ == 0
    assert ok.all() and np.array_equal(gv[ok], ov[ok])


def test_query_raises_like_reference(R, meshes):
    t = R.Octree(R.Mesh(*meshes[3]))
    with pytest.raises(R.MeshException) as e:
        t.get_closest_triangle(np.array([[0.0, 0.0, 150.0]]))
    assert e.value.status == capi.ERR_OUT_OF_BOX and "bounding box" in e.value.message
    assert len(t.get_closest_triangle(np.zeros((0, 3)))) == 0     # empty input


def test_bary_weights_bit_exact(R, oracle_built, meshes):
    xyz, tri = meshes[6]
    low = synth.rotate_sphere(meshes[5][0])
    i0, w0, n0, err = oracle_built.OracleOctree(xyz, tri).bary_weights(low)
    assert err == 0
    m = R.Mesh(xyz, tri)
    i1, w1, n1 = R.Resampler().get_barycentric_weights(R.Mesh(low, meshes[5][1]), m, R.Octree(m))
    assert np.array_equal(i0, i1) and np.array_equal(n0, n1) and np.array_equal(w0, w1)
    # resampling a mesh onto itself gives identity weights (SURVEY §4 known answer)
    i2, w2, n2 = R.Resampler().get_barycentric_weights(m, m, R.Octree(m))
    hit = np.take_along_axis(w2, np.argmax(w2, axis=1)[:, None], 1)[:, 0]
    assert np.allclose(hit, 1.0, atol=1e-9)
    assert np.array_equal(np.take_along_axis(i2, np.argmax(w2, axis=1)[:, None], 1)[:, 0], np.arange(len(xyz)))


@pytest.mark.parametrize("D", [1, 4, 7, 40])
def test_fused_bary_resample(R, oracle_built, meshes, D):
    xyz, tri = meshes[5]
    low = synth.rotate_sphere(meshes[4][0], 0.02, -0.01, 0.05)
    feat = synth.smooth_fields(xyz, D).astype(np.float32).astype(np.float64)     # FP32 payload, exactly representable
    ref = oracle_built.oracle_bary_resample(xyz, tri, low, feat)
    got = R.barycentric_resample(R.Mesh(xyz, tri), low, feat)
    # FP64 accumulation in the reference's order, rounded once to FP32 on output
    assert np.array_equal(got, ref.astype(np.float32).astype(np.float64))
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-5 * scale


def test_fused_bary_resample_linear_field(R, meshes):
    """A linear field a.x is reproduced exactly (to FP32) inside planar triangles at the projected point."""
    xyz, tri = meshes[6]
    low = synth.rotate_sphere(meshes[5][0])
    a = np.array([[0.3, -0.2, 0.5], [1.0, 0.0, 0.0]])
    feat = (a @ xyz.T).astype(np.float32).astype(np.float64)
    got = R.barycentric_resample(R.Mesh(xyz, tri), low, feat)
    m = R.Mesh(xyz, tri)
    tri_id = R.Octree(m).get_closest_triangle(low)
    v = xyz[tri[tri_id]]                                   # project the query into the triangle plane
    nrm = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    s = (nrm * v[:, 0]).sum(1) / (nrm * low).sum(1)
    expect = a @ (low * s[:, None]).T
    assert np.abs(got - expect).max() < 2e-5 * np.abs(expect).max()


@pytest.mark.parametrize("name", ["down", "up"])
def test_adaptive_weights_golden(R, name):
    g = load(f"resample_{name}.npz")
    m_in, m_low = R.Mesh(g["xyz_in"], g["tri_in"], g["feat"]), R.Mesh(g["xyz_low"], g["tri_low"])
    rowptr, col, val = R.Resampler().get_adaptive_barycentric_weights(m_in, m_low).csr()
    assert np.array_equal(rowptr, g["rowptr"]) and np.array_equal(col, g["col"])
    assert np.array_equal(val, g["val"])
    assert np.array_equal(R.metric_resample(m_in, m_low), g["metric_out"])
    got = R.barycentric_resample(m_in, g["xyz_low"])
    assert np.abs(got - g["bary_out"]).max() <= 1e-5 * np.abs(g["bary_out"]).max()
    f32out = R.metric_resample_f32(m_in, m_low, g["feat"].astype(np.float32))
    assert np.abs(f32out - g["metric_out"]).max() <= 1e-5 * np.abs(g["metric_out"]).max()


@pytest.mark.parametrize("pair", [(6, 5), (5, "j5"), ("j5", 4), (4, 5)])
def test_adaptive_weights_vs_oracle(R, oracle_built, meshes, pair):
    (xi, ti), (xl, tl) = meshes[pair[0]], meshes[pair[1]]
    xl = synth.rotate_sphere(xl, 0.05, 0.02, -0.03)
    r0, c0, v0 = oracle_built.oracle_adaptive_weights(xi, ti, xl, tl)
    m_in, m_low = R.Mesh(xi, ti), R.Mesh(xl, tl)
    W = R.Resampler().get_adaptive_barycentric_weights(m_in, m_low)
    r1, c1, v1 = W.csr()
    assert np.array_equal(r0, r1) and np.array_equal(c0, c1) and np.array_equal(v0, v1)
    assert np.array_equal(m_in.vertex_areas(), oracle_built.oracle_vertex_areas(xi, ti))
    feat = synth.smooth_fields(xi, 3)
    m_in.pvalues = feat
    assert np.array_equal(R.metric_resample(m_in, m_low), oracle_built.oracle_metric_resample(xi, ti, xl, tl, feat))
    rows = W.rows()
    assert len(rows) == len(xl) and abs(sum(rows[0].values()) - 1.0) < 1e-12


def test_adaptive_weights_batch_matches_single(R, oracle_built, meshes):
    """msmgpu_adaptive_weights_batch: subjects with DIFFERENT vertex counts in one set of launches."""
    keys = [5, "j5", 4, 6]
    xl, tl = meshes[4][0], meshes[4][1]
    xl = synth.rotate_sphere(xl, 0.01, 0.03, -0.02)
    m_low = R.Mesh(xl, tl)
    ins = [R.Mesh(*meshes[k]) for k in keys]
    Ws = R.Resampler().get_adaptive_barycentric_weights_batch(ins, m_low)
    for k, W in zip(keys, Ws):
        r0, c0, v0 = oracle_built.oracle_adaptive_weights(meshes[k][0], meshes[k][1], xl, tl)
        r1, c1, v1 = W.csr()
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1) and np.array_equal(v0, v1)
    # batched apply == the oracle's metric_resample on an FP32 payload
    import torch
    D = 8
    feats = [synth.smooth_fields(meshes[k][0], D).astype(np.float32) for k in keys]
    d_in = [torch.from_numpy(np.ascontiguousarray(f.T)).cuda() for f in feats]
    d_out = [torch.zeros(len(xl), D, device="cuda") for _ in keys]
    import ctypes as C
    n = len(keys)
    wp = (C.c_void_p * n)(*[W.h.value for W in Ws])
    ip = (C.c_void_p * n)(*[t.data_ptr() for t in d_in])
    op = (C.c_void_p * n)(*[t.data_ptr() for t in d_out])
    torch.cuda.synchronize()
    capi.check(capi.lib().msmgpu_weights_apply_batch_f32_dev(m_low.ctx.h, n, wp, D, ip, op))
    m_low.ctx.sync()
    for k, f, o in zip(keys, feats, d_out):
        ref = oracle_built.oracle_metric_resample(meshes[k][0], meshes[k][1], xl, tl, f.astype(np.float64))
        assert np.array_equal(o.cpu().numpy().T, ref.astype(np.float32))


def test_vertex_areas_high_valence_fallback(R, oracle_built):
    """Vertex areas (mesh.cpp:1275): the fixed 8-slot incidence rows overflow on a valence-12 vertex and the generic buckets take over."""
    ring = 12
    ang = 2 * np.pi * np.arange(ring) / ring
    xyz = np.concatenate([[[0.0, 0.0, 100.0]], np.stack([30 * np.cos(ang) * (1 + 0.1 * np.sin(3 * ang)), 30 * np.sin(ang), 95 + np.cos(2 * ang)], axis=1)])
    tri = np.array([[0, 1 + k, 1 + (k + 1) % ring] for k in range(ring)], np.int32)
    got = R.Mesh(xyz, tri).vertex_areas()
    assert np.array_equal(got, oracle_built.oracle_vertex_areas(xyz, tri))
    xyz6, tri6 = synth.icosphere(3)                      # valence 5 / 6: the fast path
    assert np.array_equal(R.Mesh(xyz6, tri6).vertex_areas(), oracle_built.oracle_vertex_areas(xyz6, tri6))


def test_adaptive_weights_from_kept_forward_maps(R, oracle_built, meshes):
    """msmgpu_fwd: the weight maps kept by the fused barycentric resample give the same adaptive CSR as querying again."""
    import ctypes as C
    import torch
    L = capi.lib()
    ctx = R.Context(0)
    low_xyz, low_tri = synth.rotate_sphere(meshes[4][0], 0.01, 0.02, -0.015), meshes[4][1]
    S, D = 3, 8
    src = [R.Mesh(synth.jitter_sphere(meshes[5][0], meshes[5][1], seed=70 + s), meshes[5][1], ctx=ctx) for s in range(S)]
    low = R.Mesh(low_xyz, low_tri, ctx=ctx)
    trees = R.Octree.build_batch(src + [low])
    n_low, nv = len(low_xyz), src[0].nvertices()
    dev = torch.device("cuda", 0)
    feat = [torch.randn(nv, D, device=dev) for _ in range(S)]
    out = [torch.empty(n_low, D, device=dev) for _ in range(S)]
    d_low = torch.from_numpy(low_xyz).to(dev)
    tp = (C.c_void_p * S)(*[t.h.value for t in trees[:S]])
    mp = (C.c_void_p * S)(*[m.h.value for m in src])
    fp = (C.c_void_p * S)(*[t.data_ptr() for t in feat])
    op = (C.c_void_p * S)(*[t.data_ptr() for t in out])
    fwd = C.c_void_p()
    capi.check(L.msmgpu_fwd_create(ctx.h, S, n_low, C.byref(fwd)))
    capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n_low, d_low.data_ptr(), D, fp, op, None, fwd))
    w1, w2 = (C.c_void_p * S)(), (C.c_void_p * S)()
    capi.check(L.msmgpu_adaptive_weights_batch_fwd(ctx.h, S, mp, tp, low.h, trees[-1].h, fwd, w1))
    capi.check(L.msmgpu_adaptive_weights_batch(ctx.h, S, mp, tp, low.h, trees[-1].h, w2))
    for s in range(S):
        a, b = R.Weights(L, C.c_void_p(w1[s])), R.Weights(L, C.c_void_p(w2[s]))
        for x, y in zip(a.csr(), b.csr()):
            assert np.array_equal(x, y)
        ref = oracle_built.oracle_adaptive_weights(src[s].xyz, meshes[5][1], low_xyz, low_tri)
        for x, y in zip(a.csr(), ref):
            assert np.array_equal(x, y)
        a.close(); b.close()
    # a store filled for other trees is refused
    other = R.Octree.build_batch([src[1], src[0], src[2]])
    tp2 = (C.c_void_p * S)(*[t.h.value for t in other])
    w3 = (C.c_void_p * S)()
    assert L.msmgpu_adaptive_weights_batch_fwd(ctx.h, S, mp, tp2, low.h, trees[-1].h, fwd, w3) != 0
    L.msmgpu_fwd_destroy(fwd)


def test_resident_features_match_host_payload_calls(R, meshes):
    """msmgpu_mesh_set_features_f32 + the two *_mesh_* resamplers give the same floats as the calls that upload per call."""
    import ctypes as C
    xyz, tri = meshes[5]
    low_xyz, low_tri = synth.rotate_sphere(meshes[4][0]), meshes[4][1]
    feat = synth.smooth_fields(xyz, 12).astype(np.float32)
    m, low = R.Mesh(xyz, tri), R.Mesh(low_xyz, low_tri)
    t = R.Octree(m)
    L = capi.lib()
    capi.check(L.msmgpu_mesh_set_features_f32(m.h, 12, capi.ptr(feat)))
    a = np.zeros((12, len(low_xyz)), np.float32); b = np.zeros_like(a); c = np.zeros_like(a)
    capi.check(L.msmgpu_mesh_bary_resample_f32(t.h, len(low_xyz), capi.ptr(low_xyz), capi.ptr(a)))
    capi.check(L.msmgpu_bary_resample_f32(t.h, len(low_xyz), capi.ptr(low_xyz), 12, capi.ptr(feat), capi.ptr(b)))
    assert np.array_equal(a, b)
    capi.check(L.msmgpu_mesh_metric_resample_f32(m.h, t.h, low.h, None, capi.ptr(c)))
    assert np.array_equal(c, R.metric_resample_f32(m, low, feat))


def test_blend_golden(R):
    g = load("blend.npz")
    m = R.Mesh(g["xf"], g["tf"])
    assert np.array_equal(R.sphere_project_warp(g["sph"], m, g["xto"]), g["warp"])
    assert np.array_equal(R.surface_resample(g["anat"], m, g["sph"]), g["surf"])
    assert np.array_equal(R.nearest_neighbour_interpolation(m, g["sph"], g["feat"]), g["nn"])


def test_rotation_golden(R):
    g = load("rotation.npz")
    assert np.array_equal(R.estimate_rotation_matrix(g["ci"], g["index"]), g["R"])


def test_cpp_adapter_drop_in(R):
    """The reference's own Mesh objects through include/newmsm_b200/resampler_adapter.hpp, compared in-process with
    the reference's CPU functions (integration/adapter_check.cpp, built here into oracle/_ref/ where /root/reference is)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "adapter_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_check not built (needs /root/reference: make -C oracle adapter_check)")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ADAPTER CHECK PASSED" in res.stdout, res.stdout + res.stderr


# ---------------------------------------------------------------------------------------------
# cost functions
# ---------------------------------------------------------------------------------------------
def test_patch_membership_bit_exact(R, oracle_built):
    from newmsm_b200 import discrete_cost as DC
    s = cost_setup(oracle_built, 3, 5, 1)
    cf = DC.UnivariateNonLinearSRegDiscreteCostFunction()
    cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
    for rng_ in (1.0, 0.5, 1.7):
        cf.reset_CPgrid(s["cp"], s["maxsep"], rng_)
        r1, m1 = cf.get_source_data()
        r0, m0 = oracle_built.oracle_patch_membership(s["cp"], s["src"], s["maxsep"], rng_)
        assert np.array_equal(r0, r1) and np.array_equal(m0, m1)
    # knife edge: thresholds that EQUAL a computed geodesic distance (strict '<' must exclude that vertex)
    k = np.arange(len(s["cp"]))
    chord = np.linalg.norm(s["cp"][k] - s["src"][(k * 37) % len(s["src"])], axis=1)
    chord = np.sqrt(((s["cp"][k] - s["src"][(k * 37) % len(s["src"])]) ** 2) @ np.ones(3))
    ms = 2 * 100.0 * np.arcsin(np.minimum(chord, 15.0) / (2 * 100.0))
    cf.reset_CPgrid(s["cp"], ms, 1.0)
    r1, m1 = cf.get_source_data()
    r0, m0 = oracle_built.oracle_patch_membership(s["cp"], s["src"], ms, 1.0)
    assert np.array_equal(r0, r1) and np.array_equal(m0, m1)


@pytest.mark.parametrize("kind,D", [(0, 1), (1, 6), (2, 6)])
@pytest.mark.parametrize("sim", [1, 2])
def test_unary_costs_bit_exact(R, oracle_built, kind, D, sim):
    from newmsm_b200 import discrete_cost as DC
    s = cost_setup(oracle_built, 3, 5, D)
    cls = [DC.UnivariateNonLinearSRegDiscreteCostFunction, DC.MultivariateNonLinearSRegDiscreteCostFunction,
           DC.PatchwiseMultivariateNonLinearSRegDiscreteCostFunction][kind]
    cf = cls(simmeasure=sim)
    cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
    rng = np.random.default_rng(5)
    cfw = rng.uniform(0.2, 1.0, size=(D if kind == 1 else 1, len(s["src"])))
    for weights in (None, cfw):
        cf.reset_CPgrid(s["cp"], s["maxsep"], 1.0, HIGHREScfweight=weights, AbsoluteWeights=s["absw"])
        prow, pmem = cf.get_source_data()
        got, tri_got = cf.computeUnaryCosts(s["labels"], s["rot"], want_triangles=True)
        ot = oracle_built.OracleOctree(s["xyz"], s["tri"])
        ref, tri_ref = oracle_built.oracle_unary_costs(kind, sim, ot, s["cp"], s["rot"], s["labels"], s["src"], prow, pmem,
                                                       s["src_feat"], s["ref_feat"], weights, s["absw"], want_tri=True)
        assert np.array_equal(tri_got, tri_ref)
        assert np.array_equal(got, ref)
        assert cf.computeUnaryCost(5, 2) == ref[2, 5]
    # corr(A, A) = 1 -> cost 0 when the source is the target and label 0 = "stay" (SURVEY §4 known answer)
    if sim == 2 and kind == 0:
        cf2 = cls(simmeasure=2)
        cf2.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["xyz"], s["ref_feat"], s["ref_feat"])
        cf2.reset_CPgrid(s["cp"], s["maxsep"], 1.0)
        c = cf2.computeUnaryCosts(s["labels"][:1], s["rot"])
        assert np.abs(c).max() < 1e-12


# ---------------------------------------------------------------------------------------------
# triplet costs: strain regulariser + HO likelihood (the oracle is pinned against the reference in tests/test_oracle_vs_refmr.py)
# ---------------------------------------------------------------------------------------------
def rel_close(a, b, tol):
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("kexp,rexp", [(2.0, 2.0), (2.0, 1.0), (1.5, 1.3)])
def test_triplet_strain_costs(R, oracle_built, kexp, rexp):
    from newmsm_b200 import discrete_cost as DC
    s = triplet_setup(oracle_built, 3, 5, 1)
    cf = DC.UnivariateNonLinearSRegDiscreteCostFunction()
    cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
    cf.reset_CPgrid(s["cp_now"], s["maxsep"], 1.0)
    cf.set_parameters(0.1, 0.4, 1.6, kexp, rexp, 3)
    cf.setTriplets(s["triplets"], s["labels"], s["rot_now"], s["orig"])
    rt, la, lb, lc = s["req"]
    got = cf.computeTripletCostList(rt, la, lb, lc)
    ref = oracle_built.oracle_triplet_costs(0, 2, None, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                            s["src"], None, None, s["src_feat"], s["ref_feat"], None, np.ones(len(s["cp"])), 0.1, 0.4, 1.6, kexp, rexp)
    assert np.all(np.isfinite(got))
    # bit-exact: the pow() calls of the strain energy run on the host libm (csrc/triplet.cu), everything else is IEEE + - * / sqrt
    assert np.array_equal(got, ref)
    # the 8-combination batch equals the list form (Fusion.h:181-196)
    labeling = np.random.default_rng(3).integers(0, len(s["labels"]), len(s["cp"])).astype(np.int32)
    batch = cf.computeTripletCostsForLabel(labeling, 2)
    T = len(s["triplets"])
    tt = np.repeat(np.arange(T, dtype=np.int32), 8)
    combo = np.tile(np.arange(8), T)
    pick = lambda k, bit: np.where((combo >> bit) & 1, 2, labeling[s["triplets"][tt, k]]).astype(np.int32)
    lst = cf.computeTripletCostList(tt, pick(0, 2), pick(1, 1), pick(2, 0))
    assert np.array_equal(batch.reshape(-1), lst)
    # an undeformed triangle has zero strain: label 0 = "stay" with ORIG == current grid
    cf.reset_CPgrid(s["cp"], s["maxsep"], 1.0)
    cf.setTriplets(s["triplets"], s["labels"], s["rot"], s["orig"])
    z = cf.computeTripletCostList(np.arange(T, dtype=np.int32), np.zeros(T, np.int32), np.zeros(T, np.int32), np.zeros(T, np.int32))
    assert np.abs(z).max() < 1e-12
    # folding: swapping two corners' destinations flips the normal -> FOLDING * lambda
    # folding: a control point reflected through the opposite edge flips the triangle's normal -> FOLDING * lambda (cpp:150-155)
    a, b, c = s["triplets"][0]
    folded = s["cp_now"].copy()
    mid = (folded[b] + folded[c]) / 2
    p = 2 * mid - folded[a]
    folded[a] = p / np.linalg.norm(p) * 100
    centre = np.array([0.0, 0.0, 100.0])
    rot_f = s["rot_now"].copy()
    rot_f[a] = oracle_built.oracle_rotation_matrix(centre, folded[a]).reshape(9)
    # the fold test compares ROTATIONS * label with the CURRENT grid: the grid stays unfolded, node a's rotation carries it across the edge
    cf.reset_CPgrid(s["cp_now"], s["maxsep"], 1.0)
    cf.setTriplets(s["triplets"], s["labels"], rot_f, s["orig"])
    z3 = np.zeros(1, np.int32)
    got_f = cf.computeTripletCostList(z3, z3, z3, z3)
    ref_f = oracle_built.oracle_triplet_costs(0, 2, None, s["cp_now"], s["orig"], rot_f, s["labels"], s["triplets"], z3, z3, z3, z3,
                                              s["src"], None, None, s["src_feat"], s["ref_feat"], None, np.ones(len(s["cp"])), 0.1, 0.4, 1.6, kexp, rexp)
    assert ref_f[0] == 1e7 * 0.1 and np.array_equal(got_f, ref_f)


@pytest.mark.parametrize("kind,D", [(3, 1), (4, 5)])
@pytest.mark.parametrize("sim", [1, 2])
def test_ho_triplet_likelihood(R, oracle_built, kind, D, sim):
    from newmsm_b200 import discrete_cost as DC
    s = triplet_setup(oracle_built, 3, 5, D)
    cls = DC.HOUnivariateNonLinearSRegDiscreteCostFunction if kind == 3 else DC.HOMultivariateNonLinearSRegDiscreteCostFunction
    cf = cls(simmeasure=sim)
    cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
    rng = np.random.default_rng(5)
    cfw = rng.uniform(0.2, 1.0, size=(D, len(s["src"])))
    cf.reset_CPgrid(s["cp_now"], s["cp_tri"], HIGHREScfweight=cfw, AbsoluteWeights=s["absw"])
    prow, pmem = cf.get_source_data()
    r0, m0 = oracle_built.oracle_ho_patches(s["cp_now"], s["cp_tri"], s["src"])
    assert np.array_equal(prow, r0) and np.array_equal(pmem, m0)
    assert prow[-1] == len(s["src"])                       # every source vertex belongs to exactly one CP triangle
    cf.set_parameters(0.05)
    cf.setTriplets(s["triplets"], s["labels"], s["rot_now"], s["orig"])
    rt, la, lb, lc = s["req"]
    got = cf.computeTripletCostList(rt, la, lb, lc)
    ot = oracle_built.OracleOctree(s["xyz"], s["tri"])
    ref = oracle_built.oracle_triplet_costs(kind, sim, ot, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                            s["src"], prow, pmem, s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05)
    assert np.array_equal(got, ref)


# ---------------------------------------------------------------------------------------------
# pow() of the host C library on the device (csrc/hostpow.cuh): the strain costs call it three times per request
# ---------------------------------------------------------------------------------------------
def test_device_pow_equals_host_libm_bit_for_bit(R):
    import ctypes
    L = capi.lib()
    if not L.msmgpu_device_pow_enabled():
        pytest.skip("the host C library's pow tables were not recognised: costs are finished on the host (no device pow to test)")
    libm = ctypes.CDLL("libm.so.6")
    libm.pow.restype = ctypes.c_double
    libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    rng = np.random.default_rng(5)
    n = 60000
    near1 = 1.0 + rng.uniform(-0.5, 1.0, n) * 2.0 ** -rng.integers(0, 30, n)
    small = rng.uniform(0, 1, n) * 2.0 ** -rng.integers(0, 60, n).astype(np.float64)
    wide = np.exp(rng.uniform(-20, 20, n))
    with np.errstate(invalid="ignore"):
        anyb = rng.integers(0, 2 ** 63, n, dtype=np.int64).view(np.float64) * rng.choice([-1.0, 1.0], n)
    x = np.concatenate([near1, small, wide, anyb, np.abs(anyb), -np.ldexp(1.0 + rng.uniform(0, 1, n), rng.integers(-20, 20, n)),
                        [0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1.7976931348623157e308]])
    y = np.concatenate([rng.choice([2.0, 1.0, 0.5, 3.0, 1.5, 1.3], 2 * n), rng.uniform(-4, 4, n), rng.integers(0, 2 ** 63, n, dtype=np.int64).view(np.float64),
                        rng.uniform(-1200, 1200, n), rng.integers(-20, 21, n).astype(np.float64),
                        [2.0, 2.0, np.inf, np.inf, 2.0, 3.0, 0.0, 2.0, 0.5, 2.0]])
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    assert len(x) == len(y)
    got = np.zeros(len(x))
    ctx = R.Context(0)
    capi.check(L.msmgpu_debug_device_pow(ctx.h, len(x), capi.ptr(x), capi.ptr(y), capi.ptr(got)))
    ref = np.array([libm.pow(float(a), float(b)) for a, b in zip(x, y)])
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.int64), ref[~nan].view(np.int64)), "device pow differs from the host C library's pow"
    assert np.isfinite(ref).mean() > 0.6
    ctx.close()


def test_triplet_costs_same_with_host_finish():
    """MSMGPU_DEVICE_POW=0 (the fallback when the host library is not recognised): the kernels stop at the pow arguments and the host libm
    finishes; the reference-golden triplet tests must pass on that path too (the switch is read once per process: own interpreter)."""
    import subprocess, sys
    env = dict(os.environ, MSMGPU_DEVICE_POW="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k",
                        "test_triplet_costs_vs_reference_golden or test_group_costs_vs_reference_golden or test_group_triplet_nan or test_anatomical_strain"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


# ---------------------------------------------------------------------------------------------
# groupwise (gMSM): resampled fields per (subject,label) and pair costs (oracle pinned in tests/test_oracle_vs_refmr.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sim", [2, 1])
def test_group_fields_and_pair_costs(R, oracle_built, sim):
    from newmsm_b200 import group_cost as GC
    g = group_setup()
    S, ncp = g["cps"].shape[0], g["cps"].shape[1]
    M = GC.DiscreteGroupModel(R.Mesh(g["tpl"], g["tpl_tri"]), simmeasure=sim)
    spacings = M.get_spacings(g["cps"], g["cp_tri"])
    rot = M.get_rotations(g["centre"], g["cps"])
    pairs = M.estimate_pairs(g["cps"], g["cp_tri"])
    assert len(pairs) == S * (S - 1) // 2 * ncp
    # pair partner = nearest control point of subject B (DiscreteGroupModel.cpp:51)
    a0, b0 = pairs[5]
    sb = b0 // ncp
    d = np.linalg.norm(g["cps"][sb] - g["cps"][a0 // ncp][a0 % ncp], axis=1)
    assert d[b0 % ncp] <= d.min() * 1.5
    fields = M.get_patch_data(g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], rot, spacings, 1.0)
    ref_fields = oracle_built.oracle_group_fields(g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"])
    got_fields = fields.cpu().numpy().transpose(0, 1, 3, 2)            # [S][L][D][n_tpl]
    assert np.array_equal(got_fields, ref_fields)                      # FP64 adaptive resampling: bit-exact
    rng = np.random.default_rng(9)
    n, L = 3000, len(g["labels"])
    rp, la, lb = rng.integers(0, len(pairs), n), rng.integers(0, L, n), rng.integers(0, L, n)
    got = M.computePairwiseCostList(pairs, rp, la, lb)
    ref = oracle_built.oracle_group_pair_costs(sim, ncp, g["tpl"], ref_fields, rot, g["labels"], spacings, 1.0, pairs, rp, la, lb)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ok.mean() > 0.9
    assert np.array_equal(got[ok], ref[ok])
    # Fusion's 4 combinations == the list form
    labeling = rng.integers(0, L, S * ncp).astype(np.int32)
    batch = M.computePairwiseCostsForLabel(pairs, labeling, 3)
    P = len(pairs)
    pp = np.repeat(np.arange(P), 4); combo = np.tile(np.arange(4), P)
    la4 = np.where(combo & 2, 3, labeling[pairs[pp, 0]]); lb4 = np.where(combo & 1, 3, labeling[pairs[pp, 1]])
    lst = M.computePairwiseCostList(pairs, pp, la4, lb4)
    assert np.array_equal(np.nan_to_num(batch.reshape(-1), nan=-1.0), np.nan_to_num(lst, nan=-1.0))
    # a subject compared with itself under the same label has perfectly correlated patches: cost 0 (corr) / 0 (SSD)
    self_pairs = np.array([[v, v] for v in range(ncp)], dtype=np.int32)
    z = M.computePairwiseCostList(self_pairs, np.arange(ncp), np.full(ncp, 2), np.full(ncp, 2))
    assert np.nanmax(np.abs(z)) < 1e-12


@pytest.mark.parametrize("D", [1, 3, 12])
def test_group_pair_costs_channel_counts_and_mask(R, oracle_built, D, monkeypatch):
    """The pair-cost kernels (thread per request for D <= 8 in chunks of 1 / 2 / 4 channels, warp per request above) against the oracle,
    without and with a cost mask (set_masks: weights |mask|, DiscreteGroupCostFunction.cpp:77), both similarity measures; the two
    kernels agree bit for bit where both apply."""
    from newmsm_b200 import group_cost as GC
    from cost_cases import group_mask
    g = group_setup(S=3, cp_level=2, data_level=3, tpl_level=4, D=D)
    ncp = g["cps"].shape[1]
    mask = group_mask(g)
    rng = np.random.default_rng(11)
    ref_fields = None
    for sim in (2, 1):
        M = GC.DiscreteGroupModel(R.Mesh(g["tpl"], g["tpl_tri"]), simmeasure=sim)
        spacings = M.get_spacings(g["cps"], g["cp_tri"])
        rot = M.get_rotations(g["centre"], g["cps"])
        pairs = M.estimate_pairs(g["cps"], g["cp_tri"])
        fields = M.get_patch_data(g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], rot, spacings, 1.0)
        if ref_fields is None:
            ref_fields = oracle_built.oracle_group_fields(g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"])
        assert np.array_equal(fields.cpu().numpy().transpose(0, 1, 3, 2), ref_fields)
        n, L = 2500, len(g["labels"])
        rp, la, lb = rng.integers(0, len(pairs), n), rng.integers(0, L, n), rng.integers(0, L, n)
        for mk in (None, mask):
            M.set_masks(mk)
            ref = oracle_built.oracle_group_pair_costs(sim, ncp, g["tpl"], ref_fields, rot, g["labels"], spacings, 1.0, pairs, rp, la, lb, mask=mk)
            got = {}
            for kern in ("thread", "warp"):
                monkeypatch.setenv("MSMGPU_PAIR_KERNEL", kern)
                got[kern] = M.computePairwiseCostList(pairs, rp, la, lb)
            monkeypatch.delenv("MSMGPU_PAIR_KERNEL")
            dflt = M.computePairwiseCostList(pairs, rp, la, lb)
            ok = ~np.isnan(ref)
            assert ok.mean() > 0.9
            for v in (got["thread"], got["warp"], dflt):
                assert np.array_equal(np.isnan(v), np.isnan(ref)) and np.array_equal(v[ok], ref[ok])
        M.close()


# ---------------------------------------------------------------------------------------------
# CUDA path against the outputs of the reference's OWN cost-function classes (tests/golden/costs.npz,
# generated by tests/golden/make_golden_costs.py from oracle/_ref/libref_newmeshreg.so)
# ---------------------------------------------------------------------------------------------
def test_unary_costs_vs_reference_golden(R, oracle_built):
    from newmsm_b200 import discrete_cost as DC
    from cost_cases import GOLDEN_CP, GOLDEN_DATA, golden_digest
    g = load("costs.npz")
    for kind, D in ((0, 1), (1, 4), (2, 4)):
        s = cost_setup(oracle_built, GOLDEN_CP, GOLDEN_DATA, D)
        assert np.array_equal(golden_digest(s), g[f"unary_k{kind}_digest"]), "seeded inputs drifted: regenerate the fixture"
        cls = [DC.UnivariateNonLinearSRegDiscreteCostFunction, DC.MultivariateNonLinearSRegDiscreteCostFunction,
               DC.PatchwiseMultivariateNonLinearSRegDiscreteCostFunction][kind]
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D if kind == 1 else 1, len(s["src"])))
        for sim in (1, 2):
            cf = cls(simmeasure=sim)
            cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
            cf.reset_CPgrid(s["cp"], s["maxsep"], 1.0, HIGHREScfweight=cfw, AbsoluteWeights=s["absw"])
            prow, pmem = cf.get_source_data()
            assert np.array_equal(prow, g[f"unary_k{kind}_prow"]) and np.array_equal(pmem, g[f"unary_k{kind}_pmem"])
            assert np.array_equal(cf.computeUnaryCosts(s["labels"], s["rot"]), g[f"unary_k{kind}_s{sim}"])
        for sim, pct in ((4, 0.75), (5, 0.6)):     # DICE / genDICE (similarities.cpp:201-253)
            cf = cls(simmeasure=sim, percentile=pct)
            cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
            cf.reset_CPgrid(s["cp"], s["maxsep"], 1.0, HIGHREScfweight=cfw, AbsoluteWeights=s["absw"])
            cf.get_source_data()
            assert np.array_equal(cf.computeUnaryCosts(s["labels"], s["rot"]), g[f"unary_k{kind}_s{sim}"])


def test_triplet_costs_vs_reference_golden(R, oracle_built):
    from newmsm_b200 import discrete_cost as DC
    from cost_cases import GOLDEN_CP, GOLDEN_DATA, golden_digest
    g = load("costs.npz")
    for kind, D in ((0, 1), (3, 1), (4, 3)):
        s = triplet_setup(oracle_built, GOLDEN_CP, GOLDEN_DATA, D)
        assert np.array_equal(golden_digest(s), g[f"triplet_k{kind}_digest"]), "seeded inputs drifted: regenerate the fixture"
        rt, la, lb, lc = s["req"]
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
        cls = {0: DC.UnivariateNonLinearSRegDiscreteCostFunction, 3: DC.HOUnivariateNonLinearSRegDiscreteCostFunction,
               4: DC.HOMultivariateNonLinearSRegDiscreteCostFunction}[kind]
        cf = cls(simmeasure=2)
        cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
        if kind >= 3:
            cf.reset_CPgrid(s["cp_now"], s["cp_tri"], HIGHREScfweight=cfw, AbsoluteWeights=s["absw"])
            prow, pmem = cf.get_source_data()
            assert np.array_equal(prow, g[f"triplet_k{kind}_prow"]) and np.array_equal(pmem, g[f"triplet_k{kind}_pmem"])
        else:
            cf.reset_CPgrid(s["cp_now"], s["maxsep"], 1.0)
        cf.set_parameters(0.05)
        cf.setTriplets(s["triplets"], s["labels"], s["rot_now"], s["orig"])
        got = cf.computeTripletCostList(rt, la, lb, lc)
        ref = g[f"triplet_k{kind}"]
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind,D,depth", [(0, 1, 1), (0, 1, 2), (3, 1, 2)])
def test_anatomical_strain_vs_reference_golden(R, oracle_built, kind, D, depth):
    """regoption 5 (anatomical strain, DiscreteCostFunction.cpp:169-181, 245-301): list form vs the reference's own outputs, and Fusion's
    8-combination batch vs the list form."""
    from newmsm_b200 import discrete_cost as DC
    from cost_cases import GOLDEN_CP, GOLDEN_DATA, anat_case, golden_digest
    g = load("anat.npz")
    s = triplet_setup(oracle_built, GOLDEN_CP, GOLDEN_DATA, D)
    a = anat_case(oracle_built, s, depth)
    assert np.array_equal(np.concatenate([golden_digest(s), golden_digest(a)]), g[f"k{kind}_d{depth}_digest"]), "seeded inputs drifted: regenerate"
    rt, la, lb, lc = s["req"]
    cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
    cls = {0: DC.UnivariateNonLinearSRegDiscreteCostFunction, 3: DC.HOUnivariateNonLinearSRegDiscreteCostFunction}[kind]
    cf = cls(simmeasure=2)
    cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
    if kind >= 3:
        cf.reset_CPgrid(s["cp_now"], s["cp_tri"], HIGHREScfweight=cfw, AbsoluteWeights=s["absw"])
    else:
        cf.reset_CPgrid(s["cp_now"], s["maxsep"], 1.0)
    cf.set_parameters(0.05, regularisermode=5)
    cf.setTriplets(s["triplets"], s["labels"], s["rot_now"], s["orig"])
    with pytest.raises(RuntimeError):
        cf.computeTripletCostList(rt, la, lb, lc)            # regoption 4/5 before set_anatomical: an error, not a silent spherical strain
    cf.set_anatomical(**a)
    got = cf.computeTripletCostList(rt, la, lb, lc)
    assert np.array_equal(got, g[f"k{kind}_d{depth}"])
    labeling = np.random.default_rng(3).integers(0, len(s["labels"]), len(s["cp"])).astype(np.int32)
    batch = cf.computeTripletCostsForLabel(labeling, 2)
    T = len(s["triplets"])
    tt = np.repeat(np.arange(T, dtype=np.int32), 8)
    combo = np.tile(np.arange(8), T)
    pick = lambda k, bit: np.where((combo >> bit) & 1, 2, labeling[s["triplets"][tt, k]]).astype(np.int32)
    assert np.array_equal(batch.reshape(-1), cf.computeTripletCostList(tt, pick(0, 2), pick(1, 1), pick(2, 0)))


@pytest.mark.parametrize("kexp,rexp", [(1.5, 1.3), (2.0, 1.0)])
def test_anatomical_strain_vs_oracle(R, oracle_built, kexp, rexp):
    """The same at a larger grid (control ico3, anatomical ico5) and non-default exponents vs the oracle (pinned in test_oracle_vs_refmr.py)."""
    from newmsm_b200 import discrete_cost as DC
    from cost_cases import anat_case
    s = triplet_setup(oracle_built, 3, 5, 1)
    a = anat_case(oracle_built, s, 2)
    rt, la, lb, lc = s["req"]
    cf = DC.UnivariateNonLinearSRegDiscreteCostFunction()
    cf.set_meshes(R.Mesh(s["xyz"], s["tri"]), s["src"], s["src_feat"], s["ref_feat"])
    cf.reset_CPgrid(s["cp_now"], s["maxsep"], 1.0)
    cf.set_parameters(0.1, 0.4, 1.6, kexp, rexp, 4)
    cf.setTriplets(s["triplets"], s["labels"], s["rot_now"], s["orig"])
    cf.set_anatomical(**a)
    got = cf.computeTripletCostList(rt, la, lb, lc)
    ref = oracle_built.oracle_triplet_costs(0, 2, None, s["cp_now"], s["orig"], s["rot_now"], s["labels"], s["triplets"], rt, la, lb, lc,
                                            s["src"], None, None, s["src_feat"], s["ref_feat"], None, np.ones(len(s["cp"])), 0.1, 0.4, 1.6, kexp, rexp,
                                            rmode=4, anat=a)
    assert np.array_equal(got, ref)


def test_group_costs_vs_reference_golden(R, oracle_built):
    from newmsm_b200 import group_cost as GC
    from cost_cases import golden_digest, golden_group_glue, group_triplet_case
    g = load("costs.npz")
    c = group_setup(S=2, cp_level=1, data_level=3, tpl_level=3, D=2)
    # group triplet term (DiscreteGroupCostFunction.cpp:26-52): list form vs the reference's outputs, and Fusion's 8-combination batch
    orig, trip, rot_t, (rt, ta, tb, tc) = group_triplet_case(oracle_built, c)
    Mt = GC.DiscreteGroupModel(R.Mesh(c["tpl"], c["tpl_tri"]))
    assert np.array_equal(Mt.computeTripletCostList(c["cps"], orig, rot_t, c["labels"], trip, rt, ta, tb, tc, 0.05), g["group_triplet"])
    labeling = np.random.default_rng(3).integers(0, len(c["labels"]), len(rot_t)).astype(np.int32)
    batch = Mt.computeTripletCostsForLabel(c["cps"], orig, rot_t, c["labels"], trip, labeling, 2, 0.05)
    T = len(trip)
    tt = np.repeat(np.arange(T, dtype=np.int32), 8); combo = np.tile(np.arange(8), T)
    pick = lambda k, bit: np.where((combo >> bit) & 1, 2, labeling[trip[tt, k]]).astype(np.int32)
    assert np.array_equal(batch.reshape(-1), Mt.computeTripletCostList(c["cps"], orig, rot_t, c["labels"], trip, tt, pick(0, 2), pick(1, 1), pick(2, 0), 0.05))
    # the resident plan with the costs left on the device (msmgpu_triplet_plan_batch_dev: the sharded hosts' data path) gives the same table
    Mt.reset_triplet_state(c["cps"], orig, rot_t, c["labels"], trip)
    assert np.array_equal(Mt.computeTripletCostsForLabel(None, None, None, None, None, labeling, 2, 0.05), batch)
    assert np.array_equal(Mt.computeTripletCostsForLabel(None, None, None, None, None, labeling, 2, 0.05, copy=False), batch)
    assert np.array_equal(golden_digest(c), g["group_digest"]), "seeded inputs drifted: regenerate the fixture"
    rot, spacings, pairs, (rp, la, lb) = golden_group_glue(oracle_built, c)
    for sim in (1, 2):
        M = GC.DiscreteGroupModel(R.Mesh(c["tpl"], c["tpl_tri"]), simmeasure=sim)
        fields = M.get_patch_data(c["data"], c["dtri"], c["feat"], c["labels"], c["centre"], rot, spacings, 1.0)
        got_fields = fields.cpu().numpy().transpose(0, 1, 3, 2)
        seen = ~np.isnan(g["group_fields"])
        assert np.array_equal(got_fields[seen], g["group_fields"][seen])
        got = M.computePairwiseCostList(pairs, rp, la, lb)
        ok = ~np.isnan(got)
        assert ok.mean() > 0.9 and np.array_equal(got[ok], g[f"group_pair_s{sim}"][ok])
        from cost_cases import group_mask
        M.set_masks(group_mask(c))                     # the reference's own class with set_masks (DiscreteGroupModel.cpp:164)
        assert np.array_equal(M.computePairwiseCostList(pairs, rp, la, lb)[ok], g[f"group_pair_masked_s{sim}"][ok])


def test_group_triplet_nan_energy_propagates(R, oracle_built):
    """A strain energy that comes out NaN (det(F^T F) rounding below zero, reg_tools.cpp:585-587; seen 8 times in 7.6 M costs of a groupwise
    run) must stay NaN like in the reference, not be mistaken for the 'folded' marker. Forced here with a NaN label vector."""
    from newmsm_b200 import group_cost as GC
    from cost_cases import group_triplet_case
    c = group_setup(S=2, cp_level=1, data_level=3, tpl_level=3, D=2)
    orig, trip, rot_t, (rt, ta, tb, tc) = group_triplet_case(oracle_built, c, n=400)
    labels = np.array(c["labels"], dtype=np.float64, copy=True)
    labels[3] = np.nan
    Mt = GC.DiscreteGroupModel(R.Mesh(c["tpl"], c["tpl_tri"]))
    got = Mt.computeTripletCostList(c["cps"], orig, rot_t, labels, trip, rt, ta, tb, tc, 0.05)
    ref = oracle_built.oracle_group_triplet_costs(c["cps"], orig, rot_t, labels, trip, rt, ta, tb, tc, 0.05)
    touched = (ta == 3) | (tb == 3) | (tc == 3)
    assert touched.any() and np.isnan(ref[touched]).all()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~touched], ref[~touched])
    got_fix = Mt.computeTripletCostList(c["cps"], orig, rot_t, labels, trip, rt, ta, tb, tc, 0.05, fixnan=True)
    assert np.all(got_fix[touched] == 1e7) and np.array_equal(got_fix[~touched], ref[~touched])


def test_view_meshes_match_copied_meshes(R, oracle_built, meshes):
    """msmgpu_mesh_create_view_batch: meshes that view the caller's device buffers (no copies, one table allocation) give the same forest
    and the same adaptive CSR as uploaded meshes, i.e. the oracle's; set_coords is refused on a view."""
    import ctypes as C
    import torch
    L = capi.lib()
    ctx = R.Context(0)
    dev = torch.device("cuda", 0)
    low_xyz, low_tri = synth.rotate_sphere(meshes[4][0], 0.01, 0.02, -0.015), meshes[4][1]
    S = 3
    xyz = [synth.jitter_sphere(meshes[5][0], meshes[5][1], seed=90 + s) for s in range(S)]
    tri = np.ascontiguousarray(meshes[5][1], dtype=np.int32)
    d_xyz = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in xyz]
    d_tri = torch.from_numpy(tri).to(dev)
    torch.cuda.synchronize()
    views = R.Mesh.views_from_device(ctx, len(xyz[0]), d_xyz, len(tri), d_tri)
    low = R.Mesh(low_xyz, low_tri, ctx=ctx)
    trees = R.Octree.build_batch(views + [low])
    for s in range(S):
        ref = R.Octree(R.Mesh(xyz[s], tri, ctx=ctx))
        for a, b in zip(trees[s].dump(), ref.dump()):
            assert np.array_equal(a, b)
    tp = (C.c_void_p * S)(*[t.h.value for t in trees[:S]])
    mp = (C.c_void_p * S)(*[m.h.value for m in views])
    w = (C.c_void_p * S)()
    capi.check(L.msmgpu_adaptive_weights_batch(ctx.h, S, mp, tp, low.h, trees[-1].h, w))
    for s in range(S):
        W = R.Weights(L, C.c_void_p(w[s]))
        for x, y in zip(W.csr(), oracle_built.oracle_adaptive_weights(xyz[s], tri, low_xyz, low_tri)):
            assert np.array_equal(x, y)
        W.close()
    assert L.msmgpu_mesh_set_coords(views[0].h, capi.ptr(np.ascontiguousarray(xyz[1]))) == capi.ERR_INVALID


def test_full_size_properties(R):
    """BASELINE configs[1] at full size (ico7 163 842 V -> 32 492-vertex geodesic sphere, 100 FP32 channels, 2 subjects), where the CPU
    oracle would take minutes: size-independent properties instead. (a) the fused barycentric resample equals an independent
    evaluation of its own weight maps (msmgpu_bary_weights, FP64 on the host); (b) adaptive CSR rows are ascending, non-negative and
    sum to 1; (c) constants are reproduced by both methods; (d) lane-group widths 1 and 8 give identical triangle ids; (e) the
    batched path equals the per-subject path bit for bit."""
    import ctypes as C
    import torch
    L = capi.lib()
    ctx = R.Context(0)
    dev = torch.device("cuda", 0)
    xyz7, tri7 = synth.icosphere(7)
    low_xyz, low_tri = synth.geodesic_sphere(57)
    assert len(xyz7) == 163842 and len(low_xyz) == 32492
    S, D = 2, 100
    subj = [synth.jitter_sphere(xyz7, tri7, frac=0.3, seed=1234 + s) for s in range(S)]
    rng = np.random.default_rng(5)
    feats = [rng.standard_normal((len(xyz7), D)).astype(np.float32) for _ in range(S)]
    for f in feats: f[:, 0] = 2.5                                   # channel 0: a constant
    src = [R.Mesh(x, tri7, ctx=ctx) for x in subj]
    low = R.Mesh(low_xyz, low_tri, ctx=ctx)
    trees = R.Octree.build_batch(src + [low])
    n_low = len(low_xyz)
    d_feat = [torch.from_numpy(f).to(dev) for f in feats]
    d_out = [torch.empty(n_low, D, device=dev) for _ in range(S)]
    d_low = torch.from_numpy(low_xyz).to(dev)
    tp = (C.c_void_p * S)(*[t.h.value for t in trees[:S]])
    fp = (C.c_void_p * S)(*[t.data_ptr() for t in d_feat])
    op = (C.c_void_p * S)(*[t.data_ptr() for t in d_out])
    capi.check(L.msmgpu_bary_resample_batch_f32_dev(ctx.h, S, tp, n_low, d_low.data_ptr(), D, fp, op, None))
    ctx.sync()
    for s in range(S):
        idx, w, ne = np.zeros((n_low, 3), np.int32), np.zeros((n_low, 3)), np.zeros(n_low, np.int32)
        capi.check(L.msmgpu_bary_weights(trees[s].h, n_low, capi.ptr(low_xyz), capi.ptr(idx), capi.ptr(w), capi.ptr(ne)))
        assert (ne == 3).all() and (w >= 0).all() and np.abs(w.sum(1) - 1).max() < 1e-12 and (np.diff(idx, axis=1) > 0).all()
        acc = np.zeros((n_low, D))
        for j in range(3):                                          # ascending vertex id = the reference's map order
            acc += feats[s][idx[:, j]].astype(np.float64) * w[:, j:j + 1]
        got = d_out[s].cpu().numpy()
        assert np.array_equal(got, acc.astype(np.float32))          # (a)
        assert np.abs(got[:, 0] - 2.5).max() < 1e-6                 # (c)
        # (e) per-subject fused call
        one = torch.empty(n_low, D, device=dev)
        capi.check(L.msmgpu_bary_resample_f32_dev(trees[s].h, n_low, d_low.data_ptr(), D, d_feat[s].data_ptr(), one.data_ptr(), None))
        ctx.sync()      # the library's stream is not torch's current stream
        assert torch.equal(one, d_out[s])
    # (d)
    ids1 = trees[0].get_closest_triangle(low_xyz)
    capi.check(L.msmgpu_set_query_group(8))
    try:
        ids8 = trees[0].get_closest_triangle(low_xyz)
    finally:
        capi.check(L.msmgpu_set_query_group(1))
    assert np.array_equal(ids1, ids8)
    # (b), (c) adaptive
    mp = (C.c_void_p * S)(*[m.h.value for m in src])
    wp = (C.c_void_p * S)()
    capi.check(L.msmgpu_adaptive_weights_batch(ctx.h, S, mp, tp, low.h, trees[-1].h, wp))
    d_oa = [torch.empty(n_low, D, device=dev) for _ in range(S)]
    oa = (C.c_void_p * S)(*[t.data_ptr() for t in d_oa])
    capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, wp, D, fp, oa))
    ctx.sync()
    for s in range(S):
        W = R.Weights(L, C.c_void_p(wp[s]))
        rowptr, col, val = W.csr()
        assert rowptr[0] == 0 and (np.diff(rowptr) >= 3).all() and (val >= 0).all()
        rows = np.repeat(np.arange(n_low), np.diff(rowptr))
        assert np.abs(np.bincount(rows, weights=val, minlength=n_low) - 1).max() < 1e-12
        inner = np.ones(len(col), bool); inner[rowptr[1:-1]] = False            # first entry of every row but the first
        assert (np.diff(col)[inner[1:]] > 0).all()                               # ascending columns inside each row
        assert col.min() >= 0 and col.max() < len(xyz7) and len(np.unique(col)) > 0.9 * len(xyz7)   # (nearly) every source vertex contributes
        got = d_oa[s].cpu().numpy()
        assert np.abs(got[:, 0] - 2.5).max() < 1e-6
        # independent evaluation of the CSR product (FP64, column order) on a sample of rows
        for r in rng.integers(0, n_low, 200):
            b, e = rowptr[r], rowptr[r + 1]
            acc = np.zeros(D)
            for k in range(b, e):
                acc += feats[s][col[k]].astype(np.float64) * val[k]
            assert np.array_equal(got[r], acc.astype(np.float32))
        W.close()


def test_device_buffer_functions_and_group_batch_dev(R, oracle_built):
    """msmgpu_device_malloc / _download / _copy_peer round trip; msmgpu_group_pair_batch_dev needs msmgpu_group_set_pairs first and then
    equals the host-output batch on any block of pairs."""
    import ctypes as C
    import torch
    from newmsm_b200 import group_cost as GC
    L = capi.lib()
    ctx = R.Context(0)
    a, b = C.c_void_p(), C.c_void_p()
    capi.check(L.msmgpu_device_malloc(ctx.h, 4096, C.byref(a)))
    capi.check(L.msmgpu_device_malloc(ctx.h, 4096, C.byref(b)))
    src = torch.arange(512, dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    capi.check(L.msmgpu_device_copy_peer(ctx.h, a, ctx.h, C.c_void_p(src.data_ptr()), 4096))
    capi.check(L.msmgpu_device_copy_peer(ctx.h, b, ctx.h, a, 4096))
    host = np.zeros(512)
    capi.check(L.msmgpu_device_download(ctx.h, capi.ptr(host), b, 4096))
    assert np.array_equal(host, np.arange(512.0))
    L.msmgpu_device_free(ctx.h, a); L.msmgpu_device_free(ctx.h, b)
    assert L.msmgpu_device_malloc(None, 16, C.byref(a)) == capi.ERR_INVALID

    g = group_setup()
    M = GC.DiscreteGroupModel(R.Mesh(g["tpl"], g["tpl_tri"]), simmeasure=2)
    ncp = g["cps"].shape[1]
    rot = M.get_rotations(g["centre"], g["cps"])
    spacings = M.get_spacings(g["cps"], g["cp_tri"])
    pairs = M.estimate_pairs(g["cps"], g["cp_tri"])
    M.get_patch_data(g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], rot, spacings, 1.0)
    labeling = np.random.default_rng(11).integers(0, len(g["labels"]), g["cps"].shape[0] * ncp).astype(np.int32)
    P = len(pairs)
    d_out = torch.empty((P, 4), dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    assert L.msmgpu_group_pair_batch_dev(M.g, 0, P, capi.ptr(labeling), 2, capi.ptr(d_out)) == capi.ERR_INVALID    # no resident pairs yet
    M.setPairs(pairs)
    ref = np.zeros((P, 4))
    capi.check(L.msmgpu_group_pair_batch(M.g, P, capi.ptr(np.ascontiguousarray(pairs)), capi.ptr(labeling), 2, capi.ptr(ref)))
    lo, hi = P // 3, P - 5
    capi.check(L.msmgpu_group_pair_batch_dev(M.g, lo, hi - lo, capi.ptr(labeling), 2, capi.ptr(d_out[lo:hi])))
    M.ctx.sync()
    got = d_out[lo:hi].cpu().numpy()
    assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(ref[lo:hi], nan=-1.0))
    assert L.msmgpu_group_pair_batch_dev(M.g, P - 2, 5, capi.ptr(labeling), 2, capi.ptr(d_out)) == capi.ERR_INVALID   # block beyond the list
    full = M.computePairwiseCostsForLabel(pairs, labeling, 2)
    assert np.array_equal(np.nan_to_num(full, nan=-1.0), np.nan_to_num(ref, nan=-1.0))


def test_nan_feature_under_zero_weight_entry(R, oracle_built, meshes):
    """A target exactly on a mesh vertex has the weight map {v: 1, a: 0, b: 0}; a PRESENT entry with weight 0 still multiplies
    (NaN * 0 = NaN, resampler.cpp:46-48). Every path (two-kernel bulk gather, fused register kernel, scalar fallback) must agree."""
    xyz, tri = meshes[4]
    low = xyz[[5, 17, 300]].copy()
    for D in (32, 7):                                    # bulk-copy gather (rows >= 128 B) and the scalar path
        feat = synth.smooth_fields(xyz, D).astype(np.float32).astype(np.float64)
        ref0 = oracle_built.oracle_bary_resample(xyz, tri, low, feat)
        idx, w, ne, err = oracle_built.OracleOctree(xyz, tri).bary_weights(low)
        assert err == 0 and (w == 0).any(), "the case needs an entry with weight exactly 0"
        k, j = np.argwhere((w == 0) & (idx >= 0))[0]
        feat[:, idx[k, j]] = np.nan
        ref = oracle_built.oracle_bary_resample(xyz, tri, low, feat)
        assert np.isnan(ref[:, k]).all() and not np.isnan(ref0[:, k]).any()
        for mode in (1, 0):
            capi.check(capi.lib().msmgpu_set_tuning(b"gather", mode))
            try:
                got = R.barycentric_resample(R.Mesh(xyz, tri), low, feat)
            finally:
                capi.check(capi.lib().msmgpu_set_tuning(b"gather", 1))
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            ok = ~np.isnan(ref)
            assert np.array_equal(got[ok], ref.astype(np.float32).astype(np.float64)[ok])


def test_fallback2_corner_queries(R, oracle_built):
    """Queries in the corners of the root cube of a tiny mesh reach the fallbacks of get_closest_triangle (octree.cpp:180-208); corners
    farther than 2R make the reference's asin NaN and are skipped."""
    xyz, tri = synth.icosphere(1)
    c = 100.9
    q = np.array([[sx * c, sy * c, sz * c] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)] + [[0.0, 0.0, 0.0], [100.9, 0.0, 0.0], [-100.9, 100.9, 0.3]])
    t0, v0, s0, _ = oracle_built.OracleOctree(xyz, tri).query(q)
    t1, v1, s1 = R.Octree(R.Mesh(xyz, tri)).query(q)
    assert np.array_equal(s1, np.vectorize(ST.get)(s0)) and np.array_equal(t0, t1)


def test_smooth_neighbourhoods_rejects_foreign_ids(R, meshes):
    """closest[] must index the mesh whose neighbourhoods are searched (out-of-bounds device read otherwise)."""
    import ctypes as C
    xyz, _ = meshes[3]
    n = len(xyz)
    ctx = R.Context(0)
    closest = np.arange(n, dtype=np.int32)
    rowptr = np.zeros(n + 1, np.int32)
    L = capi.lib()
    assert L.msmgpu_smooth_neighbourhoods(ctx.h, n, capi.ptr(xyz), capi.ptr(closest), 0.99, capi.ptr(rowptr), 0, None, None) == capi.OK
    closest[7] = n + 5                                   # an id of a larger `orig` mesh
    assert L.msmgpu_smooth_neighbourhoods(ctx.h, n, capi.ptr(xyz), capi.ptr(closest), 0.99, capi.ptr(rowptr), 0, None, None) == capi.ERR_INVALID
    ctx.close()


def test_kept_maps_apply_matches_resample(R, meshes):
    """msmgpu_fwd_apply_batch_f32_dev (the bulk-copy gather alone) reproduces the resample that kept the maps, and applies them to
    another feature set like a fresh resample would."""
    import ctypes as C
    import torch
    xyz, tri = meshes[5]
    low = synth.rotate_sphere(meshes[4][0], 0.01, 0.02, -0.03)
    dev = torch.device("cuda", 0)
    S, D = 3, 36
    ctx = R.Context(0)
    L = capi.lib()
    subj = [synth.jitter_sphere(xyz, tri, frac=0.2, seed=3 + s) for s in range(S)]
    ms = [R.Mesh(x, tri, ctx=ctx) for x in subj]
    trees = R.Octree.build_batch(ms)
    d_low = torch.from_numpy(low).to(dev)
    f1 = [torch.randn(len(xyz), D, device=dev) for _ in range(S)]
    f2 = [torch.randn(len(xyz), D, device=dev) for _ in range(S)]
    o1 = [torch.empty(len(low), D, device=dev) for _ in range(S)]
    o2 = [torch.empty(len(low), D, device=dev) for _ in range(S)]
    o3 = [torch.empty(len(low), D, device=dev) for _ in range(S)]
    arr = lambda ts: (C.c_void_p * S)(*[t.data_ptr() for t in ts])
    tp = (C.c_void_p * S)(*[t.h.value for t in trees])
    fwd = C.c_void_p()
    capi.check(L.msmgpu_fwd_create(ctx.h, S, len(low), C.byref(fwd)))
    capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, len(low), d_low.data_ptr(), D, arr(f1), arr(o1), None, fwd))
    capi.check(L.msmgpu_fwd_apply_batch_f32_dev(ctx.h, fwd, D, arr(f1), arr(o2)))
    capi.check(L.msmgpu_fwd_apply_batch_f32_dev(ctx.h, fwd, D, arr(f2), arr(o3)))
    capi.check(L.msmgpu_bary_resample_batch_f32_dev(ctx.h, S, tp, len(low), d_low.data_ptr(), D, arr(f2), arr(o1), None))
    ctx.sync()
    # o1 now holds the fresh resample of f2, o3 the kept maps applied to f2; o2 the kept maps applied to f1
    assert all(torch.equal(a, b) for a, b in zip(o1, o3))
    ref = R.barycentric_resample(R.Mesh(subj[1], tri), low, f1[1].T.double().cpu().numpy())
    assert np.array_equal(o2[1].T.cpu().numpy().astype(np.float64), ref)
    L.msmgpu_fwd_destroy(fwd)
    # CSR rows through the bulk gather (opt-in: measured slower than the register path) give the same bits
    W = R.Resampler().get_adaptive_barycentric_weights_batch(ms, R.Mesh(low, meshes[4][1], ctx=ctx), trees)
    wp = (C.c_void_p * S)(*[w.h.value for w in W])
    capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, wp, D, arr(f1), arr(o1)))
    capi.check(L.msmgpu_set_tuning(b"gather_csr", 1))
    try:
        capi.check(L.msmgpu_weights_apply_batch_f32_dev(ctx.h, S, wp, D, arr(f1), arr(o2)))
    finally:
        capi.check(L.msmgpu_set_tuning(b"gather_csr", 0))
    ctx.sync()
    assert all(torch.equal(a, b) for a, b in zip(o1, o2))


def test_full_size_parity_vs_reference(R):
    """BASELINE configs[1] at full size (ico7 -> 32 492 vertices) against the reference's own outputs for the same subject: the check
    bench.py runs next to its timed region (bench.parity_block), here with 12 channels to keep the CPU side short."""
    import sys
    import torch
    from oracle import bindings as B
    if not B.have_ref():
        pytest.skip("compiled reference (oracle/_ref) not present")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    dev = torch.device("cuda", 0)
    xyz7, tri7 = synth.icosphere(7)
    xyz = synth.jitter_sphere(xyz7, tri7, frac=0.3, seed=1234)
    low_xyz, low_tri = synth.geodesic_sphere(57)
    D = 12
    ctx = R.Context(0)
    L = capi.lib()
    d_feat = torch.randn(len(xyz), D, device=dev)
    m = R.Mesh(xyz, tri7, ctx=ctx)
    low = R.Mesh(low_xyz, low_tri, ctx=ctx)
    tree, low_tree = R.Octree.build_batch([m, low])
    d_low = torch.from_numpy(low_xyz).to(dev)
    ob = torch.empty(len(low_xyz), D, device=dev)
    oa = torch.empty(len(low_xyz), D, device=dev)
    capi.check(L.msmgpu_bary_resample_f32_dev(tree.h, len(low_xyz), d_low.data_ptr(), D, d_feat.data_ptr(), ob.data_ptr(), None))
    W = R.Resampler().get_adaptive_barycentric_weights_batch([m], low, [tree], low_tree)[0]
    W.apply_f32_dev(D, d_feat, oa)
    ctx.sync()
    res = bench.parity_block(R, capi, L, xyz, tri7, low_xyz, low_tri, d_feat, ob, oa)
    assert res["ok"], res
    assert res["barycentric_f32_equals_rounded_reference"] and res["adaptive_f32_equals_rounded_reference"], res


@pytest.mark.parametrize("name,D,sim", [("d1_corr", 1, 2), ("d2_corr", 2, 2), ("d3_ssd", 3, 1)])
def test_rigid_level_vs_reference_golden(R, oracle_built, name, D, sim):
    """AFFINE / RIGID level (rigid_costfunction.cpp:32-236) on the device: the cost at zero rotation and the source rotated by `run`
    equal the reference's own (tests/golden/rigid.npz, produced by the compiled reference's Rigid_cost_function) bit for bit, and the
    oracle restatement on a second case."""
    import importlib.util
    from newmsm_b200 import rigid_cost as RC
    spec = importlib.util.spec_from_file_location("make_golden_rigid", os.path.join(G, "make_golden_rigid.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = load("rigid.npz")
    xyz, tri, src, mov, ref = mod.rigid_case(3, D)
    cf = RC.Rigid_cost_function(xyz, tri, src, tri, mov, ref)
    cf.set_parameters(iters=4, simmeasure=sim, stepsize=0.01, gradsampling=0.5)
    cf.initialise()
    assert cf.rigid_cost_mesh(0.0, 0.0, 0.0) == float(g[f"{name}_cost0"])
    moved = cf.run()
    assert np.array_equal(moved, g[f"{name}_xyz"])
    # a second, larger case against the oracle restatement (ico4)
    xyz, tri, src, mov, ref = mod.rigid_case(4, D)
    want_xyz, want_cost0, _, _ = oracle_built.oracle_rigid(xyz, tri, src, tri, mov, ref, simmeasure=sim, iters=2)
    cf = RC.Rigid_cost_function(xyz, tri, src, tri, mov, ref)
    cf.set_parameters(iters=2, simmeasure=sim, stepsize=0.01, gradsampling=0.5)
    cf.initialise()
    assert cf.rigid_cost_mesh(0.0, 0.0, 0.0) == want_cost0
    assert np.array_equal(cf.run(), want_xyz)


@pytest.mark.parametrize("name,sl,ll", [("down", 4, 3), ("up", 3, 4), ("same", 3, 3)])
def test_exclusion_masks_vs_reference_golden(R, oracle_built, name, sl, ll):
    """Exclusion masks on the device path (no CPU fallback): masked adaptive weights, metric_resample + the resampled
    mask, nearest-neighbour interpolation — bit-exact against the reference's own outputs (tests/golden/excl.npz) and the oracle."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_excl", os.path.join(G, "make_golden_excl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = load("excl.npz")
    xyz, tri, low, low_tri, feat, excl = mod.excl_case(sl, ll)
    m, ml = R.Mesh(xyz, tri, feat), R.Mesh(low, low_tri)
    out, eo = R.metric_resample(m, ml, EXCL=excl)
    assert np.array_equal(out, g[f"{name}_metric"]) and np.array_equal(eo, g[f"{name}_metric_excl"])
    rp, col, val = R.adaptive_weights_excl(m, ml, excl).csr()
    assert np.array_equal(rp, g[f"{name}_rowptr"]) and np.array_equal(col, g[f"{name}_col"]) and np.array_equal(val, g[f"{name}_val"])
    assert (np.diff(rp) == 0).any(), "the case must contain targets without a row"
    n, en = R.nearest_neighbour_interpolation(m, low, EXCL=excl)
    assert np.array_equal(n, g[f"{name}_nn"]) and np.array_equal(en, g[f"{name}_nn_excl"])
    # an all-ones mask changes nothing
    ones = np.ones(len(xyz))
    o1, e1 = R.metric_resample(m, ml, EXCL=ones)
    assert np.array_equal(o1, R.metric_resample(m, ml))
    assert np.abs(e1 - 1).max() < 1e-12


def test_smooth_data_with_exclusion_mask_golden(R, oracle_built):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_excl", os.path.join(G, "make_golden_excl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = load("excl.npz")
    xyz, tri, _, _, feat, excl = mod.excl_case(4, 3)
    m = R.Mesh(xyz, tri, feat)
    for sigma in (4.0, 9.0):
        assert np.array_equal(R.smooth_data(m, m, sigma), g[f"smooth{int(sigma)}"])
        s1, e1 = R.smooth_data(m, m, sigma, EXCL=excl)
        assert np.array_equal(s1, g[f"smooth{int(sigma)}_masked"]) and np.array_equal(e1, g[f"smooth{int(sigma)}_excl"])


@pytest.mark.parametrize("masked", [False, True])
def test_variance_normalise(R, oracle_built, masked):
    """newmeshreg::variance_normalise (reg_tools.cpp:804-844) on the device vs the oracle (pinned against the reference in
    tests/test_oracle_vs_refmr.py::test_variance_normalise), bit-exact."""
    from test_oracle_vs_refmr import variance_cases
    data, excl = variance_cases()
    e = excl if masked else None
    ctx = R.Context(0)
    got = R.variance_normalise(ctx, data, e)
    assert np.array_equal(got, oracle_built.oracle_variance_normalise(data, e), equal_nan=True)
    assert np.array_equal(R.variance_normalise(ctx, data[:, :1]), oracle_built.oracle_variance_normalise(data[:, :1]), equal_nan=True)
    big = np.random.default_rng(5).normal(size=(40, 40962))
    assert np.array_equal(R.variance_normalise(ctx, big), oracle_built.oracle_variance_normalise(big))


def test_mesh_keeps_its_octree_until_the_coordinates_change(R, oracle_built, meshes):
    """A mesh handed to a mesh-taking entry point (metric_resample, sphere_project_warp, nn interpolation) builds its octree once and
    keeps it (csrc: msm::mesh_trees); msmgpu_mesh_set_coords must drop it. Repeated calls, a coordinate change in between, the same
    mesh as source AND target, all against fresh meshes / the oracle."""
    xyz, tri = meshes[4]
    low, ltri = meshes[3]
    low = synth.rotate_sphere(low)
    feat = synth.smooth_fields(xyz, 3)
    m, ml = R.Mesh(xyz, tri, feat), R.Mesh(low, ltri)
    first = R.metric_resample(m, ml)
    assert np.array_equal(first, oracle_built.oracle_metric_resample(xyz, tri, low, ltri, feat))
    assert np.array_equal(R.metric_resample(m, ml), first)                                   # cached trees, same result
    warped = synth.smooth_warp(xyz, max_disp=3.0, seed=5)
    m.set_coords(warped)                                                                      # Mesh::set_coord for every vertex
    moved = R.metric_resample(m, ml)
    fresh = R.Mesh(warped, tri, feat)
    fresh_tree_ids = R.Octree(fresh).get_closest_triangle(low)
    # the triangle ids the kept-then-dropped tree answers with are those of a fresh tree over the new coordinates ...
    assert np.array_equal(R.nearest_neighbour_interpolation(m, low), R.nearest_neighbour_interpolation(fresh, low))
    assert np.array_equal(R.sphere_project_warp(low, m, xyz), R.sphere_project_warp(low, fresh, xyz))
    assert np.array_equal(R.Octree(m).get_closest_triangle(low), fresh_tree_ids)
    assert not np.array_equal(moved, first)
    # ... and a mesh may be its own target (one tree serves both sides)
    same = R.metric_resample(m, m)
    assert same.shape == feat.shape and np.all(np.isfinite(same))


def test_page_locked_host_buffers_are_plain_inputs(R, meshes):
    """msmgpu_host_alloc: the adapter flattens Mesh::pvalues into such buffers; results do not depend on where the host bytes live."""
    import ctypes as C
    xyz, tri = meshes[4]
    low, ltri = meshes[3]
    feat = np.ascontiguousarray(synth.smooth_fields(xyz, 4))
    m, ml = R.Mesh(xyz, tri), R.Mesh(synth.rotate_sphere(low), ltri)
    L = capi.lib()
    want = np.zeros((4, len(low)))
    capi.check(L.msmgpu_metric_resample(m.h, ml.h, 4, capi.ptr(feat), capi.ptr(want)))
    pin_in, pin_out = C.c_void_p(), C.c_void_p()
    capi.check(L.msmgpu_host_alloc(m.ctx.h, feat.nbytes, C.byref(pin_in)))
    capi.check(L.msmgpu_host_alloc(m.ctx.h, want.nbytes, C.byref(pin_out)))
    try:
        C.memmove(pin_in, feat.ctypes.data, feat.nbytes)
        capi.check(L.msmgpu_metric_resample(m.h, ml.h, 4, pin_in, pin_out))
        got = np.frombuffer((C.c_char * want.nbytes).from_address(pin_out.value), dtype=np.float64).reshape(want.shape).copy()
    finally:
        L.msmgpu_host_free(m.ctx.h, pin_in)
        L.msmgpu_host_free(m.ctx.h, pin_out)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("which", ["both", "bary", "adaptive"])
def test_batch_host_call_matches_per_subject_calls(R, meshes, which):
    """msmgpu_resample_batch_host_f32 (chunks of subjects pipelined over three streams) against the per-subject host calls
    msmgpu_mesh_set_features_f32 + msmgpu_mesh_bary_resample_f32 / msmgpu_mesh_metric_resample_f32: bit-identical outputs, also when
    the subject count is not a multiple of the chunk and the stage buffers are reused (7 subjects, chunks of 2, 3 stages)."""
    import ctypes as C
    xyz0, tri = meshes[5]
    low, ltri = meshes[3]
    low = synth.rotate_sphere(low)
    S, D = 7, 12
    xyzs = [synth.jitter_sphere(xyz0, tri, frac=0.3, seed=40 + s) for s in range(S)]
    feats = [np.ascontiguousarray(synth.smooth_fields(x, D, seed0=60 + s).astype(np.float32)) for s, x in enumerate(xyzs)]
    ctx = R.Context(0)
    L = capi.lib()
    want_b, want_a = [], []
    ml = R.Mesh(low, ltri, ctx=ctx)
    tl = R.Octree(ml)
    for s in range(S):
        m = R.Mesh(xyzs[s], tri, ctx=ctx)
        capi.check(L.msmgpu_mesh_set_features_f32(m.h, D, capi.ptr(feats[s])))
        t = R.Octree(m)
        ob, oa = np.zeros((D, len(low)), np.float32), np.zeros((D, len(low)), np.float32)
        capi.check(L.msmgpu_mesh_bary_resample_f32(t.h, len(low), capi.ptr(low), capi.ptr(ob)))
        capi.check(L.msmgpu_mesh_metric_resample_f32(m.h, t.h, ml.h, tl.h, capi.ptr(oa)))
        want_b.append(ob); want_a.append(oa)
    got_b = [np.full((D, len(low)), np.nan, np.float32) for _ in range(S)] if which != "adaptive" else None
    got_a = [np.full((D, len(low)), np.nan, np.float32) for _ in range(S)] if which != "bary" else None
    R.resample_batch_host(ctx, xyzs, tri, low, ltri, feats, got_b, got_a, chunk=2)
    for s in range(S):
        if got_b is not None: assert np.array_equal(got_b[s], want_b[s]), f"barycentric, subject {s}"
        if got_a is not None: assert np.array_equal(got_a[s], want_a[s]), f"adaptive, subject {s}"
    # default chunk, one more call on the same context (streams / events are per call)
    R.resample_batch_host(ctx, xyzs[:3], tri, low, ltri, feats[:3], got_b[:3] if got_b else None, got_a[:3] if got_a else None)
    if got_b is not None: assert np.array_equal(got_b[2], want_b[2])
    if got_a is not None: assert np.array_equal(got_a[2], want_a[2])


@pytest.mark.parametrize("top", [1, 2, 3, 4, 5, 6, -1])
def test_octree_top_phase_matches_level_passes(R, oracle_built, meshes, top):
    """The one-pass construction of the first levels (octree_build.cu: k_top_count ... k_top_sort_*, knob "build_top") against the
    exact level passes from the root (knob 0) and the oracle: identical topology and identical leaf lists for every forced depth,
    single meshes and a mixed batch (per-mesh depths differ in the auto mode), a jittered mesh, and a mesh too small to split."""
    L = capi.lib()
    keys = [2, 4, "j5", 5, 6]
    try:
        capi.check(L.msmgpu_set_tuning(b"build_top", 0))
        capi.check(L.msmgpu_set_tuning(b"build_fused_levels", 0))      # the chunked level passes alone ...
        want = {k: R.Octree(R.Mesh(*meshes[k])).dump() for k in keys}
        capi.check(L.msmgpu_set_tuning(b"build_fused_levels", 1))      # ... then also the one-kernel pass for levels with short lists
        for k in keys:
            got = R.Octree(R.Mesh(*meshes[k])).dump()
            assert all(np.array_equal(a, b) for a, b in zip(want[k], got)), f"fused level passes, mesh {k}"
        for k in (3, "j5"):
            ref = oracle_built.OracleOctree(*meshes[k]).dump()
            got = want[k] if k in want else R.Octree(R.Mesh(*meshes[k])).dump()
            assert all(np.array_equal(a, b) for a, b in zip(ref, got))
        capi.check(L.msmgpu_set_tuning(b"build_top", top))
        for k in keys:
            got = R.Octree(R.Mesh(*meshes[k])).dump()
            assert all(np.array_equal(a, b) for a, b in zip(want[k], got)), f"single mesh {k}"
        tiny = (np.array([[100.0, 0, 0], [0, 100, 0], [0, 0, 100], [-100, 0, 0]]), np.array([[0, 1, 2], [1, 2, 3]], np.int32))
        ms = [R.Mesh(*meshes[k]) for k in keys] + [R.Mesh(*tiny)]
        trees = R.Octree.build_batch(ms)
        for k, t in zip(keys, trees):
            assert all(np.array_equal(a, b) for a, b in zip(want[k], t.dump())), f"batch, mesh {k}"
            q = synth.rotate_sphere(meshes[k][0])
            assert np.array_equal(t.get_closest_triangle(q), oracle_built.OracleOctree(*meshes[k]).query(q)[0])
        kinds, counts, tris = trees[-1].dump()
        assert list(kinds) == [1] and list(counts) == [2] and list(tris) == [0, 1]
    finally:
        capi.check(L.msmgpu_set_tuning(b"build_top", -1))
        capi.check(L.msmgpu_set_tuning(b"build_fused_levels", 1))
        capi.check(L.msmgpu_set_tuning(b"build_top_order", 0))
        capi.check(L.msmgpu_set_tuning(b"lazy_records", 1))


@pytest.mark.parametrize("top", [-1, 3, 6])
def test_octree_top_phase_shared_topology_order(R, oracle_built, meshes, top):
    """Five subjects that share one topology (device views, as a batch job creates them): the forests of the top phase must equal the
    exact level passes (knob 0) and the oracle — also with the optional Morton processing order of the triangles, which shared
    topologies enable (knob build_top_order) — and the batched resample in those trees (records deferred: record values built from
    the corners) must give the values of the stored-record path bit for bit."""
    import torch
    L = capi.lib()
    ctx = R.Context(0)
    dev = torch.device("cuda", 0)
    S = 5
    tri = np.ascontiguousarray(meshes[5][1], dtype=np.int32)
    xyz = [synth.jitter_sphere(meshes[5][0], tri, frac=0.3, seed=70 + s) for s in range(S)]
    d_xyz = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in xyz]
    d_tri = torch.from_numpy(tri).to(dev)
    torch.cuda.synchronize()
    q = np.concatenate([synth.rotate_sphere(meshes[4][0], 0.02, -0.01, 0.03), adversarial_points(xyz[0], tri)[:2000]])
    try:
        capi.check(L.msmgpu_set_tuning(b"build_top", 0))
        want = [R.Octree(R.Mesh(x, tri, ctx=ctx)) for x in xyz]
        want_dump = [t.dump() for t in want]
        assert all(np.array_equal(a, b) for a, b in zip(want_dump[0], oracle_built.OracleOctree(xyz[0], tri).dump()))
        capi.check(L.msmgpu_set_tuning(b"build_top", top))
        for order in (1, 0):
            capi.check(L.msmgpu_set_tuning(b"build_top_order", order))
            views = R.Mesh.views_from_device(ctx, len(xyz[0]), d_xyz, len(tri), d_tri)
            trees = R.Octree.build_batch(views)
            for s in range(S):
                assert all(np.array_equal(a, b) for a, b in zip(want_dump[s], trees[s].dump())), f"subject {s}, order {order}"
        # batched resample in the view trees with deferred records (values built from the corners) against the same call on views that
        # store their records, and the statuses against the per-mesh queries
        d_q = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
        n = len(q)
        rng = np.random.default_rng(5)
        d_feat = [torch.from_numpy(rng.normal(size=(len(xyz[0]), 4)).astype(np.float32)).to(dev) for _ in range(S)]

        def run(tree_list):
            d_out = [torch.zeros((n, 4), dtype=torch.float32, device=dev) for _ in range(S)]
            d_st = torch.zeros(S * n, dtype=torch.int32, device=dev)
            fwd = capi.C.c_void_p()
            capi.check(L.msmgpu_fwd_create(ctx.h, S, n, capi.C.byref(fwd)))
            tp = (capi.C.c_void_p * S)(*[t.h.value for t in tree_list])
            fp = (capi.C.c_void_p * S)(*[f.data_ptr() for f in d_feat])
            op = (capi.C.c_void_p * S)(*[o.data_ptr() for o in d_out])
            capi.check(L.msmgpu_bary_resample_batch_f32_dev_keep(ctx.h, S, tp, n, d_q.data_ptr(), 4, fp, op, d_st.data_ptr(), fwd))
            capi.check(L.msmgpu_ctx_sync(ctx.h))
            L.msmgpu_fwd_destroy(fwd)
            return [o.cpu().numpy() for o in d_out], d_st.cpu().numpy().reshape(S, n)
        got, st = run(trees)
        capi.check(L.msmgpu_set_tuning(b"lazy_records", 0))
        views2 = R.Mesh.views_from_device(ctx, len(xyz[0]), d_xyz, len(tri), d_tri)
        ref, st2 = run(R.Octree.build_batch(views2))
        assert np.array_equal(st, st2)
        for s in range(S):
            ok = st[s] == 0
            assert np.array_equal(got[s][ok], ref[s][ok]), f"resampled values, subject {s}"
            assert np.array_equal(st[s], want[s].query(q)[2]), f"statuses, subject {s}"
    finally:
        capi.check(L.msmgpu_set_tuning(b"build_top", -1))
        capi.check(L.msmgpu_set_tuning(b"build_top_order", 0))
        capi.check(L.msmgpu_set_tuning(b"lazy_records", 1))
