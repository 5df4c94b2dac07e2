"""Seeded input cases for the cost-function parity tests, shared by the CPU tests that pin the oracle against the
compiled reference (tests/test_oracle_vs_refmr.py), the golden-fixture generator (tests/golden/make_golden_costs.py)
and the GPU parity tests (tests/test_gpu_parity.py). `O` is oracle.bindings (only oracle_rotation_matrix is used:
get_rotations, DiscreteModel.cpp:310)."""
import numpy as np

from newmsm_b200 import synth


def cost_setup(O, cp_level, data_level, D, seed=3):
    cp, _ = synth.icosphere(cp_level)
    xyz, tri = synth.icosphere(data_level)
    src = synth.smooth_warp(xyz, max_disp=1.5, seed=seed)                 # SOURCE mesh = warped data grid
    ref_feat = synth.smooth_fields(xyz, D, seed0=100)
    src_feat = synth.smooth_fields(src, D, seed0=100, noise=0.05)
    # MAXSEP: largest distance from a CP to its mesh neighbours ~ CP spacing; any positive vector is a valid input
    cp_tri = synth.icosphere(cp_level)[1]
    e = np.zeros(len(cp))
    for a, b in ((0, 1), (1, 2), (0, 2)):
        d = np.linalg.norm(cp[cp_tri[:, a]] - cp[cp_tri[:, b]], axis=1)
        np.maximum.at(e, cp_tri[:, a], d)
        np.maximum.at(e, cp_tri[:, b], d)
    rng = np.random.default_rng(seed)
    # 7 labels: the CP itself plus 6 small displacements, expressed around the north pole like the label grid
    centre = np.array([0.0, 0.0, 100.0])
    labels = [centre]
    for k in range(6):
        ang = k * np.pi / 3
        p = centre + 0.4 * e.mean() * np.array([np.cos(ang), np.sin(ang), 0.0])
        labels.append(p / np.linalg.norm(p) * 100)
    labels = np.array(labels)
    rot = np.array([O.oracle_rotation_matrix(centre, c) for c in cp]).reshape(-1, 9)   # get_rotations (DiscreteModel.cpp:310)
    absw = rng.uniform(0.5, 1.5, size=len(cp))
    return dict(cp=cp, xyz=xyz, tri=tri, src=src, ref_feat=ref_feat, src_feat=src_feat, maxsep=e, labels=labels, rot=rot, absw=absw)


def triplet_setup(O, cp_level, data_level, D):
    s = cost_setup(O, cp_level, data_level, D)
    cp_tri = synth.icosphere(cp_level)[1]
    s["cp_tri"] = cp_tri
    s["triplets"] = np.sort(cp_tri, axis=1).astype(np.int32)          # DiscreteModel.cpp:293-303: node ids ascending
    s["orig"] = s["cp"].copy()
    s["cp_now"] = synth.smooth_warp(s["cp"], max_disp=0.3 * s["maxsep"].mean(), seed=17)   # a CP grid that has already moved
    rng = np.random.default_rng(23)
    T, L = len(cp_tri), len(s["labels"])
    n = 4000
    s["req"] = (rng.integers(0, T, n).astype(np.int32), rng.integers(0, L, n).astype(np.int32),
                rng.integers(0, L, n).astype(np.int32), rng.integers(0, L, n).astype(np.int32))
    # rotations map the label-grid centre onto the CURRENT control points (get_rotations, DiscreteModel.cpp:310)
    centre = np.array([0.0, 0.0, 100.0])
    s["rot_now"] = np.array([O.oracle_rotation_matrix(centre, c) for c in s["cp_now"]]).reshape(-1, 9)
    return s


def anat_case(O, s, depth=1):
    """Inputs of the anatomical strain regulariser (regoption 4/5, DiscreteCostFunction.cpp:169-181, 245-301) for a triplet_setup() case, built
    the way Mesh_registration::resample_anatomy does (mesh_registration.cpp:251-331): the anatomical grid is the control grid subdivided
    `depth` times, NEARESTFACES[t] = its faces inside control triangle t, _ANATbaryweights[v] = barycentric weights of vertex v in a control
    triangle that holds it (the LAST one that touches it, like the reference's loop), _aSOURCE / _aTARGET = two synthetic folded surfaces."""
    level = {12: 0, 42: 1, 162: 2, 642: 3, 2562: 4}[len(s["cp"])]
    aico, a_tri = synth.icosphere(level + depth)
    cp, cp_tri = s["cp"], s["cp_tri"]
    cen = aico[a_tri].mean(axis=1)
    cen = cen / np.linalg.norm(cen, axis=1, keepdims=True) * 100.0
    owner = O.OracleOctree(cp, cp_tri).query(cen)[0]                       # control triangle of every anatomical face
    order = np.argsort(owner, kind="stable")
    face_ptr = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=len(cp_tri)))]).astype(np.int32)
    face_ids = order.astype(np.int32)
    keys = np.zeros((len(aico), 3), np.int32)
    wts = np.zeros((len(aico), 3))

    def area(a, b, c):
        return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=-1)

    for t in range(len(cp_tri)):                                            # ascending: the last control triangle touching a vertex wins
        fs = face_ids[face_ptr[t]:face_ptr[t + 1]]
        vs = np.unique(a_tri[fs])
        v0, v1, v2 = cp[cp_tri[t, 0]], cp[cp_tri[t, 1]], cp[cp_tri[t, 2]]
        nrm = np.cross(v2 - v0, v1 - v0)
        nrm /= np.linalg.norm(nrm)
        P = aico[vs]
        PP = P - ((P - v0) @ nrm)[:, None] * nrm[None, :]
        w = np.stack([area(PP, v1, v2), area(PP, v0, v2), area(PP, v0, v1)], axis=1)
        w /= w.sum(axis=1, keepdims=True)
        o = np.argsort(cp_tri[t])
        keys[vs] = cp_tri[t][o]
        wts[vs] = w[:, o]
    f1, f2 = synth.smooth_fields(aico, 1, seed0=71)[0], synth.smooth_fields(aico, 1, seed0=83)[0]
    asource = aico * (0.62 + 0.12 * f1 / np.abs(f1).max())[:, None]
    moved = synth.smooth_warp(aico, max_disp=1.0, seed=91)
    atarget = moved * (0.60 + 0.14 * f2 / np.abs(f2).max())[:, None]
    return dict(asource_xyz=asource, asource_tri=a_tri, thi_xyz=aico, thi_tri=a_tri, atarget_xyz=atarget, face_ptr=face_ptr, face_ids=face_ids,
                bary_ptr=(3 * np.arange(len(aico) + 1)).astype(np.int32), bary_key=keys.reshape(-1), bary_w=wts.reshape(-1))


def group_setup(S=3, cp_level=2, data_level=4, tpl_level=4, D=2):
    cp0, cp_tri = synth.icosphere(cp_level)
    dxyz0, dtri = synth.icosphere(data_level)
    tpl, tpl_tri = synth.icosphere(tpl_level)
    tpl = synth.rotate_sphere(tpl, 0.004, -0.003, 0.002)
    data = np.stack([synth.smooth_warp(dxyz0, max_disp=2.0, seed=40 + s) for s in range(S)])
    cps = np.stack([synth.smooth_warp(cp0, max_disp=1.5, seed=60 + s) for s in range(S)])
    feat = np.stack([synth.smooth_fields(data[s], D, seed0=100, noise=0.1, noise_seed=7 + s) for s in range(S)])
    centre = np.array([0.0, 0.0, 100.0])
    labels = [centre]
    for k in range(6):
        p = centre + 6.0 * np.array([np.cos(k * np.pi / 3), np.sin(k * np.pi / 3), 0.0])
        labels.append(p / np.linalg.norm(p) * 100)
    return dict(cp_tri=cp_tri, dtri=dtri, tpl=tpl, tpl_tri=tpl_tri, data=data, cps=cps, feat=feat, centre=centre, labels=np.array(labels))




def group_mask(g):
    """A cost mask on the template (`set_masks`, DiscreteGroupModel.cpp:164): smooth signed values, a zero band (weight 0 vertices)
    and a region of exact ones. The cost function uses std::abs of it (DiscreteGroupCostFunction.cpp:77)."""
    t = g["tpl"] / 100.0
    m = np.sin(3.0 * t[:, 0]) * np.cos(2.0 * t[:, 1]) + 0.3 * t[:, 2]
    m[np.abs(m) < 0.15] = 0.0
    m[m > 0.9] = 1.0
    return np.ascontiguousarray(m)


def golden_group_glue(O, g, n=800):
    """Host glue of DiscreteGroupModel for a group case: rotations (get_rotations, cpp:76-86), spacings (get_spacings, 123-145),
    pairs (estimate_pairs, 37-55: partner = nearest control point of subject B) and a seeded request list."""
    S, ncp = g["cps"].shape[0], g["cps"].shape[1]
    rot = np.array([O.oracle_rotation_matrix(g["centre"], c) for c in g["cps"].reshape(-1, 3)]).reshape(-1, 9)
    spacings = np.zeros((S, ncp))
    for s_ in range(S):
        for a, b in ((0, 1), (1, 2), (0, 2)):
            d = np.sqrt(((g["cps"][s_][g["cp_tri"][:, a]] - g["cps"][s_][g["cp_tri"][:, b]]) ** 2).sum(axis=1))
            geo = 2 * 100.0 * np.arcsin(d / 200.0)
            np.maximum.at(spacings[s_], g["cp_tri"][:, a], geo)
            np.maximum.at(spacings[s_], g["cp_tri"][:, b], geo)
    near = lambda a, v, b: int(np.argmin(((g["cps"][b] - g["cps"][a][v]) ** 2).sum(axis=1)))
    pairs = np.array([[a * ncp + v, b * ncp + near(a, v, b)] for a in range(S) for v in range(ncp) for b in range(a + 1, S)], np.int32)
    rng = np.random.default_rng(9)
    L = len(g["labels"])
    req = (rng.integers(0, len(pairs), n).astype(np.int32), rng.integers(0, L, n).astype(np.int32), rng.integers(0, L, n).astype(np.int32))
    return rot, spacings, pairs, req


GOLDEN_CP, GOLDEN_DATA = 2, 4   # levels used by tests/golden/make_golden_costs.py


def golden_digest(d):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(d):
        v = d[k]
        if isinstance(v, tuple):
            v = np.concatenate([np.asarray(x).ravel() for x in v])
        h.update(np.ascontiguousarray(v).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def group_triplet_case(O, g, n=3000):
    """Inputs of DiscreteGroupCostFunction::computeTripletCost for a group case: global triplets (estimate_triplets, DiscreteGroupModel.cpp:57-74),
    undeformed grids = the shared icosphere, current grids = g["cps"], a seeded request list."""
    from newmsm_b200 import synth
    S, ncp = g["cps"].shape[0], g["cps"].shape[1]
    level = {12: 0, 42: 1, 162: 2, 642: 3, 2562: 4}[ncp]
    cp0, _ = synth.icosphere(level)
    orig = np.stack([cp0] * S)
    trip = np.concatenate([np.sort(g["cp_tri"] + s_ * ncp, axis=1) for s_ in range(S)]).astype(np.int32)
    rot = np.array([O.oracle_rotation_matrix(g["centre"], c) for c in g["cps"].reshape(-1, 3)]).reshape(-1, 9)
    rng = np.random.default_rng(31)
    L = len(g["labels"])
    req = (rng.integers(0, len(trip), n).astype(np.int32), rng.integers(0, L, n).astype(np.int32),
           rng.integers(0, L, n).astype(np.int32), rng.integers(0, L, n).astype(np.int32))
    return orig, trip, rot, req
