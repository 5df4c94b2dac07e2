"""Synthetic inputs for parity tests and bench.py.

Host-side numpy only; nothing here is on the hot path.

* :func:`icosphere` reproduces the vertex AND face order of the reference's
  ``make_mesh_from_icosa`` (mesh.cpp:1111-1196, ``retessellate`` mesh.cpp:910-1008) in
  O(N log N) instead of the reference's O(N^2) midpoint search, then ``true_rescale``
  (mesh.cpp:1210). Checked bit-for-bit against the compiled reference in
  tests/test_oracle_vs_ref.py.
* :func:`geodesic_sphere` builds the 10 f^2 + 2 vertex class-I sphere (f = 57 -> 32 492
  vertices, the "32k fs_LR"-sized target of BASELINE.json configs[1]).
* :func:`smooth_fields` / :func:`jitter_sphere` are the smooth random feature fields and the
  tangentially jittered "native" sphere of SURVEY.md §8d.
"""
from __future__ import annotations

import numpy as np

_TAU = 0.8506508084
_ONE = 0.5257311121


def _icosahedron():
    t, o = _TAU, _ONE
    v = np.array(
        [[t, o, 0], [-t, o, 0], [-t, -o, 0], [t, -o, 0],      # ZA ZB ZC ZD
         [o, 0, t], [o, 0, -t], [-o, 0, -t], [-o, 0, t],      # YA YB YC YD
         [0, t, o], [0, -t, o], [0, -t, -o], [0, t, -o]],     # XA XB XC XD
        dtype=np.float64)
    ZA, ZB, ZC, ZD, YA, YB, YC, YD, XA, XB, XC, XD = range(12)
    f = np.array(
        [[YD, XA, YA], [XB, YD, YA], [XD, YC, YB], [YC, XC, YB], [ZD, YA, ZA],
         [YB, ZD, ZA], [ZB, YD, ZC], [YC, ZB, ZC], [XD, ZA, XA], [ZB, XD, XA],
         [ZD, XC, XB], [XC, ZC, XB], [ZA, YA, XA], [YB, ZA, XD], [ZD, XB, YA],
         [XC, ZD, YB], [ZB, XA, YD], [XD, ZB, YC], [XB, ZC, YD], [ZC, XC, YC]],
        dtype=np.int64)
    f = f[:, [0, 2, 1]]  # swap_orientation (mesh.cpp:1183-1184)
    return v, f


def _normalize_rows(v):
    n = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2])
    out = v.copy()
    ok = n > 1e-8
    out[ok] = v[ok] / n[ok, None]
    return out


def _retessellate(v, f):
    """One level of mesh.cpp:910-1008: per face (v0,v1,v2) the new points are
    p0 = mid(v1,v2), p1 = mid(v0,v2), p2 = mid(v0,v1), appended in first-seen order
    (p0, p1, p2 within a face); faces (p2,p0,p1), (p1,v0,p2), (p0,v2,p1), (p2,v1,p0)."""
    nv = v.shape[0]
    v0, v1, v2 = f[:, 0], f[:, 1], f[:, 2]
    ea = np.stack([v1, v0, v0], axis=1)  # first endpoint as written in the midpoint sums
    eb = np.stack([v2, v2, v1], axis=1)
    lo = np.minimum(ea, eb).reshape(-1)
    hi = np.maximum(ea, eb).reshape(-1)
    key = lo * nv + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique edges in first-appearance order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    pid = (nv + rank[inv]).reshape(-1, 3)             # new vertex id per (face, slot)
    fa = ea.reshape(-1)[first[order]]
    fb = eb.reshape(-1)[first[order]]
    mid = (v[fa] + v[fb]) / 2                         # (a + b) / 2 as in mesh.cpp:928-936
    newv = np.concatenate([v, mid], axis=0)
    p0, p1, p2 = pid[:, 0], pid[:, 1], pid[:, 2]
    nf = np.stack([np.stack([p2, p0, p1], 1), np.stack([p1, v0, p2], 1),
                   np.stack([p0, v2, p1], 1), np.stack([p2, v1, p0], 1)], axis=1).reshape(-1, 3)
    return _normalize_rows(newv), nf


def icosphere(level: int, radius: float = 100.0):
    """(xyz float64 [V,3], tri int32 [T,3]) identical to make_mesh_from_icosa(level) + true_rescale."""
    v, f = _icosahedron()
    for _ in range(level):
        v, f = _retessellate(v, f)
    v = _normalize_rows(v) * radius  # true_rescale: normalize then * rad (mesh.cpp:1212-1217)
    return np.ascontiguousarray(v), np.ascontiguousarray(f.astype(np.int32))


def geodesic_sphere(freq: int, radius: float = 100.0):
    """Class-I geodesic sphere with 10*freq^2+2 vertices / 20*freq^2 faces (freq=57 -> 32 492)."""
    bv, bf = _icosahedron()
    keymap: dict = {}
    verts: list = []

    def vid(face, i, j):
        # lattice point i*e1 + j*e2 on the face, identified by exact integer barycentric
        # coordinates against the sorted base-vertex ids so shared edges/corners merge.
        a, b, c = (int(x) for x in bf[face])
        w = {a: freq - i - j}
        w[b] = w.get(b, 0) + i
        w[c] = w.get(c, 0) + j
        k = tuple(sorted((vv, ww) for vv, ww in w.items() if ww))
        idx = keymap.get(k)
        if idx is None:
            p = ((freq - i - j) * bv[a] + i * bv[b] + j * bv[c]) / freq
            idx = len(verts)
            keymap[k] = idx
            verts.append(p)
        return idx

    faces = []
    for fc in range(20):
        for i in range(freq):
            for j in range(freq - i):
                a = vid(fc, i, j); b = vid(fc, i + 1, j); c = vid(fc, i, j + 1)
                faces.append((a, b, c))
                if i + j < freq - 1:
                    d = vid(fc, i + 1, j + 1)
                    faces.append((b, d, c))
    v = _normalize_rows(np.asarray(verts, dtype=np.float64)) * radius
    return np.ascontiguousarray(v), np.ascontiguousarray(np.asarray(faces, dtype=np.int32))


def rotation_xyz(rx: float, ry: float, rz: float):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def rotate_sphere(xyz, rx=0.013, ry=0.021, rz=0.034, radius=100.0):
    out = xyz @ rotation_xyz(rx, ry, rz).T
    return np.ascontiguousarray(_normalize_rows(out) * radius)


def smooth_fields(xyz, D: int, seed0: int = 100, noise: float = 0.0, noise_seed: int = 7):
    """f_d(x) = sum_{m=1..16} a cos(k.x/100 + phi), a~N(0,1/m), |k|~U(1,12), phi~U(0,2pi);
    seed = seed0 + d. Returns channel-major [D, V] float64 (the reference's pvalues layout)."""
    V = xyz.shape[0]
    out = np.empty((D, V), dtype=np.float64)
    for d in range(D):
        rng = np.random.default_rng(seed0 + d)
        acc = np.zeros(V)
        for m in range(1, 17):
            a = rng.normal(0.0, 1.0 / m)
            k = rng.normal(size=3)
            k *= rng.uniform(1.0, 12.0) / np.linalg.norm(k)
            phi = rng.uniform(0.0, 2 * np.pi)
            acc += a * np.cos(xyz @ k / 100.0 + phi)
        out[d] = acc
    if noise > 0:
        rng = np.random.default_rng(noise_seed)
        out += noise * out.std() * rng.normal(size=out.shape)
    return out


def jitter_sphere(xyz, tri, frac: float = 0.3, seed: int = 1234, radius: float = 100.0):
    """Smooth-ish tangential jitter <= frac * local edge length, re-projected to the sphere."""
    rng = np.random.default_rng(seed)
    e = np.linalg.norm(xyz[tri[:, 0]] - xyz[tri[:, 1]], axis=1)
    edge = np.zeros(xyz.shape[0])
    np.maximum.at(edge, tri[:, 0], e)
    np.maximum.at(edge, tri[:, 1], e)
    edge[edge == 0] = e.mean()
    d = rng.normal(size=xyz.shape)
    n = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
    d -= (d * n).sum(1, keepdims=True) * n
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    out = xyz + d * (frac * rng.uniform(0, 1, size=(xyz.shape[0], 1)) * edge[:, None])
    return np.ascontiguousarray(_normalize_rows(out) * radius)


def smooth_warp(xyz, max_disp: float = 8.26, seed: int = 2024, n_fields: int = 8, radius: float = 100.0):
    """Sum of low-order rotational fields, scaled so the largest displacement is max_disp."""
    rng = np.random.default_rng(seed)
    n = xyz / radius
    disp = np.zeros_like(xyz)
    for _ in range(n_fields):
        axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
        k = rng.normal(size=3); k *= rng.uniform(0.5, 2.0) / np.linalg.norm(k)
        amp = np.cos(n @ k * np.pi + rng.uniform(0, 2 * np.pi))
        disp += amp[:, None] * np.cross(axis, n)
    disp *= max_disp / np.linalg.norm(disp, axis=1).max()
    return np.ascontiguousarray(_normalize_rows(xyz + disp) * radius)
