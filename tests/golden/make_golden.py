"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref
(needs /root/reference; run from the repo root: python tests/golden/make_golden.py).
The fixtures are outputs of the reference itself on seeded synthetic inputs; the reference
ships no golden vectors of its own."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from newmsm_b200 import synth  # noqa: E402
from oracle import bindings as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    B.build(ref=True)
    # --- octree + nearest triangle + barycentric weights: ico3 mesh (jittered), mixed queries
    xyz, tri = synth.icosphere(3)
    xyz = synth.jitter_sphere(xyz, tri, seed=21)
    rng = np.random.default_rng(5)
    r = rng.normal(size=(1500, 3)); r = r / np.linalg.norm(r, axis=1, keepdims=True) * 100
    planes = r[:300].copy(); planes[:100, 0] = 0.0; planes[100:200, 1] = 50.5; planes[200:, 2] = -25.25
    q = np.concatenate([synth.rotate_sphere(synth.icosphere(4)[0]), r, xyz, planes,
                        (xyz[tri[:, 0]] + xyz[tri[:, 1]]) / 2, xyz * 1.005,
                        [[101.5, 0, 0], [0, -200, 0]]])
    mesh = B.RefMesh(xyz, tri)
    tree = B.RefOctree(mesh)
    kinds, counts, leaf_tris = tree.dump()
    t, v, s = tree.query(q, nthreads=1)
    ok = s == 0
    low = B.RefMesh(q[ok], np.zeros((0, 3), np.int32))
    idx, w, ne, err = tree.bary_weights(low, nthreads=1)
    assert err == 0
    np.savez_compressed(os.path.join(OUT, "octree_ico3.npz"), xyz=xyz, tri=tri, q=q, kinds=kinds, counts=counts,
                        leaf_tris=leaf_tris, tri_id=t, vertex_id=v, status=s, w_idx=idx, w_val=w, w_n=ne,
                        vertex_area=mesh.vertex_areas())

    # --- adaptive barycentric + metric_resample, both directions (down- and up-sampling)
    for name, lin, llow in (("down", 4, 3), ("up", 3, 4)):
        xi, ti = synth.icosphere(lin); xi = synth.jitter_sphere(xi, ti, seed=31)
        xl, tl = synth.icosphere(llow); xl = synth.rotate_sphere(xl)
        feat = synth.smooth_fields(xi, 4).astype(np.float32).astype(np.float64)  # f32-representable payload
        mi, ml = B.RefMesh(xi, ti, feat=feat), B.RefMesh(xl, tl)
        rowptr, col, val = B.ref_adaptive_weights(mi, ml, nthreads=1)
        out, _ = B.ref_metric_resample(mi, ml, nthreads=1)
        bout, _ = B.ref_bary_resample(mi, ml, nthreads=1)
        np.savez_compressed(os.path.join(OUT, f"resample_{name}.npz"), xyz_in=xi, tri_in=ti, xyz_low=xl, tri_low=tl,
                            feat=feat, rowptr=rowptr, col=col, val=val, metric_out=out, bary_out=bout)

    # --- coordinate blends
    xf, tf = synth.icosphere(3)
    xto = synth.smooth_warp(xf, max_disp=8.0, seed=9)
    sph, t4 = synth.icosphere(4); sph = synth.rotate_sphere(sph)
    anat = xf * np.array([1.0, 0.8, 0.6])
    feat = synth.smooth_fields(xf, 2)
    np.savez_compressed(
        os.path.join(OUT, "blend.npz"), xf=xf, tf=tf, xto=xto, sph=sph, t4=t4, anat=anat, feat=feat,
        warp=B.ref_sphere_project_warp(B.RefMesh(sph, t4), B.RefMesh(xf, tf), B.RefMesh(xto, tf)),
        surf=B.ref_surface_resample(B.RefMesh(anat, tf), B.RefMesh(xf, tf), B.RefMesh(sph, t4)),
        nn=B.ref_nn_resample(B.RefMesh(xf, tf, feat=feat), B.RefMesh(sph, t4)))

    # --- rotation matrices (point.cpp:97), incl. the identity / antipodal special cases
    rng = np.random.default_rng(1)
    ci = rng.normal(size=(64, 3)); ix = rng.normal(size=(64, 3))
    ix[0] = ci[0]; ix[1] = -ci[1]; ix[2] = 2.5 * ci[2]
    R = np.stack([B.ref_rotation_matrix(a, b) for a, b in zip(ci, ix)])
    np.savez_compressed(os.path.join(OUT, "rotation.npz"), ci=ci, index=ix, R=R)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
