"""Generates tests/golden/excl.npz from the UNMODIFIED reference compiled into oracle/_ref/libref_newresampler.so (needs /root/reference;
run from the repo root: python tests/golden/make_golden_excl.py): the reference's resampling functions WITH an exclusion mask
(metric_resample / get_adaptive_barycentric_weights / smooth_data / nearest_neighbour_interpolation with EXCL, resampler.cpp:30-140,
169-258; used by featurespace::initialise when --excl or cut thresholds are set, featurespace.cpp:61-70), single-threaded."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from newmsm_b200 import synth  # noqa: E402
from oracle import bindings as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def excl_case(src_level=4, low_level=3, D=3):
    """Source = jittered icosphere with D smooth channels, target = a rotated coarser icosphere; mask = 0 on a polar cap (the medial wall of
    real data), graded values elsewhere (smooth_data weighs by the mask's value, resampler.cpp:209)."""
    xyz0, tri = synth.icosphere(src_level)
    xyz = synth.jitter_sphere(xyz0, tri, frac=0.2, seed=21)
    low0, low_tri = synth.icosphere(low_level)
    low = synth.rotate_sphere(low0, 0.02, -0.03, 0.01)
    feat = synth.smooth_fields(xyz, D, seed0=300)
    excl = np.where(xyz[:, 2] > 55.0, 0.0, 1.0)
    band = (xyz[:, 2] > 30.0) & (xyz[:, 2] <= 55.0)
    excl[band] = 0.5 + 0.5 * (55.0 - xyz[band, 2]) / 25.0
    return xyz, tri, low, low_tri, feat, excl


def main():
    B.build(ref=True)
    out = {}
    for name, (sl, ll) in {"down": (4, 3), "up": (3, 4), "same": (3, 3)}.items():
        xyz, tri, low, low_tri, feat, excl = excl_case(sl, ll)
        mi, ml = B.RefMesh(xyz, tri, feat=feat), B.RefMesh(low, low_tri)
        o, eo = B.ref_metric_resample_excl(mi, ml, excl)
        rp, col, val = B.ref_adaptive_weights_excl(mi, ml, excl)
        n, en = B.ref_nn_resample_excl(mi, ml, excl)
        out.update({f"{name}_metric": o, f"{name}_metric_excl": eo, f"{name}_rowptr": rp, f"{name}_col": col, f"{name}_val": val,
                    f"{name}_nn": n, f"{name}_nn_excl": en})
        print(name, "targets without a row:", int((np.diff(rp) == 0).sum()), "of", len(rp) - 1, "| mask range", eo.min(), eo.max())
    # smooth_data smooths a mesh's data over the SAME mesh (every call site passes orig == sphLow, featurespace.cpp:73)
    xyz, tri, _, _, feat, excl = excl_case(4, 3)
    m = B.RefMesh(xyz, tri, feat=feat)
    for sigma in (4.0, 9.0):
        s0, _ = B.ref_smooth_data(m, m, sigma)
        s1, e1 = B.ref_smooth_data(m, m, sigma, excl)
        out.update({f"smooth{int(sigma)}": s0, f"smooth{int(sigma)}_masked": s1, f"smooth{int(sigma)}_excl": e1})
    np.savez_compressed(os.path.join(OUT, "excl.npz"), **out)


if __name__ == "__main__":
    main()
