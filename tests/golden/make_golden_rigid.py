"""Generates tests/golden/rigid.npz from the UNMODIFIED reference compiled into oracle/_ref/libref_newmeshreg.so (needs
/root/reference; run from the repo root: python tests/golden/make_golden_rigid.py): outputs of the reference's RIGID / AFFINE level
(Rigid_cost_function::initialise + rigid_cost_mesh(0,0,0) + run, rigid_costfunction.cpp:32-236) on seeded inputs. SURVEY §8 f4: the level
is restated in the CPU oracle (orc_rigid, pinned by these vectors) and has no CUDA path yet."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from newmsm_b200 import synth  # noqa: E402
from oracle import bindings as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def rigid_case(level=3, D=2):
    """Target = icosphere, source = the same sphere rotated by a small known rotation; reference data = smooth fields on the target,
    input data = the same fields seen from the rotated frame."""
    xyz, tri = synth.icosphere(level)
    src = synth.rotate_sphere(xyz, 0.03, -0.02, 0.04)
    ref = synth.smooth_fields(xyz, D, seed0=100)
    mov = synth.smooth_fields(synth.rotate_sphere(xyz, -0.03, 0.02, -0.04), D, seed0=100)
    return xyz, tri, src, mov, ref


def main():
    B.build(ref=True)
    out = {}
    for name, D, sim in (("d2_corr", 2, 2), ("d1_corr", 1, 2), ("d3_ssd", 3, 1)):
        xyz, tri, src, mov, ref = rigid_case(3, D)
        moved, cost0, rowptr, members = B.refmr_rigid(xyz, tri, src, tri, mov, ref, simmeasure=sim, iters=4, stepsize=0.01, gradsampling=0.5, nthreads=1)
        out[f"{name}_xyz"], out[f"{name}_cost0"], out[f"{name}_rowptr"], out[f"{name}_members"] = moved, np.float64(cost0), rowptr, members
        print(name, "cost at zero rotation", cost0, "neighbours", rowptr[-1], "largest displacement", np.abs(moved - src).max())
    np.savez_compressed(os.path.join(OUT, "rigid.npz"), **out)


if __name__ == "__main__":
    main()
