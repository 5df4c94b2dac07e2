"""Generates tests/golden/anat.npz from the UNMODIFIED reference registration library compiled into
oracle/_ref/libref_newmeshreg.so (needs /root/reference; run from the repo root: python tests/golden/make_golden_anat.py).
computeTripletCost with regoption 5 (anatomical strain, DiscreteCostFunction.cpp:169-181, 245-301) of the reference's own classes on the
seeded cases of tests/cost_cases.py (inputs are regenerated from the seeds; an input checksum guards against drift)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cost_cases import GOLDEN_CP as CP, GOLDEN_DATA as DATA, anat_case, golden_digest as digest, triplet_setup  # noqa: E402
from oracle import bindings as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = ((0, 1, 1), (0, 1, 2), (3, 1, 2))   # (cost kind, feature dimension, anatomical grid levels above the control grid)


def main():
    B.build(ref=True)
    out = {}
    for kind, D, depth in CASES:
        s = triplet_setup(B, CP, DATA, D)
        a = anat_case(B, s, depth)
        rt, la, lb, lc = s["req"]
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
        c, _, _ = B.refmr_triplet(kind, 2, s["xyz"], s["tri"], s["cp_now"], s["cp_tri"], s["orig"], s["rot_now"], s["labels"], s["triplets"],
                                  rt, la, lb, lc, s["src"], s["tri"], s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05, 0.4, 1.6, 2.0, 2.0, 5,
                                  nthreads=1, anat=a)
        out[f"k{kind}_d{depth}"] = c
        out[f"k{kind}_d{depth}_digest"] = np.concatenate([digest(s), digest(a)])
    np.savez_compressed(os.path.join(OUT, "anat.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
