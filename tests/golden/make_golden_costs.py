"""Generates tests/golden/costs.npz from the UNMODIFIED reference registration library compiled into
oracle/_ref/libref_newmeshreg.so (needs /root/reference; run from the repo root:
python tests/golden/make_golden_costs.py). Outputs of the reference's own cost-function classes on the seeded cases of
tests/cost_cases.py (inputs are regenerated from the seeds; an input checksum guards against drift)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cost_cases import GOLDEN_CP as CP, GOLDEN_DATA as DATA, cost_setup, golden_digest as digest, golden_group_glue, group_mask, group_setup, group_triplet_case, triplet_setup  # noqa: E402
from newmsm_b200 import synth  # noqa: E402
from oracle import bindings as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    B.build(ref=True)
    out = {}
    cp_tri = synth.icosphere(CP)[1]
    for kind, D in ((0, 1), (1, 4), (2, 4)):
        s = cost_setup(B, CP, DATA, D)
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D if kind == 1 else 1, len(s["src"])))
        for sim in (1, 2):
            c, prow, pmem, _ = B.refmr_unary(kind, sim, s["xyz"], s["tri"], s["cp"], cp_tri, s["rot"], s["labels"], s["src"], s["tri"],
                                            s["src_feat"], s["ref_feat"], cfw, s["absw"], s["maxsep"], 1.0, nthreads=1)
            out[f"unary_k{kind}_s{sim}"] = c
        for sim, pct in ((4, 0.75), (5, 0.6)):     # DICE / genDICE (similarities.cpp:201-253)
            B.refmr_set_percentile(pct)
            c, _, _, _ = B.refmr_unary(kind, sim, s["xyz"], s["tri"], s["cp"], cp_tri, s["rot"], s["labels"], s["src"], s["tri"],
                                       s["src_feat"], s["ref_feat"], cfw, s["absw"], s["maxsep"], 1.0, nthreads=1)
            B.refmr_set_percentile(0.75)
            out[f"unary_k{kind}_s{sim}"] = c
        out[f"unary_k{kind}_prow"], out[f"unary_k{kind}_pmem"] = prow, pmem
        out[f"unary_k{kind}_digest"] = digest(s)
    for kind, D in ((0, 1), (3, 1), (4, 3)):
        s = triplet_setup(B, CP, DATA, D)
        rt, la, lb, lc = s["req"]
        cfw = np.random.default_rng(5).uniform(0.2, 1.0, size=(D, len(s["src"])))
        c, prow, pmem = B.refmr_triplet(kind, 2, s["xyz"], s["tri"], s["cp_now"], s["cp_tri"], s["orig"], s["rot_now"], s["labels"], s["triplets"],
                                        rt, la, lb, lc, s["src"], s["tri"], s["src_feat"], s["ref_feat"], cfw, s["absw"], 0.05, nthreads=1)
        out[f"triplet_k{kind}"] = c
        if kind >= 3:
            out[f"triplet_k{kind}_prow"], out[f"triplet_k{kind}_pmem"] = prow, pmem
        out[f"triplet_k{kind}_digest"] = digest(s)
    g = group_setup(S=2, cp_level=1, data_level=3, tpl_level=3, D=2)
    rot, spacings, pairs, (rp, la, lb) = golden_group_glue(B, g)
    ncp = g["cps"].shape[1]
    for sim in (1, 2):
        c, fields = B.refmr_group_pair_costs(sim, g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"], ncp, rot, spacings, 1.0,
                                             pairs, rp, la, lb, want_fields=True)
        out[f"group_pair_s{sim}"] = c
        out[f"group_pair_masked_s{sim}"] = B.refmr_group_pair_costs(sim, g["data"], g["dtri"], g["feat"], g["labels"], g["centre"], g["tpl"], g["tpl_tri"], ncp,
                                                                    rot, spacings, 1.0, pairs, rp, la, lb, mask=group_mask(g))
    out["group_fields"] = fields          # NaN where no patch of any (CP, label) contains the template vertex
    orig, trip, rot_t, (rt, ta, tb, tc) = group_triplet_case(B, g)
    out["group_triplet"] = B.refmr_group_triplet_costs(g["cps"], orig, g["cp_tri"], rot_t, g["labels"], trip, 0.05, 0.4, 1.6, 2.0, 2.0, rt, ta, tb, tc)
    out["group_digest"] = digest(g)
    np.savez_compressed(os.path.join(OUT, "costs.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
