"""Writes a synthetic pairwise-registration case for the newmsm CLI (docs/guide.md "Example B"):
--inmesh = --refmesh = an icosphere of the given level (R = 100, FreeSurfer ASCII, mesh.cpp:455), the misalignment lives in
the data: refdata(x) = f(x) + 5 % noise, indata(x) = f(warp(x)) with a smooth random tangential warp. Data as ASCII matrices
([V][D] text, mesh.cpp:517). Also writes the reference's own configs with the AFFINE level removed (config/README:3).

    python tools/make_newmsm_case.py --out /tmp/case --level 6 --D 1
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from newmsm_b200 import synth  # noqa: E402

CONFIGS = {
    # config/basic_configs/config_standard_MSMpair without the AFFINE level (FastPD, pairwise regulariser)
    "MSMpair": ["--sigma_in=6,4,2", "--sigma_ref=6,4,2", "--lambda=0.1,0.2,0.3", "--it=5,10,10", "--opt=DISCRETE,DISCRETE,DISCRETE",
                "--CPgrid=2,3,4", "--SGgrid=4,5,6", "--datagrid=5,5,6", "--regoption=1", "--dopt=FastPD"],
    # config/basic_configs/config_standard_MSMpair AS SHIPPED, i.e. with its AFFINE first level (rigid_costfunction.cpp: not accelerated,
    # it runs on the reference's host code inside the same program; the three DISCRETE levels use the GPU paths)
    "MSMpairAffine": ["--sigma_in=6,6,4,2", "--sigma_ref=6,6,4,2", "--lambda=0,0.1,0.2,0.3", "--it=50,5,10,10", "--opt=AFFINE,DISCRETE,DISCRETE,DISCRETE",
                      "--CPgrid=0,2,3,4", "--SGgrid=0,4,5,6", "--datagrid=5,5,5,6", "--regoption=1"],
    # config/HCP_multimodal_alignment/MSMAllStrainFinalconf1to1_1to3_2 (HOCR, triclique likelihood, strain regulariser)
    "MSMAllStrain": ["--simval=2,2,2", "--sigma_in=0,0,0", "--sigma_ref=0,0,0", "--lambda=0.00001,0.0075,0.01", "--it=10,15,15",
                     "--opt=DISCRETE,DISCRETE,DISCRETE", "--CPgrid=2,3,4", "--SGgrid=4,5,6", "--datagrid=4,5,6", "--regoption=3", "--regexp=2",
                     "--dopt=HOCR", "--VN", "--rescaleL", "--triclique", "--k_exponent=2", "--bulkmod=1.6", "--shearmod=0.4"],
    # config/basic_configs/config_standard_MSM_strain without the AFFINE level (HOCR, univariate unary costs + strain triplets)
    "MSMstrain": ["--simval=2,2,2", "--sigma_in=4,2,1", "--sigma_ref=4,2,1", "--lambda=0.2,0.2,0.2", "--it=20,25,25", "--opt=DISCRETE,DISCRETE,DISCRETE",
                  "--CPgrid=2,3,4", "--SGgrid=4,5,6", "--datagrid=5,5,6", "--regoption=3", "--regexp=2", "--dopt=HOCR", "--VN", "--k_exponent=2",
                  "--bulkmod=1.6", "--shearmod=0.4", "--rescaleL"],
    # BASELINE configs[3]: NeuroImage2017 sMSM_STR semantics on ONE level with an ico5 control grid, ico6 data, ico7 sampling grid
    # (SURVEY §8d cfg4): 10 242 control points, 20 480 triplets
    # docs/guide.md:390-407, the example config of groupwise (gMSM) registration
    "gMSM": ["--simval=2,2,2", "--sigma_in=0,0,0", "--sigma_ref=0,0,0", "--lambda=0.2,0.2,0.2", "--it=9,9,9", "--opt=DISCRETE,DISCRETE,DISCRETE",
             "--CPgrid=2,3,4", "--SGgrid=4,5,6", "--datagrid=4,5,6", "--regoption=3", "--regexp=2", "--dopt=HOCR", "--k_exponent=2",
             "--bulkmod=1.6", "--shearmod=0.4"],
    # config/NeuroImage2017_configs/aMSM_STR_longitudinal_alignment (anatomical strain, regoption 5: run with --inanat / --refanat, the
    # anatomical grid two levels above the control grid like the shipped --anatgrid=4,5,6 for --CPgrid=2,3,4)
    "aMSMSTR": ["--simval=2,2,2", "--sigma_in=6,4,2", "--sigma_ref=6,4,2", "--lambda=0.025,0.025,0.025", "--it=40,40,40",
                "--opt=DISCRETE,DISCRETE,DISCRETE", "--CPgrid=2,3,4", "--SGgrid=4,5,6", "--datagrid=4,5,6", "--anatgrid=4,5,6", "--regoption=5",
                "--regexp=2", "--dopt=HOCR", "--VN", "--rescaleL", "--triclique", "--k_exponent=2", "--bulkmod=1.6", "--shearmod=0.4"],
    "sMSMSTRcp5": ["--simval=2", "--sigma_in=2", "--sigma_ref=2", "--lambda=0.025", "--it=40", "--opt=DISCRETE", "--CPgrid=5", "--SGgrid=7",
                   "--datagrid=6", "--regoption=3", "--regexp=2", "--dopt=HOCR", "--VN", "--rescaleL", "--triclique", "--k_exponent=2",
                   "--bulkmod=1.6", "--shearmod=0.4"],
}


def write_asc(path, xyz, tri):
    with open(path, "w") as f:
        f.write("#!ascii version of synthetic sphere\n%d %d\n" % (len(xyz), len(tri)))
        for p in xyz:
            f.write("%.17g %.17g %.17g 0\n" % tuple(p))
        for t in tri:
            f.write("%d %d %d 0\n" % tuple(t))


def scaled(cfg_lines, levels_drop, it_scale):
    """drop the finest `levels_drop` levels / shrink the iteration counts (for quick smoke cases)"""
    out = []
    for ln in cfg_lines:
        if "=" in ln and ("," in ln or ln.startswith("--it=")):
            k, v = ln.split("=")
            vals = v.split(",")
            if levels_drop and len(vals) > levels_drop:
                vals = vals[:-levels_drop]
            if k == "--it":
                vals = [str(max(1, int(round(int(x) * it_scale)))) for x in vals]
            ln = k + "=" + ",".join(vals)
        out.append(ln)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--D", type=int, default=1)
    ap.add_argument("--levels-drop", type=int, default=0)
    ap.add_argument("--it-scale", type=float, default=1.0)
    ap.add_argument("--max-disp", type=float, default=8.26)
    ap.add_argument("--group", type=int, default=0, help="also write a groupwise case: this many subjects (meshes.txt, data.txt, template.asc)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    xyz, tri = synth.icosphere(a.level)
    write_asc(os.path.join(a.out, "sphere.asc"), xyz, tri)
    warped = synth.smooth_warp(xyz, max_disp=a.max_disp, seed=2024)
    ref = synth.smooth_fields(xyz, a.D, seed0=100, noise=0.05, noise_seed=7)        # [D][V]
    mov = synth.smooth_fields(warped, a.D, seed0=100)
    np.savetxt(os.path.join(a.out, "refdata.txt"), ref.T, fmt="%.9g")
    np.savetxt(os.path.join(a.out, "indata.txt"), mov.T, fmt="%.9g")
    # anatomical surfaces (--inanat / --refanat, mesh_registration.cpp:434): smooth folded surfaces over the same topology, the reference one
    # grown and displaced (a "later time point")
    f1, f2 = synth.smooth_fields(xyz, 1, seed0=71)[0], synth.smooth_fields(xyz, 1, seed0=83)[0]
    write_asc(os.path.join(a.out, "inanat.asc"), xyz * (0.62 + 0.10 * f1 / np.abs(f1).max())[:, None], tri)
    write_asc(os.path.join(a.out, "refanat.asc"), synth.smooth_warp(xyz, max_disp=2.0, seed=91) * (0.66 + 0.12 * f2 / np.abs(f2).max())[:, None], tri)
    for name, lines in CONFIGS.items():
        with open(os.path.join(a.out, "conf_" + name), "w") as f:
            f.write("\n".join(scaled(lines, a.levels_drop, a.it_scale)) + "\n")
    if a.group:
        # gMSM inputs (src/newmsm.cpp:14-28): a list of subject spheres, a list of their data files, a template sphere
        tpl = synth.rotate_sphere(xyz, 0.004, -0.003, 0.002)
        write_asc(os.path.join(a.out, "template.asc"), tpl, tri)
        meshes, datas = [], []
        for s_ in range(a.group):
            sx = synth.smooth_warp(xyz, max_disp=a.max_disp * 0.5, seed=300 + s_)
            write_asc(os.path.join(a.out, f"subject{s_}.asc"), sx, tri)
            f = synth.smooth_fields(synth.smooth_warp(xyz, max_disp=a.max_disp * 0.5, seed=400 + s_), a.D, seed0=100, noise=0.05, noise_seed=20 + s_)
            np.savetxt(os.path.join(a.out, f"subject{s_}.txt"), f.T, fmt="%.9g")
            meshes.append(os.path.join(a.out, f"subject{s_}.asc"))
            datas.append(os.path.join(a.out, f"subject{s_}.txt"))
        # a cost mask on the template (newmsm --mask, groupwise mode only: msmOptions.h:85): signed smooth values with a zero band
        t = tpl / 100.0
        mk = np.sin(3.0 * t[:, 0]) * np.cos(2.0 * t[:, 1]) + 0.3 * t[:, 2]
        mk[np.abs(mk) < 0.15] = 0.0
        np.savetxt(os.path.join(a.out, "mask.txt"), mk[:, None], fmt="%.9g")
        open(os.path.join(a.out, "meshes.txt"), "w").write("\n".join(meshes) + "\n")
        open(os.path.join(a.out, "data.txt"), "w").write("\n".join(datas) + "\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
