"""End-to-end drop-in check on the GPU box: the UNMODIFIED reference `newmsm` program with libmsmgpu.so bound in at link time
(integration/_build/newmsm_gpu, integration/newmsm_gpu_hooks.cpp) against the same program on the CPU (oracle/_ref/newmsm_ref_trace,
single-threaded: the reference is only deterministic at --numthreads=1). Compared after every discrete
iteration: the solver's labeling (bit-exact), the control-point grid and the warped source mesh (FNV hash of the doubles), and the
final sphere.reg. Both binaries are built in the container (they contain compiled reference code) and travel with the snapshot."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINS = [os.path.join(ROOT, "integration", "_build", "newmsm_gpu"), os.path.join(ROOT, "oracle", "_ref", "newmsm_ref_trace")]


@pytest.mark.parametrize("config,D,extra", [
    ("MSMpair", 1, ["--levels-drop", "1", "--it-scale", "0.4"]),            # FastPD, univariate unary table, pairwise regulariser, smoothing
    ("MSMpairAffine", 1, ["--levels-drop", "1", "--it-scale", "0.2"]),     # the shipped basic config: AFFINE level (csrc/rigid.cu through integration/newmsm_gpu_rigid_hooks.cpp) + DISCRETE levels
    ("MSMAllStrain", 3, ["--levels-drop", "1", "--it-scale", "0.1"]),       # HOCR, HO multivariate triplet likelihood, strain regulariser
    ("MSMstrain", 1, ["--levels-drop", "2", "--it-scale", "0.1"]),         # HOCR, per-call unary costs from the device table + strain-only triplets
    ("aMSMSTR", 1, ["--levels-drop", "2", "--it-scale", "0.05"]),           # aMSM: anatomical strain (regoption 5) with --inanat / --refanat, triclique likelihood
    ("gMSM", 1, ["--levels-drop", "2", "--it-scale", "0.25", "--group", "3"]),   # groupwise driver: estimate_pairs, get_patch_data and the pair / triplet costs on the device (integration/newmsm_gpu_group_hooks.cpp)
])
def test_newmsm_labels_bit_exact(config, D, extra):
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("integration/_build/newmsm_gpu not built (needs /root/reference at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "newmsm_e2e.py"), "--level", "4", "--config", config, "--D", str(D),
                          "--threads", "4", "--skip-timing-cpu", *extra], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["discrete_iterations"] >= 2
    assert res["trace_calls"][0] == res["trace_calls"][1]
    assert res["labels_bit_exact"], res["label_mismatch_per_iteration"]
    assert res["all_meshes_bit_exact"], (res["hashes_equal"], res["trace_calls"])
    assert res["final_sphere_max_abs_diff"] == 0.0


def test_affine_level_costs_match_reference_in_process():
    """MSMGPU_VERIFY=1 on the shipped basic config: every Rigid_cost_function::rigid_cost_mesh call of the AFFINE level is evaluated by the
    device path AND by the reference's own member in the same process (rigid_costfunction.cpp:130-141) and compared bit for bit."""
    import re
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("integration/_build/newmsm_gpu not built (needs /root/reference at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "newmsm_e2e.py"), "--level", "4", "--config", "MSMpairAffine", "--D", "1", "--threads", "4",
                          "--skip-cpu", "--verify", "--levels-drop", "2", "--it-scale", "0.2"], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    line = [ln for ln in res["gpu_split"] if "rigid_cost_mesh:" in ln]
    assert line, res["gpu_split"]
    m = re.search(r"rigid_cost_mesh: (\d+) of (\d+) costs differ", line[0])
    assert m, line[0]
    bad, n = map(int, m.groups())
    assert n >= 8 and bad == 0, line[0]


@pytest.mark.parametrize("mask", [False, True])
def test_gmsm_costs_match_reference_in_process(mask):
    """MSMGPU_VERIFY=1: inside the groupwise run every pair / triplet cost Fusion::optimize asks for is ALSO evaluated by the reference's own
    computePairwiseCost / computeTripletCost on the reference's own patch maps, in the same process, and compared bit for bit.
    mask=True: `newmsm --mask` (cost mask on the template, DiscreteGroupModel.cpp:164, DiscreteGroupCostFunction.cpp:77) on the device path."""
    import re
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("integration/_build/newmsm_gpu not built (needs /root/reference at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "newmsm_e2e.py"), "--level", "4", "--config", "gMSM", "--D", "2", "--threads", "4",
                          "--skip-cpu", "--verify", "--levels-drop", "2", "--it-scale", "0.25", "--group", "3"] + (["--mask"] if mask else []),
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    line = [ln for ln in res["gpu_split"] if "group pair costs" in ln]
    assert line, res["gpu_split"]
    m = re.search(r"group pair costs: (\d+) of (\d+) differ.*triplet costs: (\d+) of (\d+) differ", line[0])
    assert m, line[0]
    bad_p, n_p, bad_t, n_t = map(int, m.groups())
    assert n_p > 1000 and n_t > 1000, line[0]
    assert bad_p == 0 and bad_t == 0, line[0]
    used = [ln for ln in res["gpu_split"] if ln.startswith("[msmgpu group]")]
    assert used and "pair batches" in used[0], res["gpu_split"]


def test_gmsm_two_devices_same_result():
    """MSMGPU_DEVICES=2: subjects (fields) and pair blocks sharded over two GPUs inside the one reference process; the result must not
    depend on the device count (same traces as the single-thread reference). Needs 2 GPUs: skipped on a 1-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("integration/_build/newmsm_gpu not built (needs /root/reference at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "newmsm_e2e.py"), "--level", "4", "--config", "gMSM", "--D", "2", "--threads", "4",
                          "--skip-timing-cpu", "--levels-drop", "2", "--it-scale", "0.25", "--group", "5", "--devices", "2"],
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["trace_calls"][0] == res["trace_calls"][1] and res["all_meshes_bit_exact"], (res["hashes_equal"], res["trace_calls"])
