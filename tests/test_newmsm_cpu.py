"""CPU-side checks of the end-to-end harness (no GPU needed; skipped when the binaries were not built because /root/reference is absent):
the reference `newmsm` program compiled against the FSL shim runs the reference's own config format and is deterministic single-threaded,
and the GPU-bound program refuses to run without a CUDA device instead of falling back to the CPU."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "newmsm_ref_trace")
GPU = os.path.join(ROOT, "integration", "_build", "newmsm_gpu")


def _case(tmp, config):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_newmsm_case.py"), "--out", tmp, "--level", "3", "--D", "1",
                    "--levels-drop", "2", "--it-scale", "0.1"], check=True, stdout=subprocess.DEVNULL)
    conf = os.path.join(tmp, "conf_" + config)
    # a level-3 sphere cannot carry the configs' ico5 data grid: keep the reference's syntax, shrink the grids
    with open(conf, "w") as f:
        f.write("--simval=2\n--sigma_in=2\n--sigma_ref=2\n--lambda=0.1\n--it=2\n--opt=DISCRETE\n--CPgrid=1\n--SGgrid=3\n--datagrid=3\n"
                "--regoption=1\n--dopt=FastPD\n--numthreads=1\n")
    return conf


def _run(binary, tmp, conf, out, trace):
    os.makedirs(os.path.join(tmp, out), exist_ok=True)
    env = dict(os.environ, MSMGPU_TRACE=os.path.join(tmp, trace), OMP_NUM_THREADS="1")
    return subprocess.run([binary, "--inmesh=" + os.path.join(tmp, "sphere.asc"), "--refmesh=" + os.path.join(tmp, "sphere.asc"),
                           "--indata=" + os.path.join(tmp, "indata.txt"), "--refdata=" + os.path.join(tmp, "refdata.txt"), "--conf=" + conf,
                           "--out=" + os.path.join(tmp, out) + "/", "-f", "ASCII"], env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/newmsm_ref_trace not built (needs /root/reference)")
def test_reference_cli_runs_and_is_deterministic():
    with tempfile.TemporaryDirectory() as tmp:
        conf = _case(tmp, "tiny")
        a = _run(REF, tmp, conf, "out_a", "trace_a.txt")
        b = _run(REF, tmp, conf, "out_b", "trace_b.txt")
        assert a.returncode == 0 and b.returncode == 0, a.stdout[-1500:] + a.stderr[-1500:]
        ta, tb = open(os.path.join(tmp, "trace_a.txt")).read(), open(os.path.join(tmp, "trace_b.txt")).read()
        assert ta == tb and ta.count("\nL ") + ta.startswith("L ") >= 2          # two discrete iterations traced, bit-identical runs
        assert os.path.exists(os.path.join(tmp, "out_a", "sphere.reg.asc"))


@pytest.mark.skipif(not os.path.exists(GPU), reason="integration/_build/newmsm_gpu not built (needs /root/reference)")
def test_gpu_program_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with tempfile.TemporaryDirectory() as tmp:
        conf = _case(tmp, "tiny")
        r = _run(GPU, tmp, conf, "out_g", "trace_g.txt")
        assert r.returncode != 0
        assert "CUDA" in r.stdout + r.stderr
        assert not os.path.exists(os.path.join(tmp, "out_g", "sphere.reg.asc"))      # no CPU fallback produced a result


@pytest.mark.skipif(not os.path.exists(GPU), reason="integration/_build/newmsm_gpu not built (needs /root/reference)")
def test_group_members_are_bound_at_link_time():
    """The four groupwise members resolve to the hooks (strong definitions) and the reference's own code stays reachable as `__real_`
    aliases (weakened copies of the compiled reference objects, oracle/Makefile: GROUPSYMS); the reference CPU binary has neither."""
    import shutil
    if not shutil.which("nm"):
        pytest.skip("binutils nm not available")
    syms = ["_ZN10newmeshreg18DiscreteGroupModel14estimate_pairsEv", "_ZN10newmeshreg18DiscreteGroupModel14get_patch_dataEv",
            "_ZN10newmeshreg25DiscreteGroupCostFunction19computePairwiseCostEiii", "_ZN10newmeshreg25DiscreteGroupCostFunction18computeTripletCostEiiii"]
    table = {}
    for line in subprocess.run(["nm", GPU], capture_output=True, text=True, check=True).stdout.splitlines():
        p = line.split()
        if len(p) == 3:
            table[p[2]] = (p[1], p[0])
    for s_ in syms:
        assert table.get(s_, ("?",))[0] == "T", (s_, table.get(s_))                     # the hook: a strong text symbol
        assert table.get("__real_" + s_, ("?",))[0] == "T", s_                          # the reference's body, still linked in
        assert table[s_][1] != table["__real_" + s_][1]                                 # and they are different functions
    ref_syms = subprocess.run(["nm", REF], capture_output=True, text=True, check=True).stdout if os.path.exists(REF) else ""
    assert "__real_" + syms[0] not in ref_syms
