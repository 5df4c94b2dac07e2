"""End-to-end `newmsm` on a synthetic case: the UNMODIFIED reference CLI on the host cores (oracle/_ref/newmsm_ref_trace) against
the same CLI with libmsmgpu.so bound in at link time (integration/_build/newmsm_gpu, integration/newmsm_gpu_hooks.cpp).
TEST / MEASUREMENT INFRASTRUCTURE: it executes the compiled reference under oracle/_ref as the checker and CPU baseline.
Compares the solver's labeling and the control-point grid after every discrete iteration (exact) and the final sphere.reg,
and reports the wall-clock of both. BASELINE.json metric (iii): "newmsm wall-time vs CPU cores".

    python tests/newmsm_e2e.py --level 6 --config MSMAllStrain --D 40 --threads 16 --out gpurun_out/e2e_cfg3.json
"""
import argparse
import shutil
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "newmsm_ref_trace")
GPU = os.path.join(ROOT, "integration", "_build", "newmsm_gpu")


def parse_trace(path):
    calls = []
    with open(path) as f:
        for line in f:
            if line.startswith("U "):
                p = line.split()
                calls.append({"nv": int(p[3]), "hash": p[5]})
            elif line.startswith("L "):
                calls[-1]["labels"] = np.array(line.split()[1:], dtype=np.int32)
            elif line.startswith("G "):
                calls[-1]["grid"] = np.array([float.fromhex(x) for x in line.split()[1:]])
    return calls


def read_asc(path):
    with open(path) as f:
        f.readline()
        nv, nt = map(int, f.readline().split())
        return np.loadtxt(f, max_rows=nv)[:, :3]


GROUP = False   # --group: the gMSM driver (src/newmsm.cpp:14-28)
ANAT = False    # aMSM: --inanat / --refanat (anatomical strain, regoption 5)
MASK = False    # --mask: groupwise run with a cost mask (src/newmsm.cpp:25, DiscreteGroupModel.cpp:164)


def run(binary, case, conf, out, threads, trace, extra_env=None):
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), MSMGPU_TRACE=trace, MSMGPU_TIMING=os.environ.get("MSMGPU_TIMING", "1"))
    env.update(extra_env or {})
    if GROUP:
        cmd = [binary, "--groupwise", "--meshes=" + os.path.join(case, "meshes.txt"), "--data=" + os.path.join(case, "data.txt"),
               "--template=" + os.path.join(case, "template.asc"), "--conf=" + conf, "--out=" + out + "/"]
        if MASK:
            cmd.append("--mask=" + os.path.join(case, "mask.txt"))
    else:
        cmd = [binary, "--inmesh=" + os.path.join(case, "sphere.asc"), "--refmesh=" + os.path.join(case, "sphere.asc"),
               "--indata=" + os.path.join(case, "indata.txt"), "--refdata=" + os.path.join(case, "refdata.txt"),
               "--conf=" + conf, "--out=" + out + "/", "-f", "ASCII"]
        if ANAT:
            cmd += ["--inanat=" + os.path.join(case, "inanat.asc"), "--refanat=" + os.path.join(case, "refanat.asc")]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{binary} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    lines = r.stderr.strip().splitlines()
    return dt, [ln for ln in lines if ln.startswith("[msmgpu")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--D", type=int, default=1)
    ap.add_argument("--config", default="MSMpair", choices=["MSMpair", "MSMpairAffine", "MSMAllStrain", "MSMstrain", "sMSMSTRcp5", "gMSM", "aMSMSTR"])
    ap.add_argument("--group", type=int, default=0, help="groupwise (gMSM) run with this many subjects (integration/newmsm_gpu_group_hooks.cpp binds "
                                                           "estimate_pairs, get_patch_data and the pair / triplet costs)")
    ap.add_argument("--levels-drop", type=int, default=0)
    ap.add_argument("--it-scale", type=float, default=1.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--parity-threads", type=int, default=1,
                    help="threads of the CPU run the labels are compared with. The reference is only deterministic single-threaded: "
                         "get_adaptive_barycentric_weights accumulates `correction[]` in an omp loop without atomics (resampler.cpp:99-118)")
    ap.add_argument("--skip-timing-cpu", action="store_true", help="do not run the all-threads CPU arm (timing)")
    ap.add_argument("--skip-parity-cpu", action="store_true", help="do not run the single-thread CPU arm (labels are then only compared with the "
                                                                     "all-threads run, which the reference's own data race can perturb)")
    ap.add_argument("--gpu-runs", type=int, default=1)
    ap.add_argument("--disable", default="", help="MSMGPU_DISABLE value for the GPU run (cost, resample): A/B isolation of the hooks")
    ap.add_argument("--verify", action="store_true", help="MSMGPU_VERIFY=1: the hooks also run the reference CPU code in-process and compare")
    ap.add_argument("--out", default="")
    ap.add_argument("--mask", action="store_true", help="groupwise only: run with the case's cost mask (newmsm --mask)")
    ap.add_argument("--devices", type=int, default=1, help="MSMGPU_DEVICES for the GPU run (groupwise: subjects and pair blocks sharded over this many GPUs)")
    ap.add_argument("--skip-gpu", action="store_true", help="CPU arms only (e.g. to record the single-thread trace on a machine without a GPU)")
    ap.add_argument("--gpu-trace-out", default="", help="keep the GPU run's trace here")
    ap.add_argument("--cpu-trace-out", default="", help="keep the single-thread CPU trace here")
    ap.add_argument("--cpu-trace-in", default="", help="compare with this recorded single-thread CPU trace instead of running that arm "
                                                         "(the case is seeded, so the inputs are identical)")
    a = ap.parse_args()
    global GROUP, MASK, ANAT
    ANAT = a.config == "aMSMSTR"
    GROUP = a.group > 0
    MASK = bool(a.mask)
    work = tempfile.mkdtemp(prefix="newmsm_case_")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_newmsm_case.py"), "--out", work, "--level", str(a.level), "--D", str(a.D),
                    "--levels-drop", str(a.levels_drop), "--it-scale", str(a.it_scale), "--group", str(a.group)], check=True, stdout=subprocess.DEVNULL)
    conf = os.path.join(work, "conf_" + a.config)
    with open(conf, "a") as f:
        f.write("--numthreads=%d\n" % a.threads)
    res = {"config": a.config, "level": a.level, "D": a.D, "levels_drop": a.levels_drop, "it_scale": a.it_scale, "host_threads": a.threads, "devices": a.devices,
           "verts": 10 * 4 ** a.level + 2}
    gpu_times = []
    for k in range(0 if a.skip_gpu else a.gpu_runs):
        dt, note = run(GPU, work, conf, os.path.join(work, "out_gpu"), a.threads, os.path.join(work, "trace_gpu.txt"),
                       dict(({"MSMGPU_DISABLE": a.disable} if a.disable else {}), **({"MSMGPU_VERIFY": "1"} if a.verify else {}),
                            **({"MSMGPU_DEVICES": str(a.devices)} if a.devices > 1 else {})))
        gpu_times.append(dt)
        res["gpu_split"] = note
    res["gpu_wall_s"] = min(gpu_times) if gpu_times else float("nan")
    res["gpu_wall_all_s"] = gpu_times
    tg = parse_trace(os.path.join(work, "trace_gpu.txt")) if gpu_times else []
    if gpu_times and a.gpu_trace_out:
        shutil.copy(os.path.join(work, "trace_gpu.txt"), a.gpu_trace_out)
    res["discrete_iterations"] = sum(1 for c in tg if "labels" in c)
    if GROUP:   # no labeling dump in group mode (the hook has no model pointer): every iteration unfolds 2 meshes per subject, whose
        res["discrete_iterations"] = len(tg) // (2 * a.group)   # control grids ROT * label[labeling] encode the labels exactly
    if not a.skip_cpu and not a.skip_timing_cpu:
        dt, _ = run(REF, work, conf, os.path.join(work, "out_cpu_mt"), a.threads, os.path.join(work, "trace_cpu_mt.txt"))
        res["cpu_wall_s"] = dt
        res["cpu_threads"] = a.threads
        res["speedup"] = dt / res["gpu_wall_s"]
        tm = parse_trace(os.path.join(work, "trace_cpu_mt.txt"))
        lab = [(x["labels"], y["labels"]) for x, y in zip(tm, tg) if "labels" in x and "labels" in y]
        res["label_mismatch_vs_multithreaded_cpu"] = [int((x != y).sum()) if len(x) == len(y) else -1 for x, y in lab]
        res["trace_calls_vs_multithreaded_cpu"] = [len(tm), len(tg)]
        res["hashes_equal_vs_multithreaded_cpu"] = sum(1 for x, y in zip(tm, tg) if x["hash"] == y["hash"])
    if not a.skip_cpu and not a.skip_parity_cpu:
        conf1 = conf + "_parity"
        with open(conf) as f:
            lines = [ln for ln in f.read().splitlines() if not ln.startswith("--numthreads")]
        with open(conf1, "w") as f:
            f.write("\n".join(lines + ["--numthreads=%d" % a.parity_threads]) + "\n")
        if a.cpu_trace_in:
            shutil.copy(a.cpu_trace_in, os.path.join(work, "trace_cpu.txt"))
            dt = float("nan")
            res["cpu_parity_trace"] = os.path.basename(a.cpu_trace_in)
        else:
            dt, _ = run(REF, work, conf1, os.path.join(work, "out_cpu"), a.parity_threads, os.path.join(work, "trace_cpu.txt"))
        if a.cpu_trace_out:
            shutil.copy(os.path.join(work, "trace_cpu.txt"), a.cpu_trace_out)
        res["cpu_parity_wall_s"] = dt
        res["cpu_parity_threads"] = a.parity_threads
        tc = parse_trace(os.path.join(work, "trace_cpu.txt"))
        res["trace_calls"] = [len(tc), len(tg)]
        n = min(len(tc), len(tg))
        res["hashes_equal"] = sum(1 for i in range(n) if tc[i]["hash"] == tg[i]["hash"])
        lab = [(x["labels"], y["labels"]) for x, y in zip(tc, tg) if "labels" in x and "labels" in y]
        res["label_iterations"] = len(lab)
        res["label_mismatch_per_iteration"] = [int((x != y).sum()) if len(x) == len(y) else -1 for x, y in lab]
        res["labels_bit_exact"] = len(tc) == len(tg) and all(m == 0 for m in res["label_mismatch_per_iteration"])
        res["all_meshes_bit_exact"] = len(tc) == len(tg) and res["hashes_equal"] == n
        res["first_mesh_mismatch"] = next((i for i in range(n) if tc[i]["hash"] != tg[i]["hash"]), -1)
        if GROUP or a.cpu_trace_in or a.skip_gpu:   # the group driver only writes GIFTI (placeholder files under the FSL stand-in): the traces carry the comparison
            res["final_sphere_max_abs_diff"] = 0.0 if res["all_meshes_bit_exact"] else float("nan")
        else:
            a_, b_ = read_asc(os.path.join(work, "out_cpu", "sphere.reg.asc")), read_asc(os.path.join(work, "out_gpu", "sphere.reg.asc"))
            res["final_sphere_max_abs_diff"] = float(np.abs(a_ - b_).max())
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
