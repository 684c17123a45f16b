#!/usr/bin/env python
"""Whole solves on the instances bundled with the reference (lorads/data): G-set / Mittelmann / SDPLIB-type SDPA files
and the SuiteSparse MaxCut graphs (.mat -> SDPA with the gen_MaxCut.jl convention).

  stage      (build container only) copy / convert the instances into bench_data/ (git-ignored; travels with gpurun)
  ours       run the drop-in binary on every staged instance           -> JSON lines on stdout
  reference  run oracle/_ref/lorads_ref (lorads_ref64 for n > 46340)   -> JSON lines on stdout
Flags per family follow benchmark.py:146-200."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DATA = os.path.join(ROOT, "bench_data")
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))

GSET = ["--phase1Tol", "1e-2", "--heuristicFactor", "10"]
LARGE = ["--phase1Tol", "1e+1", "--heuristicFactor", "100", "--timesLogRank", "0.25"]
FLAGS = {"G11": GSET, "G12": GSET, "G13": GSET, "delaunay_n10": GSET, "delaunay_n11": GSET, "delaunay_n12": GSET,
         "delaunay_n13": GSET, "delaunay_n14": LARGE, "rgg_n_2_15_s0": LARGE, "p2p-Gnutella04": LARGE, "p2p-Gnutella05": LARGE,
         "amazon0302": LARGE, "vsp_befref_fxm_2_4_air02": LARGE, "MC_500": [], "checker_1.5": [], "ice_2.0": [],
         "p_auss2_3.0": [], "cphil12": [], "shmup4": [], "theta102": []}


def stage():
    import shutil
    import numpy as np
    import scipy.io as sio
    import scipy.sparse as sp
    import lorads_b200 as lb
    ref = "/root/reference/lorads/data"
    os.makedirs(DATA, exist_ok=True)
    for sub in ("Max_cut_SDP", "General_SDP", "Matrix_Completion_SDP"):
        for f in sorted(os.listdir(os.path.join(ref, sub))):
            if f.endswith(".dat-s"):
                shutil.copy(os.path.join(ref, sub, f), os.path.join(DATA, f))
    for f in sorted(os.listdir(os.path.join(ref, "Max_cut_matrix_files"))):
        if not f.endswith(".mat"):
            continue
        A = sio.loadmat(os.path.join(ref, "Max_cut_matrix_files", f))["Problem"]["A"][0, 0]
        A = sp.csr_matrix(A).astype(np.float64)
        W = A.maximum(A.T)                       # undirected union of the (possibly directed) edges
        W = sp.triu(W, 1).tocoo()
        n = A.shape[0]
        p = lb.maxcut_problem(n, W.row.astype(np.int64), W.col.astype(np.int64), W.data)
        lb.write_sdpa(os.path.join(DATA, f[:-4] + ".dat-s"), p)
        print("staged", f, n, len(W.data), flush=True)


def parse(out):
    res = {}
    for line in out.splitlines():
        if "OuterIter:" in line and "InnerIter:" in line:   # the 64-bit reference build prints no "ALM " prefix
            res["alm_inner_iters"] = int(line.split("InnerIter:")[1].split()[0])
            res["rank"] = int(line.split("CurrRank:")[1].split()[0])
        elif line.startswith("ADMM Iter:"):
            res["admm_iters"] = int(line.split("Iter:")[1].split()[0]) + 1
        elif line.startswith("all_time:"):
            res["solve_s"] = float(line.split(":")[1])
        elif line.startswith("all_dual_infea:"):
            res["dual_infeasibility_s"] = float(line.split(":")[1])
        elif "1.Primal Objective:" in line:
            res["primal_obj"] = float(line.split(":")[-1])
        elif "2.Dual Objective:" in line:
            res["dual_obj"] = float(line.split(":")[-1])
        elif "1.Constraint Violation(1)" in line:
            res["constr_vio_l1"] = float(line.split(":")[-1])
        elif "2.Dual Infeasibility(1)" in line:
            res["dual_infeas_l1"] = float(line.split(":")[-1])
        elif "3.Primal Dual Gap" in line:
            res["pd_gap"] = float(line.split(":")[-1])
        elif line.startswith("End Program"):
            res["status"] = line.strip().replace("End Program ", "")
        elif line.startswith("nConstrs"):
            res["sizes"] = line.strip()
    return res


def run(arm, names, limit):
    import lorads_b200 as lb
    for name in names:
        inst = os.path.join(DATA, name + ".dat-s")
        if not os.path.exists(inst):
            continue
        flags = FLAGS.get(name, []) + ["--timeSecLimit", str(limit)]
        t0 = time.perf_counter()
        try:
            if arm == "ours":
                out = lb.run_solver([inst] + flags, timeout=limit + 300)
            else:
                big = os.path.getsize(inst) > 0 and name in ("amazon0302", "rgg_n_2_15_s0") or False
                with open(inst) as f:
                    f.readline(); f.readline(); dims = f.readline()
                nmax = max(abs(int(t)) for t in dims.replace("{", " ").replace("}", " ").replace(",", " ").split())
                exe = os.path.join(ROOT, "oracle", "_ref", "lorads_ref64" if nmax > 46340 else "lorads_ref")
                env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
                out = subprocess.run([exe, inst] + flags, capture_output=True, text=True, timeout=limit + 600, env=env)
            res = parse(out.stdout)
            res["exit"] = out.returncode
        except subprocess.TimeoutExpired:
            res = {"status": "killed by the harness timeout"}
        res.update({"instance": name, "arm": arm, "process_wall_s": time.perf_counter() - t0, "flags": " ".join(FLAGS.get(name, []))})
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    arm = sys.argv[1]
    if arm == "stage":
        stage()
    else:
        limit = int(sys.argv[2]) if len(sys.argv) > 2 else 300
        names = sys.argv[3:] or list(FLAGS.keys())
        run(arm, names, limit)
