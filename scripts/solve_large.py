#!/usr/bin/env python
"""Whole solve of a large synthetic MaxCut SDP through the drop-in binary with benchmark.py's large-MaxCut flags
(benchmark.py:160-179): --phase1Tol 1e+1 --heuristicFactor 100 --timesLogRank 0.25 --reoptLevel 0.
Usage: python scripts/solve_large.py [n] [out_degree] ; prints one JSON line (setup and solve times, iterations)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
import lorads_b200 as lb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    t0 = time.perf_counter()
    ei, ej, w = lb.random_graph(n, deg, 0)
    p = lb.maxcut_problem(n, ei, ej, w)
    d = tempfile.mkdtemp(prefix="lorads_large_")
    inst = os.path.join(d, f"rand_{n}.dat-s")
    lb.write_sdpa(inst, p)
    t_write = time.perf_counter() - t0
    jf = os.path.join(d, "out.json")
    flags = ["--phase1Tol", "1e+1", "--heuristicFactor", "100", "--timesLogRank", "0.25", "--reoptLevel", "0",
             "--timeSecLimit", "1800", "--jsonfile", jf]
    t0 = time.perf_counter()
    out = lb.run_solver([inst] + flags, timeout=3600)
    wall = time.perf_counter() - t0
    res = {"n": n, "edges": len(ei), "file_mb": os.path.getsize(inst) / 1e6, "generate_write_s": t_write, "process_wall_s": wall,
           "exit": out.returncode}
    for line in out.stdout.splitlines():
        if line.startswith("ALM OuterIter:"):
            res["alm_inner_iters"] = int(line.split("InnerIter:")[1].split()[0])
            res["rank"] = int(line.split("CurrRank:")[1].split()[0])
        elif line.startswith("ADMM Iter:"):
            res["admm_iters"] = int(line.split("Iter:")[1].split()[0]) + 1
            res["cg_iters_avg"] = int(line.split("cgIter:")[1].split()[0])
        elif line.startswith("Reading SDPA file in"):
            res["read_s"] = float(line.split()[4])
        elif line.startswith("all_time:"):
            res["solve_s"] = float(line.split(":")[1])
        elif "1.Primal Objective:" in line:
            res["primal_obj"] = float(line.split(":")[-1])
        elif "2.Dual Objective:" in line:
            res["dual_obj"] = float(line.split(":")[-1])
        elif "1.Constraint Violation(1)" in line:
            res["constr_vio_l1"] = float(line.split(":")[-1])
        elif "3.Primal Dual Gap" in line:
            res["pd_gap"] = float(line.split(":")[-1])
        elif line.startswith("End Program"):
            res["status"] = line.strip()
    if os.path.exists(jf):
        res["json_solve_time_sec"] = json.load(open(jf))["metrics"]["solve_time_sec"]
    print(json.dumps(res))
    sys.stderr.write(out.stdout[-3000:])


if __name__ == "__main__":
    main()
