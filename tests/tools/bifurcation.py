#!/usr/bin/env python
"""Rounding sensitivity of the REFERENCE's own trajectories (build container only; reads oracle/_ref built from
/root/reference).  Runs two legitimate builds of the unmodified reference sources -- `lorads_ref` (gcc -O2) and
`lorads_ref_fma` (gcc -O3 -mfma -ffp-contract=fast, oracle/Makefile) -- on the same file, flags and seed, and reports
where their logs first differ, their iteration counts and their objectives.  An instance on which the reference
disagrees with ITSELF by more than north-star's +-5 % / 1e-6 cannot be held to that tolerance by any other
implementation; the GPU parity tests (tests/test_gpu_solves.py) read the verdicts this script commits to
tests/golden/bifurcation.json.

usage: bifurcation.py [limit_s] [instance ...]      -> profiles/r2_bifurcation.md, tests/golden/bifurcation.json
"""
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DATA = os.path.join(ROOT, "bench_data")
INST = os.path.join(ROOT, "tests", "golden", "instances")
REF = os.path.join(ROOT, "oracle", "_ref")

GSET = ["--phase1Tol", "1e-2", "--heuristicFactor", "10"]
# (name, file, flags)
CASES = [
    ("G1", os.path.join(INST, "G1.dat-s"), GSET + ["--reoptLevel", "0"]),
    ("G11", os.path.join(DATA, "G11.dat-s"), GSET),
    ("control_like_12_6", os.path.join(INST, "control_like_12_6.dat-s"), []),
    ("multiblock_sdp", os.path.join(INST, "multiblock_sdp.dat-s"), []),
    ("torus_100x200", None, GSET + ["--reoptLevel", "0"]),
    ("MC_500", os.path.join(DATA, "MC_500.dat-s"), []),
    ("checker_1.5", os.path.join(DATA, "checker_1.5.dat-s"), []),
    ("ice_2.0", os.path.join(DATA, "ice_2.0.dat-s"), []),
    ("p_auss2_3.0", os.path.join(DATA, "p_auss2_3.0.dat-s"), []),
    ("cphil12", os.path.join(DATA, "cphil12.dat-s"), []),
]

# builds / BLAS kernel sets of the UNMODIFIED reference: all legitimate, all the same arithmetic up to rounding
VARIANTS = [("O2", "lorads_ref", {}), ("FMA", "lorads_ref_fma", {}),
            ("O2, OpenBLAS Haswell kernels", "lorads_ref", {"OPENBLAS_CORETYPE": "Haswell"}),
            ("O2, OpenBLAS SkylakeX kernels", "lorads_ref", {"OPENBLAS_CORETYPE": "SkylakeX"}),
            ("O2, OpenBLAS Nehalem kernels", "lorads_ref", {"OPENBLAS_CORETYPE": "Nehalem"})]


def summary(out):
    res = {"alm_inner": None, "admm": None, "obj": None, "status": None, "rank_first": None, "rank_last": None}
    for line in out.splitlines():
        if "OuterIter:" in line and "InnerIter:" in line:
            res["alm_inner"] = int(line.split("InnerIter:")[1].split()[0])
            res["rank_last"] = int(line.split("CurrRank:")[1].split()[0])
            if res["rank_first"] is None:
                res["rank_first"] = res["rank_last"]
        elif "2.Dual Objective:" in line:
            res["dobj"] = float(line.split(":")[-1])
        elif "1.Constraint Violation(1)" in line:
            res["constr_vio_l1"] = float(line.split(":")[-1])
        elif "3.Primal Dual Gap" in line:
            res["pd_gap"] = float(line.split(":")[-1])
        elif line.startswith("ADMM Iter:"):
            res["admm"] = int(line.split("Iter:")[1].split()[0]) + 1
        elif "1.Primal Objective:" in line:
            res["obj"] = float(line.split(":")[-1])
        elif line.startswith("End Program"):
            res["status"] = line.strip()
        elif line.startswith("all_time:"):
            res["solve_s"] = float(line.split(":")[1])
    return res


def iteration_lines(out):
    """the per-iteration log lines without their wall-clock field"""
    return [re.sub(r"\s*Time:\s*\S+", "", ln) for ln in out.splitlines() if "OuterIter:" in ln or ln.startswith("ADMM Iter:")]


def first_difference(a, b):
    la, lb_ = iteration_lines(a), iteration_lines(b)
    for k, (x, y) in enumerate(zip(la, lb_)):
        if x != y:
            return k, x, y
    if len(la) != len(lb_):
        return min(len(la), len(lb_)), "(end)", "(end)"
    return None


def main():
    limit = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    names = set(sys.argv[2:])
    sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
    rows, verdicts = [], {}
    for name, path, flags in CASES:
        if names and name not in names:
            continue
        if path is None:
            import lorads_b200 as lb
            ei, ej, w = lb.torus_graph(100, 200, 81)
            path = "/tmp/bif_torus_100x200.dat-s"
            lb.write_sdpa(path, lb.maxcut_problem(20000, ei, ej, w))
        if not os.path.exists(path):
            print("missing", path)
            continue
        outs, sums = {}, {}
        for vname, exe, extra in VARIANTS:
            env = dict(os.environ, OPENBLAS_NUM_THREADS="1", **extra)
            t0 = time.perf_counter()
            r = subprocess.run([os.path.join(REF, exe), path] + flags + ["--timeSecLimit", str(limit)], capture_output=True,
                               text=True, env=env, timeout=limit + 600)
            outs[vname] = r.stdout
            sums[vname] = summary(r.stdout)
            print(name, vname, f"{time.perf_counter() - t0:.1f}s", sums[vname], flush=True)
        a, b = sums["O2"], sums["FMA"]
        diff = first_difference(outs["O2"], outs["FMA"])
        inner = [v["alm_inner"] for v in sums.values() if v["alm_inner"]]
        objs = [v["obj"] for v in sums.values() if v["obj"] is not None]
        rel_it = (max(inner) - min(inner)) / max(a["alm_inner"], 1) if inner and a["alm_inner"] else None
        rel_obj = (max(objs) - min(objs)) / max(abs(a["obj"]), 1.0) if objs else None
        stable = rel_it is not None and rel_it <= 0.05 and rel_obj <= 1e-6 and len({v["status"] for v in sums.values()}) == 1
        verdicts[name] = {"stable": bool(stable), "ref": a, "ref_fma": b, "variants": sums, "rel_iter_diff": rel_it, "rel_obj_diff": rel_obj,
                          "inner_min": min(inner) if inner else None, "inner_max": max(inner) if inner else None,
                          "obj_min": min(objs) if objs else None, "obj_max": max(objs) if objs else None,
                          "first_different_log_line": None if diff is None else diff[0], "flags": " ".join(flags)}
        rows.append((name, a, b, rel_it, rel_obj, diff, stable))
    jpath = os.path.join(ROOT, "tests", "golden", "bifurcation.json")
    if names and os.path.exists(jpath):     # a partial re-run keeps the other instances' records
        with open(jpath) as f:
            old = json.load(f)
        old.update(verdicts)
        verdicts = old
    with open(jpath, "w") as f:
        json.dump(verdicts, f, indent=1, sort_keys=True)
    with open(os.path.join(ROOT, "profiles", "r2_bifurcation.md"), "w") as f:
        f.write("# Round 2 -- the reference against ITSELF\n\n"
                "The unmodified reference sources in five legitimate forms (`tests/tools/bifurcation.py`): `gcc -O2` (`lorads_ref`), "
                "`gcc -O3 -mfma -ffp-contract=fast` (`lorads_ref_fma`, `oracle/Makefile`), and the `-O2` binary with OpenBLAS told to use "
                "its Haswell / SkylakeX / Nehalem kernels (`OPENBLAS_CORETYPE`: other summation orders inside ddot / daxpy / dsymm).  "
                "Same file, flags, seed, one core of the build container.\n\n"
                "An instance is *stable* when all five runs agree within north-star's own tolerance (ALM inner iterations +-5 %, "
                "objective 1e-6 relative, same status).  The GPU whole-solve tests (`tests/test_gpu_solves.py`) hold our solver to "
                "north-star's tolerance on the stable instances, and on the others to: same status, an objective inside the range the "
                "reference itself spans (widened by that range), an iteration count within [1/2 min, 2 max] of the reference's.\n\n"
                "| instance | flags | ALM inner its: O2 / FMA / min..max over the five | objective: O2 / min..max | iteration spread | objective spread | stable |\n"
                "|---|---|---|---|---|---|---|\n")
        for name, v in verdicts.items():
            a, b, rel_it, rel_obj = v["ref"], v["ref_fma"], v["rel_iter_diff"], v["rel_obj_diff"]
            f.write(f"| {name} | `{v['flags']}` | {a['alm_inner']} / {b['alm_inner']} / {v.get('inner_min')}..{v.get('inner_max')} | "
                    f"{a['obj']} / {v.get('obj_min')}..{v.get('obj_max')} | {'' if rel_it is None else f'{100 * rel_it:.1f} %'} | "
                    f"{'' if rel_obj is None else f'{rel_obj:.1e}'} | {'yes' if v['stable'] else 'NO'} |\n")
        f.write("\nPer variant (ALM inner / ADMM iterations, objective):\n\n")
        for name, v in verdicts.items():
            f.write(f"* {name}: " + "; ".join(f"{k}: {x['alm_inner']} / {x['admm']}, {x['obj']}" for k, x in v.get("variants", {}).items()) + "\n")


if __name__ == "__main__":
    main()
