#!/usr/bin/env python
"""Gather locality: the fused ALM inner iteration on a graph WITH locality whose vertex labels were scrambled (rows x cols
torus, n = rows * cols), with and without the library's breadth-first row relabelling.  The factor (8 n r bytes) is far larger
than the L2, so without the relabelling every CSR entry pulls its factor row from DRAM; with it the rows a CSR row gathers
sit within a narrow window.  usage: reorder_bench.py <0|1> [rows] [cols] [rank] [steps]   (GPU box; one JSON line)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
sys.path.insert(0, ROOT)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "1"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    cols = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
    r = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    os.environ["LORADS_REORDER"] = mode
    import lorads_b200 as lb
    import bench
    n = rows * cols
    ei, ej, w = lb.torus_graph(rows, cols, 7)
    lab = np.random.default_rng(7).permutation(n)
    p = lb.maxcut_problem(n, lab[ei], lab[ej], w)
    H = lb.host_lib()
    rng = np.random.default_rng(925)
    R0 = np.asfortranarray(rng.random((n, r)) - rng.random((n, r)))
    rho = 1.0 / np.sqrt(n)
    with lb.Context(0) as ctx:
        t0 = time.perf_counter()
        ctx.load(p)
        t_load = time.perf_counter() - t0
        info = ctx.reorder_info(0)
        ci = ctx.cone_info(0)
        ctx.alloc_vars([r], 2)
        ctx.set_factor(lb.R, 0, R0)
        ctx.set_vec(lb.VEC_DUAL, np.zeros(n))
        ctx.init_constr_val(lb.PAIR_RR)
        ctx.alm_cal_grad(rho)
        k = 0
        for _ in range(3):
            out = bench.alm_iteration(ctx, lb, H, rho, k); k += 1
        ctx.sync()
        ctx.timer_record(0)
        for _ in range(steps):
            out = bench.alm_iteration(ctx, lb, H, rho, k); k += 1
        ctx.timer_record(1)
        ctx.sync()
        ms = ctx.timer_elapsed_ms(0, 1) / steps
        ctx.profile_enable(True)
        for _ in range(steps):
            out = bench.alm_iteration(ctx, lb, H, rho, k); k += 1
        prof = ctx.profile_read()
        ctx.profile_enable(False)
    ld = (r + 3) // 4 * 4
    nnzF = 2 * ci["nnzP"] - n
    alg = 12.0 * nnzF + 4.0 * (n + 1) + 2 * 8.0 * n * ld
    print(json.dumps({"workload": f"{rows}x{cols} torus MaxCut with scrambled labels, n={n}, rank {r}", "LORADS_REORDER": mode,
                      "reorder": info, "layout_upload_s": t_load, "ms_per_iteration": ms, "iterations_per_s": 1e3 / ms,
                      "classes_ms": {c: v[0] / steps for c, v in prof.items() if v[1]},
                      "k_mc_spmm_algorithmic_bytes": alg, "last": list(out)}), flush=True)


if __name__ == "__main__":
    main()
