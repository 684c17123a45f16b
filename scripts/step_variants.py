#!/usr/bin/env python
"""A/B of the fused-path kernel variants on the bench workload (GPU box): per-class device time per ALM inner iteration
(CUDA events around every launch, lgpu_profile_*) for each setting of the library's tuning switches.

usage: step_variants.py [n] [rank] [steps]        -> JSON lines on stdout
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
sys.path.insert(0, ROOT)

VARIANTS = [
    ("r1 kernels (register step, separate <D,T>, direction reads R and CR)", {"LORADS_STEP_BULK": "0", "LORADS_SPMM_DOT": "0", "LORADS_ROWDOTS": "0"}),
    ("bulk step, dot mode 3, direction reads R and CR", {"LORADS_ROWDOTS": "0"}),
    ("default: bulk step, dot mode 3, direction from carried row products", {}),
    ("default with separate <D,T>", {"LORADS_SPMM_DOT": "0"}),
]
KEYS = ["LORADS_STEP_BULK", "LORADS_SPMM_DOT", "LORADS_STEP_TILE", "LORADS_STEP_STAGES", "LORADS_STEP_VARIANT", "LORADS_ROWDOTS"]


def main():
    import lorads_b200 as lb
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    H = lb.host_lib()
    p, _ = bench.build_problem(lb, n, 5, 0)
    rng = np.random.default_rng(925)
    R0 = np.asfortranarray(rng.random((n, r)) - rng.random((n, r)))
    rho = 1.0 / np.sqrt(n)
    ref = None
    for name, env in VARIANTS:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        with lb.Context(0) as ctx:
            ctx.load(p)
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
            sig = (out[0], out[1], out[2])
            if ref is None:
                ref = sig
            rel = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(sig, ref))
            print(json.dumps({"variant": name, "env": env, "n": n, "rank": r, "ms_per_step": ms,
                              "classes_ms_per_step": {c: v[0] / steps for c, v in prof.items() if v[1]},
                              "launches_per_step": sum(v[1] for v in prof.values()) / steps,
                              "last_step_scalars_rel_diff_vs_first_variant": rel}), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(json.dumps({"total_s": time.time() - t0}))
