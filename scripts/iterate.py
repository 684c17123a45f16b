#!/usr/bin/env python
"""A short run of the bench workload for ncu: build the synthetic MaxCut problem, `warm` untimed ALM inner iterations, then
`iters` more (4 kernel launches each on the fused path).  usage: iterate.py [n] [rank] [warm] [iters]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
sys.path.insert(0, ROOT)


def main():
    import lorads_b200 as lb
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    H = lb.host_lib()
    p, _ = bench.build_problem(lb, n, 5, 0)
    rng = np.random.default_rng(925)
    R0 = np.asfortranarray(rng.random((n, r)) - rng.random((n, r)))
    rho = 1.0 / np.sqrt(n)
    with lb.Context(0) as ctx:
        ctx.load(p)
        ctx.alloc_vars([r], 2)
        ctx.set_factor(lb.R, 0, R0)
        ctx.set_vec(lb.VEC_DUAL, np.zeros(n))
        ctx.init_constr_val(lb.PAIR_RR)
        ctx.alm_cal_grad(rho)
        for k in range(warm + iters):
            out = bench.alm_iteration(ctx, lb, H, rho, k)
        print("iterate ok", n, r, out)


if __name__ == "__main__":
    main()
