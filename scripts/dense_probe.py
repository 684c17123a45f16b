#!/usr/bin/env python
"""Dense-aggregate cone probe: theta-type SDP (C dense) at n = 3000, r = 32; times LORADSUVt (SYR2K) and
(C + A^*(w)) X (SYMM) on the DMMA path and on the pattern-gather path through the host-buffer operator entry points'
device kernels (per-launch CUDA events).  Used under ncu for the FP64 tensor-pipe counters (profiles/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
import lorads_b200 as lb  # noqa: E402


def theta_problem(n, p_edge, seed):
    rng = np.random.default_rng(seed)
    iu, ju = np.triu_indices(n, 1)
    keep = rng.random(len(iu)) < p_edge
    ei, ej = iu[keep], ju[keep]
    m = 1 + len(ei)
    ti, tj = np.tril_indices(n)
    obj_idx = np.sort(lb.pack_idx(n, ti.astype(np.int64), tj.astype(np.int64)))
    k = np.arange(n, dtype=np.int64)
    idx = np.concatenate([obj_idx, lb.pack_idx(n, k, k), lb.pack_idx(n, ej.astype(np.int64), ei.astype(np.int64))])
    val = np.concatenate([-np.ones(len(obj_idx)), np.ones(n), np.ones(len(ei))])
    beg = np.concatenate([[0, len(obj_idx), len(obj_idx) + n], len(obj_idx) + n + 1 + np.arange(len(ei), dtype=np.int64)])
    b = np.zeros(m)
    b[0] = 1.0
    return lb.SdpaProblem(m, [n], b, [beg], [idx], [val])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    r, reps = 32, 5
    p = theta_problem(n, 0.05, 1)
    rng = np.random.default_rng(0)
    Um, Vm = rng.normal(size=(n, r)), rng.normal(size=(n, r))
    w = rng.normal(size=p.m)
    out = {"n": n, "r": r, "m": p.m}
    res = {}
    for tensor in (True, False):
        ctx = lb.Context(0).load(p)
        ctx.set_dense_tensor_path(tensor)
        ctx.profile_enable(True)
        for _ in range(reps):
            t = ctx.op_uvt(0, Um, Vm)
            Y = ctx.op_wsum_mulrk(0, w, Vm, True)
        prof = ctx.profile_read()
        res[tensor] = (t, Y)
        out["tensor" if tensor else "gather"] = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in prof.items() if v[1]}
        ctx.close()
    out["uvt_rel_diff"] = float(np.max(np.abs(res[True][0] - res[False][0])) / np.max(np.abs(res[False][0])))
    out["symm_rel_diff"] = float(np.max(np.abs(res[True][1] - res[False][1])) / np.max(np.abs(res[False][1])))
    flops_syr2k, flops_symm = 4.0 * n * (n + 1) / 2 * 32, 2.0 * n * n * 32
    d = out["tensor"].get("k_dense_dmma")
    if d:
        out["note"] = f"k_dense_dmma averages one SYR2K ({flops_syr2k:.3g} flop) and one SYMM ({flops_symm:.3g} flop) launch"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
