#!/usr/bin/env python
"""By-cone partition: ALM inner iterations per second of a multi-block problem on 1 .. P GPUs (torchrun, one rank per GPU).
Workload: `nb` blocks, each the MaxCut SDP of a random graph with `n` vertices (average degree 10), rank `r`; every
constraint touches one block, the blocks couple through the length-m constraint vector and the scalar reductions only
(lorads_alg_common.c:221-229).  usage (under torchrun or plain python): bycone_scaling.py [nb] [n] [r] [steps]"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))


def main():
    import torch
    import lorads_b200 as lb
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="cpu:gloo,cuda:nccl")
    p = lb.block_maxcut_problem([(n,) + lb.random_graph(n, 5, s) for s in range(nb)])
    ctx = lb.Context(local)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(lb.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.numpy().tobytes()), rank, world)
    ctx.load(p)
    ctx.alloc_vars([r] * nb, 2)
    rng = np.random.default_rng(925)
    for c in range(nb):
        ctx.set_factor(lb.R, c, rng.random((n, r)) - rng.random((n, r)))
    rho = 1.0 / np.sqrt(n * nb)
    H = lb.host_lib()
    ctx.init_constr_val(lb.PAIR_RR)
    ctx.alm_cal_grad(rho)

    def it(k):
        ctx.lbfgs_direction(k)
        terms = ctx.alm_linesearch_terms(rho)
        tau = ctypes.c_double(0.0)
        H.lh_line_search(float(rho), terms.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(tau))
        return (tau.value,) + ctx.alm_inner_update(rho, tau.value)

    k = 0
    for _ in range(3):
        out = it(k); k += 1
    ctx.sync()
    if dist is not None:
        dist.barrier()
    ctx.timer_record(0)
    for _ in range(steps):
        out = it(k); k += 1
    ctx.timer_record(1)
    ctx.sync()
    ms = ctx.timer_elapsed_ms(0, 1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        print(json.dumps({"workload": f"{nb} blocks x MaxCut(n={n}, avg degree 10), rank {r}: general per-cone path", "n_gpus": world,
                          "partition": "by cone" if world > 1 else "single GPU", "iterations_per_s": steps / (ms * 1e-3),
                          "ms_per_iteration": ms / steps, "last": {"tau": out[0], "grad_norm_sq": out[1], "pinf": out[2]}}), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
