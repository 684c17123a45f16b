"""Worker for tests/test_gpu_multi.py::test_by_cone_partition_vs_reference_goldens: torchrun, one rank per GPU.
Problems that are not the single MaxCut-type cone are partitioned BY CONE (each cone's operator work on its owner, the
length-m constraint values and the objective scalars all-reduced).  Every rank runs the exact call sequence of
tests/test_gpu_parity.py::test_alm_admm_sequence_vs_reference -- the reference's own trace: gradient, five ALM inner
iterations, objective, oracle rank, dual update, hand-off, ADMM sweep with CG, rank augmentation -- on a context joined
to the ranks' communicator, and must meet the same tolerances against the reference-generated golden vectors."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lorads_b200 as lb  # noqa: E402
import pytest  # noqa: E402
import test_gpu_parity as T  # noqa: E402

FIXTURES = ["multiblock_sdp", "multiblock_lp", "control_like_12_6", "general_sparse_n60", "theta_n30", "dense_constraint_n24"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl")

    def factory():
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(lb.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(uid, 0)
        ctx = lb.Context(local)
        ctx.comm_init(bytes(uid.numpy().tobytes()), rank, world)
        return ctx

    T.CTX_FACTORY[0] = factory
    ok, ran = True, 0
    for name in FIXTURES:
        for mode in ("split_calls", "inner_update"):
            try:
                T.test_alm_admm_sequence_vs_reference(lb, name, mode)
                ran += 1
                if rank == 0:
                    print("BY_CONE_PASS", name, mode, flush=True)
            except pytest.skip.Exception:
                pass
            except BaseException as e:  # noqa: BLE001 -- an assertion on one rank must still reach the verdict below
                ok = False
                print(f"BY_CONE_FAIL rank {rank} {name} {mode}: {type(e).__name__}: {str(e)[:600]}", flush=True)
                break
        if not ok:
            break
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"BY_CONE_OK {ran} sequences" if int(flag[0]) == 1 else "BY_CONE_FAILED", flush=True)
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
