"""Worker for tests/test_partition_gloo.py (CPU, gloo, world_size 2): the host-side logic of the N > 1 paths -- row
ownership from the C ABI's lgpu_partition_rows, CSR slicing with global column ids, the all-gather of the direction's
rows and the all-reduce of the scalar packs, the addressing of the peer-memory halo PUT (dst_off), the rank-order scalar
all-reduce, and the by-cone partition's owner map and m-vector sum -- emulated in numpy and checked against the
single-process oracle.  No GPU work happens here."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lorads_b200 as lb          # noqa: E402
import lorads_oracle as orc       # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. ownership: disjoint, ordered, covering, equal allocation
    for n in (7, 80, 600, 10_000_001):
        lo, hi, rpr = lb.partition_rows(n, world, rank)
        t = torch.tensor([lo, hi, rpr], dtype=torch.int64)
        allr = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, t)
        assert allr[0][0] == 0 and allr[-1][1] == n
        for a, b in zip(allr[:-1], allr[1:]):
            assert a[1] == b[0] and a[2] == b[2]
        assert all(int(a[1] - a[0]) <= int(a[2]) for a in allr) and world * int(allr[0][2]) >= n
    # 2. the 128-byte communicator id travels from rank 0 to everybody
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    assert bytes(uid.numpy().tobytes()) == bytes(range(128))
    # 3. one partitioned ALM inner iteration's quantities against the oracle
    n, r = 600, 13
    ei, ej, w = lb.torus_graph(20, 30, 81)
    p = lb.maxcut_problem(n, ei, ej, w)
    path = f"/tmp/_part_{os.getpid()}.dat-s"
    lb.write_sdpa(path, p)
    q = orc.read_sdpa(path)
    os.remove(path)
    cone = orc.build_cone(q.blocks[0], q.m)
    rng = np.random.default_rng(11)
    R, D = rng.normal(size=(n, r)), rng.normal(size=(n, r))
    lam, cvs = rng.normal(size=n), rng.normal(size=n)
    rho = 0.3
    q1o, q2o, p1o, p2o = orc.q12p12([cone], [R], [D])
    coef_o = orc.line_search_coeffs(rho, lam, p1o, p2o, q.b - cvs, q1o, q2o)
    # local rows of the symmetric CSR of C (global column ids), as lgpu_cone_upload slices them
    lo, hi, rpr = lb.partition_rows(n, world, rank)
    Cd = np.zeros((n, n))
    Cd[cone.pat_row[cone.c_slot], cone.pat_col[cone.c_slot]] = cone.c_val
    Cd = Cd + Cd.T - np.diag(np.diag(Cd))
    Cl = Cd[lo:hi]
    # all-gather of the direction's rows (equal counts: rows_per_rank, zero padded)
    send = torch.zeros((rpr, r), dtype=torch.float64)
    send[:hi - lo] = torch.from_numpy(D[lo:hi])
    parts = [torch.zeros((rpr, r), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, send)
    Dg = torch.cat(parts).numpy()[:n]
    assert np.array_equal(Dg, D)
    T = Cl @ Dg
    rd = np.einsum("ij,ij->i", R[lo:hi], D[lo:hi])
    dd = np.einsum("ij,ij->i", D[lo:hi], D[lo:hi])
    q1, q2 = 2.0 * rd, dd                      # a_k = 1 for MaxCut, constraint k <-> row k
    q0p = (q.b[lo:hi] - cvs[lo:hi]) + lam[lo:hi] / rho
    pack = torch.tensor([2.0 * np.sum(R[lo:hi] * T), np.sum(D[lo:hi] * T), q2 @ q2, q1 @ q2, q0p @ q2, q1 @ q1, q0p @ q1])
    dist.all_reduce(pack)                      # the seven line-search terms, summed over the ranks
    pk = pack.numpy()
    assert np.allclose(q1, q1o[lo:hi], rtol=1e-12) and np.allclose(q2, q2o[lo:hi], rtol=1e-12)
    assert abs(pk[0] - p1o) <= 1e-11 * np.sum(np.abs(R)) and abs(pk[1] - p2o) <= 1e-11 * np.sum(np.abs(D))
    a = rho * pk[2] / 2
    b = rho * pk[3]
    c = pk[1] - rho * pk[4] + rho * pk[5] / 2
    d = pk[0] - rho * pk[6]
    assert np.allclose([a, b, c, d], coef_o, rtol=1e-10)
    # every rank derives the same step length from the same reduced scalars
    tau = orc.line_search_tau(a, b, c, d)[1]
    taus = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(taus, torch.tensor([tau], dtype=torch.float64))
    assert all(float(t[0]) == tau for t in taus)
    # 4. halo PUT addressing (peer-memory exchange): every rank writes the rows its peers reference straight into THEIR halo
    #    buffers at dst_off (derived locally, never communicated); emulated with an all-to-all of the packed rows.  What lands
    #    in my halo must be, row for row, the global rows halo_gid names.
    os.environ["LORADS_HALO"] = "1"
    L = lb.cone_layout(p, 0, ["send_idx", "send_off", "send_cnt", "recv_off", "recv_cnt", "dst_off", "halo_gid"], world, rank)
    halo = np.full((max(len(L["halo_gid"]), 1), r), np.nan)
    for q_ in range(world):
        if q_ == rank:
            continue
        s0, sc = int(L["send_off"][q_]), int(L["send_cnt"][q_])
        rows = torch.from_numpy(np.ascontiguousarray(D[lo + L["send_idx"][s0:s0 + sc].astype(np.int64)]))
        meta = torch.tensor([int(L["dst_off"][q_]), sc], dtype=torch.int64)
        other_meta = torch.zeros(2, dtype=torch.int64)
        # world 2: a symmetric exchange with the one peer
        reqs = [dist.isend(meta, q_), dist.irecv(other_meta, q_)]
        for rq in reqs:
            rq.wait()
        got = torch.zeros((int(other_meta[1]), r), dtype=torch.float64)
        reqs = [dist.isend(rows, q_), dist.irecv(got, q_)]
        for rq in reqs:
            rq.wait()
        off = int(other_meta[0])
        assert off == int(L["recv_off"][q_]) and int(other_meta[1]) == int(L["recv_cnt"][q_])
        halo[off:off + int(other_meta[1])] = got.numpy()
    if len(L["halo_gid"]):
        assert np.array_equal(halo[:len(L["halo_gid"])], D[L["halo_gid"].astype(np.int64)])
    # 5. one-shot scalar all-reduce in RANK ORDER: every rank adds the same numbers in the same order -> same bits
    mine = torch.tensor(np.random.default_rng(100 + rank).normal(size=18) * 1e8, dtype=torch.float64)
    inbox = [torch.zeros(18, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(inbox, mine)
    tot = np.zeros(18)
    for src in range(world):
        tot = tot + inbox[src].numpy()
    allt = [torch.zeros(18, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allt, torch.from_numpy(tot))
    assert all(np.array_equal(t.numpy(), tot) for t in allt)
    # 6. by-cone partition of a multi-block problem: the owners' constraint values, summed over the ranks, are the whole
    #    A(RR^T); gradient segments travel as "owner's value + zeros" and arrive bit for bit
    qm = orc.read_sdpa(os.path.join(ROOT, "tests", "golden", "instances", "multiblock_lp.dat-s"))
    cones = [orc.build_cone(bk, qm.m) for bk in qm.blocks]
    cost = [float((c_.nnzP + len(c_.a_slot) + c_.n) * 8) for c_ in cones]
    owner = lb.cone_owner_map(cost, world)
    assert set(owner.tolist()) == set(range(world))
    rngc = np.random.default_rng(5)
    Rs = [rngc.normal(size=(c_.n, 4)) for c_ in cones]
    whole = np.zeros(qm.m)
    part = np.zeros(qm.m)
    for k_, c_ in enumerate(cones):
        cv = orc.cone_auv(c_, orc.uvt(c_, Rs[k_], Rs[k_]))
        whole += cv
        if owner[k_] == rank:
            part += cv
    tpart = torch.from_numpy(part.copy())
    dist.all_reduce(tpart)
    assert np.allclose(tpart.numpy(), whole, rtol=1e-13, atol=1e-13 * np.max(np.abs(whole)))
    seg = torch.from_numpy(Rs[0].copy() if owner[0] == rank else np.zeros_like(Rs[0]))
    dist.all_reduce(seg)
    assert np.array_equal(seg.numpy(), Rs[0])
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("PARTITION_OK")


if __name__ == "__main__":
    main()
