"""Device layout of a cone, built on the host (csrc/lgpu_layout.h) and read back through lgpu_cone_layout_* without a
GPU, against an independent numpy construction from the definitions: the union pattern and its (row, col), C on the
pattern, the constraint CSR and its transpose by pattern slot, the full symmetric CSR, the fused MaxCut-type arrays,
and for row-block partitioned runs the per-rank slices and the halo-exchange plan.  SURVEY 8f-3: this preprocessing
replaces AConeProcData / AConePresolveData (lorads_sdp_conic.c:1185-1393) and is multi-threaded; the arrays must not
depend on the thread count."""
import numpy as np
import pytest

from conftest import FIXTURES, inst_path


def _unpack(n, packed):
    """packed lower-triangular index (column-major) -> (row, col)   PACK_IDX, lorads_utils.h:167"""
    j = np.arange(n + 1, dtype=np.int64)
    start = j * (2 * n - j + 1) // 2
    col = np.searchsorted(start, packed, side="right") - 1
    return (packed - start[col] + col).astype(np.int64), col.astype(np.int64)


def expected_layout(p, c):
    n, m = int(p.dims[c]), p.m
    beg = p.mat_beg[c]
    idx, val = p.mat_idx[c].copy(), p.mat_elem[c].copy()
    for k in range(m + 1):  # per-column ascending, ties in input order
        o = np.argsort(idx[beg[k]:beg[k + 1]], kind="stable")
        idx[beg[k]:beg[k + 1]] = idx[beg[k]:beg[k + 1]][o]
        val[beg[k]:beg[k + 1]] = val[beg[k]:beg[k + 1]][o]
    tri = n * (n + 1) // 2
    lens = np.diff(beg)
    any_dense = bool(np.any(lens > 0.1 * tri))
    pat = np.unique(idx)
    dense = n < 20 or any_dense or len(pat) / tri >= 0.1
    if dense:
        pat = np.arange(tri, dtype=np.int64)
    E = {}
    prow, pcol = _unpack(n, pat)
    E["pat_row"], E["pat_col"] = prow, pcol
    slot = np.searchsorted(pat, idx)
    dg = prow[slot] == pcol[slot]
    no = int(beg[1])
    cval = np.zeros(len(pat))
    np.add.at(cval, slot[:no], val[:no])
    E["cval"] = cval
    E["c_slot"] = slot[:no]
    E["c_coef"] = np.where(dg[:no], val[:no], 2.0 * val[:no])
    E["_c_val"] = val[:no]
    nz = np.nonzero(lens[1:])[0]
    E["con_gid"] = nz
    E["a_ptr"] = np.concatenate([[0], np.cumsum(lens[1:][nz])])
    E["a_slot"] = slot[no:]
    E["a_coef"] = np.where(dg[no:], val[no:], 2.0 * val[no:])
    ent_con = np.repeat(np.arange(len(nz)), lens[1:][nz])  # local constraint of every entry
    o = np.argsort(E["a_slot"], kind="stable")
    E["t_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(E["a_slot"], minlength=len(pat)))])
    E["t_loc"], E["t_gid"], E["t_val"] = ent_con[o], nz[ent_con[o]], val[no:][o]
    k = np.arange(len(pat))
    off = prow != pcol
    fr = np.concatenate([prow, pcol[off]])
    fc = np.concatenate([pcol, prow[off]])
    fs = np.concatenate([k, k[off]])
    diag_only = len(nz) > 0 and bool(np.all(lens[1:][nz] == 1)) and bool(np.all(dg[no:]))
    # rows sorted by column; in the fused (diagonal-constraint) layout a row's diagonal entry comes last
    o = np.lexsort((fc, (fc == fr) if diag_only else np.zeros(len(fc), bool), fr))
    E["f_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(fr, minlength=n))])
    E["f_col"], E["f_slot"] = fc[o], fs[o]
    S = dict(mA=len(nz), nnzP=len(pat), nnzA=len(idx) - no, nnzC=no, nnzF=len(fc), dense=float(dense),
             diag_only=float(diag_only), sparse_container=float(not (len(nz) > 0.3 * m)),
             max_con_len=int(lens[1:].max()) if m else 0, max_slot_len=int(np.diff(E["t_ptr"]).max()) if len(pat) else 0,
             c_nrminf=float(np.max(np.abs(val[:no]))) if no else 0.0)
    if diag_only:
        E["d_row"], E["d_val"] = prow[E["a_slot"]], val[no:]
        E["mc_val"] = cval[E["f_slot"]]
        o = np.argsort(E["d_row"], kind="stable")
        E["rc_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(E["d_row"], minlength=n))])
        E["rc_gid"], E["rc_a"] = nz[o], E["d_val"][o]
    else:
        for nm in ("d_row", "d_val", "mc_val", "rc_ptr", "rc_gid", "rc_a"):
            E[nm] = np.zeros(0)
    return E, S


def check_single(lb, p, c):
    got = lb.cone_layout(p, c, list(lb.LAYOUT_ARRAYS) + ["scalars"])
    E, S = expected_layout(p, c)
    for nm in lb.LAYOUT_ARRAYS:
        assert len(got[nm]) == len(E[nm]), nm
        assert np.array_equal(got[nm], E[nm].astype(got[nm].dtype)), nm
    for k, v in S.items():
        assert got["scalars"][k] == v, (k, got["scalars"][k], v)
    # norms of C with off-diagonal entries counted twice
    w = np.where(E["pat_row"][E["c_slot"]] == E["pat_col"][E["c_slot"]], 1.0, 2.0)
    v = E["_c_val"]
    assert np.isclose(got["scalars"]["c_nrm1"], np.sum(w * np.abs(v)), rtol=1e-13)
    assert np.isclose(got["scalars"]["c_nrm2sq"], np.sum(w * v * v), rtol=1e-13)
    return got


@pytest.mark.parametrize("name", FIXTURES)
def test_layout_of_golden_instances(built, name):
    p = built.read_sdpa(inst_path(name))
    for c in range(p.ncones):
        check_single(built, p, c)


def _random_problem(lb, rng, n, m, per_con, nobj, shuffle):
    """general sparse cone: objective with duplicated positions, constraints with a few entries, some empty"""
    tri = n * (n + 1) // 2
    cols_idx, cols_val = [], []
    oi = rng.integers(0, tri, size=nobj)
    oi = np.concatenate([oi, oi[: nobj // 7]])  # duplicates of one position add up
    cols_idx.append(oi)
    cols_val.append(rng.normal(size=len(oi)))
    for k in range(m):
        cnt = 0 if k % 5 == 3 else int(rng.integers(1, per_con + 1))
        ci = rng.choice(tri, size=cnt, replace=False)
        cols_idx.append(ci)
        cols_val.append(rng.normal(size=cnt))
    if not shuffle:
        for k in range(m + 1):
            o = np.argsort(cols_idx[k], kind="stable")
            cols_idx[k], cols_val[k] = cols_idx[k][o], cols_val[k][o]
    beg = np.concatenate([[0], np.cumsum([len(a) for a in cols_idx])]).astype(np.int64)
    return lb.SdpaProblem(m, [n], rng.normal(size=m), [beg], [np.concatenate(cols_idx).astype(np.int64)],
                          [np.concatenate(cols_val)])


@pytest.mark.parametrize("threads", ["1", "4", "7"])
@pytest.mark.parametrize("shuffle", [False, True])
def test_layout_general_sparse_random(built, monkeypatch, threads, shuffle):
    monkeypatch.setenv("LORADS_HOST_THREADS", threads)
    rng = np.random.default_rng(11)
    p = _random_problem(built, rng, n=300, m=400, per_con=6, nobj=900, shuffle=shuffle)
    got = check_single(built, p, 0)
    assert not got["scalars"]["dense"] and not got["scalars"]["diag_only"]


def test_layout_dense_and_tiny_cones(built):
    rng = np.random.default_rng(5)
    # n < 20: dense aggregate by rule; n = 40 with a pattern above 10 % of the triangle: dense by fill
    for n, nobj in ((12, 20), (40, 400)):
        p = _random_problem(built, rng, n=n, m=30, per_con=3, nobj=nobj, shuffle=True)
        got = check_single(built, p, 0)
        assert got["scalars"]["dense"] and got["scalars"]["nnzP"] == n * (n + 1) // 2


def _partition_expect(lb, E, n, world, rank, force):
    lo, hi, rpr = lb.partition_rows(n, world, rank)
    fp, fc = E["f_ptr"], E["f_col"]
    own = lambda q: lb.partition_rows(n, world, q)[:2]
    ref = lambda q: np.unique(fc[fp[own(q)[0]]:fp[own(q)[1]]])
    mine = ref(rank)
    halo = mine[(mine < lo) | (mine >= hi)]  # owners are ascending row ranges: rank order = ascending global row
    X = dict(lf_ptr=fp[lo:hi + 1] - fp[lo], lmc_val=E["mc_val"][fp[lo]:fp[hi]], lrc_ptr=E["rc_ptr"][lo:hi + 1] - E["rc_ptr"][lo],
             lrc_gid=E["rc_gid"][E["rc_ptr"][lo]:E["rc_ptr"][hi]], lrc_a=E["rc_a"][E["rc_ptr"][lo]:E["rc_ptr"][hi]])
    recv_cnt = np.array([0 if q == rank else int(np.sum((halo >= own(q)[0]) & (halo < own(q)[1]))) for q in range(world)])
    send = [np.zeros(0, np.int64) if q == rank else (lambda r: r[(r >= lo) & (r < hi)] - lo)(ref(q)) for q in range(world)]
    X["recv_cnt"], X["recv_off"] = recv_cnt, np.concatenate([[0], np.cumsum(recv_cnt)[:-1]])
    X["send_cnt"] = np.array([len(s) for s in send])
    X["send_off"] = np.concatenate([[0], np.cumsum(X["send_cnt"])[:-1]])
    all_halo = sum(int(np.sum((ref(q) < own(q)[0]) | (ref(q) >= own(q)[1]))) for q in range(world))
    use_halo = all_halo < 0.85 * n * (world - 1) if force is None else bool(int(force))
    cols = fc[fp[lo]:fp[hi]]
    if use_halo:
        remap = np.full(n, -1, np.int64)
        remap[lo:hi] = np.arange(hi - lo)
        remap[halo] = rpr + np.arange(len(halo))
        X["lf_col"], X["halo_gid"], X["send_idx"] = remap[cols], halo, np.concatenate(send)
    else:
        X["lf_col"] = cols
    return X, use_halo, len(halo)


@pytest.mark.parametrize("graph", ["torus", "random"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("force", [None, "0", "1"])
def test_layout_partitioned_maxcut(built, monkeypatch, graph, world, force):
    lb = built
    if force is None:
        monkeypatch.delenv("LORADS_HALO", raising=False)
    else:
        monkeypatch.setenv("LORADS_HALO", force)
    n = 600
    ei, ej, w = lb.torus_graph(20, 30, 3) if graph == "torus" else lb.random_graph(n, 4, 3)
    p = lb.maxcut_problem(n, ei, ej, w)
    E, S = expected_layout(p, 0)
    assert S["diag_only"]
    for rank in range(world):
        got = lb.cone_layout(p, 0, list(lb.LAYOUT_ARRAYS_PARTITIONED) + ["scalars"], world, rank)
        X, use_halo, nhalo = _partition_expect(lb, E, n, world, rank, force)
        assert bool(got["scalars"]["use_halo"]) == use_halo
        if force is None:
            assert use_halo == (graph == "torus")  # a grid has a thin boundary, a random graph references everything
        assert got["scalars"]["halo_rows"] == nhalo
        for nm, want in X.items():
            assert np.array_equal(got[nm], np.asarray(want).astype(got[nm].dtype)), (nm, rank)


def test_layout_does_not_depend_on_thread_count(built, monkeypatch):
    lb = built
    n = 40000
    ei, ej, w = lb.random_graph(n, 5, 1)
    p = lb.maxcut_problem(n, ei, ej, w)
    names = list(lb.LAYOUT_ARRAYS) + list(lb.LAYOUT_ARRAYS_PARTITIONED)
    ref = None
    for thr in ("1", "3", "8"):
        monkeypatch.setenv("LORADS_HOST_THREADS", thr)
        got = lb.cone_layout(p, 0, names, 4, 2)
        if ref is None:
            ref = got
            E, _ = expected_layout(p, 0)
            for nm in lb.LAYOUT_ARRAYS:
                assert np.array_equal(got[nm], E[nm].astype(got[nm].dtype)), nm
        else:
            for nm in names:
                assert np.array_equal(got[nm], ref[nm]), (nm, thr)


def test_general_cones_are_not_row_sliced(built):
    """only the single MaxCut-type cone is partitioned by rows; any other cone keeps its whole layout on every rank (the
    solver then partitions BY CONE): with world = 2 the arrays equal the one-GPU ones and no row-block slice exists"""
    rng = np.random.default_rng(2)
    p = _random_problem(built, rng, n=60, m=40, per_con=3, nobj=50, shuffle=False)
    names = ["f_ptr", "f_col", "f_slot", "a_ptr", "a_slot", "t_ptr", "pat_row", "pat_col", "lf_ptr", "send_idx"]
    one = built.cone_layout(p, 0, names, world=1, rank=0)
    for r in range(2):
        two = built.cone_layout(p, 0, names, world=2, rank=r)
        for nm in names:
            assert np.array_equal(one[nm], two[nm]), nm
        assert len(two["lf_ptr"]) == 0 and len(two["send_idx"]) == 0


@pytest.mark.parametrize("graph,world", [("torus", 2), ("torus", 4), ("random", 3), ("random", 8)])
def test_halo_plan_is_consistent_across_ranks(built, monkeypatch, graph, world):
    """What rank p packs for rank q must be, row for row and in the same order, what q expects to receive from p: the
    two sides derive their lists independently (each from the whole CSR) and never talk about them."""
    lb = built
    monkeypatch.setenv("LORADS_HALO", "1")
    n = 1200
    ei, ej, w = lb.torus_graph(30, 40, 5) if graph == "torus" else lb.random_graph(n, 4, 5)
    p = lb.maxcut_problem(n, ei, ej, w)
    L = [lb.cone_layout(p, 0, ["send_idx", "send_off", "send_cnt", "recv_off", "recv_cnt", "dst_off", "halo_gid", "lf_col", "lf_ptr"],
                        world, r) for r in range(world)]
    lo = [lb.partition_rows(n, world, r)[0] for r in range(world)]
    rpr = lb.partition_rows(n, world, 0)[2]
    for src in range(world):
        for dst in range(world):
            if src == dst:
                assert L[src]["send_cnt"][dst] == 0 and L[dst]["recv_cnt"][src] == 0
                continue
            s0, sc = int(L[src]["send_off"][dst]), int(L[src]["send_cnt"][dst])
            r0, rc = int(L[dst]["recv_off"][src]), int(L[dst]["recv_cnt"][src])
            assert sc == rc
            # peer-memory PUT: the sender computes by itself where its block starts inside the receiver's halo
            assert int(L[src]["dst_off"][dst]) == r0
            assert np.array_equal(L[src]["send_idx"][s0:s0 + sc].astype(np.int64) + lo[src], L[dst]["halo_gid"][r0:r0 + rc])
    # every remapped column is an own row or a halo row that exists
    for r in range(world):
        nl = len(L[r]["lf_ptr"]) - 1
        cols = L[r]["lf_col"]
        assert np.all((cols < nl) | ((cols >= rpr) & (cols < rpr + len(L[r]["halo_gid"]))))


def _csr_rows(ptr, col):
    return [np.sort(col[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


@pytest.mark.parametrize("graph", ["scrambled_torus", "random_forced"])
def test_row_relabelling_is_a_consistent_permutation(built, monkeypatch, graph):
    """bfs_relabel (lgpu_layout.h): the relabelled fused layout is the original one under a row permutation -- same graph,
    same values per entry, diagonal last, constraints renumbered in row order -- and on a graph with locality it pulls the
    entries to the diagonal."""
    lb = built
    if graph == "scrambled_torus":
        rows, cols = 40, 50
        ei, ej, w = lb.torus_graph(rows, cols, 3)
        lab = np.random.default_rng(3).permutation(rows * cols)
        n, ei, ej = rows * cols, lab[ei], lab[ej]
    else:
        n = 1500
        ei, ej, w = lb.random_graph(n, 4, 11)
        w = np.random.default_rng(11).choice([-1.0, 1.0], size=len(ei))
    p = lb.maxcut_problem(n, ei, ej, w)
    names = ["f_ptr", "f_col", "f_slot", "mc_val", "rc_ptr", "rc_gid", "rc_a", "d_row", "pat_row", "pat_col", "perm", "iperm", "cperm",
             "dev_pat_row", "dev_pat_col", "con_gid", "t_gid"]
    monkeypatch.setenv("LORADS_REORDER", "0")
    A = lb.cone_layout(p, 0, names)
    monkeypatch.setenv("LORADS_REORDER", "1")
    B = lb.cone_layout(p, 0, names)
    assert len(A["perm"]) == 0 and len(B["perm"]) == n
    perm, iperm = B["perm"].astype(np.int64), B["iperm"].astype(np.int64)
    assert np.array_equal(np.sort(perm), np.arange(n)) and np.array_equal(perm[iperm], np.arange(n))
    # the pattern (slot numbering) is untouched; the device copy carries the new labels
    assert np.array_equal(A["pat_row"], B["pat_row"]) and np.array_equal(A["pat_col"], B["pat_col"])
    assert np.array_equal(B["dev_pat_row"], perm[B["pat_row"]]) and np.array_equal(B["dev_pat_col"], perm[B["pat_col"]])
    # CSR: new row perm[i] holds row i's entries with relabelled columns; values travel with the entries
    ra, rb = _csr_rows(A["f_ptr"], A["f_col"]), _csr_rows(B["f_ptr"], B["f_col"])
    for i in range(0, n, 7):
        assert np.array_equal(np.sort(perm[ra[i]]), rb[perm[i]])
        ea = {int(perm[c]): v for c, v in zip(A["f_col"][A["f_ptr"][i]:A["f_ptr"][i + 1]], A["mc_val"][A["f_ptr"][i]:A["f_ptr"][i + 1]])}
        a = int(perm[i])
        eb = {int(c): v for c, v in zip(B["f_col"][B["f_ptr"][a]:B["f_ptr"][a + 1]], B["mc_val"][B["f_ptr"][a]:B["f_ptr"][a + 1]])}
        assert ea == eb
        cols_b = B["f_col"][B["f_ptr"][a]:B["f_ptr"][a + 1]]
        if a in cols_b:
            assert cols_b[-1] == a                       # diagonal entry last
            assert np.all(np.diff(cols_b[:-1]) > 0)      # the others sorted by (new) column
    # constraints renumbered in row order: the row -> constraint lists are the identity sequence
    cperm = B["cperm"].astype(np.int64)
    assert np.array_equal(np.sort(cperm), np.arange(p.m))
    assert np.array_equal(B["rc_gid"], np.arange(p.m))
    assert np.array_equal(cperm[A["con_gid"]], B["con_gid"]) and np.array_equal(cperm[A["t_gid"]], B["t_gid"])
    # row of the caller's constraint k, in new labels, lists device constraint cperm[k]
    for k in range(0, p.m, 11):
        a = int(perm[A["d_row"][k]])
        assert cperm[k] in B["rc_gid"][B["rc_ptr"][a]:B["rc_ptr"][a + 1]]
    if graph == "scrambled_torus":
        band_a = np.abs(np.repeat(np.arange(n), np.diff(A["f_ptr"])) - A["f_col"]).mean()
        band_b = np.abs(np.repeat(np.arange(n), np.diff(B["f_ptr"])) - B["f_col"]).mean()
        assert band_b < 0.1 * band_a, (band_a, band_b)


def test_relabelling_gives_up_on_an_expander(built, monkeypatch):
    """default mode: a uniform random graph above the size threshold is probed (the BFS frontier passes n/16 within a few
    levels) and keeps its order; a torus of the same size with scrambled labels is relabelled"""
    lb = built
    monkeypatch.delenv("LORADS_REORDER", raising=False)
    n = 1 << 18
    ei, ej, w = lb.random_graph(n, 5, 0)
    L = lb.cone_layout(lb.maxcut_problem(n, ei, ej, w), 0, ["perm"])
    assert len(L["perm"]) == 0
    ei, ej, w = lb.torus_graph(512, 512, 1)
    lab = np.random.default_rng(1).permutation(n)
    L = lb.cone_layout(lb.maxcut_problem(n, lab[ei], lab[ej], w), 0, ["perm"])
    assert len(L["perm"]) == n


def test_cone_owner_map(built):
    """by-cone partition: largest cost first onto the least loaded rank; deterministic, every rank used when there is work"""
    lb = built
    cost = [5.0, 1.0, 9.0, 3.0, 3.0, 2.0]
    own = lb.cone_owner_map(cost, 2)
    assert own.tolist() == [1, 0, 0, 0, 1, 1] or sorted(np.bincount(own, weights=cost, minlength=2).tolist()) == [11.0, 12.0]
    loads = np.bincount(own, weights=cost, minlength=2)
    assert abs(loads[0] - loads[1]) <= max(cost)
    assert np.array_equal(lb.cone_owner_map(cost, 2), own)              # deterministic
    assert set(lb.cone_owner_map([1.0] * 8, 4).tolist()) == {0, 1, 2, 3}
    assert lb.cone_owner_map([4.0], 8).tolist() == [0]
