"""CPU: the numpy restatement (oracle/lorads_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py -> tests/golden/*.npz).  This is what pins the oracle."""
import ctypes

import numpy as np
import pytest

from conftest import FIXTURES, golden, inst_path, rel
import lorads_oracle as orc

TOL = 1e-12


def _setup(name):
    g = golden(name)
    prob = orc.read_sdpa(inst_path(name))
    cones = [orc.build_cone(bk, prob.m) for bk in prob.blocks]
    return g, prob, cones


def _fac(g, key, nc):
    return [g[f"{key}_{c}"] for c in range(nc)]


@pytest.mark.parametrize("name", FIXTURES)
def test_reader_storage_rank_rules(name):
    g, prob, cones = _setup(name)
    assert prob.m == int(g["m"]) and len(cones) == int(g["ncones"]) and prob.nlp == int(g["nlp"])
    assert np.array_equal(np.array(prob.dims), g["dims"])
    assert np.array_equal(prob.b, g["b"])
    for c, cone in enumerate(cones):
        assert cone.sparse_container == bool(g["sparse_container"][c])
        assert cone.dense_aggregate == bool(g["dense_aggregate"][c])
        rk, rmax = orc.determine_rank(cone, len(cones))
        assert (rk, rmax) == (int(g["rank"][c]), int(g["rank_max"][c]))


@pytest.mark.parametrize("name", FIXTURES)
def test_glibc_seeded_initial_point(name):
    g, prob, cones = _setup(name)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(925)
    for c in range(len(cones)):
        Rg = orc.glibc_rand_factor(libc, int(g["dims"][c]), int(g["rank"][c]))
        assert np.array_equal(Rg, g[f"R0_{c}"])  # bit-exact: same libc, same arithmetic


@pytest.mark.parametrize("name", FIXTURES)
def test_wsum_and_mulrk(name):
    g, prob, cones = _setup(name)
    for c, cone in enumerate(cones):
        S = orc.wsum(cone, g["w"], True)
        assert rel(S, g[f"S_{c}"]) < TOL
        assert rel(orc.mul_rk(cone, S, g[f"R0_{c}"]), g[f"SR_{c}"]) < TOL


@pytest.mark.parametrize("name", [f for f in FIXTURES if f != "multiblock_lp"])
def test_constraint_operator_and_gradient(name):
    g, prob, cones = _setup(name)
    nc = len(cones)
    R0, U0, V0 = _fac(g, "R0", nc), _fac(g, "U0", nc), _fac(g, "V0", nc)
    for nm, A, B in (("RR", R0, R0), ("UV", U0, V0)):
        cv, obj = orc.constr_val_all(cones, A, B)
        for c in range(nc):
            assert rel(cv[c], g[f"cv_{nm}"][c]) < TOL
        assert abs(obj - float(g[f"obj_{nm}"])) <= TOL * max(1.0, abs(float(g[f"obj_{nm}"])))
    G, lag = orc.alm_grad(cones, R0, prob.b, np.zeros(prob.m), g["cvs0"], float(g["rho0"]))
    for c in range(nc):
        assert rel(G[c], g[f"G0_{c}"]) < TOL
    assert abs(lag - float(g["lag0"])) <= TOL * float(g["lag0"])


@pytest.mark.parametrize("name", [f for f in FIXTURES if f != "multiblock_lp"])
def test_five_alm_inner_iterations(name):
    g, prob, cones = _setup(name)
    nc = len(cones)
    R, G, cvs = _fac(g, "R0", nc), _fac(g, "G0", nc), g["cvs0"].copy()
    hist = orc.LbfgsHistory(2, sum(r.size for r in R))
    rho = float(g["rho0"])
    sc = g["inner_scalars"]
    for it in range(sc.shape[0]):
        o = orc.alm_inner_iter(cones, R, G, hist, prob.b, np.zeros(prob.m), cvs, rho, it)
        if it == 0:
            assert rel(orc.flat(o["D"]), g["D_first"]) < TOL
            assert rel(o["q1"], g["q1_first"]) < TOL and rel(o["q2"], g["q2_first"]) < TOL
        R, G, cvs = o["R"], o["G"], o["cvs"]
        got = np.array([o["rootNum"], o["tau"], o["p1"], o["p2"], o["lag"], o["pinf"]])
        assert np.all(np.abs(got - sc[it]) <= 1e-10 * np.maximum(np.abs(sc[it]), 1e-300)), (it, got, sc[it])
    for c in range(nc):
        assert rel(R[c], g[f"R5_{c}"]) < 1e-10 and rel(G[c], g[f"G5_{c}"]) < 1e-9
    assert rel(cvs, g["cvs5"]) < 1e-10


@pytest.mark.parametrize("name", [f for f in FIXTURES if f != "multiblock_lp"])
def test_admm_sweep(name):
    g, prob, cones = _setup(name)
    nc = len(cones)
    R5 = _fac(g, "R5", nc)
    U, V, cvs, it = orc.admm_sweep(cones, [r.copy() for r in R5], [r.copy() for r in R5], prob.b, g["lam1"],
                                   float(g["rho_admm"]), float(g["cg_tol"]), 800)
    assert abs(it - int(g["cg_iter"])) <= max(2, 0.05 * int(g["cg_iter"]))
    for c in range(nc):
        assert rel(U[c], g[f"Ua_{c}"]) < 1e-7 and rel(V[c], g[f"Va_{c}"]) < 1e-7
    assert rel(cvs, g["cvs_admm"]) < 1e-7


def test_line_search_cubic_branches():
    # the three discriminant branches of Shengjin's formulas against numpy's polynomial roots
    for (a, b, c, d) in [(1.0, -6.0, 11.0, -6.0), (2.0, 0.0, 1.0, -3.0), (1.0, -3.0, 3.0, -1.0), (4.0, 1.0, -2.0, 0.3)]:
        n, roots = orc.cubic_equation(a, b, c, d)
        true = np.roots([a, b, c, d])
        true = np.sort(true[np.abs(true.imag) < 1e-7].real)
        for k in range(n):
            if a == 1.0 and b == -3.0:  # triple root: A == B == 0 branch returns max(0, -c/b)
                continue
            assert np.min(np.abs(true - roots[k])) < 1e-6
