#!/usr/bin/env python
"""oracle/make_golden.py -- generate tests/golden/*.npz from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference to build
oracle/_ref/liblorads_ref.so via oracle/Makefile).  For every fixture under
tests/golden/instances it drives the reference's own functions through oracle/ref_harness.c and
stores their inputs/outputs; it also prints how far the numpy restatement (lorads_oracle.py) is
from the reference on the same inputs.  The .npz files are committed; this script is how they
were made.  Each fixture runs in a fresh subprocess because the reference keeps global state.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "_ref", "liblorads_ref.so")
INST = os.path.join(ROOT, "tests", "golden", "instances")
OUT = os.path.join(ROOT, "tests", "golden")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)


def dp(a):
    return a.ctypes.data_as(c_dp)


def load_lib():
    lib = ctypes.CDLL(LIB)
    lib.rh_load.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_int]
    lib.rh_rho0.restype = ctypes.c_double
    lib.rh_obj_constr_val_all.restype = ctypes.c_double
    lib.rh_obj_constr_val_all.argtypes = [ctypes.c_int, c_dp]
    lib.rh_alm_cal_grad.restype = ctypes.c_double
    lib.rh_alm_cal_grad.argtypes = [ctypes.c_double]
    lib.rh_wsum.argtypes = [ctypes.c_int, c_dp, ctypes.c_int, c_dp, ctypes.c_int]
    lib.rh_pattern.argtypes = [ctypes.c_int, c_ip, c_ip, ctypes.c_int]
    lib.rh_mul_rk.argtypes = [ctypes.c_int, ctypes.c_int, c_dp]
    lib.rh_alm_inner_iter.argtypes = [ctypes.c_double, ctypes.c_int, c_dp]
    lib.rh_get_factor.argtypes = [ctypes.c_int, ctypes.c_int, c_dp]
    lib.rh_set_factor.argtypes = [ctypes.c_int, ctypes.c_int, c_dp]
    lib.rh_get_vec.argtypes = [ctypes.c_int, c_dp]
    lib.rh_set_vec.argtypes = [ctypes.c_int, c_dp]
    lib.rh_get_lp.argtypes = [ctypes.c_int, c_dp]
    lib.rh_get_b.argtypes = [c_dp]
    lib.rh_constants.argtypes = [c_dp]
    lib.rh_update_dual_var.argtypes = [ctypes.c_double]
    lib.rh_admm_sweep.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    lib.rh_cal_obj_admm.restype = ctypes.c_double
    lib.rh_cal_obj_alm.restype = ctypes.c_double
    lib.rh_dimacs_admm.restype = ctypes.c_double
    lib.rh_aug_rank.argtypes = [ctypes.c_double]
    lib.rh_dual_infeasibility.restype = ctypes.c_double
    lib.rh_line_search.argtypes = [ctypes.c_double, ctypes.c_int, c_dp, ctypes.c_double, ctypes.c_double,
                                   c_dp, c_dp, c_dp, c_dp]
    lib.rh_cubic.argtypes = [ctypes.c_double] * 4 + [c_dp]
    return lib


def get_factor(lib, which, c):
    n, r = lib.rh_dim(c), lib.rh_rank(c)
    a = np.zeros(n * r)
    lib.rh_get_factor(which, c, dp(a))
    return a.reshape((n, r), order="F").copy()


def get_vec(lib, which):
    a = np.zeros(lib.rh_m())
    lib.rh_get_vec(which, dp(a))
    return a


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / den if a.size else 0.0


def one(name):
    sys.path.insert(0, HERE)
    import lorads_oracle as orc
    lib = load_lib()
    path = os.path.join(INST, name + ".dat-s")
    # silence the reference's stdout chatter
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    rc = lib.rh_load(path.encode(), 2.0, -1)
    os.dup2(saved, 1)
    assert rc == 0
    m, nc, nlp = lib.rh_m(), lib.rh_ncones(), lib.rh_nlp()
    g = {}
    g["m"], g["ncones"], g["nlp"] = m, nc, nlp
    g["dims"] = np.array([lib.rh_dim(c) for c in range(nc)])
    g["rank"] = np.array([lib.rh_rank(c) for c in range(nc)])
    g["rank_max"] = np.array([lib.rh_rank_max(c) for c in range(nc)])
    g["sparse_container"] = np.array([lib.rh_cone_is_sparse_container(c) for c in range(nc)])
    g["dense_aggregate"] = np.array([lib.rh_cone_aggregate_is_dense(c) for c in range(nc)])
    k = np.zeros(6); lib.rh_constants(dp(k)); g["constants"] = k
    b = np.zeros(m); lib.rh_get_b(dp(b)); g["b"] = b
    rho0 = lib.rh_rho0(); g["rho0"] = rho0
    R0 = [get_factor(lib, 0, c) for c in range(nc)]
    U0 = [get_factor(lib, 1, c) for c in range(nc)]
    V0 = [get_factor(lib, 2, c) for c in range(nc)]
    for c in range(nc):
        g[f"R0_{c}"], g[f"U0_{c}"], g[f"V0_{c}"] = R0[c], U0[c], V0[c]
    if nlp:
        for i, nm in enumerate(("rLp0", "uLp0", "vLp0")):
            a = np.zeros(nlp); lib.rh_get_lp(i, dp(a)); g[nm] = a

    # numpy-oracle side
    prob = orc.read_sdpa(path)
    cones = [orc.build_cone(bk, m) for bk in prob.blocks]
    report = {}
    for c in range(nc):
        assert cones[c].sparse_container == bool(g["sparse_container"][c]), "container class differs"
        assert cones[c].dense_aggregate == bool(g["dense_aggregate"][c]), "aggregate class differs"
        rk, rmax = orc.determine_rank(cones[c], nc)
        assert rk == g["rank"][c] and rmax == g["rank_max"][c], ("rank rule differs", rk, rmax, g["rank"][c], g["rank_max"][c])
        prow = np.zeros(cones[c].nnzP, np.int32); pcol = np.zeros(cones[c].nnzP, np.int32)
        nn = lib.rh_pattern(c, prow.ctypes.data_as(c_ip), pcol.ctypes.data_as(c_ip), len(prow))
        if nn >= 0:
            assert nn == cones[c].nnzP and np.all(prow == cones[c].pat_row) and np.all(pcol == cones[c].pat_col)
    # glibc seed parity of the initial point (srand(925), lorads_solver.c:625)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(925)
    Rg = [orc.glibc_rand_factor(libc, int(g["dims"][c]), int(g["rank"][c])) for c in range(nc)]
    report["R0_glibc"] = max(rel(Rg[c], R0[c]) for c in range(nc))

    # (1) A(UV^T) and <C,UV^T> on (R,R) and on (U,V)
    for pair, nm, A, B in ((0, "RR", R0, R0), (3, "UV", U0, V0)):
        cv = np.zeros(nc * m)
        obj = lib.rh_obj_constr_val_all(pair, dp(cv))
        cv = cv.reshape(nc, m)
        g[f"cv_{nm}"], g[f"obj_{nm}"] = cv, obj
        if not nlp:
            ocv, oobj = orc.constr_val_all(cones, A, B)
            report[f"cv_{nm}"] = max(rel(ocv[c], cv[c]) for c in range(nc))
            report[f"obj_{nm}"] = abs(oobj - obj) / max(abs(obj), 1e-300)
    # (2) A^*(w) + C and S R
    rng = np.random.default_rng(123)
    w = rng.normal(size=m)
    g["w"] = w
    for c in range(nc):
        nn = lib.rh_wsum(c, dp(w), 1, None, 0)
        S = np.zeros(nn)
        lib.rh_wsum(c, dp(w), 1, dp(S), nn)
        Y = np.zeros(R0[c].size)
        lib.rh_mul_rk(c, 0, dp(Y))
        Y = Y.reshape(R0[c].shape, order="F")
        g[f"S_{c}"], g[f"SR_{c}"] = S, Y
        oS = orc.wsum(cones[c], w, True)
        report[f"S_{c}"] = rel(oS, S)
        report[f"SR_{c}"] = rel(orc.mul_rk(cones[c], oS, R0[c]), Y)
    # (3) gradient at rho0, lambda = 0
    lib.rh_init_constr_val_sum_RR()
    cvs0 = get_vec(lib, 1)
    g["cvs0"] = cvs0
    lag = lib.rh_alm_cal_grad(rho0)
    G0 = [get_factor(lib, 3, c) for c in range(nc)]
    g["lag0"] = lag
    for c in range(nc):
        g[f"G0_{c}"] = G0[c]
    if not nlp:
        oG, olag = orc.alm_grad(cones, R0, b, np.zeros(m), cvs0, rho0)
        report["G0"] = max(rel(oG[c], G0[c]) for c in range(nc))
        report["lag0"] = abs(olag - lag) / lag
    # (4) five ALM inner iterations (lorads_alm.c:1302-1378)
    niter = 5
    sc = np.zeros((niter, 6))
    if not nlp:
        N = sum(r.size for r in R0)
        hist = orc.LbfgsHistory(2, N)
        oR, oG, ocvs = [r.copy() for r in R0], [x.copy() for x in G0], cvs0.copy()
        osc = np.zeros((niter, 6))
    for it in range(niter):
        lib.rh_alm_inner_iter(rho0, it, dp(sc[it]))
        if it == 0:
            g["D_first"] = np.concatenate([get_factor(lib, 1, c).ravel(order="F") for c in range(nc)])
            g["q1_first"], g["q2_first"] = get_vec(lib, 2), get_vec(lib, 3)
        if not nlp:
            o = orc.alm_inner_iter(cones, oR, oG, hist, b, np.zeros(m), ocvs, rho0, it)
            oR, oG, ocvs = o["R"], o["G"], o["cvs"]
            osc[it] = (o["rootNum"], o["tau"], o["p1"], o["p2"], o["lag"], o["pinf"])
    g["inner_scalars"] = sc
    R5 = [get_factor(lib, 0, c) for c in range(nc)]
    G5 = [get_factor(lib, 3, c) for c in range(nc)]
    for c in range(nc):
        g[f"R5_{c}"], g[f"G5_{c}"] = R5[c], G5[c]
    g["cvs5"] = get_vec(lib, 1)
    if nlp:
        a = np.zeros(nlp); lib.rh_get_lp(0, dp(a)); g["rLp5"] = a
    if not nlp:
        report["inner_scalars"] = float(np.max(np.abs(osc - sc) / np.maximum(np.abs(sc), 1e-300)))
        report["R5"] = max(rel(oR[c], R5[c]) for c in range(nc))
        report["G5"] = max(rel(oG[c], G5[c]) for c in range(nc))
    g["obj_alm5"] = lib.rh_cal_obj_alm()
    g["oracle_rank5"] = lib.rh_oracle_rank(1)
    # (5) dual update then one ADMM sweep from (U,V) = (R5,R5)
    lib.rh_update_dual_var(rho0)
    lam1 = get_vec(lib, 0)
    g["lam1"] = lam1
    lib.rh_alm_to_admm_copy()
    rho_admm = 10.0 * rho0
    g["rho_admm"] = rho_admm
    cg_tol, cg_max = 1e-8, 800
    g["cg_tol"] = cg_tol
    cgit = lib.rh_admm_sweep(rho_admm, cg_tol, cg_max, 1)
    g["cg_iter"] = cgit
    Ua = [get_factor(lib, 1, c) for c in range(nc)]
    Va = [get_factor(lib, 2, c) for c in range(nc)]
    for c in range(nc):
        g[f"Ua_{c}"], g[f"Va_{c}"] = Ua[c], Va[c]
    g["cvs_admm"] = get_vec(lib, 1)
    if nlp:
        for i, nm in ((1, "uLpa"), (2, "vLpa")):
            a = np.zeros(nlp); lib.rh_get_lp(i, dp(a)); g[nm] = a
    g["obj_admm"] = lib.rh_cal_obj_admm()
    g["pinf_admm"] = lib.rh_dimacs_admm()
    if not nlp:
        oU, oV, ocvs2, oit = orc.admm_sweep(cones, [r.copy() for r in R5], [r.copy() for r in R5], b, lam1,
                                            rho_admm, cg_tol, cg_max)
        report["Ua"] = max(rel(oU[c], Ua[c]) for c in range(nc))
        report["Va"] = max(rel(oV[c], Va[c]) for c in range(nc))
        report["cg_iter(ref,oracle)"] = (cgit, oit)
    g["oracle_rank_admm"] = lib.rh_oracle_rank(2)
    g["dual_infeas"] = lib.rh_dual_infeasibility()
    # (6) rank augmentation x1.5 (lorads_solver.c:1154-1254)
    lib.rh_aug_rank(1.5)
    g["rank_aug"] = np.array([lib.rh_rank(c) for c in range(nc)])
    for c in range(nc):
        g[f"Raug_{c}"] = get_factor(lib, 0, c)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in report.items()})


def main():
    subprocess.check_call(["make", "-C", HERE, "_ref/liblorads_ref.so"], stdout=subprocess.DEVNULL)
    names = sorted(f[:-6] for f in os.listdir(INST) if f.endswith(".dat-s"))
    if len(sys.argv) > 1:
        one(sys.argv[1])
        return
    for nme in names:
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), nme], env=env)


if __name__ == "__main__":
    main()
