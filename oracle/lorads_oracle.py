"""oracle/lorads_oracle.py -- CPU restatement of the LoRADS inner-loop operators (numpy).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product path (ltr-lowrank-sdp_b200) never does.

Every function restates one reference function and cites it (paths relative to
/root/reference/lorads/src/src_semi).  Factors are numpy arrays of shape (n, r) holding the SAME numbers as
the reference's column-major `matElem` (element (row, k) = matElem[row + n*k]).

Pinning: oracle/make_golden.py runs these functions next to the UNMODIFIED reference objects
(oracle/_ref/liblorads_ref.so, built from /root/reference by oracle/Makefile) on the fixtures in
tests/golden/instances and stores the reference's outputs in tests/golden/*.npz;
tests/test_oracle_golden.py re-checks the restatement against those vectors on every run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

ZERO, SPARSE, DENSE = 0, 1, 2


# ------------------------------------------------------------------------------------------------
# SDPA reader  (io/lorads_file_io.c:59-455)
# ------------------------------------------------------------------------------------------------
@dataclass
class BlockData:
    """One SDP block as the reference reader hands it on: for column c (0 = objective, 1..m =
    constraints) the packed-lower indices and values, sorted ascending by packed index."""
    n: int
    cols_idx: List[np.ndarray]   # len m+1, int64 packed idx  (PACK_IDX, lorads_utils.h:167)
    cols_val: List[np.ndarray]


@dataclass
class Problem:
    m: int
    dims: List[int]
    nlp: int
    b: np.ndarray
    blocks: List[BlockData]
    lp_obj: Optional[np.ndarray] = None          # length nlp (already negated)
    lp_cols: Optional[list] = None               # per LP column: (rows int array, vals)


def pack_idx(n, i, j):
    """lower-triangular column-major packed index of (row i >= col j)  (lorads_utils.h:167)"""
    return (2 * n - j - 1) * j // 2 + i


def unpack_idx(n, idx):
    """inverse of pack_idx (linalg/lorads_sparse_opts.c:82-96 tsp_decompress)"""
    idx = np.asarray(idx, dtype=np.int64)
    # column j is the largest j with j*(2n-j+1)/2 <= idx
    j = np.floor(((2 * n + 1) - np.sqrt((2.0 * n + 1) ** 2 - 8.0 * idx)) / 2.0).astype(np.int64)
    j = np.clip(j, 0, n - 1)
    start = j * (2 * n - j + 1) // 2
    j = np.where(start > idx, j - 1, j)
    start = j * (2 * n - j + 1) // 2
    nxt = (j + 1) * (2 * n - j) // 2
    j = np.where(idx >= nxt, j + 1, j)
    start = j * (2 * n - j + 1) // 2
    i = idx - start + j
    return i, j


def read_sdpa(path) -> Problem:
    with open(path) as f:
        lines = f.readlines()
    k = 0
    while lines[k][0] in '*"':
        k += 1
    m = int(lines[k].split()[0]); k += 1
    nblk = int(lines[k].split()[0]); k += 1
    toks = lines[k].replace("{", " ").replace("}", " ").replace("(", " ").replace(")", " ") \
        .replace("'", " ").replace(",", " ").split()
    dims_all = [int(t) for t in toks[:nblk]]; k += 1
    nlp = 0
    if dims_all[-1] < 0:
        nlp = -dims_all[-1]
        dims = dims_all[:-1]
    else:
        dims = dims_all
    b = []
    while len(b) < m:
        b += [float(t) for t in lines[k].replace(",", " ").split()]
        k += 1
    b = np.array(b[:m], dtype=np.float64)
    nb = len(dims)
    trip = [[[] for _ in range(m + 1)] for _ in range(nb)]   # per block, per column: list of (packed, val)
    lp_trip = [[] for _ in range(m + 1)]
    for line in lines[k:]:
        t = line.split()
        if len(t) < 5:
            if line.startswith("BEGIN.COMMENT"):
                break
            continue
        con, blk, i, j, v = int(t[0]), int(t[1]) - 1, int(t[2]) - 1, int(t[3]) - 1, float(t[4])
        if abs(v) < 1e-12:            # lorads_file_io.c:288-294
            continue
        if con == 0:                   # objective negated on read (:317-319)
            v = -v
        if nlp > 0 and blk == nb:
            lp_trip[con].append((i, v))
        else:
            if i > j:
                i, j = j, i            # (i <= j) then stored as lower (row j, col i)
            trip[blk][con].append((pack_idx(dims[blk], j, i), v))
    blocks = []
    for bi in range(nb):
        ci, cv = [], []
        for c in range(m + 1):
            e = trip[bi][c]
            if e:
                idx = np.array([x[0] for x in e], dtype=np.int64)
                val = np.array([x[1] for x in e], dtype=np.float64)
                o = np.argsort(idx, kind="stable")
                ci.append(idx[o]); cv.append(val[o])
            else:
                ci.append(np.zeros(0, np.int64)); cv.append(np.zeros(0, np.float64))
        blocks.append(BlockData(dims[bi], ci, cv))
    prob = Problem(m, dims, nlp, b, blocks)
    if nlp > 0:
        obj = np.zeros(nlp)
        for (i, v) in lp_trip[0]:
            obj[i] = v                                    # lorads_lp_conic.c:38-40 (last write wins)
        cols = [([], []) for _ in range(nlp)]
        for c in range(1, m + 1):
            for (i, v) in sorted(lp_trip[c], key=lambda x: x[0]):
                cols[i][0].append(c - 1); cols[i][1].append(v)
        prob.lp_obj = obj
        prob.lp_cols = [(np.array(r, dtype=np.int64), np.array(v, dtype=np.float64)) for (r, v) in cols]
    return prob


# ------------------------------------------------------------------------------------------------
# cone construction (data/lorads_sdp_data.c:1180-1197, data/lorads_sdp_conic.c:1185-1393,
#                    io/lorads_user_data.c:97-122)
# ------------------------------------------------------------------------------------------------
@dataclass
class Cone:
    n: int
    m: int
    obj_type: int
    con_type: np.ndarray            # per constraint: ZERO/SPARSE/DENSE
    sparse_container: bool          # SPARSE_CONE container (<= 30% of the constraints touch the block)
    dense_aggregate: bool
    pat_row: np.ndarray             # aggregated lower pattern, sorted by (col,row)
    pat_col: np.ndarray
    # objective on the pattern
    c_slot: np.ndarray
    c_val: np.ndarray
    # constraints in CSR-by-constraint over pattern slots (all m constraints; empty rows allowed)
    a_ptr: np.ndarray
    a_slot: np.ndarray
    a_val: np.ndarray
    nnz_rows: int = 0               # number of non-zero A_i (sdpDenseConeNnzStat, lorads_sdp_conic.c:497-504)

    @property
    def nnzP(self):
        return len(self.pat_row)

    @property
    def pat_diag(self):
        return self.pat_row == self.pat_col


def mat_type(n, nnz):
    if nnz == 0:
        return ZERO
    if float(nnz) > 0.1 * float(n * (n + 1) // 2):
        return DENSE
    return SPARSE


def build_cone(blk: BlockData, m: int) -> Cone:
    n = blk.n
    tri = n * (n + 1) // 2
    types = np.array([mat_type(n, len(blk.cols_idx[c])) for c in range(m + 1)])
    nz_con = int(np.sum(types[1:] != ZERO))
    sparse_container = not (nz_con > 0.3 * m)
    dense = n < 20 or bool(np.any(types == DENSE))
    if not dense:
        allidx = np.unique(np.concatenate(blk.cols_idx)) if blk.cols_idx else np.zeros(0, np.int64)
        if float(len(allidx)) / float(tri) >= 0.1:
            dense = True
    if dense:
        allidx = np.arange(tri, dtype=np.int64)
    # packed index is column-major lower => sorted packed order == sorted by (col,row) (cmpfunc :1075-1088)
    pr, pc = unpack_idx(n, allidx)
    slot_of = lambda idx: np.searchsorted(allidx, idx)
    a_ptr = np.zeros(m + 1, dtype=np.int64)
    for c in range(1, m + 1):
        a_ptr[c] = a_ptr[c - 1] + len(blk.cols_idx[c])
    a_slot = slot_of(np.concatenate(blk.cols_idx[1:])) if m > 0 else np.zeros(0, np.int64)
    a_val = np.concatenate(blk.cols_val[1:]) if m > 0 else np.zeros(0)
    return Cone(n, m, int(types[0]), types[1:], sparse_container, dense, pr, pc,
                slot_of(blk.cols_idx[0]), blk.cols_val[0].copy(), a_ptr, a_slot.astype(np.int64), a_val,
                nnz_rows=nz_con)


def determine_rank(cone: Cone, ncones: int, times_log_rank=2.0, fixed_rank=-1, init_rank=-1):
    """LORADSDetermineRank (data/lorads_solver.c:406-459) -> (rank, rank_max)"""
    n = cone.n
    calc_max = min(int(math.sqrt(2 * cone.nnz_rows)) + 1, n)
    if fixed_rank > 0:
        r = max(1, min(fixed_rank, n))
        return r, r
    if init_rank > 0:
        return max(1, min(init_rank, n)), calc_max
    if times_log_rank <= 1e-6:
        r = calc_max
    elif cone.nnz_rows // n >= 20 and n <= 400 and ncones <= 3:
        r = calc_max
    else:
        r = min(math.ceil(times_log_rank * math.log(n)), calc_max)
    return max(1, int(r)), calc_max


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------
def uvt(cone: Cone, U, V):
    """LORADSUVt (lorads_alg/lorads_alg_common.c:43-90): samples of (UV^T + VU^T)/2 on the pattern."""
    r, c = cone.pat_row, cone.pat_col
    a = np.einsum("ij,ij->i", U[r], V[c])
    bb = np.einsum("ij,ij->i", U[c], V[r])
    out = 0.5 * a + 0.5 * bb
    d = r == c
    out[d] = a[d]
    return out


def cone_auv(cone: Cone, uvt_vals):
    """coneAUV -> mul_inner_rk_double (data/lorads_sdp_conic.c:378-385,681-688;
    data/lorads_sdp_data.c:803-876): y_i = sum_k 2 a_k UVt[slot_k], halved on the diagonal.
    Returns the dense m-vector (zero where A_i is zero in this block)."""
    diag = cone.pat_row[cone.a_slot] == cone.pat_col[cone.a_slot]
    t = 2.0 * cone.a_val * uvt_vals[cone.a_slot]
    t = np.where(diag, t - 0.5 * t, t)
    y = np.zeros(cone.m)
    rows = np.repeat(np.arange(cone.m), np.diff(cone.a_ptr))
    np.add.at(y, rows, t)
    return y


def obj_auv(cone: Cone, uvt_vals):
    """objAUV (data/lorads_sdp_conic.c:395-402): <C, UVt>"""
    diag = cone.pat_row[cone.c_slot] == cone.pat_col[cone.c_slot]
    t = 2.0 * cone.c_val * uvt_vals[cone.c_slot]
    t = np.where(diag, t - 0.5 * t, t)
    return float(np.sum(t))


def wsum(cone: Cone, w, add_obj=True):
    """zeros + addObjCoeff + sdpDataWSum (data/lorads_sdp_conic.c:448-460,608-616;
    data/lorads_sdp_data.c:878-939): S = C + sum_i w_i A_i on the pattern."""
    S = np.zeros(cone.nnzP)
    if add_obj:
        np.add.at(S, cone.c_slot, cone.c_val)
    rows = np.repeat(np.arange(cone.m), np.diff(cone.a_ptr))
    np.add.at(S, cone.a_slot, w[rows] * cone.a_val)
    return S


def mul_rk(cone: Cone, S, X):
    """mul_rk (data/lorads_sdp_data.c:750-763,948-973): Y = S X, S symmetric given on the lower pattern."""
    Y = np.zeros_like(X)
    r, c = cone.pat_row, cone.pat_col
    np.add.at(Y, r, S[:, None] * X[c])
    off = r != c
    np.add.at(Y, c[off], S[off, None] * X[r[off]])
    return Y


def alm_m1(b, lam, cvs, rho):
    """M1 = -lambda - rho b + rho A(RR^T)  (lorads_alg/lorads_alm.c:38-50)"""
    return -lam - rho * b + rho * cvs


def alm_grad(cones: List[Cone], R: List[np.ndarray], b, lam, cvs, rho):
    """ALMCalGrad (lorads_alg/lorads_alm.c:32-87): Grad_c = 2 (C_c + A_c^*(M1)) R_c ; returns (grads, sum ||Grad||^2)"""
    M1 = alm_m1(b, lam, cvs, rho)
    G, nrm = [], 0.0
    for cone, Rc in zip(cones, R):
        g = 2.0 * mul_rk(cone, wsum(cone, M1, True), Rc)
        G.append(g)
        nrm += float(np.linalg.norm(g.ravel())) ** 2
    return G, nrm


def constr_val_all(cones, U, V):
    """LORADSObjConstrValAll (lorads_alg/lorads_alg_common.c:169-176) -> (per-cone constrVal, sum_c <C_c, U_cV_c^T>)"""
    cv, obj = [], 0.0
    for cone, Uc, Vc in zip(cones, U, V):
        t = uvt(cone, Uc, Vc)
        obj += obj_auv(cone, t)
        cv.append(cone_auv(cone, t))
    return cv, obj


def q12p12(cones, R, D):
    """ALMCalq12p12 (lorads_alg/lorads_alm.c:714-734)"""
    cv, p1 = constr_val_all(cones, R, D)
    q1 = 2.0 * sum(cv)
    cv2, p2 = constr_val_all(cones, D, D)
    q2 = sum(cv2)
    return q1, q2, 2.0 * p1, p2


def nthroot(base, n):
    if base < 0 and n % 2 == 0:
        return float("nan")
    return base ** (1.0 / n) if base > 0 else -((-base) ** (1.0 / n))


def cubic_equation(a, b, c, d):
    """LORADScubic_equation (lorads_alg/lorads_alm.c:191-231): Shengjin's formulas. -> (rootNum, roots[3])"""
    A = b * b - 3 * a * c
    B = b * c - 9 * a * d
    C = c * c - 3 * b * d
    delta = B * B - 4 * A * C
    res = [0.0, 0.0, 0.0]
    if A == 0 and B == 0:
        res[0] = max(res[0], -c / b)
        return 1, res
    elif delta > 0:
        Y1 = A * b + 1.5 * a * (-B + math.sqrt(delta))
        Y2 = A * b + 1.5 * a * (-B - math.sqrt(delta))
        res[0] = max(res[0], (-b - nthroot(Y1, 3) - nthroot(Y2, 3)) / 3 / a)
        return 1, res
    elif delta == 0 and A != 0 and B != 0:
        K = B / A
        res[0] = -b / a + K
        res[1] = -K / 2
        return 2, res
    elif delta < 0:
        sqA = math.sqrt(A)
        T = (A * b - 1.5 * a * B) / (A * sqA)
        theta = math.acos(T)
        csth = math.cos(theta / 3)
        sn3th = math.sqrt(3) * math.sin(theta / 3)
        res[0] = (-b - 2 * sqA * csth) / 3 / a
        res[1] = (-b + sqA * (csth + sn3th)) / 3 / a
        res[2] = (-b + sqA * (csth - sn3th)) / 3 / a
        return 3, res
    return 0, res


def line_search_coeffs(rho, lam, p1, p2, q0, q1, q2):
    """quartic coefficients of ALMLineSearch (lorads_alg/lorads_alm.c:266-279); q0 is b - A(RR^T)"""
    a = rho * float(np.linalg.norm(q2)) ** 2 / 2
    bb = rho * float(q1 @ q2)
    q0p = q0 + lam / rho
    c = p2 - rho * float(q0p @ q2) + rho * float(np.linalg.norm(q1)) ** 2 / 2
    d = p1 - rho * float(q0p @ q1)
    return a, bb, c, d


def line_search_tau(a, b, c, d):
    """root selection of ALMLineSearch (lorads_alg/lorads_alm.c:280-333) -> (rootNum, tau)"""
    f = lambda x: a * x ** 4 + b * x ** 3 + c * x ** 2 + d * x
    nroot, roots = cubic_equation(4 * a, 3 * b, 2 * c, d)
    f0, f1 = 0.0, f(1.0)
    fr = [1e30, 1e30, 1e30]
    for k in range(3):
        if nroot >= k + 1 and (k < 2 or nroot == 3):
            if roots[k] > 1e-20 and roots[k] <= 1.0:
                fr[k] = f(roots[k])
    mn = min(f0, f1, fr[0], fr[1], fr[2])
    tau = 0.0
    if abs(mn - f0) < 1e-10:
        tau = 0.0
    if abs(mn - f1) < 1e-10:
        tau = 1.0
    for k in range(3):
        if abs(mn - fr[k]) < 1e-10:
            tau = roots[k]
    return nroot, tau


@dataclass
class LbfgsHistory:
    """circular list of `h` (s, y, beta) nodes; `head` = oldest slot, overwritten next
    (data/def_lorads_lbfgs.h:16-34, data/lorads_solver.c:680-706)"""
    h: int
    N: int
    s: List[np.ndarray] = field(default_factory=list)
    y: List[np.ndarray] = field(default_factory=list)
    beta: List[float] = field(default_factory=list)
    head: int = 0

    def __post_init__(self):
        self.s = [np.zeros(self.N) for _ in range(self.h)]
        self.y = [np.zeros(self.N) for _ in range(self.h)]
        self.beta = [0.0] * self.h


def lbfgs_direction(hist: LbfgsHistory, grad_flat, inner_iter):
    """LBFGSDirection + LBFGSDirectionUseGrad (lorads_alg/lorads_alm.c:347-508,607-627), SDP-only variant
    (node count test `innerIter <= h-1`)."""
    if inner_iter == 0:
        D = -grad_flat
    else:
        q = grad_flat.copy()
        nn = inner_iter if inner_iter <= hist.h - 1 else hist.h
        alpha = [0.0] * hist.h
        node = (hist.head - 1) % hist.h
        for _ in range(nn):
            alpha[node] = hist.beta[node] * float(hist.s[node] @ q)
            q -= alpha[node] * hist.y[node]
            node = (node - 1) % hist.h
        node = (node + 1) % hist.h
        for _ in range(nn):
            wgt = alpha[node] - hist.beta[node] * float(hist.y[node] @ q)
            q += wgt * hist.s[node]
            node = (node + 1) % hist.h
        D = -q
    if float(D @ grad_flat) >= 0:
        D = -grad_flat
    return D


def lbfgs_push(hist: LbfgsHistory, neg_old_grad, new_grad, D, tau):
    """SetyAsNegGrad + setlbfgsHisTwo (lorads_alg/lorads_alm.c:768-783,842-863)"""
    k = hist.head
    hist.y[k] = neg_old_grad + new_grad
    hist.s[k] = tau * D
    hist.beta[k] = 1.0 / float(hist.y[k] @ hist.s[k])
    hist.head = (k + 1) % hist.h


def flat(mats):
    """concatenate cone factors in the reference's memory order (column-major per cone)"""
    return np.concatenate([m.ravel(order="F") for m in mats])


def unflat(vec, like):
    out, o = [], 0
    for m in like:
        n, r = m.shape
        out.append(vec[o:o + n * r].reshape((n, r), order="F").copy())
        o += n * r
    return out


def primal_infeasibility(cones, R, b):
    """primalInfeasibility (lorads_alg/lorads_alg_common.c:386-394): recompute A(RR^T); ||b - A||_2 / (1 + ||b||_1)"""
    cv, _ = constr_val_all(cones, R, R)
    cvs = sum(cv)
    return float(np.linalg.norm(b - cvs)) / (1.0 + float(np.sum(np.abs(b)))), cvs


def alm_inner_iter(cones, R, G, hist, b, lam, cvs, rho, inner_iter):
    """one pass of the loop body lorads_alg/lorads_alm.c:1302-1378 (SDP cones only).
    Returns dict with D, q1, q2, p1, p2, tau, rootNum, R, G, cvs, lagNormSquare, pinf."""
    D = unflat(lbfgs_direction(hist, flat(G), inner_iter), R)
    q0 = b - cvs
    q1, q2, p1, p2 = q12p12(cones, R, D)
    a, bb, c, d = line_search_coeffs(rho, lam, p1, p2, q0, q1, q2)
    nroot, tau = line_search_tau(a, bb, c, d)
    negG = -flat(G)
    Rn = [r + tau * dd for r, dd in zip(R, D)]
    cvs_n = cvs + tau * q1 + tau * tau * q2
    Gn, lag = alm_grad(cones, Rn, b, lam, cvs_n, rho)
    lbfgs_push(hist, negG, flat(Gn), flat(D), tau)
    pinf, cvs_re = primal_infeasibility(cones, Rn, b)
    return dict(D=D, q1=q1, q2=q2, p1=p1, p2=p2, tau=tau, rootNum=nroot, R=Rn, G=Gn, cvs=cvs_re,
                lag=lag, pinf=pinf, coeffs=(a, bb, c, d))


def cg_solve(mvec, x, bvec, tol, maxiter, restart=20):
    """CGSolve (linalg/lorads_cgs.c:128-287): plain CG, ||r||_2 / ||b||_1 stopping rule, residual recomputed
    when k % 20 == 0.  Returns (x, iters) ; iters = None when the initial residual already passes."""
    bnorm = float(np.sum(np.abs(bvec)))
    r = bvec - mvec(x)
    res = float(np.linalg.norm(r))
    if res / bnorm < tol:
        return x, None
    p = r.copy()
    q = r.copy()
    qTr = float(q @ r)
    it = 0
    for k in range(maxiter):
        it += 1
        Q = mvec(p)
        qTr = float(q @ r)
        alpha = qTr / float(p @ Q)
        x = x + alpha * p
        r = r - alpha * Q
        res = float(np.linalg.norm(r))
        if res / bnorm < tol:
            break
        if k % restart == 0:
            r = bvec - mvec(x)
            p = r.copy()
            q = r.copy()
            qTr = float(q @ r)
        qn = r.copy()
        qTrN = float(qn @ r)
        beta = qTrN / qTr
        p = beta * p + r
        qTr = qTrN
        q = qn
    return x, it


def admm_update_one(cone: Cone, upd, fixed, b, lam, cvs, cv_cone, rho, tol, maxiter):
    """LORADSUpdateSDPVarOne (lorads_alg/lorads_admm.c:564-616) for one cone; returns (new upd, cg iters)"""
    M1 = rho * (-b + cvs - cv_cone) - lam
    M2 = mul_rk(cone, wsum(cone, M1, True), fixed) - rho * fixed
    rhs = (-1.0 / rho) * M2
    n, r = fixed.shape

    def mvec(xf):
        X = xf.reshape((n, r), order="F")
        w = cone_auv(cone, uvt(cone, X, fixed))
        Y = mul_rk(cone, wsum(cone, w, False), fixed) + X
        return Y.ravel(order="F")

    x, it = cg_solve(mvec, upd.ravel(order="F").copy(), rhs.ravel(order="F"), tol, maxiter)
    return x.reshape((n, r), order="F"), it


def admm_sweep(cones, U, V, b, lam, rho, tol, maxiter):
    """LORADSUpdateSDPVar (lorads_alg/lorads_alg_common.c:298-326): Gauss-Seidel over cones, U then V."""
    cv = [cone_auv(c, uvt(c, u, v)) for c, u, v in zip(cones, U, V)]
    cvs = sum(cv)
    iters = 0
    for k, cone in enumerate(cones):
        for which in (0, 1):
            if which == 0:
                U[k], it = admm_update_one(cone, U[k], V[k], b, lam, cvs, cv[k], rho, tol, maxiter)
            else:
                V[k], it = admm_update_one(cone, V[k], U[k], b, lam, cvs, cv[k], rho, tol, maxiter)
            iters += it or 0
            cvs = cvs - cv[k]
            cv[k] = cone_auv(cone, uvt(cone, U[k], V[k]))
            cvs = cvs + cv[k]
    return U, V, cvs, iters


def gram_oracle_rank(Rm, eps=1e-6):
    """oracle_rank_from_factor (lorads_logging.c:216-240,272-370): #eig(R^T R) > eps * lambda_max"""
    w = np.linalg.eigvalsh(Rm.T @ Rm)
    if w[-1] <= 0:
        return 0
    return int(np.sum(w > eps * w[-1]))


def glibc_rand_factor(libc, n, r):
    """LORADS_RANDOM_rk_MAT (data/lorads_solver.c:529-539): element k = rand()/RAND_MAX - rand()/RAND_MAX in memory
    (column-major) order; `libc` is ctypes.CDLL('libc.so.6') with srand() already called."""
    RAND_MAX = 2147483647
    flat_ = np.empty(n * r)
    for k in range(n * r):
        a = libc.rand()
        b2 = libc.rand()
        flat_[k] = a / RAND_MAX - b2 / RAND_MAX
    return flat_.reshape((n, r), order="F")
