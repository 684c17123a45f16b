#!/usr/bin/env python
"""Generate the small synthetic SDPA (.dat-s) parity fixtures under tests/golden/instances/.

TEST INFRASTRUCTURE.  Every file is produced from a fixed numpy seed so the fixtures can be
regenerated bit-for-bit; they cover the storage classes the reference distinguishes
(SURVEY.md Appendix B): sparse aggregate with diagonal-only constraints (MaxCut), sparse aggregate
with general multi-entry constraints, dense aggregate (Lovasz-theta: dense C), a dense A_i,
multi-block problems whose blocks become SPARSE_CONE containers, and a trailing LP block.

File convention (reference reader lorads/src/src_semi/io/lorads_file_io.c:104-331): line 1 m,
line 2 #blocks, line 3 block dims (negative = LP, last only), line 4 b, then
`con blk i j val` 1-based; constraint 0 is the objective F0 and is negated on read.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "instances")


def write_sdpa(path, m, dims, b, entries):
    """entries: list of (con, blk, i, j, val), 1-based blk/i/j, con 0 = objective."""
    with open(path, "w") as f:
        f.write(f"{m}\n{len(dims)}\n")
        f.write(" ".join(str(d) for d in dims) + "\n")
        f.write(" ".join(repr(float(x)) for x in b) + "\n")
        for (c, k, i, j, v) in entries:
            f.write(f"{c} {k} {i} {j} {float(v)!r}\n")


def maxcut_entries(n, edges, weights, blk=1):
    """MaxCut SDP in the gen_MaxCut.jl convention: F0 = -L/2 (F0_ij = w_ij/2, F0_ii = -sum_j w_ij/2),
    A_k = e_k e_k^T, b = 1."""
    deg = np.zeros(n)
    ent = []
    for (i, j), w in zip(edges, weights):
        deg[i] += w
        deg[j] += w
        ent.append((0, blk, min(i, j) + 1, max(i, j) + 1, 0.5 * w))
    for i in range(n):
        if deg[i] != 0:
            ent.append((0, blk, i + 1, i + 1, -0.5 * deg[i]))
    for i in range(n):
        ent.append((i + 1, blk, i + 1, i + 1, 1.0))
    return ent


def torus_graph(rows, cols, rng, pm1=True):
    edges, w = [], []
    for r in range(rows):
        for c in range(cols):
            v = r * cols + c
            for (rr, cc) in ((r, (c + 1) % cols), ((r + 1) % rows, c)):
                u = rr * cols + cc
                if u != v and (min(u, v), max(u, v)) not in edges:
                    edges.append((min(u, v), max(u, v)))
                    w.append(float(rng.choice([-1.0, 1.0])) if pm1 else 1.0)
    return edges, w


def gen_maxcut_torus(path, rows, cols, seed):
    rng = np.random.default_rng(seed)
    n = rows * cols
    edges, w = torus_graph(rows, cols, rng)
    write_sdpa(path, n, [n], np.ones(n), maxcut_entries(n, edges, w))


def gen_g1(path, mtx="/root/reference/hallar/py/graphs/G1.mtx"):
    """BASELINE configs[0]: G-set G1 (n = m = 800, 19176 unit-weight edges) from the MatrixMarket pattern file the
    reference ships, in the gen_MaxCut.jl convention (reference LoRADS: objective -24166.3)."""
    edges = []
    with open(mtx) as f:
        for line in f:
            if line.startswith("%"):
                continue
            t = line.split()
            if len(t) == 3 and not edges and int(t[0]) == int(t[1]):   # size line: rows cols nnz
                n = int(t[0])
                continue
            i, j = int(t[0]) - 1, int(t[1]) - 1
            if i != j:
                edges.append((min(i, j), max(i, j)))
    edges = sorted(set(edges))
    # gen_MaxCut.jl / bundled G11.dat-s sign: the FILE holds F0 = +L/2 (the reader negates it, lorads_file_io.c:317-319),
    # i.e. the opposite sign of maxcut_entries() above, whose +-1-weight fixtures do not depend on it
    ent = [(c, k, i, j, (-v if c == 0 else v)) for (c, k, i, j, v) in maxcut_entries(n, edges, [1.0] * len(edges))]
    write_sdpa(path, n, [n], np.ones(n), ent)
    return n, len(edges)


def gen_general_sparse(path, n, m, seed):
    """Sparse aggregate (< 10% of the triangle), every constraint has several off-diagonal entries."""
    rng = np.random.default_rng(seed)
    ent = []
    # objective: banded
    for i in range(n):
        ent.append((0, 1, i + 1, i + 1, -(2.0 + float(rng.random()))))
        if i + 1 < n:
            ent.append((0, 1, i + 1, i + 2, 0.5 * float(rng.normal())))
    # a planted PSD X0 = Z Z^T gives a feasible b
    Z = rng.normal(size=(n, 3))
    X0 = Z @ Z.T
    b = np.zeros(m)
    for k in range(m):
        c = int(rng.integers(0, n))
        vals = {}
        vals[(c, c)] = 1.0 + float(rng.random())
        for _ in range(3):
            d = int(rng.integers(1, 3))
            i, j = c, (c + d) % n
            vals[(min(i, j), max(i, j))] = float(rng.normal())
        for (i, j), v in vals.items():
            ent.append((k + 1, 1, i + 1, j + 1, v))
            b[k] += v * X0[i, j] * (1.0 if i == j else 2.0)
    write_sdpa(path, m, [n], b, ent)


def gen_theta(path, n, p, seed):
    """Lovasz theta: max <J,X> s.t. tr X = 1, X_ij = 0 on edges.  Dense objective => dense aggregate."""
    rng = np.random.default_rng(seed)
    ent = []
    for i in range(n):
        for j in range(i, n):
            ent.append((0, 1, i + 1, j + 1, 1.0))
    for i in range(n):
        ent.append((1, 1, i + 1, i + 1, 1.0))
    k = 1
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() < p:
                k += 1
                ent.append((k, 1, i + 1, j + 1, 1.0))
    b = np.zeros(k)
    b[0] = 1.0
    write_sdpa(path, k, [n], b, ent)


def gen_dense_constraint(path, n, m, seed):
    """Sparse objective, some sparse constraints and ONE dense A_i (> 10% of the triangle)."""
    rng = np.random.default_rng(seed)
    ent = []
    Z = rng.normal(size=(n, 2))
    X0 = Z @ Z.T
    b = np.zeros(m)
    for i in range(n):
        ent.append((0, 1, i + 1, i + 1, -(1.0 + float(rng.random()))))
    for k in range(m):
        if k == 1:
            for i in range(n):
                for j in range(i, n):
                    if rng.random() < 0.5:
                        v = float(rng.normal())
                        ent.append((k + 1, 1, i + 1, j + 1, v))
                        b[k] += v * X0[i, j] * (1.0 if i == j else 2.0)
        else:
            i = int(rng.integers(0, n))
            j = int(rng.integers(0, n))
            i, j = min(i, j), max(i, j)
            ent.append((k + 1, 1, i + 1, i + 1, 1.0))
            b[k] += X0[i, i]
            if i != j:
                v = float(rng.normal())
                ent.append((k + 1, 1, i + 1, j + 1, v))
                b[k] += 2.0 * v * X0[i, j]
    write_sdpa(path, m, [n], b, ent)


def gen_multiblock(path, dims, nlp, m, seed, lp=True):
    """Several SDP blocks; each constraint touches ONE block (so every block sees < 30% of the
    constraints and becomes a SPARSE_CONE container, lorads_user_data.c:105-109) plus, optionally, a
    trailing LP block of nlp columns."""
    rng = np.random.default_rng(seed)
    ent = []
    nb = len(dims)
    X0 = []
    for n in dims:
        Z = rng.normal(size=(n, 2))
        X0.append(Z @ Z.T)
    x_lp = rng.random(nlp) + 0.1
    for k, n in enumerate(dims):
        for i in range(n):
            ent.append((0, k + 1, i + 1, i + 1, -(1.0 + float(rng.random()))))
            if i + 1 < n and rng.random() < 0.5:
                ent.append((0, k + 1, i + 1, i + 2, float(rng.normal()) * 0.3))
    if lp:
        for c in range(nlp):
            ent.append((0, nb + 1, c + 1, c + 1, -(0.5 + float(rng.random()))))
    b = np.zeros(m)
    for q in range(m):
        k = q % (nb + 1) if nb >= 3 else q % nb
        if k >= nb:
            k = int(rng.integers(0, nb))
        n = dims[k]
        i = int(rng.integers(0, n))
        j = int(rng.integers(0, n))
        i, j = min(i, j), max(i, j)
        v = 1.0 + float(rng.random())
        ent.append((q + 1, k + 1, i + 1, i + 1, v))
        b[q] += v * X0[k][i, i]
        if i != j:
            v = float(rng.normal())
            ent.append((q + 1, k + 1, i + 1, j + 1, v))
            b[q] += 2.0 * v * X0[k][i, j]
        if lp and (q < nlp or rng.random() < 0.4):
            # every LP column appears in >= 1 constraint (the reference aborts on an all-zero LP column)
            c = q if q < nlp else int(rng.integers(0, nlp))
            v = float(rng.random()) + 0.2
            ent.append((q + 1, nb + 1, c + 1, c + 1, v))
            b[q] += v * x_lp[c]
    d = list(dims) + ([-nlp] if lp else [])
    write_sdpa(path, m, d, b, ent)


def gen_control_like(path, dims, m, seed):
    """SDPLIB `control`-type stand-in (BASELINE configs[1]): two small dense blocks [2k, k] in which EVERY constraint
    couples both blocks through dense-ish symmetric pieces (an LMI in SDPA form), so both cones are dense containers
    (every A_i non-zero in each block) with dense aggregates (n < 20 / nnzP >= 0.1 tri); b from a feasible X0."""
    rng = np.random.default_rng(seed)
    ent, X0 = [], []
    for n in dims:
        Z = rng.normal(size=(n, 3))
        X0.append(Z @ Z.T + 0.1 * np.eye(n))
    for k, n in enumerate(dims):
        for i in range(n):
            ent.append((0, k + 1, i + 1, i + 1, -(1.0 + float(rng.random()))))
            for j in range(i + 1, n):
                if rng.random() < 0.3:
                    ent.append((0, k + 1, i + 1, j + 1, 0.2 * float(rng.normal())))
    b = np.zeros(m)
    for q in range(m):
        for k, n in enumerate(dims):
            dens = 0.6 if q % 3 else 0.25
            for i in range(n):
                for j in range(i, n):
                    if rng.random() < dens or (i == j == q % n):
                        v = 0.25 * float(rng.normal())
                        ent.append((q + 1, k + 1, i + 1, j + 1, v))
                        b[q] += v * X0[k][i, j] * (1.0 if i == j else 2.0)
    write_sdpa(path, m, dims, b, ent)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_maxcut_torus(os.path.join(OUT, "maxcut_torus_8x10.dat-s"), 8, 10, 11)
    gen_maxcut_torus(os.path.join(OUT, "maxcut_torus_20x30.dat-s"), 20, 30, 81)
    gen_general_sparse(os.path.join(OUT, "general_sparse_n60.dat-s"), 60, 40, 5)
    gen_theta(os.path.join(OUT, "theta_n30.dat-s"), 30, 0.3, 7)
    gen_dense_constraint(os.path.join(OUT, "dense_constraint_n24.dat-s"), 24, 12, 9)
    gen_multiblock(os.path.join(OUT, "multiblock_sdp.dat-s"), [25, 30, 22, 28], 0, 40, 13, lp=False)
    gen_multiblock(os.path.join(OUT, "multiblock_lp.dat-s"), [25, 30, 22, 28], 12, 40, 17, lp=True)
    gen_control_like(os.path.join(OUT, "control_like_12_6.dat-s"), [12, 6], 21, 23)
    if os.path.exists("/root/reference/hallar/py/graphs/G1.mtx"):
        print("G1:", gen_g1(os.path.join(OUT, "G1.dat-s")))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
