"""CPU: host-side logic of the drop-in (C reader, scalar line search, Jacobi eigenvalues, CLI) and the C-ABI
surface: the library loads and exports every symbol include/lorads_b200.h declares.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import FIXTURES, ROOT, inst_path
import lorads_oracle as orc


def test_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "lorads_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(lgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    L = ctypes.CDLL(built.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in ctypes.c_char_p(ctypes.cast(L.lgpu_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value
    built.lib()  # the binding's own signature table resolves too


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", built.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("name", FIXTURES)
def test_c_reader_matches_oracle_reader(built, name):
    p = built.read_sdpa(inst_path(name))
    q = orc.read_sdpa(inst_path(name))
    assert p.m == q.m and list(p.dims) == list(q.dims) and p.nlp == q.nlp
    assert np.array_equal(p.b, q.b)
    for k, blk in enumerate(q.blocks):
        beg = p.mat_beg[k]
        assert len(beg) == q.m + 2
        for c in range(q.m + 1):
            idx, val = p.mat_idx[k][beg[c]:beg[c + 1]], p.mat_elem[k][beg[c]:beg[c + 1]]
            o = np.argsort(idx, kind="stable")
            assert np.array_equal(idx[o], blk.cols_idx[c]) and np.array_equal(val[o], blk.cols_val[c]), (k, c)
    if q.nlp:
        obj = np.zeros(q.nlp)
        obj[p.lp_idx[p.lp_beg[0]:p.lp_beg[1]]] = p.lp_elem[p.lp_beg[0]:p.lp_beg[1]]
        assert np.array_equal(obj, q.lp_obj)
        cols = [([], []) for _ in range(q.nlp)]
        for c in range(q.m):
            for e in range(p.lp_beg[c + 1], p.lp_beg[c + 2]):
                cols[p.lp_idx[e]][0].append(c)
                cols[p.lp_idx[e]][1].append(p.lp_elem[e])
        for j in range(q.nlp):
            assert np.array_equal(np.array(cols[j][0]), q.lp_cols[j][0])
            assert np.array_equal(np.array(cols[j][1]), q.lp_cols[j][1])


def test_reader_edge_cases(built, tmp_path):
    # comments, braces/commas in the dims line, upper-triangular entries, tiny entries dropped, trailing comment block
    f = tmp_path / "edge.dat-s"
    f.write_text('"a comment\n* another\n2\n2\n{3, -2}\n1.0, 2.0\n'
                 "0 1 1 2 0.5\n0 1 3 3 -1\n1 1 2 1 1.0\n1 1 1 1 1e-13\n2 1 3 3 2.0\n1 2 1 1 3.0\n2 2 2 2 -4.0\n0 2 2 2 7\n"
                 "BEGIN.COMMENT\n9 9 9 9 9\n")
    p = built.read_sdpa(str(f))
    q = orc.read_sdpa(str(f))
    assert p.m == 2 and list(p.dims) == [3] and p.nlp == 2
    assert np.array_equal(p.b, [1.0, 2.0])
    beg = p.mat_beg[0]
    assert beg[-1] == 4  # the 1e-13 entry is dropped
    for c in range(3):
        assert np.array_equal(np.sort(p.mat_idx[0][beg[c]:beg[c + 1]]), q.blocks[0].cols_idx[c])
    assert p.mat_elem[0][beg[0]:beg[1]].tolist() == [-0.5, 1.0]  # objective negated on read
    # missing file: the binding raises, the binary exits 0 like the reference
    with pytest.raises(built.LoradsError):
        built.read_sdpa(str(tmp_path / "nope.dat-s"))
    assert built.run_solver([str(tmp_path / "nope.dat-s")]).returncode == 0


def test_line_search_matches_oracle(built):
    H = built.host_lib()
    rng = np.random.default_rng(7)
    dp = ctypes.POINTER(ctypes.c_double)
    for trial in range(2000):
        rho = float(10 ** rng.uniform(-2, 3))
        m = 6
        lam, q0, q1, q2 = (rng.normal(size=m) for _ in range(4))
        if trial % 7 == 0:
            q2 *= 0.0  # degenerate quartic -> quadratic
        p1, p2 = float(rng.normal()), float(rng.normal())
        a, b, c, d = orc.line_search_coeffs(rho, lam, p1, p2, q0, q1, q2)
        q0p = q0 + lam / rho
        terms = np.array([p1, p2, q2 @ q2, q1 @ q2, q0p @ q2, q1 @ q1, q0p @ q1])
        tau = ctypes.c_double(-7.0)
        n = H.lh_line_search(rho, terms.ctypes.data_as(dp), ctypes.byref(tau))
        try:
            with np.errstate(all="ignore"):
                on, otau = orc.line_search_tau(a, b, c, d)
        except ZeroDivisionError:  # python raises where C yields inf/nan (0/0 in the degenerate cubic)
            continue
        assert n == on
        if not (np.isnan(otau) or np.isnan(tau.value)):
            assert abs(tau.value - otau) <= 1e-9 * max(1.0, abs(otau)), (trial, tau.value, otau)


def test_cubic_equation(built):
    H = built.host_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    for coef in [(1.0, -6.0, 11.0, -6.0), (2.0, 0.0, 1.0, -3.0), (4.0, 1.0, -2.0, 0.3), (1.0, -3.0, 3.0, -1.0)]:
        res = np.zeros(3)
        n = H.lh_cubic_equation(*coef, res.ctypes.data_as(dp))
        on, ores = orc.cubic_equation(*coef)
        assert n == on and np.allclose(res, ores, rtol=1e-13, atol=0)


def test_jacobi_eigenvalues(built):
    H = built.host_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 14, 41):
        A = rng.normal(size=(n + 3, n))
        G = A.T @ A
        w = np.zeros(n)
        H.lh_sym_eigvals(n, G.copy().ctypes.data_as(dp), w.ctypes.data_as(dp))
        assert np.allclose(w, np.linalg.eigvalsh(G), rtol=1e-10, atol=1e-12 * np.max(np.abs(G)))


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly (nonzero exit, message), never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(built.LoradsError, match="no CPU fallback"):
        built.Context(0)
    r = built.run_solver([inst_path("G11")])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_product_never_imports_oracle():
    """only tests/, smoke() and bench.py's baseline legs may touch oracle/: not the package, not scripts/, not include/"""
    for top in ("ltr-lowrank-sdp_b200", "scripts", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith((".py", ".c", ".cpp", ".h", ".cu", ".cuh", ".sh")) or fn == "Makefile":
                    txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                    assert "lorads_oracle" not in txt and "oracle/" not in txt and "lorads_ref" not in txt, fn


@pytest.mark.parametrize("name", FIXTURES)
def test_storage_rules_without_gpu(built, name):
    """lgpu_cone_classify (pure host) reproduces the reference's container / aggregate classes recorded in the golden
    vectors, and the pattern size / non-zero constraint count of the oracle's cone."""
    from conftest import golden
    g = golden(name)
    p = built.read_sdpa(inst_path(name))
    q = orc.read_sdpa(inst_path(name))
    for c, blk in enumerate(q.blocks):
        cone = orc.build_cone(blk, q.m)
        info = built.cone_classify(p, c)
        assert info["sparse_container"] == bool(g["sparse_container"][c])
        assert info["dense_aggregate"] == bool(g["dense_aggregate"][c])
        assert info["nnz_rows"] == cone.nnz_rows and info["nnzP"] == cone.nnzP
        assert info["nnzA"] == len(cone.a_slot)
        is_maxcut = name in ("G11", "maxcut_torus_8x10", "maxcut_torus_20x30")
        assert info["diag_only"] == is_maxcut


def test_synthetic_builders(built):
    lb = built
    ei, ej, w = lb.random_graph(5000, 5, 0)
    assert np.all(ei < ej) and len(np.unique(ei * 5000 + ej)) == len(ei)       # unique undirected edges
    p = lb.maxcut_problem(5000, ei, ej, w)
    info = lb.cone_classify(p, 0)
    assert info["diag_only"] and not info["dense_aggregate"] and info["nnz_rows"] == 5000
    assert info["nnzP"] == len(ei) + 5000
    ti, tj, tw = lb.torus_graph(100, 200, 81)
    assert len(ti) == 40000 and set(np.unique(tw)) == {-1.0, 1.0}              # G81-like: 2 edges per vertex


def test_cli_surface_matches_reference_options(built, tmp_path):
    """argv contract (main.c:125-154, 264-350): positional instance first, the 27 long options parsed with atof/atoi and
    echoed, rhoCellingADMM always reset to 200 x rhoMax, unknown options ignored, plus the three options benchmark.py
    passes.  Runs on CPU: the echo happens before the device is touched (the run then stops for lack of a GPU)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the end-to-end tests")
    except Exception:
        pass
    sched = tmp_path / "r_sched.json"
    sched.write_text('{"rank_schedule": [5, 8, 12], "schedule_length": 3}')
    r = built.run_solver([inst_path("G11"), "--phase1Tol", "1e-2", "--heuristicFactor", "10", "--rhoMax", "4000", "--rhoCellingADMM", "7",
                          "--timeSecLimit", "77", "--reoptLevel", "0", "--fixedRank", "9", "--initRank", "6", "--maxALMIter", "12",
                          "--lbfgsListLength", "3", "--oracleRankNaive", "--rankSchedule", str(sched), "--nearStallFactor", "0.7",
                          "--disableOracle", "--noSuchOption", "--jsonfile", str(tmp_path / "o.json")])
    out = r.stdout
    for needle in ("phase1Tol = 0.010000", "heuristicFactor = 10.000000", "rhoMax = 4000.000000", "rhoCellingADMM = 800000.000000",
                   "timeSecLimit = 77.000000", "reoptLevel = 0", "fixedRank = 9", "initRank = 6", "maxALMIter = 12",
                   "lbfgsListLength = 3", "oracleRankMethod = 1", "disableOracle = 1", "nearStallFactor = 0.700000",
                   "nConstrs = 800, sdp nBlks = 1, lp Cols = 0", "Pre-solver starts"):
        assert needle in out, needle
    assert "unrecognized option" in r.stderr          # glibc getopt reports it and the run goes on, like the reference
    assert r.returncode == 3 and not (tmp_path / "o.json").exists()


@pytest.mark.parametrize("ranks", [2, 4])
def test_ranks_are_forked_after_the_file_is_read(built, ranks):
    """`--ranks P` (main.c): the file is read once, by the parent, and the ranks are forked afterwards; rank 0 alone prints.
    Without a GPU every rank stops at the device with the no-fallback message, the parent reaps them and exits 3 -- no
    rank is left waiting for the communicator id."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the multi-GPU tests")
    except Exception:
        pass
    r = built.run_solver([inst_path("G11"), "--ranks", str(ranks)], timeout=60)
    assert r.returncode == 3
    assert r.stdout.count("Reading SDPA file in") == 1 and r.stdout.count("Input parameters:") == 1
    assert r.stdout.count("nConstrs = 800, sdp nBlks = 1, lp Cols = 0") == 1 and r.stdout.count("Pre-solver starts") == 1
    assert r.stderr.count("no CPU fallback") == ranks


def test_partition_and_classify_need_no_gpu(built):
    assert built.partition_rows(10_000_000, 8, 7) == (8_750_000, 10_000_000, 1_250_000)
    p = built.read_sdpa(inst_path("multiblock_lp"))
    infos = [built.cone_classify(p, c) for c in range(p.ncones)]
    assert all(not i["diag_only"] for i in infos) and len(infos) == 4


def test_rand_stream_is_glibc_rand(built):
    """The starting point must be the reference's srand(925) stream (lorads_solver.c:529-539): the lock-free
    restatement in csrc/host/glibc_rand.c against the C library, number for number, and the factor fill."""
    H = built.host_lib()
    libc = ctypes.CDLL("libc.so.6")
    H.lh_random_fill.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    for seed in (925, 0, 1, 20000, 2 ** 31 + 5):
        libc.srand(ctypes.c_uint(seed))
        H.lh_srand(ctypes.c_uint(seed))
        assert [libc.rand() for _ in range(5000)] == [H.lh_rand() for _ in range(5000)], seed
    libc.srand(925)
    H.lh_srand(925)
    n = 20011
    got = np.empty(n)
    H.lh_random_fill(got.ctypes.data, n)
    want = np.empty(n)
    for i in range(n):
        x = libc.rand() / 2147483647
        x -= libc.rand() / 2147483647
        want[i] = x
    assert np.array_equal(got, want)
    assert libc.rand() == H.lh_rand()  # the fill leaves the stream where the library's would be


def _random_sdpa_text(rng, m, dims, nlp, count, long_objective=0):
    """Entry lines in random order, both triangles, duplicates, assorted number spellings."""
    spell = [lambda v: repr(v), lambda v: "%.17g" % v, lambda v: "%.3e" % v, lambda v: "%+.6f" % v, lambda v: "%.12E" % v,
             lambda v: ("%.4f" % v).replace("0.", ".", 1), lambda v: "%d." % round(v * 10), lambda v: "%d" % round(v * 10)]
    lines, want = [], []
    nblk = len(dims) + (1 if nlp else 0)
    for k in range(count):
        con = int(rng.integers(0, m + 1))
        if k < long_objective:
            con = 0
        blk = int(rng.integers(0, nblk)) if k >= long_objective else 0
        v = float(rng.normal()) * 10 ** int(rng.integers(-3, 4))
        txt = spell[int(rng.integers(0, len(spell)))](v)
        if abs(float(txt)) < 1e-12:
            txt = "1.5"
        if blk < len(dims):
            n = dims[blk]
            i, j = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
        else:
            i = j = int(rng.integers(1, nlp + 1))
        sep = ["  ", " ", "\t"][int(rng.integers(0, 3))]
        lines.append(sep.join(str(x) for x in (con, blk + 1, i, j)) + sep + txt + ("\r" if k % 11 == 0 else ""))
        want.append((con, blk, i - 1, j - 1, -float(txt) if con == 0 else float(txt)))
    head = '"generated\n%d\n%d\n%s\n%s\n' % (m, nblk, " ".join(str(d) for d in list(dims) + ([-nlp] if nlp else [])),
                                              " ".join(repr(float(x)) for x in rng.normal(size=m)))
    return head + "\n".join(lines) + "\n", want


@pytest.mark.parametrize("threads", [1, 2, 5, 8])
def test_reader_threads_and_number_formats(built, tmp_path, monkeypatch, threads):
    """the multi-threaded reader gives the single-threaded answer -- column order = file order then
    ascending packed index, every value bit-identical to float(text) -- for any thread count, including columns long
    enough for the cooperative sort."""
    rng = np.random.default_rng(100 + threads)
    m, dims, nlp = 7, [400, 3], 5
    text, want = _random_sdpa_text(rng, m, dims, nlp, count=90000, long_objective=70000)
    f = tmp_path / "rnd.dat-s"
    f.write_text(text)
    monkeypatch.setenv("LORADS_READ_THREADS", str(threads))
    p = built.read_sdpa(str(f))
    assert p.m == m and list(p.dims) == dims and p.nlp == nlp
    # expectation, built straight from the generated lines
    cols = [[[] for _ in range(m + 1)] for _ in range(len(dims) + 1)]
    for con, blk, i, j, v in want:
        if blk < len(dims):
            n = dims[blk]
            lo, hi = min(i, j), max(i, j)
            cols[blk][con].append(((2 * n - lo - 1) * lo // 2 + hi, v))
        else:
            cols[blk][con].append((i, v))
    for k in range(len(dims)):
        beg = p.mat_beg[k]
        assert beg[-1] == sum(len(c) for c in cols[k])
        for c in range(m + 1):
            exp = sorted(cols[k][c], key=lambda t: t[0])  # python's sort is stable: ties stay in file order
            assert p.mat_idx[k][beg[c]:beg[c + 1]].tolist() == [t[0] for t in exp], (k, c)
            got = p.mat_elem[k][beg[c]:beg[c + 1]]
            assert got.tobytes() == np.array([t[1] for t in exp], dtype=np.float64).tobytes(), (k, c)
    for c in range(m + 1):
        exp = cols[len(dims)][c]  # LP: file order inside the column
        assert p.lp_idx[p.lp_beg[c]:p.lp_beg[c + 1]].tolist() == [t[0] for t in exp]
        assert p.lp_elem[p.lp_beg[c]:p.lp_beg[c + 1]].tobytes() == np.array([t[1] for t in exp]).tobytes()


def test_reader_number_conversion_is_strtod(built, tmp_path, monkeypatch):
    """the fast decimal path and the library path agree with float() on awkward spellings"""
    toks = ["1", "-1", "+1", "0.5", ".5", "5.", "1e0", "1E+3", "1e-3", "123456789012345678", "1234567890123456789012",
            "0.1", "0.30000000000000004", "9007199254740993", "9007199254740992", "1e22", "1e23", "1e-22", "1e-23",
            "4.9e-324", "1.7976931348623157e308", "2.2250738585072014e-308", "0.000000000000000000000001234",
            "123456.789e-5", "00012.5000", "-0.25e+1", "7.", "3.14159265358979323846264338327950288"]
    f = tmp_path / "num.dat-s"
    n = len(toks)
    f.write_text("1\n1\n%d\n1.0\n" % n + "".join("1 1 %d %d %s\n" % (k + 1, k + 1, t) for k, t in enumerate(toks)))
    for thr in ("1", "3"):
        monkeypatch.setenv("LORADS_READ_THREADS", thr)
        p = built.read_sdpa(str(f))
        beg = p.mat_beg[0]
        keep = [float(t) for t in toks if abs(float(t)) >= 1e-12]
        assert p.mat_elem[0][beg[1]:beg[2]].tobytes() == np.array(keep).tobytes()


@pytest.mark.parametrize("name", ["G11", "multiblock_lp", "theta_n30"])
def test_binary_image_round_trip(built, tmp_path, name):
    """SURVEY 8d side channel: the binary image of a parsed problem reads back array for array; a damaged or truncated
    image is refused, never half-read."""
    lb = built
    p = lb.read_sdpa(inst_path(name))
    f = tmp_path / "img.lbin"
    lb.write_sdpa_binary(str(f), p)
    q = lb.read_sdpa(str(f))
    assert q.m == p.m and list(q.dims) == list(p.dims) and q.nlp == p.nlp and np.array_equal(q.b, p.b)
    for k in range(p.ncones):
        assert np.array_equal(q.mat_beg[k], p.mat_beg[k]) and np.array_equal(q.mat_idx[k], p.mat_idx[k])
        assert q.mat_elem[k].tobytes() == p.mat_elem[k].tobytes()
    if p.nlp:
        assert np.array_equal(q.lp_beg, p.lp_beg) and np.array_equal(q.lp_idx, p.lp_idx) and np.array_equal(q.lp_elem, p.lp_elem)
    raw = f.read_bytes()
    (tmp_path / "short.lbin").write_bytes(raw[:-9])
    (tmp_path / "long.lbin").write_bytes(raw + b"\0" * 8)
    bad = bytearray(raw)
    off = 8 + 32 + 8 * p.ncones + 8 * p.m + 8 * (p.m + 2)   # first packed index of block 0
    bad[off:off + 8] = (2 ** 40).to_bytes(8, "little")
    (tmp_path / "bad.lbin").write_bytes(bytes(bad))
    for nm in ("short.lbin", "long.lbin", "bad.lbin"):
        with pytest.raises(lb.LoradsError):
            lb.read_sdpa(str(tmp_path / nm))


@pytest.mark.parametrize("seed", range(6))
def test_reader_layout_variations(built, tmp_path, monkeypatch, seed):
    """Text-layout variations the reference's per-line sscanf accepts: CRLF line ends, blank lines, leading blanks, extra
    tokens after the value, no newline at the end of the file, a comment block after the entries -- for several thread
    counts, so that piece boundaries fall inside all of them."""
    rng = np.random.default_rng(1000 + seed)
    m, dims, nlp = 5, [30, 7], (4 if seed % 2 else 0)
    text, want = _random_sdpa_text(rng, m, dims, nlp, count=3000)
    head, body = text.split("\n", 5)[:5], text.split("\n", 5)[5]
    lines = body.rstrip("\n").split("\n")
    out = []
    for k, ln in enumerate(lines):
        ln = ln.rstrip("\r")
        if k % 7 == 0:
            ln = "   " + ln
        if k % 5 == 0:
            ln = ln + "  trailing 1 2 3"
        out.append(ln)
        if k % 13 == 0:
            out.append("")
    eol = "\r\n" if seed % 3 == 0 else "\n"
    text2 = "\n".join(head) + "\n" + eol.join(out)
    if seed % 2 == 0:
        text2 += eol + "BEGIN.COMMENT" + eol + "1 1 1 1 99.0" + eol
    f = tmp_path / "var.dat-s"
    f.write_bytes(text2.encode())
    ref = None
    for thr in ("1", "3", "8"):
        monkeypatch.setenv("LORADS_READ_THREADS", thr)
        p = built.read_sdpa(str(f))
        got = ([a.tobytes() for a in p.mat_beg + p.mat_idx + p.mat_elem], None if not nlp else
               (p.lp_beg.tobytes(), p.lp_idx.tobytes(), p.lp_elem.tobytes()))
        if ref is None:
            ref = got
            total = sum(int(b[-1]) for b in p.mat_beg) + (int(p.lp_beg[-1]) if nlp else 0)
            assert total == len(want)          # nothing dropped, nothing read from the comment block
            vals = np.sort(np.concatenate(p.mat_elem + ([p.lp_elem] if nlp else [])))
            assert vals.tobytes() == np.sort(np.array([w[4] for w in want])).tobytes()
        else:
            assert got == ref, thr


def test_oracle_rank_count_matches_dense_eigenvalues(built):
    """lh_sym_rank (Householder + Sturm counts) = #{eigenvalues > 1e-6 lambda_max}, the reference's oracle-rank rule
    (lorads_logging.c:503-543), on full-rank, low-rank + noise, threshold-straddling, diagonal, zero and negative
    matrices; a disagreement is accepted only when an eigenvalue sits on the cut to 1e-7 relative."""
    H = built.host_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    H.lh_sym_rank.restype = ctypes.c_int64
    H.lh_sym_rank.argtypes = [ctypes.c_int, dp, ctypes.c_double]
    rng = np.random.default_rng(0)
    for trial in range(800):
        n = int(rng.integers(1, 60)) if trial % 50 else int(rng.integers(150, 280))
        k = int(rng.integers(0, n + 1))
        kind = trial % 5
        if kind == 0:
            A = rng.normal(size=(n + 3, n))
            G = A.T @ A
        elif kind == 1:
            A = rng.normal(size=(n, k))
            G = A @ A.T + 1e-9 * np.eye(n) * rng.random()
        elif kind == 2:
            Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
            G = (Q * 10.0 ** rng.uniform(-10, 2, size=n)) @ Q.T
        elif kind == 3:
            G = np.diag(rng.random(n) * (rng.random(n) > 0.5))
        else:
            G = -np.eye(n) * rng.random() if trial % 2 else np.zeros((n, n))
        G = np.ascontiguousarray(0.5 * (G + G.T))
        w = np.linalg.eigvalsh(G)
        want = int(np.sum(w > 1e-6 * w[-1])) if w[-1] > 0 else -1
        got = H.lh_sym_rank(n, G.copy().ctypes.data_as(dp), 1e-6)
        if got != want:
            cut = 1e-6 * w[-1]
            assert np.min(np.abs(w - cut)) <= 1e-7 * abs(cut) + 1e-14 * abs(w[-1]), (trial, n, kind, got, want)


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/lorads_b200.h is the drop-in boundary: it must compile on its own as C99 and as C++ (no torch types, no
    CUDA headers), so a maintainer can include it from the reference's C sources."""
    hdr = os.path.join(ROOT, "include", "lorads_b200.h")
    for cc, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "cpp")):
        src = tmp_path / f"inc.{ext}"
        src.write_text('#include "lorads_b200.h"\nint main(void) { return lgpu_version() == 0; }\n')
        r = subprocess.run([cc, std, "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.dirname(hdr), str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("name", ["multiblock_lp", "control_like_12_6", "theta_n30", "G11"])
def test_constraint_stats_match_the_feature_extractor_formulas(built, name):
    """lh_constraint_stats / lh_constraint_rows (hand-off to dataset/processor.py) against a scipy restatement
    of FeatureExtractor._precompute_constraint_stats and _build_pattern_matrix (processor.py:246-345): all SDP blocks as
    one block-diagonal symmetric matrix per constraint."""
    import scipy.sparse as sp
    lb = built
    p = lb.read_sdpa(inst_path(name))
    stats, obj, ptr, rows = lb.constraint_stats(p)
    n = int(np.sum(p.dims))
    offs = np.concatenate([[0], np.cumsum(p.dims)])

    def matrix(c):
        ri, ci, vv = [], [], []
        for k in range(p.ncones):
            nk = int(p.dims[k])
            e = slice(int(p.mat_beg[k][c]), int(p.mat_beg[k][c + 1]))
            idx, val = p.mat_idx[k][e], p.mat_elem[k][e]
            j = np.floor(((2 * nk + 1) - np.sqrt((2.0 * nk + 1) ** 2 - 8.0 * idx)) / 2.0).astype(np.int64)
            j = np.where(j * (2 * nk - j + 1) // 2 > idx, j - 1, j)
            j = np.where((j + 1) * (2 * nk - j) // 2 <= idx, j + 1, j)
            i = idx - j * (2 * nk - j + 1) // 2 + j
            off = i != j
            ri += [offs[k] + i, offs[k] + j[off]]
            ci += [offs[k] + j, offs[k] + i[off]]
            vv += [val, val[off]]
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(ri), np.concatenate(ci))), shape=(n, n))

    for c in list(range(0, p.m + 1, max(1, p.m // 40))) + [p.m]:
        A = matrix(c) if c else -matrix(0)           # the parsed objective is -F0; the hand-off is in the file's sign
        got = obj if c == 0 else stats[c - 1]
        diag = A.diagonal()
        row_sums = np.abs(A).sum(axis=1).A1
        ur = np.unique(A.tocoo().row)
        blocks = int(np.sum((offs[:-1] <= ur.max()) & (offs[1:] > ur.min()))) if len(ur) and p.ncones > 1 else (1 if len(ur) else 0)
        want = [np.sqrt(np.sum(A.data ** 2)), A.nnz, diag.sum(), np.linalg.norm(diag), row_sums.max() if A.nnz else 0.0, len(ur)]
        assert np.allclose(got[:6], want, rtol=1e-13, atol=1e-13), (c, got, want)
        assert got[6] == blocks
        if c > 0:
            assert np.array_equal(rows[ptr[c - 1]:ptr[c]], ur)


FEATURE_FIXTURES = ["theta_n30", "multiblock_lp", "multiblock_sdp", "control_like_12_6", "general_sparse_n60", "dense_constraint_n24",
                    "maxcut_torus_8x10", "G11"]


@pytest.mark.parametrize("name", FEATURE_FIXTURES)
def test_feature_handoff_reproduces_the_reference_extractor(built, name):
    """SURVEY 8f-4 against the REFERENCE's own output: tests/golden/features.npz holds what dataset/processor.py
    (SDPAParser + FeatureExtractor, run unmodified by tests/golden/make_feature_golden.py) produces for the fixture.  The
    17 global, 16 node and 5 edge features are assembled here from nothing but the hand-off arrays (lh_constraint_stats,
    lh_constraint_rows, lh_constraint_cost_alignment, lh_constraint_pairs) with the extractor's closed-form expressions
    (processor.py:296-316, :368-505, :666-700); float64 internals to 1e-12, the float32 features to float32 rounding."""
    lb = built
    G = np.load(os.path.join(ROOT, "tests", "golden", "features.npz"))
    g = lambda k: G[name + "/" + k]
    p = lb.read_sdpa(inst_path(name))
    stats, obj, ptr, rows = lb.constraint_stats(p)
    cp = lb.constraint_couplings(p)
    m, n, eps = p.m, int(np.sum(p.dims)), 1e-8
    assert n == int(g("n"))
    norms, nnz, traces, dnorm, gersh, rsz, blk = (stats[:, k] for k in range(7))
    for ours, key in ((norms, "norms"), (nnz, "nnz_counts"), (traces, "traces"), (dnorm, "diag_norms"), (gersh, "gershgorin_bounds"),
                      (blk, "blocks_touched"), (rsz, "row_sizes")):
        assert np.allclose(ours, g(key), rtol=1e-12, atol=1e-12), key
    c_fro = obj[0] if obj[1] > 0 else eps
    assert np.isclose(c_fro, float(g("C_frob")), rtol=1e-13)
    cos = np.where(nnz > 0, cp["cost_inner"] / (norms * c_fro + eps), 0.0) if obj[1] > 0 else np.zeros(m)
    assert np.allclose(cos, g("cos_with_C"), rtol=1e-12, atol=1e-13)

    # edges: every pair with a common row whose Jaccard index reaches the threshold (processor.py:666-712, m < 1000)
    log_norms, log_nnz = np.log(1.0 + norms), np.log(1.0 + nnz)
    pi = np.repeat(np.arange(m), np.diff(cp["ptr"]))
    pj, ov = cp["col"], cp["overlap"].astype(np.float64)
    jac = ov / (rsz[pi] + rsz[pj] - ov)
    keep = jac >= 0.05
    pi, pj, ov, jac, inner = pi[keep], pj[keep], ov[keep], jac[keep], cp["inner"][keep]
    want_ei, want_ea = g("edge_index"), g("edge_attr")
    if len(pi):
        feat = np.stack([jac, ov / (np.minimum(rsz[pi], rsz[pj]) + eps), np.abs(inner) / (norms[pi] * norms[pj] + eps),
                         np.minimum(log_norms[pi], log_norms[pj]), np.abs(log_norms[pi] - log_norms[pj])], axis=1)
        ei = np.stack([np.stack([pi, pj]), np.stack([pj, pi])], axis=2).reshape(2, -1)     # (i, j), (j, i), pair after pair
        assert np.array_equal(ei, want_ei)
        assert np.allclose(np.repeat(feat, 2, axis=0).astype(np.float32), want_ea, rtol=2e-6, atol=1e-7)
    else:
        # no two constraints share a row (MaxCut): the extractor falls back to nearest neighbours in log-norm, which needs
        # nothing from the problem data; its marker is a zero coupling column
        assert len(cp["col"]) == 0 and np.all(want_ea[:, 2] == 0.0)
        ei = want_ei
    deg = np.bincount(ei[0], minlength=m) if ei.shape[1] else np.zeros(m)

    # node features (processor.py:440-505)
    x = np.zeros((m, 16))
    nrhs = np.clip(p.b / (norms + eps), -100.0, 100.0)
    x[:, 0], x[:, 1] = log_norms, log_nnz
    x[:, 2] = np.clip(traces / (norms + eps), -100.0, 100.0)
    x[:, 3] = dnorm / (norms + eps)
    x[:, 4] = nrhs
    x[:, 5] = np.log(1.0 + gersh)
    x[:, 6] = cos
    x[:, 7] = np.where(cos > 0.01, 1.0, np.where(cos < -0.01, -1.0, 0.0))
    x[:, 8] = (log_norms - log_norms.mean()) / (log_norms.std() + eps)
    x[:, 9] = (log_nnz - log_nnz.mean()) / (log_nnz.std() + eps)
    x[:, 10] = (np.abs(nrhs) - np.abs(nrhs).mean()) / (np.abs(nrhs).std() + eps)
    x[:, 11] = np.digitize(log_norms, np.percentile(log_norms, [25, 50, 75])) / 3.0
    x[:, 12] = np.log(1.0 + rsz)
    if obj[1] > 0:
        x[:, 13] = np.where(rsz > 0, cp["rows_shared_with_cost"] / np.maximum(rsz, 1), 0.0)
    x[:, 14] = np.log(1.0 + deg) if ei.shape[1] else 0.0
    x[:, 15] = np.log(1.0 + blk)
    want_x = g("node")
    for k in range(16):
        assert np.allclose(x[:, k].astype(np.float32), want_x[:, k], rtol=3e-6, atol=2e-6), (k, x[:3, k], want_x[:3, k])

    # global features (processor.py:368-437 + the degree summary of :803-808)
    nsq = float(n * n) + eps
    dens = nnz / nsq
    glob = [np.log(1.0 + n), np.log(1.0 + m), np.log(1.0 + n / max(m, 1)), np.log(1.0 + c_fro), np.log(1.0 + norms.mean()),
            dens.mean(), dens.var(), obj[1] / nsq, log_norms.mean(), log_norms.std(), np.median(log_norms),
            cos.mean(), cos.std(), cos.max(), cos.min(), deg.mean() if ei.shape[1] else 0.0, deg.std() if ei.shape[1] else 0.0]
    assert np.allclose(np.array(glob, dtype=np.float32), g("global"), rtol=3e-6, atol=1e-7)


@pytest.mark.parametrize("seed", [0, 1])
def test_constraint_pairs_against_scipy_on_random_blocks(built, seed):
    """lh_constraint_pairs / lh_constraint_cost_alignment on a random three-block problem with m > 1000 (the size at which
    the extractor switches to `pattern @ pattern.T`, processor.py:563-583): overlap counts = the strict upper triangle of
    P P^T, inner = sum(A_i .* A_j), cost terms = sum(A_i .* F0) -- including empty constraints, a position listed twice in
    one constraint, and constraints that share a row but no position."""
    import scipy.sparse as sp
    lb = built
    rng = np.random.default_rng(seed)
    dims, m = [17, 40, 9], 1100
    offs = np.concatenate([[0], np.cumsum(dims)])
    n = int(offs[-1])
    begs, idxs, vals = [], [], []
    A = [sp.lil_matrix((n, n)) for _ in range(m + 1)]
    for k, nk in enumerate(dims):
        tri = nk * (nk + 1) // 2
        ci, cv = [], []
        for c in range(m + 1):
            cnt = 0 if c % 7 == 3 else int(rng.integers(0, 4)) + (6 if c == 0 else 0)
            ii = np.sort(rng.choice(tri, size=min(cnt, tri), replace=False))
            if c % 11 == 5 and len(ii):
                ii = np.sort(np.concatenate([ii, ii[:1]]))          # the same position twice: the values add up
            vv = rng.normal(size=len(ii))
            ci.append(ii)
            cv.append(vv)
            for t, v in zip(ii, vv):
                j = int(np.floor(((2 * nk + 1) - np.sqrt((2.0 * nk + 1) ** 2 - 8.0 * t)) / 2.0))
                while j * (2 * nk - j + 1) // 2 > t:
                    j -= 1
                while (j + 1) * (2 * nk - j) // 2 <= t:
                    j += 1
                i = int(t - j * (2 * nk - j + 1) // 2 + j)
                A[c][offs[k] + i, offs[k] + j] += v
                if i != j:
                    A[c][offs[k] + j, offs[k] + i] += v
        begs.append(np.concatenate([[0], np.cumsum([len(a) for a in ci])]))
        idxs.append(np.concatenate(ci))
        vals.append(np.concatenate(cv))
    p = lb.SdpaProblem(m, dims, rng.normal(size=m), begs, idxs, vals)
    cp = lb.constraint_couplings(p)
    A = [a.tocsr() for a in A]
    F0 = -A[0]                                                       # column 0 of the parsed arrays is -F0
    rows = [np.unique(a.tocoo().row) for a in A]
    # a position listed twice may cancel to an explicit zero only with probability 0; the row sets are those of the entries
    P = sp.lil_matrix((m, n))
    for c in range(1, m + 1):
        P[c - 1, rows[c]] = 1
    P = P.tocsr()
    O = sp.triu(P @ P.T, k=1).tocsr()
    O.sort_indices()
    assert np.array_equal(cp["ptr"], O.indptr) and np.array_equal(cp["col"], O.indices)
    assert np.array_equal(cp["overlap"], O.data.astype(np.int64))
    pi = np.repeat(np.arange(m), np.diff(cp["ptr"]))
    pick = rng.choice(len(pi), size=min(len(pi), 4000), replace=False)
    want = np.array([A[pi[t] + 1].multiply(A[cp["col"][t] + 1]).sum() for t in pick])
    assert np.allclose(cp["inner"][pick], want, rtol=1e-12, atol=1e-13)
    assert np.count_nonzero(want) > 0 and np.count_nonzero(want == 0) > 0     # both kinds of pair are present
    assert np.allclose(cp["cost_inner"], [A[c].multiply(F0).sum() for c in range(1, m + 1)], rtol=1e-12, atol=1e-13)
    rc = set(rows[0].tolist())
    assert np.array_equal(cp["rows_shared_with_cost"], [len(rc & set(rows[c].tolist())) for c in range(1, m + 1)])
