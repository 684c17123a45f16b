/*
 * lgpu_layout.h -- host-side construction of one cone's device layout (pure C++, no CUDA: it runs and is
 * tested without a GPU through lgpu_cone_classify / lgpu_cone_layout_query).
 *
 * Input: the reference reader's CSC over packed lower-triangular indices (column 0 = objective, columns 1..m =
 * constraints; LReadSDPA, io/lorads_file_io.c:59).  Output: every array lgpu_cone_upload sends to HBM.  This
 * replaces the reference's per-constraint preprocessing (AConeProcData / AConePresolveData,
 * lorads_sdp_conic.c:1185-1393: four mallocs per constraint and a (row+col) % size chained hash for the pattern
 * slots) with flat sorts, merges and counting sorts spread over the host cores (at n = 1e7 the
 * preprocessing, not the solve, was the start-up cost).  Every stage is deterministic: threads own disjoint output
 * ranges and visit their inputs in index order, so the arrays do not depend on the thread count
 * (LORADS_HOST_THREADS overrides it; default: hardware threads, at most 32, 1 for small cones).
 */
#ifndef LGPU_LAYOUT_H
#define LGPU_LAYOUT_H

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

extern "C" int lgpu_partition_rows(int64_t n, int world, int rank, int64_t *lo, int64_t *hi, int64_t *rows_per_rank);

static inline void unpack_lower(int64_t n, int64_t idx, int64_t *row, int64_t *col)
{
    /* column-major packed lower triangle: start(j) = j (2n - j + 1) / 2  (PACK_IDX, lorads_utils.h:167) */
    double t = (2.0 * (double)n + 1.0);
    int64_t j = (int64_t)floor((t - sqrt(t * t - 8.0 * (double)idx)) / 2.0);
    if (j < 0) j = 0;
    if (j > n - 1) j = n - 1;
    while (j > 0 && j * (2 * n - j + 1) / 2 > idx) --j;
    while (j + 1 < n && (j + 1) * (2 * n - j) / 2 <= idx) ++j;
    *col = j;
    *row = idx - j * (2 * n - j + 1) / 2 + j;
}

/* ------------------------------------------------------------------------------------------------
 * The reference's storage rules for one cone:
 *   coefficient class  ZERO / SPARSE / DENSE by nnz > 10 % of n(n+1)/2        sdpDataMatSetData, lorads_sdp_data.c:1180-1197
 *   container          SPARSE_CONE iff #non-zero A_i <= 0.3 m                  LUserDataChooseCone, lorads_user_data.c:105-109
 *   aggregate          dense iff n < 20, or a dense member, or |union pattern| >= 10 % of the triangle
 *                                                                              AConePresolveData, lorads_sdp_conic.c:1185-1393
 * plus what this library derives: the union pattern (sorted packed indices; empty when dense), nnzA and whether every
 * non-zero constraint is one diagonal entry (MaxCut-type, fused path).
 * ------------------------------------------------------------------------------------------------*/

/* ---- small thread helpers ---------------------------------------------------------------------*/
static int layout_threads(int64_t work)
{
    if (const char *s = getenv("LORADS_HOST_THREADS")) {
        const int v = atoi(s);
        if (v >= 1) return std::min(v, 256);
    }
    if (work < 200000) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    /* one process per GPU: the ranks of a node preprocess at the same time and share the cores (LORADS_LOCAL_RANKS is set
     * by the binary's --ranks, LOCAL_WORLD_SIZE by torchrun) */
    for (const char *name : {"LORADS_LOCAL_RANKS", "LOCAL_WORLD_SIZE"})
        if (const char *s = getenv(name)) {
            const int p = atoi(s);
            if (p > 1) { hw = std::max(1u, hw / (unsigned)p); break; }
        }
    return (int)std::min<unsigned>(hw, 32u);
}
template <class F> static void par_run(int T, F f)
{
    if (T <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)T - 1);
    for (int t = 1; t < T; ++t) th.emplace_back(f, t);
    f(0);
    for (auto &x : th) x.join();
}
/* f(t, lo, hi) over T contiguous pieces of [0, n) */
template <class F> static void par_ranges(int T, int64_t n, F f)
{
    par_run(T, [&](int t) { f(t, n * t / T, n * (t + 1) / T); });
}
/* in-place inclusive prefix sum of v[1..n] (v[0] stays): the usual ptr[i + 1] += ptr[i] loop, in three passes */
static void par_prefix(int T, std::vector<int32_t> &v)
{
    const int64_t n = (int64_t)v.size() - 1;
    if (T <= 1 || n < 100000) {
        for (int64_t i = 0; i < n; ++i) v[i + 1] += v[i];
        return;
    }
    std::vector<int64_t> part((size_t)T + 1, 0);
    par_ranges(T, n, [&](int t, int64_t lo, int64_t hi) {
        int64_t s = 0;
        for (int64_t i = lo; i < hi; ++i) s += v[i + 1];
        part[(size_t)t + 1] = s;
    });
    for (int t = 0; t < T; ++t) part[(size_t)t + 1] += part[(size_t)t];
    par_ranges(T, n, [&](int t, int64_t lo, int64_t hi) {
        int64_t run = part[(size_t)t] + v[0];
        for (int64_t i = lo; i < hi; ++i) { run += v[i + 1]; v[i + 1] = (int32_t)run; }
    });
}

/* sorted, duplicate-free copy of src[0..N): pieces sorted per thread (skipped when already ascending, the usual
 * case: every CSC column is ascending and MaxCut-type constraints come in row order), then rounds of two-way
 * merges in which every thread takes a slice of the output */
static void sorted_unique(const int64_t *src, int64_t N, int T, std::vector<int64_t> &out)
{
    out.clear();
    if (N <= 0) return;
    if (T > N) T = (int)std::max<int64_t>(1, N);
    std::unique_ptr<int64_t[]> A(new int64_t[(size_t)N]), B;
    std::vector<int64_t> bnd((size_t)T + 1);
    for (int t = 0; t <= T; ++t) bnd[(size_t)t] = N * t / T;
    par_run(T, [&](int t) {
        int64_t *a = A.get() + bnd[(size_t)t], *e = A.get() + bnd[(size_t)t + 1];
        std::copy(src + bnd[(size_t)t], src + bnd[(size_t)t + 1], a);
        /* ascending already, or two ascending runs (the piece that holds a column boundary), or anything */
        int64_t *d = std::is_sorted_until(a, e);
        if (d == e) return;
        if (std::is_sorted(d, e)) std::inplace_merge(a, d, e);
        else std::sort(a, e);
    });
    if (T > 1) B.reset(new int64_t[(size_t)N]);
    for (int width = 1; width < T; width *= 2) {
        const int groups = (T + 2 * width - 1) / (2 * width);
        const int per = std::max(1, T / groups); /* threads per merge */
        par_run(groups * per, [&](int id) {
            const int g = id / per, q = id % per;
            const int c0 = g * 2 * width, c1 = std::min(T, c0 + width), c2 = std::min(T, c0 + 2 * width);
            const int64_t *x0 = A.get() + bnd[(size_t)c0], *x1 = A.get() + bnd[(size_t)c1];
            const int64_t *y0 = x1, *y1 = A.get() + bnd[(size_t)c2];
            int64_t *dst = B.get() + bnd[(size_t)c0];
            /* slice q of the merge: a cut of the first run and the matching cut of the second (pieces are never empty) */
            const int64_t nx = x1 - x0;
            const int64_t *xa = x0 + nx * q / per, *xb = (q == per - 1) ? x1 : x0 + nx * (q + 1) / per;
            const int64_t *ya = (q == 0) ? y0 : std::lower_bound(y0, y1, *xa);
            const int64_t *yb = (q == per - 1) ? y1 : std::lower_bound(y0, y1, *xb);
            std::merge(xa, xb, ya, yb, dst + (xa - x0) + (ya - y0));
        });
        A.swap(B);
    }
    /* unique: count per piece, then compact */
    std::vector<int64_t> cnt((size_t)T + 1, 0);
    const int64_t *a = A.get();
    par_ranges(T, N, [&](int t, int64_t lo, int64_t hi) {
        int64_t c = 0;
        for (int64_t k = lo; k < hi; ++k) c += (k == 0 || a[k] != a[k - 1]);
        cnt[(size_t)t + 1] = c;
    });
    for (int t = 0; t < T; ++t) cnt[(size_t)t + 1] += cnt[(size_t)t];
    out.resize((size_t)cnt[(size_t)T]);
    par_ranges(T, N, [&](int t, int64_t lo, int64_t hi) {
        int64_t w = cnt[(size_t)t];
        for (int64_t k = lo; k < hi; ++k)
            if (k == 0 || a[k] != a[k - 1]) out[(size_t)w++] = a[k];
    });
}

struct ConeRules {
    int obj_type = 0;
    int64_t mA = 0, nnzA = 0, nnzP = 0;
    bool sparse_container = false, dense = false, diag_only = false;
    std::vector<int64_t> pat;
};
static void cone_rules(int64_t n, int64_t m, const int64_t *beg, const int64_t *idx, ConeRules &r)
{
    const int64_t tri = n * (n + 1) / 2;
    const int T = layout_threads(beg[m + 1] + m);
    auto mtype = [&](int64_t nnz) -> int {
        if (nnz == 0) return 0;
        if ((double)nnz > 0.1 * (double)tri) return 2;
        return 1;
    };
    r.obj_type = mtype(beg[1] - beg[0]);
    struct Part { int64_t mA = 0; bool any_dense = false, diag_only = true; };
    std::vector<Part> part((size_t)T);
    par_ranges(T, m, [&](int t, int64_t lo, int64_t hi) {
        Part p;
        for (int64_t col = lo + 1; col <= hi; ++col) {
            const int64_t nnz = beg[col + 1] - beg[col];
            if (nnz > 0) ++p.mA;
            if (mtype(nnz) == 2) p.any_dense = true;
            if (nnz > 1) p.diag_only = false;
            if (nnz == 1 && p.diag_only) {
                int64_t rr, cc;
                unpack_lower(n, idx[beg[col]], &rr, &cc);
                if (rr != cc) p.diag_only = false;
            }
        }
        part[(size_t)t] = p;
    });
    bool any_dense = r.obj_type == 2, diag_only = true;
    r.mA = 0;
    for (const Part &p : part) { r.mA += p.mA; any_dense = any_dense || p.any_dense; diag_only = diag_only && p.diag_only; }
    r.diag_only = diag_only && r.mA > 0;
    r.nnzA = beg[m + 1] - beg[1];
    r.sparse_container = !((double)r.mA > 0.3 * (double)m);
    r.dense = (n < 20) || any_dense;
    r.pat.clear();
    if (!r.dense) {
        sorted_unique(idx, beg[m + 1], T, r.pat);
        if ((double)r.pat.size() / (double)tri >= 0.1) {
            r.dense = true;
            std::vector<int64_t>().swap(r.pat);
        }
    }
    r.nnzP = r.dense ? tri : (int64_t)r.pat.size();
}

/* every host array of one cone, named like the DevCone member it is uploaded to */
struct ConeLayout {
    int obj_type = 0;
    bool dense = false, sparse_container = false, diag_only = false;
    int64_t mA = 0, nnzP = 0, nnzA = 0, nnzC = 0, nnzF = 0, max_con_len = 0, max_slot_len = 0;
    double c_nrm1 = 0, c_nrm2sq = 0, c_nrminf = 0;
    std::vector<int32_t> pat_row, pat_col;              /* [nnzP] aggregated lower pattern, sorted by (col, row) */
    std::vector<double> cval;                           /* [nnzP] C on the pattern */
    std::vector<int32_t> c_slot; std::vector<double> c_coef; /* [nnzC] (2 - delta) c */
    std::vector<int32_t> a_ptr, a_slot, con_gid; std::vector<double> a_coef; /* CSR by non-zero constraint */
    std::vector<int32_t> t_ptr, t_loc, t_gid; std::vector<double> t_val;    /* the same entries by pattern slot */
    std::vector<int32_t> f_ptr, f_col, f_slot;          /* full symmetric CSR: row -> (col, slot) */
    std::vector<int32_t> d_row; std::vector<double> d_val; /* diag_only: row and value per constraint */
    std::vector<double> mc_val;                         /* diag_only: C per full-CSR entry */
    std::vector<int32_t> rc_ptr, rc_gid; std::vector<double> rc_a; /* diag_only: row -> constraints */
    /* row relabelling for gather locality (fused layout only): device row perm[i] holds the caller's row i; the arrays
     * above that carry ROW labels (f_ptr/f_col, d_row, rc_*) are in the new labels, slot numbers are never touched, and
     * dev_pat_row/dev_pat_col are the pattern's labels as the device kernels need them (pat_row/pat_col stay the caller's) */
    bool reordered = false;
    double window_hits_before = 0.0, window_hits_after = 0.0; /* share of CSR entries within +-65536 rows of the diagonal */
    std::vector<int32_t> perm, iperm, dev_pat_row, dev_pat_col;
    /* with the rows, the constraints are renumbered in the order the rows list them (cperm[k] = device id of the caller's
     * constraint k), so that the row -> constraint walk of the fused kernels addresses the m-vectors sequentially */
    std::vector<int32_t> cperm;
    /* row-block partition (world > 1): this rank's slices and the exchange plan */
    bool partitioned = false, use_halo = false;
    int64_t lo = 0, hi = 0, rows_per_rank = 0, halo_rows = 0, send_rows = 0;
    std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt; /* per peer, in rows */
    std::vector<int64_t> dst_off;  /* per peer q: first row of THIS rank's block inside q's halo (= q's recv_off[rank]) */
    std::vector<int32_t> send_idx, halo_gid;
    std::vector<int32_t> lf_ptr, lf_col, lrc_ptr, lrc_gid;
    std::vector<double> lmc_val, lrc_a;
};

/* Breadth-first relabelling of the rows of a symmetric CSR (Cuthill-McKee without the degree sort): vertices that are
 * adjacent in the graph get nearby labels, so the factor rows a CSR row gathers lie within a narrow window -- L2 hits
 * instead of DRAM reads for the sparse product, thin halos for the row-block partition.  Gives up early on graphs without
 * locality (expander-like: the frontier passes 1/16 of all vertices within a few levels), where no labelling helps.
 * reference: none (the reference walks its entries in file order, lorads_sdp_data.c:750-763). */
static bool bfs_relabel(int64_t n, const std::vector<int32_t> &fp, const std::vector<int32_t> &fc, std::vector<int32_t> &perm)
{
    perm.assign((size_t)n, -1);
    std::vector<int32_t> order;
    order.reserve((size_t)n);
    int64_t next_seed = 0;
    while ((int64_t)order.size() < n) {
        while (perm[next_seed] >= 0) ++next_seed;
        size_t level_begin = order.size();
        perm[next_seed] = (int32_t)order.size();
        order.push_back((int32_t)next_seed);
        while (level_begin < order.size()) {
            const size_t level_end = order.size();
            for (size_t q = level_begin; q < level_end; ++q) {
                const int32_t v = order[q];
                for (int32_t e = fp[v]; e < fp[v + 1]; ++e) {
                    const int32_t u = fc[e];
                    if (perm[u] < 0) { perm[u] = (int32_t)order.size(); order.push_back(u); }
                }
            }
            if ((int64_t)(order.size() - level_end) > std::max<int64_t>(n / 16, 1024)) return false; /* no locality to find */
            level_begin = level_end;
        }
    }
    return true;
}
static double window_hit_share(int64_t n, const std::vector<int32_t> &fp, const std::vector<int32_t> &fc, const int32_t *perm)
{
    const int64_t W = 65536;
    int64_t hit = 0, tot = 0;
    const int64_t stride = std::max<int64_t>(1, n / 200000); /* a sample of the rows is enough */
    for (int64_t i = 0; i < n; i += stride)
        for (int32_t e = fp[i]; e < fp[i + 1]; ++e) {
            const int64_t a = perm ? perm[i] : i, b = perm ? perm[fc[e]] : fc[e];
            hit += (a - b <= W && b - a <= W);
            ++tot;
        }
    return tot ? (double)hit / (double)tot : 1.0;
}

/* returns 0, or 1 with a message in err */
static int build_cone_layout(int64_t n, int64_t m, const int64_t *beg, const int64_t *idx_in, const double *val_in, int world,
                             int rank, bool single_cone_no_lp, ConeLayout &L, std::string &err, bool force_halo = false)
{
    const int64_t tri = n * (n + 1) / 2;
    const int64_t total = beg[m + 1];
    const int T = layout_threads(total + m + n);
    const bool timing = getenv("LORADS_LAYOUT_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "lorads_b200: layout %-12s %.3f s (threads %d)\n", what, std::chrono::duration<double>(now - t_prev).count(), T);
        t_prev = now;
    };
    /* per-column ascending order (dataMatCreateSparseImpl sorts when needed, lorads_sdp_data.c:100-102); the reader
     * already delivers it, so the copy is made only when some column needs sorting */
    const int64_t *idx = idx_in;
    const double *val = val_in;
    std::vector<int64_t> idx_own;
    std::vector<double> val_own;
    {
        std::vector<char> bad((size_t)T, 0);
        par_ranges(T, total, [&](int t, int64_t lo, int64_t hi) {
            /* entry e and e - 1 are compared unless e starts a column */
            int64_t col = std::upper_bound(beg, beg + m + 2, lo) - beg - 1; /* column of entry lo */
            for (int64_t e = lo; e < hi; ++e) {
                while (col + 1 <= m + 1 && beg[col + 1] <= e) ++col;
                if (e > beg[col] && idx_in[e] < idx_in[e - 1]) { bad[(size_t)t] = 1; break; }
            }
        });
        bool any_bad = false;
        for (char b2 : bad) any_bad = any_bad || b2;
        if (any_bad) {
            idx_own.assign(idx_in, idx_in + total);
            val_own.assign(val_in, val_in + total);
            for (int64_t col = 0; col <= m; ++col) {
                const int64_t e0 = beg[col], e1 = beg[col + 1];
                if (std::is_sorted(idx_own.begin() + e0, idx_own.begin() + e1)) continue;
                std::vector<int64_t> o(e1 - e0);
                std::iota(o.begin(), o.end(), (int64_t)0);
                std::stable_sort(o.begin(), o.end(), [&](int64_t a, int64_t b2) { return idx_in[e0 + a] < idx_in[e0 + b2]; });
                for (size_t q = 0; q < o.size(); ++q) { idx_own[e0 + q] = idx_in[e0 + o[q]]; val_own[e0 + q] = val_in[e0 + o[q]]; }
            }
            idx = idx_own.data();
            val = val_own.data();
        }
    }
    lap("column order");
    /* storage rules of the reference, shared with the GPU-less lgpu_cone_classify */
    ConeRules rules;
    cone_rules(n, m, beg, idx, rules);
    lap("rules+pattern");
    L.obj_type = rules.obj_type;
    const int64_t mA = rules.mA;
    L.mA = mA;
    L.sparse_container = rules.sparse_container;
    const bool dense = rules.dense;
    std::vector<int64_t> pat;
    pat.swap(rules.pat);
    L.dense = dense;
    if (dense && tri >= (int64_t)1 << 30) { err = "dense aggregate too large for this build (n=" + std::to_string(n) + ")"; return 1; }
    const int64_t nnzP = dense ? tri : (int64_t)pat.size();
    if (nnzP >= (int64_t)1 << 31 || total >= (int64_t)1 << 31) { err = "cone too large for 32-bit device indices"; return 1; }
    L.nnzP = nnzP;
    L.pat_row.resize(nnzP);
    L.pat_col.resize(nnzP);
    int64_t ndiag = 0;
    if (dense) {
        /* the whole packed triangle, column by column */
        ndiag = n;
        par_ranges(T, n, [&](int, int64_t lo, int64_t hi) {
            for (int64_t j = lo; j < hi; ++j) {
                const int64_t k0 = j * (2 * n - j + 1) / 2;
                for (int64_t i = j; i < n; ++i) { L.pat_row[k0 + i - j] = (int32_t)i; L.pat_col[k0 + i - j] = (int32_t)j; }
            }
        });
    } else {
        std::vector<int64_t> nd((size_t)T, 0);
        par_ranges(T, nnzP, [&](int t, int64_t lo, int64_t hi) {
            int64_t c = 0;
            for (int64_t k = lo; k < hi; ++k) {
                int64_t r, q;
                unpack_lower(n, pat[k], &r, &q);
                L.pat_row[k] = (int32_t)r;
                L.pat_col[k] = (int32_t)q;
                c += r == q;
            }
            nd[(size_t)t] = c;
        });
        for (int64_t c : nd) ndiag += c;
    }
    lap("unpack");
    /* pattern slot of every entry.  Columns are ascending, so the search gallops forward from the previous hit:
     * a long column degenerates into a merge, a run of one-entry constraints in row order too */
    std::vector<int32_t> slot((size_t)total);
    if (dense) {
        par_ranges(T, total, [&](int, int64_t lo, int64_t hi) { for (int64_t e = lo; e < hi; ++e) slot[e] = (int32_t)idx[e]; });
    } else {
        const int64_t *P0 = pat.data();
        par_ranges(T, total, [&](int, int64_t lo, int64_t hi) {
            int64_t pos = 0;
            for (int64_t e = lo; e < hi; ++e) {
                const int64_t key = idx[e];
                if (e == lo || key < P0[pos]) {
                    pos = std::lower_bound(P0, P0 + (e == lo ? nnzP : pos), key) - P0;
                } else {
                    int64_t step = 1, base = pos;
                    while (base + step < nnzP && P0[base + step] <= key) { base += step; step *= 2; }
                    pos = std::upper_bound(P0 + base, P0 + std::min(base + step, nnzP), key) - P0 - 1;
                }
                slot[e] = (int32_t)pos;
            }
        });
    }
    std::vector<int64_t>().swap(pat);
    lap("slots");
    auto is_diag = [&](int32_t s) { return L.pat_row[s] == L.pat_col[s]; };
    /* objective: values on the pattern (duplicates add up in file order) and the norms, accumulated in entry order */
    std::vector<double> &cval = L.cval;
    cval.assign(nnzP, 0.0);
    L.nnzC = beg[1] - beg[0];
    L.c_slot.assign(slot.begin() + beg[0], slot.begin() + beg[1]);
    L.c_coef.resize(L.nnzC);
    par_ranges(T, L.nnzC, [&](int, int64_t lo, int64_t hi) {
        for (int64_t e = lo; e < hi; ++e) L.c_coef[e] = is_diag(L.c_slot[e]) ? val[beg[0] + e] : 2.0 * val[beg[0] + e];
    });
    {
        /* C on the pattern: duplicates of one position add up in file order; without any (the usual case) every slot is
         * written once and the scatter can be spread over the threads */
        std::vector<char> dup((size_t)T, 0);
        par_ranges(T, L.nnzC, [&](int t, int64_t lo, int64_t hi) {
            for (int64_t e = std::max<int64_t>(lo, 1); e < hi; ++e)
                if (L.c_slot[e] == L.c_slot[e - 1]) { dup[(size_t)t] = 1; break; }
        });
        bool any_dup = false;
        for (char d2 : dup) any_dup = any_dup || d2;
        if (any_dup) {
            for (int64_t e = 0; e < L.nnzC; ++e) cval[L.c_slot[e]] += val[beg[0] + e];
        } else {
            par_ranges(T, L.nnzC, [&](int, int64_t lo, int64_t hi) {
                for (int64_t e = lo; e < hi; ++e) cval[L.c_slot[e]] = 0.0 + val[beg[0] + e];
            });
        }
    }
    /* the norms, accumulated in entry order */
    L.c_nrm1 = L.c_nrm2sq = L.c_nrminf = 0.0;
    for (int64_t e = 0; e < L.nnzC; ++e) {
        const double v = val[beg[0] + e];
        const double w = is_diag(L.c_slot[e]) ? 1.0 : 2.0;
        L.c_nrm1 += w * fabs(v);
        L.c_nrm2sq += w * v * v;
        L.c_nrminf = std::max(L.c_nrminf, fabs(v));
    }
    lap("objective");
    /* constraints: CSR over non-zero constraints -- the entries of columns 1..m, in place */
    const int64_t a0 = beg[1];
    L.nnzA = total - a0;
    L.con_gid.reserve(mA);
    L.a_ptr.reserve(mA + 1);
    L.a_ptr.push_back(0);
    L.max_con_len = 0;
    bool diag_only = mA > 0;
    for (int64_t col = 1; col <= m; ++col) {
        const int64_t len = beg[col + 1] - beg[col];
        if (len == 0) continue;
        L.con_gid.push_back((int32_t)(col - 1));
        L.a_ptr.push_back((int32_t)(beg[col + 1] - a0));
        if (len != 1) diag_only = false;
        L.max_con_len = std::max(L.max_con_len, len);
    }
    L.a_slot.assign(slot.begin() + a0, slot.end());
    L.a_coef.resize(L.nnzA);
    {
        std::vector<char> offd((size_t)T, 0);
        par_ranges(T, L.nnzA, [&](int t, int64_t lo, int64_t hi) {
            for (int64_t e = lo; e < hi; ++e) {
                const bool dg = is_diag(L.a_slot[e]);
                L.a_coef[e] = dg ? val[a0 + e] : 2.0 * val[a0 + e];
                if (!dg) offd[(size_t)t] = 1;
            }
        });
        for (char o : offd) if (o) diag_only = false;
    }
    L.diag_only = diag_only;
    if (diag_only) {
        L.d_row.resize(mA);
        L.d_val.resize(mA);
        par_ranges(T, mA, [&](int, int64_t lo, int64_t hi) {
            for (int64_t t = lo; t < hi; ++t) { L.d_row[t] = L.pat_row[L.a_slot[t]]; L.d_val[t] = val[a0 + t]; }
        });
    }
    std::vector<int32_t>().swap(slot);
    lap("constraints");
    /* the same entries by pattern slot, stable in constraint order (the reference's accumulation order): every thread
     * owns a range of slots and walks ALL entries in order */
    L.t_ptr.assign(nnzP + 1, 0);
    L.t_loc.resize(L.nnzA);
    L.t_gid.resize(L.nnzA);
    L.t_val.resize(L.nnzA);
    {
        const int Tt = (int)std::max<int64_t>(1, std::min<int64_t>(T, nnzP / 4096 + 1));
        par_ranges(Tt, nnzP, [&](int, int64_t lo, int64_t hi) {
            for (int64_t e = 0; e < L.nnzA; ++e) {
                const int32_t s = L.a_slot[e];
                if (s >= lo && s < hi) L.t_ptr[s + 1]++;
            }
        });
        std::vector<int64_t> mx((size_t)Tt, 0);
        par_ranges(Tt, nnzP, [&](int t, int64_t lo, int64_t hi) {
            int64_t v = 0;
            for (int64_t k = lo; k < hi; ++k) v = std::max<int64_t>(v, L.t_ptr[k + 1]);
            mx[(size_t)t] = v;
        });
        L.max_slot_len = 0;
        for (int64_t v : mx) L.max_slot_len = std::max(L.max_slot_len, v);
        par_prefix(T, L.t_ptr);
        par_ranges(Tt, nnzP, [&](int, int64_t lo, int64_t hi) {
            if (lo >= hi) return;
            std::vector<int32_t> fill(L.t_ptr.begin() + lo, L.t_ptr.begin() + hi);
            for (int64_t t = 0; t < mA; ++t)
                for (int32_t e = L.a_ptr[t]; e < L.a_ptr[t + 1]; ++e) {
                    const int32_t s = L.a_slot[e];
                    if (s < lo || s >= hi) continue;
                    const int32_t p = fill[s - lo]++;
                    L.t_loc[p] = (int32_t)t;
                    L.t_gid[p] = L.con_gid[t];
                    L.t_val[p] = val[a0 + e];
                }
        });
    }
    lap("by slot");
    /* full symmetric CSR: row -> (col, slot).  The pattern is sorted by (col, row): row i first receives the entries
     * whose pattern row is i (columns <= i, ascending k = ascending column), then those whose pattern column is i
     * (rows > i), so every CSR row is sorted by column.  Threads own row ranges. */
    L.nnzF = 2 * nnzP - ndiag;
    if (L.nnzF >= (int64_t)1 << 31) { err = "cone too large for 32-bit device indices"; return 1; }
    std::vector<int32_t> &f_ptr = L.f_ptr, &f_col = L.f_col, &f_slot = L.f_slot;
    f_ptr.assign(n + 1, 0);
    f_col.resize(L.nnzF);
    f_slot.resize(L.nnzF);
    {
        /* first pattern entry of every column */
        std::vector<int32_t> cp((size_t)n + 1, 0);
        par_ranges(T, nnzP, [&](int, int64_t lo, int64_t hi) {
            for (int64_t k = lo; k < hi; ++k) {
                const int32_t j = L.pat_col[k], jp = k > 0 ? L.pat_col[k - 1] : -1;
                for (int32_t c = jp + 1; c <= j; ++c) cp[c] = (int32_t)k;
            }
        });
        for (int64_t c = (nnzP > 0 ? L.pat_col[nnzP - 1] + 1 : 0); c <= n; ++c) cp[c] = (int32_t)nnzP;
        const int Tr = (int)std::max<int64_t>(1, std::min<int64_t>(T, n / 1024 + 1));
        std::vector<int32_t> low((size_t)n, 0); /* entries of row i that come from its pattern row */
        par_ranges(Tr, n, [&](int, int64_t lo, int64_t hi) {
            for (int64_t k = 0; k < nnzP; ++k) {
                const int32_t i = L.pat_row[k];
                if (i >= lo && i < hi) low[i]++;
            }
            for (int64_t j = lo; j < hi; ++j) {
                int32_t up = cp[j + 1] - cp[j];
                if (up > 0 && L.pat_row[cp[j]] == j) --up; /* the diagonal is the first entry of its column */
                f_ptr[j + 1] = low[j] + up;
            }
        });
        par_prefix(T, f_ptr);
        par_ranges(Tr, n, [&](int, int64_t lo, int64_t hi) {
            if (lo >= hi) return;
            std::vector<int32_t> fill(f_ptr.begin() + lo, f_ptr.begin() + hi);
            for (int64_t k = 0; k < nnzP; ++k) {
                const int32_t i = L.pat_row[k];
                if (i < lo || i >= hi) continue;
                const int32_t p = fill[i - lo]++;
                f_col[p] = L.pat_col[k];
                f_slot[p] = (int32_t)k;
            }
            for (int64_t j = lo; j < hi; ++j)
                for (int32_t k = cp[j]; k < cp[j + 1]; ++k) {
                    const int32_t i = L.pat_row[k];
                    if (i == j) continue;
                    const int32_t p = fill[j - lo]++;
                    f_col[p] = i;
                    f_slot[p] = k;
                }
        });
    }
    lap("full CSR");
    /* Row relabelling (fused MaxCut-type layout, factors larger than the L2): LORADS_REORDER=1 forces it, =0 forbids it,
     * default: applied when it raises the share of near-diagonal entries by more than 0.15. */
    {
        int want = -1;
        if (const char *rv = getenv("LORADS_REORDER")) want = atoi(rv);
        const bool candidate = diag_only && mA == m && single_cone_no_lp && n >= 4 && !L.dense; /* dense kernels address rows by position */
        if (candidate && want != 0 && (want == 1 || n >= (int64_t)1 << 18)) {
            std::vector<int32_t> perm;
            const bool found = bfs_relabel(n, f_ptr, f_col, perm);
            if (found || want == 1) {
                if (!found) { /* forced on a graph without locality: finish the labelling in index order */
                    int32_t nxt = 0;
                    std::vector<uint8_t> used((size_t)n, 0);
                    for (int64_t i = 0; i < n; ++i) if (perm[i] >= 0) used[perm[i]] = 1;
                    for (int64_t i = 0; i < n; ++i)
                        if (perm[i] < 0) { while (used[nxt]) ++nxt; perm[i] = nxt; used[nxt] = 1; }
                }
                L.window_hits_before = window_hit_share(n, f_ptr, f_col, nullptr);
                L.window_hits_after = window_hit_share(n, f_ptr, f_col, perm.data());
                if (want == 1 || L.window_hits_after > L.window_hits_before + 0.15) {
                    L.reordered = true;
                    L.perm.swap(perm);
                    L.iperm.resize((size_t)n);
                    for (int64_t i = 0; i < n; ++i) L.iperm[L.perm[i]] = (int32_t)i;
                    /* CSR in the new labels: new row a = perm[i] takes row i's entries, columns relabelled and sorted */
                    std::vector<int32_t> np((size_t)n + 1, 0), nc((size_t)L.nnzF), ns((size_t)L.nnzF);
                    for (int64_t a = 0; a < n; ++a) np[a + 1] = f_ptr[L.iperm[a] + 1] - f_ptr[L.iperm[a]];
                    par_prefix(T, np);
                    par_ranges(T, n, [&](int, int64_t lo, int64_t hi) {
                        std::vector<std::pair<int32_t, int32_t>> row;
                        for (int64_t a = lo; a < hi; ++a) {
                            const int32_t i = L.iperm[a];
                            row.clear();
                            for (int32_t e = f_ptr[i]; e < f_ptr[i + 1]; ++e) row.emplace_back(L.perm[f_col[e]], f_slot[e]);
                            std::sort(row.begin(), row.end());
                            int32_t o = np[a];
                            for (auto &pr : row) { nc[o] = pr.first; ns[o] = pr.second; ++o; }
                        }
                    });
                    f_ptr.swap(np); f_col.swap(nc); f_slot.swap(ns);
                    for (int64_t t = 0; t < mA; ++t) L.d_row[t] = L.perm[L.d_row[t]];
                    L.dev_pat_row.resize((size_t)nnzP);
                    L.dev_pat_col.resize((size_t)nnzP);
                    par_ranges(T, nnzP, [&](int, int64_t lo, int64_t hi) {
                        for (int64_t k = lo; k < hi; ++k) { L.dev_pat_row[k] = L.perm[L.pat_row[k]]; L.dev_pat_col[k] = L.perm[L.pat_col[k]]; }
                    });
                }
            }
        }
    }
    lap("relabel");
    if (diag_only) {
        /* fused path layout.  The diagonal entry of every CSR row goes LAST (the others stay sorted by column): the
         * sparse product's walk then ends on the row's own factor row, which is exactly what its <X_i, (C X)_i> epilogue
         * needs -- no extra load, no compare inside the walk (k_mc_spmm). */
        par_ranges(T, n, [&](int, int64_t lo, int64_t hi) {
            for (int64_t j = lo; j < hi; ++j) {
                const int32_t b = f_ptr[j], e = f_ptr[j + 1];
                for (int32_t q = b; q < e - 1; ++q)
                    if (f_col[q] == (int32_t)j) {
                        const int32_t sl = f_slot[q];
                        for (int32_t t = q; t < e - 1; ++t) { f_col[t] = f_col[t + 1]; f_slot[t] = f_slot[t + 1]; }
                        f_col[e - 1] = (int32_t)j;
                        f_slot[e - 1] = sl;
                        break;
                    }
            }
        });
        /* C's value per full-CSR entry; row -> constraints, stable in constraint order */
        L.mc_val.resize(L.nnzF);
        par_ranges(T, L.nnzF, [&](int, int64_t lo, int64_t hi) { for (int64_t e = lo; e < hi; ++e) L.mc_val[e] = cval[f_slot[e]]; });
        L.rc_ptr.assign(n + 1, 0);
        L.rc_gid.resize(mA);
        L.rc_a.resize(mA);
        for (int64_t t = 0; t < mA; ++t) L.rc_ptr[L.d_row[t] + 1]++;
        par_prefix(T, L.rc_ptr);
        std::vector<int32_t> fill(L.rc_ptr.begin(), L.rc_ptr.end() - 1);
        for (int64_t t = 0; t < mA; ++t) {
            const int32_t q = fill[L.d_row[t]]++;
            L.rc_gid[q] = L.con_gid[t];
            L.rc_a[q] = L.d_val[t];
        }
        if (L.reordered) {
            /* relabelled rows: renumber the constraints in row order too (mA == m here), else the m-vector accesses of the
             * row-streaming kernels become random 8-byte gathers (measured: the step pass 36 % slower) */
            L.cperm.assign((size_t)m, -1);
            for (int64_t q = 0; q < mA; ++q) { L.cperm[L.rc_gid[q]] = (int32_t)q; L.rc_gid[q] = (int32_t)q; }
            for (auto &g : L.con_gid) g = L.cperm[g];
            for (auto &g : L.t_gid) g = L.cperm[g];
        }
    }
    lap("fused arrays");
    L.partitioned = false;
    if (world <= 1) return 0;
    /* row-block partition: this rank keeps the CSR rows, the row -> constraint lists and the vector rows of
     * [lo, hi); column indices and constraint ids stay global (they index the all-gathered factor and the
     * replicated-length m-vectors).  Only the fused MaxCut-type layout is partitioned by rows; every other problem
     * (several cones, an LP block, general constraints) keeps whole cones and is partitioned BY CONE (lgpu_api.cu). */
    if (!(diag_only && mA == m && single_cone_no_lp)) return 0;
    L.partitioned = true;
    int64_t lo, hi, rpr;
    lgpu_partition_rows(n, world, rank, &lo, &hi, &rpr);
    L.lo = lo; L.hi = hi; L.rows_per_rank = rpr;
    const int64_t nl = hi - lo;
    const std::vector<int32_t> &rc_ptr = L.rc_ptr;
    const int32_t e0 = f_ptr[lo], e1 = f_ptr[hi], k0 = rc_ptr[lo], k1 = rc_ptr[hi];
    L.lf_ptr.resize(nl + 1);
    L.lf_col.assign(f_col.begin() + e0, f_col.begin() + e1);
    L.lrc_ptr.resize(nl + 1);
    L.lrc_gid.assign(L.rc_gid.begin() + k0, L.rc_gid.begin() + k1);
    /* halo plan.  Every rank holds the whole CSR on the host during upload, so it can derive without any
     * communication both what it needs from each peer and what each peer needs from it (same lists, same
     * ascending order on both sides). */
    const int P = world;
    L.send_off.assign(P, 0); L.send_cnt.assign(P, 0); L.recv_off.assign(P, 0); L.recv_cnt.assign(P, 0);
    std::vector<int32_t> remap(n, -1);
    {
        std::vector<uint8_t> mark(n, 0);
        for (int32_t e = e0; e < e1; ++e) mark[f_col[e]] = 1;
        int64_t pos = 0;
        for (int q = 0; q < P; ++q) {
            int64_t qlo, qhi, qr;
            lgpu_partition_rows(n, P, q, &qlo, &qhi, &qr);
            L.recv_off[q] = pos;
            if (q != rank)
                for (int64_t j = qlo; j < qhi; ++j)
                    if (mark[j]) remap[j] = (int32_t)(rpr + pos++);
            L.recv_cnt[q] = pos - L.recv_off[q];
        }
        L.halo_rows = pos;
        for (int64_t j = lo; j < hi; ++j) remap[j] = (int32_t)(j - lo);
        std::vector<uint8_t> want(nl);
        for (int q = 0; q < P; ++q) {
            L.send_off[q] = (int64_t)L.send_idx.size();
            if (q == rank) continue;
            int64_t qlo, qhi, qr;
            lgpu_partition_rows(n, P, q, &qlo, &qhi, &qr);
            std::fill(want.begin(), want.end(), 0);
            for (int32_t e = f_ptr[qlo]; e < f_ptr[qhi]; ++e) {
                const int32_t j = f_col[e];
                if (j >= lo && j < hi) want[j - lo] = 1;
            }
            for (int64_t j = 0; j < nl; ++j)
                if (want[j]) L.send_idx.push_back((int32_t)j);
            L.send_cnt[q] = (int64_t)L.send_idx.size() - L.send_off[q];
        }
        L.send_rows = (int64_t)L.send_idx.size();
    }
    /* exchange only what is referenced when that is clearly less than everything (structured graphs: a thin
     * boundary; uniform random graphs at P = 2: nearly all rows, where the plain all-gather is the better tool).
     * The choice must be the same on every rank, so it is made from the halo sizes of ALL ranks (each rank can
     * count them: it holds the whole CSR) */
    {
        int64_t all_halo = 0;
        std::vector<uint8_t> seen(n);
        L.dst_off.assign(P, 0);
        for (int q = 0; q < P; ++q) {
            int64_t qlo, qhi, qr;
            lgpu_partition_rows(n, P, q, &qlo, &qhi, &qr);
            std::fill(seen.begin(), seen.end(), 0);
            for (int32_t e = f_ptr[qlo]; e < f_ptr[qhi]; ++e) {
                const int32_t j = f_col[e];
                if ((j < qlo || j >= qhi) && !seen[j]) { seen[j] = 1; ++all_halo; }
            }
            /* q's halo lists its sources in rank order: my block starts after the rows q references below my first row
             * (q's own rows are not marked) */
            if (q != rank) {
                int64_t before = 0;
                for (int64_t j = 0; j < lo; ++j) before += seen[j];
                L.dst_off[q] = before;
            }
        }
        L.use_halo = (double)all_halo < 0.85 * (double)n * (double)(P - 1);
    }
    if (force_halo) L.use_halo = true;
    if (const char *hv = getenv("LORADS_HALO")) L.use_halo = atoi(hv) != 0;
    if (L.use_halo) {
        for (auto &cj : L.lf_col) cj = remap[cj];
        /* halo row k holds global row halo_gid[k] (owners in rank order, ascending inside an owner) */
        L.halo_gid.resize((size_t)L.halo_rows);
        for (int64_t jg = 0; jg < n; ++jg)
            if ((jg < lo || jg >= hi) && remap[jg] >= 0) L.halo_gid[remap[jg] - rpr] = (int32_t)jg;
    }
    L.lmc_val.resize((size_t)(e1 - e0));
    L.lrc_a.assign(L.rc_a.begin() + k0, L.rc_a.begin() + k1);
    for (int64_t i = 0; i <= nl; ++i) { L.lf_ptr[i] = f_ptr[lo + i] - e0; L.lrc_ptr[i] = rc_ptr[lo + i] - k0; }
    for (int32_t e = e0; e < e1; ++e) L.lmc_val[e - e0] = cval[f_slot[e]];
    return 0;
}

#endif
