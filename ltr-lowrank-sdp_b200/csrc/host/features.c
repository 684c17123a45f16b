/*
 * features.c -- hand-off to the reference's feature extractor.
 *
 * dataset/processor.py re-parses the same .dat-s in Python (SDPAParser, processor.py:40-200) and then loops over the m
 * constraint matrices to collect per-constraint statistics (FeatureExtractor._precompute_constraint_stats,
 * processor.py:246-295) and the constraint x row incidence pattern (_build_pattern_matrix, :318-345).  Both are
 * available here for free once lh_read_sdpa has parsed the file: this module computes them from the parsed arrays (one
 * pass over the entries, no matrix objects), so the Python side can take them instead of parsing again.
 *
 * Conventions of processor.py that are kept: all SDP blocks form ONE block-diagonal n x n matrix per constraint
 * (n = sum of block dimensions, block k starts at row offset[k]); the .dat-s stores one triangle, the extractor works on
 * the symmetrised matrix, so an off-diagonal entry counts twice; the LP block is not part of these statistics (the
 * extractor keeps the positive block dimensions only, processor.py:75-76, and drops the entries of a trailing LP block).
 * The reader negates the objective (lorads_file_io.c:317-319): column 0 of the parsed arrays holds -F0, the extractor
 * works on the file's F0, so everything that depends on the sign of C (its trace, <A_i, C>) is handed over in the
 * FILE's sign.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lorads_host.h"

static void unpack_lower(int64_t n, int64_t idx, int64_t *row, int64_t *col)
{
    /* packed column-major lower triangle: idx = (2n - j - 1) j / 2 + i  (PACK_IDX, lorads_utils.h:167) */
    int64_t j = (int64_t)(((2.0 * n + 1.0) - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * (double)idx)) / 2.0);
    while (j > 0 && (2 * n - j + 1) * j / 2 > idx) --j;
    while ((2 * n - j - 1) * (j + 1) / 2 + (j + 1) <= idx) ++j;
    *col = j;
    *row = idx - (2 * n - j - 1) * j / 2;
}

/* out: m x 7 doubles, row i = {frobenius norm, nnz, trace, norm of the diagonal, Gershgorin bound (max absolute row
 * sum), number of distinct rows touched, number of blocks spanned} of the symmetrised constraint matrix A_i
 * (processor.py:254-290).  `objective` != 0: the same seven numbers for C (constraint column 0) in out_obj[7]. */
int lh_constraint_stats(const lh_sdpa *d, double *out, double *out_obj)
{
    if (!d || !out) return 1;
    const int64_t m = d->m, nb = d->nBlks;
    int64_t n = 0, nmax = 0;
    for (int64_t k = 0; k < nb; ++k) { n += d->blkDims[k]; if (d->blkDims[k] > nmax) nmax = d->blkDims[k]; }
    double *rowsum = (double *)calloc((size_t)(nmax > 0 ? nmax : 1), sizeof(double));
    int64_t *touched = (int64_t *)malloc(sizeof(int64_t) * (size_t)(2 * (nmax > 0 ? nmax : 1)));
    if (!rowsum || !touched) { free(rowsum); free(touched); return 1; }
    for (int64_t c = (out_obj ? 0 : 1); c <= m; ++c) {
        double fro2 = 0.0, trace = 0.0, diag2 = 0.0, gersh = 0.0;
        int64_t nnz = 0, rows = 0, first_blk = -1, last_blk = -1;
        for (int64_t k = 0; k < nb; ++k) {
            const int64_t nk = d->blkDims[k];
            const int64_t e0 = d->matBeg[k][c], e1 = d->matBeg[k][c + 1];
            if (e1 <= e0) continue;
            int64_t nt = 0;
            for (int64_t e = e0; e < e1; ++e) {
                int64_t i, j;
                unpack_lower(nk, d->matIdx[k][e], &i, &j);
                const double v = c == 0 ? -d->matElem[k][e] : d->matElem[k][e], a = fabs(v); /* column 0 holds -F0 */
                if (i == j) {
                    nnz += 1; fro2 += v * v; trace += v; diag2 += v * v;
                    if (rowsum[i] == 0.0) touched[nt++] = i;
                    rowsum[i] += a;
                } else {
                    nnz += 2; fro2 += 2.0 * v * v;
                    if (rowsum[i] == 0.0) touched[nt++] = i;
                    rowsum[i] += a;
                    if (rowsum[j] == 0.0) touched[nt++] = j;
                    rowsum[j] += a;
                }
            }
            /* an entry stored as exactly 0.0 does not occur: the reader drops |v| < 1e-12 (lorads_file_io.c:288-294) */
            for (int64_t t = 0; t < nt; ++t) {
                if (rowsum[touched[t]] > gersh) gersh = rowsum[touched[t]];
                rowsum[touched[t]] = 0.0;
            }
            rows += nt;
            if (nt > 0) { if (first_blk < 0) first_blk = k; last_blk = k; }
        }
        double *o = c == 0 ? out_obj : out + 7 * (c - 1);
        o[0] = sqrt(fro2); o[1] = (double)nnz; o[2] = trace; o[3] = sqrt(diag2); o[4] = gersh; o[5] = (double)rows;
        /* processor.py:281-288 counts the blocks that intersect [first touched row, last touched row] */
        o[6] = (double)(rows > 0 ? (nb > 1 ? last_blk - first_blk + 1 : 1) : 0);
    }
    free(rowsum);
    free(touched);
    return 0;
}

/* constraint x row incidence (processor.py:318-345) as a CSR: ptr[m + 1], rows = the distinct global rows constraint i
 * touches, ascending.  Call with rows == NULL to size it (returns the count in *count). */
static int cmp_i64(const void *a, const void *b)
{
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}
int lh_constraint_rows(const lh_sdpa *d, int64_t *ptr, int64_t *rows, int64_t *count)
{
    if (!d || !ptr || !count) return 1;
    const int64_t m = d->m, nb = d->nBlks;
    int64_t cap = 0;
    for (int64_t k = 0; k < nb; ++k)
        for (int64_t c = 1; c <= m; ++c) { const int64_t l = 2 * (d->matBeg[k][c + 1] - d->matBeg[k][c]); if (l > cap) cap = l; }
    int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap * (nb > 0 ? nb : 1) + 1));
    if (!buf) return 1;
    int64_t total = 0;
    ptr[0] = 0;
    for (int64_t c = 1; c <= m; ++c) {
        int64_t nt = 0, off = 0;
        for (int64_t k = 0; k < nb; ++k) {
            for (int64_t e = d->matBeg[k][c]; e < d->matBeg[k][c + 1]; ++e) {
                int64_t i, j;
                unpack_lower(d->blkDims[k], d->matIdx[k][e], &i, &j);
                buf[nt++] = off + i;
                if (i != j) buf[nt++] = off + j;
            }
            off += d->blkDims[k];
        }
        qsort(buf, (size_t)nt, sizeof(int64_t), cmp_i64);
        int64_t u = 0;
        for (int64_t t = 0; t < nt; ++t)
            if (t == 0 || buf[t] != buf[t - 1]) { if (rows) rows[total + u] = buf[t]; ++u; }
        total += u;
        ptr[c] = total;
    }
    free(buf);
    *count = total;
    return 0;
}

/* ---- couplings: what the extractor's cost-alignment and edge features are made of ----------------------------------
 * processor.py:347-366 takes <A_i, C> for every constraint, :497-505 the share of A_i's rows that C touches too;
 * :580-600 multiplies the incidence pattern with its transpose (overlap = P P^T) and, for every pair with a common row,
 * :640-643 / :691-693 takes <A_i, A_j>.  All of it follows from the parsed arrays: positions are matched on
 * (block, packed index), an off-diagonal position counts twice (symmetrised matrices). */
typedef struct { int64_t key, con; double val; } pos_entry;

static int cmp_pos(const void *a, const void *b)
{
    const pos_entry *x = (const pos_entry *)a, *y = (const pos_entry *)b;
    if (x->key != y->key) return (x->key > y->key) - (x->key < y->key);
    return (x->con > y->con) - (x->con < y->con);
}

/* entries of the constraint columns [c0, c1] of all blocks, sorted by (position, constraint).  key = 2 * position +
 * (1 if off-diagonal): the low bit carries the weight of the position (1 diagonal, 2 off-diagonal) */
static pos_entry *sorted_positions(const lh_sdpa *d, int64_t c0, int64_t c1, int64_t *count)
{
    int64_t E = 0;
    for (int64_t k = 0; k < d->nBlks; ++k) E += d->matBeg[k][c1 + 1] - d->matBeg[k][c0];
    pos_entry *P = (pos_entry *)malloc(sizeof(pos_entry) * (size_t)(E > 0 ? E : 1));
    if (!P) return NULL;
    int64_t q = 0, base = 0;
    for (int64_t k = 0; k < d->nBlks; ++k) {
        const int64_t nk = d->blkDims[k];
        for (int64_t c = c0; c <= c1; ++c)
            for (int64_t e = d->matBeg[k][c]; e < d->matBeg[k][c + 1]; ++e) {
                int64_t i, j;
                unpack_lower(nk, d->matIdx[k][e], &i, &j);
                P[q].key = 2 * (base + d->matIdx[k][e]) + (i != j);
                P[q].con = c;
                P[q].val = d->matElem[k][e];
                ++q;
            }
        base += nk * (nk + 1) / 2;
    }
    qsort(P, (size_t)E, sizeof(pos_entry), cmp_pos);
    *count = E;
    return P;
}

/* inner[i] = <A_i, F0> (the cost matrix as the file states it), rows_shared[i] = |rows(A_i) & rows(F0)| (either may be
 * NULL), i = 0 .. m-1 */
int lh_constraint_cost_alignment(const lh_sdpa *d, double *inner, int64_t *rows_shared)
{
    if (!d) return 1;
    const int64_t m = d->m, nb = d->nBlks;
    if (inner) {
        int64_t nc = 0;
        pos_entry *C = sorted_positions(d, 0, 0, &nc);
        if (!C) return 1;
        int64_t base = 0;
        for (int64_t i = 0; i < m; ++i) inner[i] = 0.0;
        for (int64_t k = 0; k < nb; ++k) {
            const int64_t nk = d->blkDims[k];
            for (int64_t c = 1; c <= m; ++c)
                for (int64_t e = d->matBeg[k][c]; e < d->matBeg[k][c + 1]; ++e) {
                    int64_t i, j;
                    unpack_lower(nk, d->matIdx[k][e], &i, &j);
                    const int64_t key = 2 * (base + d->matIdx[k][e]) + (i != j);
                    int64_t lo = 0, hi = nc; /* first entry of C at this position (duplicates in the file add up) */
                    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (C[mid].key < key) lo = mid + 1; else hi = mid; }
                    for (; lo < nc && C[lo].key == key; ++lo) inner[c - 1] -= (i != j ? 2.0 : 1.0) * d->matElem[k][e] * C[lo].val; /* -: the file's F0 */
                }
            base += nk * (nk + 1) / 2;
        }
        free(C);
    }
    if (rows_shared) {
        int64_t n = 0;
        for (int64_t k = 0; k < nb; ++k) n += d->blkDims[k];
        unsigned char *mark = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
        int64_t *ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m + 1));
        int64_t cnt = 0, off = 0;
        if (!mark || !ptr || lh_constraint_rows(d, ptr, NULL, &cnt) != 0) { free(mark); free(ptr); return 1; }
        int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cnt > 0 ? cnt : 1));
        if (!rows || lh_constraint_rows(d, ptr, rows, &cnt) != 0) { free(mark); free(ptr); free(rows); return 1; }
        for (int64_t k = 0; k < nb; ++k) {
            for (int64_t e = d->matBeg[k][0]; e < d->matBeg[k][1]; ++e) {
                int64_t i, j;
                unpack_lower(d->blkDims[k], d->matIdx[k][e], &i, &j);
                mark[off + i] = 1;
                mark[off + j] = 1;
            }
            off += d->blkDims[k];
        }
        for (int64_t c = 0; c < m; ++c) {
            int64_t s = 0;
            for (int64_t t = ptr[c]; t < ptr[c + 1]; ++t) s += mark[rows[t]];
            rows_shared[c] = s;
        }
        free(mark); free(ptr); free(rows);
    }
    return 0;
}

/* The pairs i < j of constraints with a common row (the non-zeros of P P^T above its diagonal) as a CSR over i:
 * ptr[m + 1], col = j ascending, overlap = |rows(A_i) & rows(A_j)|, inner = <A_i, A_j>.  Call with col == NULL to size
 * (fills ptr and *count only). */
int lh_constraint_pairs(const lh_sdpa *d, int64_t *ptr, int64_t *col, int64_t *overlap, double *inner, int64_t *count)
{
    if (!d || !ptr || !count) return 1;
    const int64_t m = d->m, nb = d->nBlks;
    int64_t n = 0, nrows = 0;
    for (int64_t k = 0; k < nb; ++k) n += d->blkDims[k];
    int rc = 1;
    int64_t *cptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m + 1));
    int64_t *crow = NULL, *rptr = NULL, *rcon = NULL, *cur = NULL, *acc = NULL, *touched = NULL;
    pos_entry *P = NULL;
    if (!cptr || lh_constraint_rows(d, cptr, NULL, &nrows) != 0) goto done;
    crow = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nrows > 0 ? nrows : 1));
    rptr = (int64_t *)calloc((size_t)(n + 2), sizeof(int64_t));
    rcon = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nrows > 0 ? nrows : 1));
    cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    acc = (int64_t *)calloc((size_t)(m > 0 ? m : 1), sizeof(int64_t));
    touched = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
    if (!crow || !rptr || !rcon || !cur || !acc || !touched || lh_constraint_rows(d, cptr, crow, &nrows) != 0) goto done;
    /* row -> constraints, ascending constraint ids */
    for (int64_t t = 0; t < nrows; ++t) rptr[crow[t] + 1]++;
    for (int64_t r = 0; r < n; ++r) rptr[r + 1] += rptr[r];
    for (int64_t r = 0; r < n; ++r) cur[r] = rptr[r];
    for (int64_t c = 0; c < m; ++c)
        for (int64_t t = cptr[c]; t < cptr[c + 1]; ++t) rcon[cur[crow[t]]++] = c;
    /* overlap counts: constraint i meets, through each of its rows, the later constraints on that row.  The cursor of a
     * row only moves forward (i ascends), so every incidence is skipped once */
    for (int64_t r = 0; r < n; ++r) cur[r] = rptr[r];
    int64_t total = 0;
    ptr[0] = 0;
    for (int64_t i = 0; i < m; ++i) {
        int64_t nt = 0;
        for (int64_t t = cptr[i]; t < cptr[i + 1]; ++t) {
            const int64_t r = crow[t];
            while (cur[r] < rptr[r + 1] && rcon[cur[r]] <= i) ++cur[r];
            for (int64_t q = cur[r]; q < rptr[r + 1]; ++q) {
                const int64_t j = rcon[q];
                if (acc[j]++ == 0) touched[nt++] = j;
            }
        }
        qsort(touched, (size_t)nt, sizeof(int64_t), cmp_i64);
        for (int64_t t = 0; t < nt; ++t) {
            if (col) { col[total + t] = touched[t]; if (overlap) overlap[total + t] = acc[touched[t]]; if (inner) inner[total + t] = 0.0; }
            acc[touched[t]] = 0;
        }
        total += nt;
        ptr[i + 1] = total;
    }
    *count = total;
    if (col && inner && total > 0) {
        /* <A_i, A_j>: the constraints that share a POSITION meet in that position's group; a pair with a common position has
         * a common row, so it is in the CSR */
        int64_t E = 0;
        P = sorted_positions(d, 1, m, &E);
        if (!P) goto done;
        for (int64_t g0 = 0; g0 < E;) {
            int64_t g1 = g0 + 1;
            while (g1 < E && P[g1].key == P[g0].key) ++g1;
            const double w = (P[g0].key & 1) ? 2.0 : 1.0;
            for (int64_t a = g0; a < g1; ++a)
                for (int64_t b2 = a + 1; b2 < g1; ++b2) {
                    const int64_t i = P[a].con - 1, j = P[b2].con - 1;
                    if (i == j) continue; /* a position listed twice for one constraint */
                    int64_t lo = ptr[i], hi = ptr[i + 1];
                    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (col[mid] < j) lo = mid + 1; else hi = mid; }
                    if (lo < ptr[i + 1] && col[lo] == j) inner[lo] += w * P[a].val * P[b2].val;
                }
            g0 = g1;
        }
    }
    rc = 0;
done:
    free(cptr); free(crow); free(rptr); free(rcon); free(cur); free(acc); free(touched); free(P);
    return rc;
}
