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
 * the symmetrised matrix, so an off-diagonal entry counts twice; the LP block is not part of these statistics.
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
                const double v = d->matElem[k][e], a = fabs(v);
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
