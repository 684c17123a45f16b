/*
 * sdpa_reader.c -- SDPA sparse (.dat-s) reader, 64-bit clean, single pass over an in-memory copy.
 *
 * Produces the same logical data the reference reader hands to the solver (LReadSDPA,
 * lorads/src/src_semi/io/lorads_file_io.c:59-455): per SDP block a CSC over PACKED lower-triangular
 * indices whose column 0 is the (negated) objective and columns 1..m the constraints, plus an LP CSC for
 * a trailing negative-dimension block.  Differences by design: indices are int64 (the reference's
 * INT32 build overflows for n > 46340), the file is tokenised from one buffer instead of per-line
 * sscanf, and entries are bucketed with a counting sort instead of a growing triplet store.
 *
 * File rules mirrored from the reference: comment lines start with '*' or '"' (:104-108); the block
 * dimension line may carry { } ( ) ' , (:143-175); only the LAST block may be LP (:159-190); b is free
 * form with commas (:203-220); entries are `con blk i j val`, 1-based, either triangle (:260-331);
 * |val| < 1e-12 is dropped with one warning (:288-294); objective entries are negated (:317-319).
 */
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lorads_host.h"

typedef struct {
    int64_t col;    /* constraint column (0 = objective) */
    int64_t idx;    /* packed lower index (SDP) or LP column id */
    double val;
    int64_t seq;
} trip_t;

typedef struct {
    trip_t *t;
    int64_t n, cap;
} trip_vec;

static int push_trip(trip_vec *v, int64_t col, int64_t idx, double val)
{
    if (v->n == v->cap) {
        int64_t nc = v->cap ? v->cap * 2 : 1024;
        trip_t *nt = (trip_t *)realloc(v->t, sizeof(trip_t) * (size_t)nc);
        if (!nt) return 1;
        v->t = nt;
        v->cap = nc;
    }
    v->t[v->n].col = col;
    v->t[v->n].idx = idx;
    v->t[v->n].val = val;
    v->t[v->n].seq = v->n;
    v->n++;
    return 0;
}

static int cmp_trip_sdp(const void *a, const void *b)
{
    const trip_t *x = (const trip_t *)a, *y = (const trip_t *)b;
    if (x->col != y->col) return x->col < y->col ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
static int cmp_trip_lp(const void *a, const void *b)
{
    const trip_t *x = (const trip_t *)a, *y = (const trip_t *)b;
    if (x->col != y->col) return x->col < y->col ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

/* to CSC with ncols columns */
static int to_csc(trip_vec *v, int64_t ncols, int lp, int64_t **beg, int64_t **idx, double **val)
{
    qsort(v->t, (size_t)v->n, sizeof(trip_t), lp ? cmp_trip_lp : cmp_trip_sdp);
    *beg = (int64_t *)calloc((size_t)ncols + 1, sizeof(int64_t));
    *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(v->n > 0 ? v->n : 1));
    *val = (double *)malloc(sizeof(double) * (size_t)(v->n > 0 ? v->n : 1));
    if (!*beg || !*idx || !*val) return 1;
    for (int64_t k = 0; k < v->n; ++k) (*beg)[v->t[k].col + 1]++;
    for (int64_t c = 0; c < ncols; ++c) (*beg)[c + 1] += (*beg)[c];
    for (int64_t k = 0; k < v->n; ++k) {
        (*idx)[k] = v->t[k].idx;
        (*val)[k] = v->t[k].val;
    }
    return 0;
}

static char *next_line(char *p, char *end)
{
    while (p < end && *p != '\n') ++p;
    return p < end ? p + 1 : end;
}

int lh_read_sdpa(const char *fname, lh_sdpa *out, int quiet)
{
    memset(out, 0, sizeof(*out));
    FILE *f = fopen(fname, "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 2);
    if (!buf) { fclose(f); return 1; }
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return 1; }
    fclose(f);
    buf[sz] = '\n';
    buf[sz + 1] = '\0';
    char *p = buf, *end = buf + sz + 1;
    int rc = 1;
    trip_vec *sdp = NULL, lpv = {0};
    /* comments */
    while (p < end && (*p == '*' || *p == '"')) p = next_line(p, end);
    char *q;
    int64_t m = strtoll(p, &q, 10);
    if (q == p || m <= 0) goto done;
    p = next_line(p, end);
    int64_t nblk = strtoll(p, &q, 10);
    if (q == p || nblk <= 0) goto done;
    p = next_line(p, end);
    /* block dimensions: numbers separated by anything that is not part of a number */
    int64_t *dims = (int64_t *)calloc((size_t)nblk, sizeof(int64_t));
    {
        int64_t got = 0;
        while (got < nblk && p < end) {
            while (p < end && !(isdigit((unsigned char)*p) || *p == '-' || *p == '+')) ++p;
            if (p >= end) break;
            int64_t v = strtoll(p, &q, 10);
            if (q == p) { ++p; continue; }
            dims[got++] = v;
            p = q;
        }
        if (got != nblk) { free(dims); goto done; }
        p = next_line(p, end);
    }
    int64_t nlp = 0, nsdp = nblk;
    for (int64_t k = 0; k < nblk; ++k) {
        if (dims[k] <= 0 && k != nblk - 1) { free(dims); goto done; } /* only the last block may be diagonal */
    }
    if (dims[nblk - 1] < 0) {
        nlp = -dims[nblk - 1];
        nsdp = nblk - 1;
    }
    out->m = m;
    out->nBlks = nsdp;
    out->nLpCols = nlp;
    out->blkDims = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nsdp > 0 ? nsdp : 1));
    for (int64_t k = 0; k < nsdp; ++k) out->blkDims[k] = dims[k];
    free(dims);
    /* right-hand side */
    out->b = (double *)calloc((size_t)m, sizeof(double));
    {
        int64_t got = 0;
        while (got < m && p < end) {
            while (p < end && !(isdigit((unsigned char)*p) || *p == '-' || *p == '+' || *p == '.')) ++p;
            if (p >= end) break;
            double v = strtod(p, &q);
            if (q == p) { ++p; continue; }
            out->b[got++] = v;
            p = q;
        }
        if (got != m) goto done;
        p = next_line(p, end);
    }
    /* entries */
    sdp = (trip_vec *)calloc((size_t)(nsdp > 0 ? nsdp : 1), sizeof(trip_vec));
    int warned = 0;
    while (p < end) {
        char *line = p;
        while (line < end && (*line == ' ' || *line == '\t' || *line == '\r')) ++line;
        if (line >= end) break;
        if (*line == '\n') { p = line + 1; continue; }
        int64_t con = strtoll(line, &q, 10);
        if (q == line) break;
        line = q;
        int64_t blk = strtoll(line, &q, 10);
        if (q == line) break;
        line = q;
        int64_t i = strtoll(line, &q, 10);
        if (q == line) break;
        line = q;
        int64_t j = strtoll(line, &q, 10);
        if (q == line) break;
        line = q;
        double v = strtod(line, &q);
        if (q == line) break;
        p = next_line(q, end);
        blk -= 1; i -= 1; j -= 1;
        if (con < 0 || con > m || blk < 0 || blk >= nblk) goto done;
        if (fabs(v) < 1e-12) {
            if (!warned && !quiet) printf("[Warning] Entry smaller than 1e-12 is ignored. \n");
            warned = 1;
            continue;
        }
        if (con == 0) v = -v;
        if (nlp > 0 && blk == nsdp) {
            if (i < 0 || i >= nlp) goto done;
            if (push_trip(&lpv, con, i, v)) goto done;
        } else {
            const int64_t n = out->blkDims[blk];
            if (i > j) { int64_t t = i; i = j; j = t; }
            if (i < 0 || j >= n) goto done;
            /* lower-triangular entry (row j, col i), column-major packed */
            const int64_t packed = (2 * n - i - 1) * i / 2 + j;
            if (push_trip(&sdp[blk], con, packed, v)) goto done;
        }
        out->nElems++;
    }
    out->matBeg = (int64_t **)calloc((size_t)(nsdp > 0 ? nsdp : 1), sizeof(int64_t *));
    out->matIdx = (int64_t **)calloc((size_t)(nsdp > 0 ? nsdp : 1), sizeof(int64_t *));
    out->matElem = (double **)calloc((size_t)(nsdp > 0 ? nsdp : 1), sizeof(double *));
    for (int64_t k = 0; k < nsdp; ++k)
        if (to_csc(&sdp[k], m + 1, 0, &out->matBeg[k], &out->matIdx[k], &out->matElem[k])) goto done;
    if (nlp > 0)
        if (to_csc(&lpv, m + 1, 1, &out->lpBeg, &out->lpIdx, &out->lpElem)) goto done;
    rc = 0;
done:
    if (sdp) {
        for (int64_t k = 0; k < nsdp; ++k) free(sdp[k].t);
        free(sdp);
    }
    free(lpv.t);
    free(buf);
    if (rc) lh_free_sdpa(out);
    return rc;
}

void lh_free_sdpa(lh_sdpa *d)
{
    if (d->matBeg)
        for (int64_t k = 0; k < d->nBlks; ++k) {
            free(d->matBeg[k]);
            free(d->matIdx[k]);
            free(d->matElem[k]);
        }
    free(d->matBeg);
    free(d->matIdx);
    free(d->matElem);
    free(d->lpBeg);
    free(d->lpIdx);
    free(d->lpElem);
    free(d->blkDims);
    free(d->b);
    memset(d, 0, sizeof(*d));
}
