/*
 * logging.c -- stdout/logfile tee, per-iteration rank trajectory, oracle rank, JSON result file.
 *
 * Output contract restated from lorads/src/src_semi/lorads_logging.c: log header lines `problem:` and
 * `oracle_method:` (:160-170), oracle rank = #eig(Gram) > 1e-6 * lambda_max per cone (:272-370,503-543),
 * JSON schema and number formats (:618-712).  The r x r Gram matrix comes from the device (lgpu_gram); its
 * eigenvalues are computed here with a cyclic Jacobi sweep (the reference calls LAPACK dsyevr).
 */
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "lorads_host.h"

void lh_log(lh_solver *S, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    if (S && S->logFp) {
        va_list ap2;
        va_start(ap2, fmt);
        vfprintf(S->logFp, fmt, ap2);
        va_end(ap2);
        fflush(S->logFp);
    }
}

void lh_logging_init(lh_solver *S, const lh_params *p, double solve_start)
{
    S->oracleMethod = p->oracleRankMethod;
    S->disableOracle = p->disableOracle;
    S->solveStartTime = solve_start;
    S->logFp = NULL;
    S->problemName[0] = S->inputPath[0] = S->jsonPath[0] = '\0';
    if (p->fname) {
        snprintf(S->inputPath, sizeof(S->inputPath), "%s", p->fname);
        const char *slash = strrchr(p->fname, '/');
        snprintf(S->problemName, sizeof(S->problemName), "%s", slash ? slash + 1 : p->fname);
        char *dot = strrchr(S->problemName, '.'); /* "G11.dat-s" -> "G11" (lorads_logging.c:41-46) */
        if (dot) *dot = '\0';
    }
    S->p1Count = S->p2Count = 0;
    S->p1Cap = S->p2Cap = 128;
    S->p1Curr = (int64_t *)calloc((size_t)S->p1Cap, sizeof(int64_t));
    S->p1Oracle = (int64_t *)calloc((size_t)S->p1Cap, sizeof(int64_t));
    S->p2Curr = (int64_t *)calloc((size_t)S->p2Cap, sizeof(int64_t));
    S->p2Oracle = (int64_t *)calloc((size_t)S->p2Cap, sizeof(int64_t));
    if (p->jsonFile) snprintf(S->jsonPath, sizeof(S->jsonPath), "%s", p->jsonFile);
    if (p->logFile) {
        S->logFp = fopen(p->logFile, "w");
        if (S->logFp) {
            fprintf(S->logFp, "problem:%s\n", S->problemName);
            fprintf(S->logFp, "oracle_method:%d epsilon:%g\n", S->oracleMethod, 1e-6);
            fflush(S->logFp);
        }
    }
}

void lh_logging_close(lh_solver *S)
{
    if (S->logFp) fclose(S->logFp);
    S->logFp = NULL;
    free(S->p1Curr); free(S->p1Oracle); free(S->p2Curr); free(S->p2Oracle);
    S->p1Curr = S->p1Oracle = S->p2Curr = S->p2Oracle = NULL;
}

int64_t lh_sum_rank(const lh_solver *S)
{
    int64_t t = 0;
    for (int64_t c = 0; c < S->nCones; ++c) t += S->rank[c];
    return t;
}

/* eigenvalues of a symmetric n x n matrix (row-major, destroyed) by cyclic Jacobi rotations, ascending in w */
static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y);
}
int lh_sym_eigvals(int n, double *a, double *w)
{
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < n; ++i) {
            dia += a[i * n + i] * a[i * n + i];
            for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
        }
        if (off <= 1e-30 * (dia + off) || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q];
                if (apq == 0.0) continue;
                const double app = a[p * n + p], aqq = a[q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = c * akp - s * akq;
                    a[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = c * apk - s * aqk;
                    a[q * n + k] = s * apk + c * aqk;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = a[i * n + i];
    qsort(w, (size_t)n, sizeof(double), cmp_double);
    return 0;
}

/* Number of eigenvalues of the symmetric n x n matrix a (row-major, destroyed) above eps * lambda_max -- all the oracle
 * rank needs (lorads_compute_oracle_rank, lorads_logging.c:503-543, calls dsyevr for the whole spectrum).  Householder
 * reduction to tridiagonal form, then Sturm counts: lambda_max by bisection, the rank as n - #{lambda <= cut}.  O(n^3)
 * with a small constant instead of the cyclic Jacobi sweeps of lh_sym_eigvals (93 ms -> ~3 ms at n = 158); returns -1
 * if lambda_max <= 0 (rank 0 by the reference's rule). */
static int sturm_count_below(int n, const double *d, const double *e2, double x, double tiny)
{
    /* number of eigenvalues of tridiag(d, e) that are < x: negative pivots of the LDL^T of T - x I */
    int cnt = 0;
    double q = 1.0;
    for (int i = 0; i < n; ++i) {
        q = d[i] - x - (i > 0 ? e2[i - 1] / q : 0.0);
        if (fabs(q) < tiny) q = -tiny;
        if (q < 0.0) ++cnt;
    }
    return cnt;
}

int64_t lh_sym_rank(int n, double *a, double eps)
{
    if (n <= 0) return 0;
    double *d = (double *)malloc(sizeof(double) * (size_t)n * 4);
    if (!d) return -1;
    double *e2 = d + n, *v = e2 + n, *p = v + n;
    /* Householder: after step k, row/column k is (.., e[k-1], d[k], e[k], 0, ..) */
    for (int k = 0; k + 2 < n; ++k) {
        double scale = 0.0;
        for (int i = k + 1; i < n; ++i) scale = fmax(scale, fabs(a[i * n + k]));
        d[k] = a[k * n + k];
        if (scale == 0.0) { e2[k] = 0.0; continue; }
        double nrm2 = 0.0;
        for (int i = k + 1; i < n; ++i) { v[i] = a[i * n + k] / scale; nrm2 += v[i] * v[i]; }
        const double alpha = (v[k + 1] >= 0.0 ? -1.0 : 1.0) * sqrt(nrm2);
        e2[k] = (alpha * scale) * (alpha * scale);
        v[k + 1] -= alpha;
        const double vtv = nrm2 - 2.0 * alpha * (v[k + 1] + alpha) + alpha * alpha; /* |v|^2 after the shift */
        if (vtv == 0.0) continue;
        /* B <- H B H on the trailing block, H = I - 2 v v^T / (v^T v):  B - v q^T - q v^T,  q = p - (v^T p / v^T v) v */
        double vtp = 0.0;
        for (int i = k + 1; i < n; ++i) {
            double t = 0.0;
            for (int j = k + 1; j < n; ++j) t += a[i * n + j] * v[j];
            p[i] = 2.0 * t / vtv;
            vtp += v[i] * p[i];
        }
        const double kk = vtp / vtv;
        for (int i = k + 1; i < n; ++i) p[i] -= kk * v[i];
        for (int i = k + 1; i < n; ++i)
            for (int j = k + 1; j < n; ++j) a[i * n + j] -= v[i] * p[j] + p[i] * v[j];
    }
    if (n >= 2) {
        d[n - 2] = a[(n - 2) * n + (n - 2)];
        e2[n - 2] = a[(n - 1) * n + (n - 2)] * a[(n - 1) * n + (n - 2)];
    }
    d[n - 1] = a[(n - 1) * n + (n - 1)];
    /* Gershgorin interval, then lambda_max by bisection on the Sturm count */
    double lo = d[0], hi = d[0];
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? sqrt(e2[i - 1]) : 0.0) + (i + 1 < n ? sqrt(e2[i]) : 0.0);
        lo = fmin(lo, d[i] - r);
        hi = fmax(hi, d[i] + r);
    }
    const double span = fmax(fabs(lo), fabs(hi));
    const double tiny = fmax(span * 1e-30, 1e-300); /* a pivot this small is taken as negative (as in LAPACK's dstebz) */
    double a0 = lo, b0 = hi;
    for (int it = 0; it < 200 && b0 - a0 > 4e-16 * fmax(fabs(a0), fabs(b0)); ++it) {
        const double mid = 0.5 * (a0 + b0);
        if (mid <= a0 || mid >= b0) break;
        if (sturm_count_below(n, d, e2, mid, tiny) >= n) b0 = mid; /* all eigenvalues < mid */
        else a0 = mid;
    }
    const double lmax = 0.5 * (a0 + b0);
    int64_t rk = -1;
    if (lmax > 0.0) {
        /* eigenvalues > cut  =  n - #{lambda < cut} - #{lambda == cut}; equality has measure zero, the count just above
         * the cut is what dsyevr's rounded eigenvalues would give as well */
        const double cut = eps * lmax;
        rk = n - sturm_count_below(n, d, e2, nextafter(cut, INFINITY), tiny);
    }
    free(d);
    return rk;
}

int64_t lh_oracle_rank(lh_solver *S, int phase)
{
    if (S->disableOracle) return 0;
    const double eps = 1e-6;
    int64_t total = 0;
    for (int64_t c = 0; c < S->nCones; ++c) {
        const int r = (int)S->rank[c];
        double *g = (double *)malloc(sizeof(double) * (size_t)r * r);
        if (!g || lgpu_gram(S->gpu, phase, (int)c, g) != 0) {
            free(g);
            return -1;
        }
        int64_t rk = lh_sym_rank(r, g, eps);
        if (rk < 0) rk = 0; /* lambda_max <= 0: rank 0 (lorads_logging.c:535-541) */
        total += rk;
        free(g);
    }
    return total;
}

static void push_pair(int64_t **cur, int64_t **orc, int64_t *count, int64_t *cap, int64_t a, int64_t b)
{
    if (*count >= *cap) {
        const int64_t nc = *cap * 2;
        int64_t *n1 = (int64_t *)realloc(*cur, sizeof(int64_t) * (size_t)nc);
        int64_t *n2 = (int64_t *)realloc(*orc, sizeof(int64_t) * (size_t)nc);
        if (!n1 || !n2) return;
        *cur = n1; *orc = n2; *cap = nc;
    }
    (*cur)[*count] = a;
    (*orc)[*count] = b;
    (*count)++;
}

void lh_append_trajectory(lh_solver *S, int phase, int64_t cur_rank, int64_t oracle_rank)
{
    if (phase == 1) push_pair(&S->p1Curr, &S->p1Oracle, &S->p1Count, &S->p1Cap, cur_rank, oracle_rank);
    else if (phase == 2) push_pair(&S->p2Curr, &S->p2Oracle, &S->p2Count, &S->p2Cap, cur_rank, oracle_rank);
}

static void json_int_array(FILE *f, const int64_t *a, int64_t n)
{
    fprintf(f, "[");
    for (int64_t i = 0; i < n; ++i) {
        if (i > 0) fprintf(f, ", ");
        fprintf(f, "%lld", (long long)a[i]);
    }
    fprintf(f, "]");
}

void lh_write_json(lh_solver *S, int64_t final_oracle_rank, double pobj, double dobj, double l1, double linf, double gap,
                   double solve_time, double rho_max, double heuristic_factor)
{
    if (S->jsonPath[0] == '\0') return;
    FILE *f = fopen(S->jsonPath, "w");
    if (!f) {
        lh_log(S, "Warning: Could not open JSON file for writing: %s\n", S->jsonPath);
        return;
    }
    fprintf(f, "{\n");
    fprintf(f, "  \"problem_id\": \"%s\",\n", S->problemName);
    fprintf(f, "  \"file_path\": \"%s\",\n", S->inputPath);
    fprintf(f, "  \"metrics\": {\n");
    fprintf(f, "    \"oracle_rank\": %lld,\n", (long long)final_oracle_rank);
    fprintf(f, "    \"primal_obj\": %.16e,\n", pobj);
    fprintf(f, "    \"dual_obj\": %.16e,\n", dobj);
    fprintf(f, "    \"constr_violation_l1\": %.16e,\n", l1);
    fprintf(f, "    \"constr_violation_inf\": %.16e,\n", linf);
    fprintf(f, "    \"primal_dual_gap\": %.16e,\n", gap);
    fprintf(f, "    \"solve_time_sec\": %.16e,\n", solve_time);
    fprintf(f, "    \"rho_max\": %.16e,\n", rho_max);
    fprintf(f, "    \"heuristic_factor\": %.16e\n", heuristic_factor);
    fprintf(f, "  },\n");
    fprintf(f, "  \"trajectory\": {\n");
    fprintf(f, "    \"phase_1\": {\n      \"curr_rank\": ");
    json_int_array(f, S->p1Curr, S->p1Count);
    fprintf(f, ",\n      \"oracle_rank\": ");
    json_int_array(f, S->p1Oracle, S->p1Count);
    fprintf(f, "\n    },\n");
    fprintf(f, "    \"phase_2\": {\n      \"curr_rank\": ");
    json_int_array(f, S->p2Curr, S->p2Count);
    fprintf(f, ",\n      \"oracle_rank\": ");
    json_int_array(f, S->p2Oracle, S->p2Count);
    fprintf(f, "\n    }\n");
    fprintf(f, "  }\n");
    fprintf(f, "}\n");
    fclose(f);
    lh_log(S, "JSON output written to: %s\n", S->jsonPath);
}
