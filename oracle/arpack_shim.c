/*
 * oracle/arpack_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference LoRADS solver links against ARPACK (dsaupd_/dseupd_) for ONE
 * purpose: the smallest eigenvalue of the dual slack C - A*(lambda) in
 * dual_infeasible() (reference: lorads/src/src_semi/data/lorads_sdp_conic.c:1636-1699;
 * CMake links an un-pinned system "arpack", lorads/CMakeLists.txt:113).  ARPACK
 * is absent from /root/reference and from this image, so the oracle build
 * satisfies the two symbols with this shim: a plain Lanczos iteration with full
 * re-orthogonalisation driven through ARPACK's reverse-communication protocol
 * (ido = 1: caller applies y = A x with x = workd[ipntr[0]-1], y = workd[ipntr[1]-1];
 * ido = 99: done).  Only mode 1, bmat = 'I', which = "SA", nev = 1 is supported,
 * which is the only way the reference calls it.  The smallest Ritz value is
 * returned by dseupd_ in d[0].  ARPACK's own tolerance at that call site is 1e-2,
 * so dual-infeasibility parity is loose by construction ("parity unpinned" for
 * that one number).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef UNIX_INT64
typedef long shim_int;
#else
typedef int shim_int;
#endif

typedef struct {
    int active;
    int n;
    int k;        /* number of Lanczos vectors built so far            */
    int kmax;     /* cap on the Krylov dimension                       */
    double *Q;    /* n x (kmax+1) Lanczos basis                        */
    double *alpha;
    double *beta; /* beta[j] couples q_j and q_{j+1}                   */
    double theta; /* current smallest Ritz value                       */
    double theta_prev;
    double tol;
} lanczos_state;

static lanczos_state S = {0};

/* smallest eigenvalue of the k x k symmetric tridiagonal (alpha, beta) by
 * Sturm-sequence bisection */
static double tridiag_min_eig(const double *a, const double *b, int k)
{
    double lo = a[0], hi = a[0];
    for (int i = 0; i < k; ++i) {
        double r = 0.0;
        if (i > 0) r += fabs(b[i - 1]);
        if (i < k - 1) r += fabs(b[i]);
        if (a[i] - r < lo) lo = a[i] - r;
        if (a[i] + r > hi) hi = a[i] + r;
    }
    for (int it = 0; it < 200; ++it) {
        double mid = 0.5 * (lo + hi);
        /* count eigenvalues < mid */
        int cnt = 0;
        double d = a[0] - mid;
        if (d < 0) cnt++;
        for (int i = 1; i < k; ++i) {
            double dd = (d == 0.0) ? 1e-300 : d;
            d = a[i] - mid - b[i - 1] * b[i - 1] / dd;
            if (d < 0) cnt++;
        }
        if (cnt >= 1) hi = mid; else lo = mid;
        if (hi - lo <= 1e-15 * (fabs(lo) + fabs(hi)) + 1e-300) break;
    }
    return 0.5 * (lo + hi);
}

void dsaupd_(int *ido, char *bmat, shim_int *n_, char *which, int *nev, double *tol,
             double *resid, int *ncv, double *v, shim_int *ldv, int *iparam, int *ipntr,
             double *workd, double *workl, int *lworkl, int *info)
{
    (void)bmat; (void)which; (void)nev; (void)v; (void)ldv; (void)workl; (void)lworkl; (void)ncv;
    int n = (int)(*n_);
    if (*ido == 0) {
        /* first call: set up and ask for A*q0 */
        if (S.active) { free(S.Q); free(S.alpha); free(S.beta); }
        S.active = 1;
        S.n = n;
        S.kmax = n < 300 ? n : 300;
        S.k = 0;
        S.Q = (double *)calloc((size_t)n * (size_t)(S.kmax + 1), sizeof(double));
        S.alpha = (double *)calloc((size_t)S.kmax + 1, sizeof(double));
        S.beta = (double *)calloc((size_t)S.kmax + 1, sizeof(double));
        S.theta = 0.0;
        S.theta_prev = 1e300;
        S.tol = (*tol > 0 ? *tol : 1e-8) * 1e-4; /* tighter than ARPACK's request */
        /* deterministic start vector */
        double nrm = 0.0;
        unsigned long long s = 88172645463325252ULL;
        for (int i = 0; i < n; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            S.Q[i] = ((double)(s % 2000001ULL) / 1000000.0) - 1.0;
            nrm += S.Q[i] * S.Q[i];
        }
        nrm = sqrt(nrm);
        for (int i = 0; i < n; ++i) S.Q[i] /= nrm;
        memcpy(workd, S.Q, sizeof(double) * (size_t)n);
        ipntr[0] = 1;
        ipntr[1] = n + 1;
        *ido = 1;
        *info = 0;
        (void)resid; (void)iparam;
        return;
    }
    /* returning from a mat-vec: w = A q_k sits in workd[n .. 2n) */
    {
        double *w = workd + n;
        int k = S.k;
        double *qk = S.Q + (size_t)k * n;
        double a = 0.0;
        for (int i = 0; i < n; ++i) a += qk[i] * w[i];
        S.alpha[k] = a;
        /* full re-orthogonalisation (twice) against all previous vectors */
        for (int pass = 0; pass < 2; ++pass) {
            for (int j = 0; j <= k; ++j) {
                double *qj = S.Q + (size_t)j * n;
                double c = 0.0;
                for (int i = 0; i < n; ++i) c += qj[i] * w[i];
                for (int i = 0; i < n; ++i) w[i] -= c * qj[i];
            }
        }
        double b = 0.0;
        for (int i = 0; i < n; ++i) b += w[i] * w[i];
        b = sqrt(b);
        S.beta[k] = b;
        S.k = k + 1;
        S.theta_prev = S.theta;
        S.theta = tridiag_min_eig(S.alpha, S.beta, S.k);
        int converged = 0;
        if (S.k >= 2 && fabs(S.theta - S.theta_prev) <= S.tol * (fabs(S.theta) + 1e-12)) converged = 1;
        if (b <= 1e-14 * (fabs(a) + 1.0)) converged = 1; /* invariant subspace */
        if (S.k >= S.kmax) converged = 1;
        if (converged) {
            *ido = 99;
            *info = 0;
            iparam[4] = 1; /* number of converged Ritz values */
            return;
        }
        double *qn = S.Q + (size_t)(k + 1) * n;
        for (int i = 0; i < n; ++i) qn[i] = w[i] / b;
        memcpy(workd, qn, sizeof(double) * (size_t)n);
        ipntr[0] = 1;
        ipntr[1] = n + 1;
        *ido = 1;
        *info = 0;
    }
}

void dseupd_(int *rvec, char *HowMny, int *select, double *d, double *z, shim_int *ldz, double *sigma,
             char *bmat, shim_int *n, char *which, int *nev, double *tol, double *resid, int *ncv,
             double *v, shim_int *ldv, int *iparam, int *ipntr, double *workd, double *workl,
             int *lworkl, int *info)
{
    (void)rvec; (void)HowMny; (void)select; (void)z; (void)ldz; (void)sigma; (void)bmat; (void)n;
    (void)which; (void)nev; (void)tol; (void)resid; (void)ncv; (void)v; (void)ldv; (void)iparam;
    (void)ipntr; (void)workd; (void)workl; (void)lworkl;
    d[0] = S.theta;
    *info = 0;
    if (S.active) {
        free(S.Q); free(S.alpha); free(S.beta);
        S.Q = NULL; S.alpha = NULL; S.beta = NULL;
        S.active = 0;
    }
}
