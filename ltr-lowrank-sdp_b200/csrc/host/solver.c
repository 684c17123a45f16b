/*
 * solver.c -- host-side solver lifecycle: problem upload, rank heuristic, seeded initial point,
 * rank augmentation, ALM->ADMM hand-off, re-optimisation round, dual infeasibility.
 *
 * Reference behaviour restated (lorads/src/src_semi/data/lorads_solver.c): LORADSDetermineRank :406-459,
 * LORADSInitALMVars / LORADSInitADMMVars :616-708,851-946 (srand(925), rand() stream order), CheckAllRankMax
 * :1066-1083, AUG_RANK :1154-1254, LORADS_ALMtoADMM :1351-1387, calculate_dual_infeasibility_solver :1396-1426,
 * reopt :1497-1539, initial_solver_state :1592-1614.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "lorads_host.h"

double lh_time(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + (double)tv.tv_usec * 1e-6;
}

int lh_time_is_up(lh_solver *S, const lh_params *p, double timeSolveStart, int strict)
{
    const double el = lh_time() - timeSolveStart;
    int up = strict ? (el > p->timeSecLimit) : (el >= p->timeSecLimit);
    if (S->gpu != NULL && lgpu_agree_flag(S->gpu, &up) != 0)
        fprintf(stderr, "lorads_b200: device error: %s\n", lgpu_last_error(S->gpu));
    return up;
}

#define GPU_TRY(S, call)                                                                   \
    do {                                                                                   \
        if ((call) != 0) {                                                                 \
            fprintf(stderr, "lorads_b200: device error: %s\n", lgpu_last_error((S)->gpu)); \
            return LH_RET_DEVICE;                                                          \
        }                                                                                  \
    } while (0)

int lh_setup_problem(lh_solver *S, const lh_sdpa *d, const lh_params *p)
{
    S->m = d->m;
    S->nCones = d->nBlks;
    S->nLpCols = d->nLpCols;
    S->blkDims = (int64_t *)malloc(sizeof(int64_t) * (size_t)(d->nBlks > 0 ? d->nBlks : 1));
    memcpy(S->blkDims, d->blkDims, sizeof(int64_t) * (size_t)d->nBlks);
    S->rank = (int64_t *)calloc((size_t)(d->nBlks > 0 ? d->nBlks : 1), sizeof(int64_t));
    S->rankMax = (int64_t *)calloc((size_t)(d->nBlks > 0 ? d->nBlks : 1), sizeof(int64_t));
    S->nnzRows = (int64_t *)calloc((size_t)(d->nBlks > 0 ? d->nBlks : 1), sizeof(int64_t));
    if (lgpu_create(&S->gpu, p->device) != 0) {
        fprintf(stderr, "lorads_b200: %s\n", lgpu_last_error(NULL));
        return LH_RET_DEVICE;
    }
    if (p->ranks > 1) GPU_TRY(S, lgpu_comm_init(S->gpu, p->ncclId, p->rank, p->ranks));
    GPU_TRY(S, lgpu_set_problem(S->gpu, d->m, d->b, (int)d->nBlks, d->blkDims, d->nLpCols));
    if (d->nLpCols > 0) GPU_TRY(S, lgpu_lp_upload(S->gpu, d->lpBeg, d->lpIdx, d->lpElem));
    for (int64_t c = 0; c < d->nBlks; ++c) {
        GPU_TRY(S, lgpu_cone_upload(S->gpu, (int)c, d->matBeg[c], d->matIdx[c], d->matElem[c]));
        int64_t info[6];
        GPU_TRY(S, lgpu_cone_info(S->gpu, (int)c, info));
        S->nnzRows[c] = info[0];
    }
    return LH_RET_OK;
}

void lh_determine_rank(lh_solver *S, const lh_params *p)
{
    const int use_fixed = p->fixedRank > 0, use_init = p->initRank > 0;
    for (int64_t c = 0; c < S->nCones; ++c) {
        const int64_t n = S->blkDims[c];
        const int64_t nnzRows = S->nnzRows[c];
        int64_t calc_max = (int64_t)sqrt((double)(2 * nnzRows)) + 1;
        if (calc_max > n) calc_max = n;
        if (use_fixed) {
            int64_t r = p->fixedRank < n ? p->fixedRank : n;
            if (r < 1) r = 1;
            S->rank[c] = r;
            S->rankMax[c] = r;
            continue;
        }
        S->rankMax[c] = calc_max;
        if (use_init) {
            int64_t r = p->initRank < n ? p->initRank : n;
            S->rank[c] = r < 1 ? 1 : r;
            continue;
        }
        int64_t r;
        if (p->timesLogRank <= 1e-6) r = calc_max;
        else if (nnzRows / n >= 20 && n <= 400 && S->nCones <= 3) r = calc_max;
        else {
            const double lg = ceil(p->timesLogRank * log((double)n));
            r = (lg < (double)calc_max) ? (int64_t)lg : calc_max;
        }
        S->rank[c] = r < 1 ? 1 : r;
    }
    /* rank schedule: entry 0 is the starting rank of every cone (capped by n and the sqrt(2m) bound) */
    if (S->scheduleLen > 0 && !use_fixed) {
        for (int64_t c = 0; c < S->nCones; ++c) {
            int64_t r = S->schedule[0];
            if (r > S->blkDims[c]) r = S->blkDims[c];
            if (r > S->rankMax[c]) r = S->rankMax[c];
            S->rank[c] = r < 1 ? 1 : r;
        }
        S->schedulePos = 0;
    }
}

int lh_init_variables(lh_solver *S, const lh_params *p)
{
    GPU_TRY(S, lgpu_alloc_vars(S->gpu, S->rank, (int)p->lbfgsListLength));
    int64_t maxel = S->nLpCols;
    for (int64_t c = 0; c < S->nCones; ++c)
        if (S->blkDims[c] * S->rank[c] > maxel) maxel = S->blkDims[c] * S->rank[c];
    double *buf = (double *)malloc(sizeof(double) * (size_t)(maxel > 0 ? maxel : 1));
    if (!buf) return LH_RET_DEVICE;
    /* the reference draws from ONE glibc rand() stream in this order: R of every cone, rLp, then uLp, vLp,
     * then per cone U, V (lorads_solver.c:625-669, 864-906) */
    lh_srand(925); /* glibc_rand.c: the C library's stream without its per-call lock */
    for (int64_t c = 0; c < S->nCones; ++c) {
        lh_random_fill(buf, S->blkDims[c] * S->rank[c]);
        GPU_TRY(S, lgpu_set_factor(S->gpu, LGPU_R, (int)c, buf));
    }
    if (S->nLpCols > 0) {
        lh_random_fill(buf, S->nLpCols);
        GPU_TRY(S, lgpu_set_lp(S->gpu, LGPU_R, buf));
        lh_random_fill(buf, S->nLpCols);
        GPU_TRY(S, lgpu_set_lp(S->gpu, LGPU_U, buf));
        lh_random_fill(buf, S->nLpCols);
        GPU_TRY(S, lgpu_set_lp(S->gpu, LGPU_V, buf));
    }
    for (int64_t c = 0; c < S->nCones; ++c) {
        lh_random_fill(buf, S->blkDims[c] * S->rank[c]);
        GPU_TRY(S, lgpu_set_factor(S->gpu, LGPU_U, (int)c, buf));
        lh_random_fill(buf, S->blkDims[c] * S->rank[c]);
        GPU_TRY(S, lgpu_set_factor(S->gpu, LGPU_V, (int)c, buf));
    }
    free(buf);
    return LH_RET_OK;
}

void lh_initial_state(lh_solver *S, const lh_params *p, lh_alm_state *alm, lh_admm_state *admm)
{
    double k[6];
    lgpu_constants(S->gpu, k);
    S->cObjNrm1 = k[0]; S->cObjNrm2 = k[1]; S->cObjNrmInf = k[2];
    S->bRHSNrm1 = k[3]; S->bRHSNrm2 = k[4]; S->bRHSNrmInf = k[5];
    double rho;
    if (p->initRho == 0) {
        int64_t sum = 0;
        for (int64_t c = 0; c < S->nCones; ++c) sum += S->blkDims[c];
        rho = 1 / sqrt((double)sum);
    } else {
        rho = p->initRho;
    }
    memset(alm, 0, sizeof(*alm));
    memset(admm, 0, sizeof(*admm));
    alm->dual_objective_value = alm->primal_objective_value = 1e+30;
    alm->l_1_dual_infeasibility = alm->l_1_primal_infeasibility = 1e+30;
    alm->l_inf_dual_infeasibility = alm->l_inf_primal_infeasibility = 1e+30;
    alm->rho = rho;
    admm->dual_objective_value = admm->primal_objective_value = admm->primal_dual_gap = 1e+30;
    admm->l_1_dual_infeasibility = admm->l_1_primal_infeasibility = 1e+30;
    admm->l_inf_dual_infeasibility = admm->l_inf_primal_infeasibility = 1e+30;
    admm->l_2_dual_infeasibility = admm->l_2_primal_infeasibility = 1e+30;
    admm->rho = rho;
    admm->iter = 0;
    admm->nBlks = S->nCones;
    S->scaleObjHis = 1;
}

int lh_all_rank_max(const lh_solver *S, double aug_factor)
{
    int64_t hit = 0;
    for (int64_t c = 0; c < S->nCones; ++c) {
        double nr = ceil((double)S->rank[c] * aug_factor);
        if (nr > (double)S->rankMax[c]) nr = (double)S->rankMax[c];
        if ((int64_t)nr >= S->rankMax[c]) ++hit;
    }
    return hit == S->nCones;
}

/* grows every cone: default r' = min(ceil(r * factor), rank_max); with a rank schedule the next entry is used
 * instead of the factor.  Returns (through *is_max) what the reference's AUG_RANK returns. */
int lh_aug_rank(lh_solver *S, double aug_factor, const lh_params *p)
{
    (void)p;
    if (lh_all_rank_max(S, 1.0)) return 1;
    int64_t *nr = (int64_t *)malloc(sizeof(int64_t) * (size_t)S->nCones);
    int use_sched = (S->scheduleLen > 0 && S->schedulePos + 1 < S->scheduleLen);
    if (use_sched) S->schedulePos++;
    for (int64_t c = 0; c < S->nCones; ++c) {
        double v = ceil((double)S->rank[c] * aug_factor);
        if (use_sched) {
            v = (double)S->schedule[S->schedulePos];
            if (v < (double)S->rank[c]) v = (double)S->rank[c];
        }
        if (v > (double)S->rankMax[c]) v = (double)S->rankMax[c];
        nr[c] = (int64_t)v;
        /* the reference prints "**Rank truncated to sqrt(2m)..." only under an #ifdef that is never defined on
         * Linux (quirk Q3), so nothing is printed here either */
    }
    if (lgpu_aug_rank(S->gpu, nr) != 0) {
        fprintf(stderr, "lorads_b200: device error: %s\n", lgpu_last_error(S->gpu));
        free(nr);
        return -1;
    }
    for (int64_t c = 0; c < S->nCones; ++c) S->rank[c] = nr[c];
    free(nr);
    if (S->scheduleLen > 0 && S->schedulePos + 1 >= S->scheduleLen) return 1; /* schedule exhausted: no further growth */
    return lh_all_rank_max(S, aug_factor);
}

void lh_alm_to_admm(lh_solver *S, lh_params *p, lh_alm_state *alm, lh_admm_state *admm)
{
    lgpu_alm_to_admm(S->gpu);
    admm->l_1_dual_infeasibility = alm->l_1_dual_infeasibility;
    admm->l_1_primal_infeasibility = alm->l_1_primal_infeasibility;
    admm->l_2_dual_infeasibility = alm->l_2_dual_infeasibility;
    admm->l_inf_dual_infeasibility = alm->l_inf_dual_infeasibility;
    admm->l_inf_primal_infeasibility = alm->l_inf_primal_infeasibility;
    admm->l_2_primal_infeasibility = alm->l_2_primal_infeasibility;
    admm->primal_dual_gap = alm->primal_dual_gap;
    admm->rho = alm->rho * p->heuristicFactor;
    if (alm->rho > p->rhoMax) {
        const double mx = p->rhoMax > alm->rho ? p->rhoMax : alm->rho;
        const double v = sqrt(mx / p->rhoMax) * p->rhoMax;
        admm->rho = v < alm->rho ? v : alm->rho;
        p->rhoMax = admm->rho;
    }
}

int lh_dual_infeasibility(lh_solver *S)
{
    double sum = 0.0;
    GPU_TRY(S, lgpu_dual_infeasibility(S->gpu, &sum));
    S->dimacDualInf = sum;
    S->dimacDualInf /= S->scaleObjHis;
    S->dimacDualInf /= (S->cObjNrm1 + 1);
    return LH_RET_OK;
}

double lh_reopt(lh_params *p, lh_solver *S, lh_alm_state *alm, lh_admm_state *admm, double *reopt_param,
                int64_t *reopt_alm_iter, int64_t *reopt_admm_iter, double timeSolveStart, int *admm_bad_iter_flag,
                int reopt_level)
{
    const int64_t old_maxALMIter = p->maxALMIter, old_maxADMMIter = p->maxADMMIter;
    const double old_rhoMax = p->rhoMax;
    p->maxALMIter = reopt_alm_iter[0] - 1 + alm->outerIter;
    p->maxADMMIter = reopt_admm_iter[0];
    /* objScale_dualvar: C *= s, lambda *= s, history *= s */
    S->scaleObjHis *= reopt_param[0];
    lgpu_obj_scale(S->gpu, reopt_param[0]);
    if (admm->rho <= p->rhoMax) alm->rho = admm->rho > alm->rho ? admm->rho : alm->rho;
    const double t0 = lh_time();
    lh_alm_optimize_reopt(p, S, alm, 1, sqrt(p->ALMRhoFactor), timeSolveStart);
    {
        const double mx = admm->rho > alm->rho ? admm->rho : alm->rho;
        const double v = sqrt(mx / admm->rho) * admm->rho;
        p->rhoMax = v > p->rhoMax ? v : p->rhoMax;
    }
    lh_alm_to_admm(S, p, alm, admm);
    if (*admm_bad_iter_flag == 0 || reopt_level < 2) {
        int64_t ceil_it = admm->iter * 4 < admm->iter + old_maxADMMIter ? admm->iter * 4 : admm->iter + old_maxADMMIter;
        const int rc = lh_admm_optimize_reopt(p, S, admm, ceil_it, timeSolveStart);
        *admm_bad_iter_flag = (rc == LH_RET_BAD_ITER) ? 1 : 0;
    }
    const double t1 = lh_time();
    p->maxALMIter = old_maxALMIter;
    p->maxADMMIter = old_maxADMMIter;
    p->rhoMax = old_rhoMax;
    return t1 - t0;
}

void lh_free_solver(lh_solver *S)
{
    if (S->gpu) lgpu_destroy(S->gpu);
    free(S->blkDims);
    free(S->rank);
    free(S->rankMax);
    free(S->nnzRows);
    free(S->schedule);
    memset(S, 0, sizeof(*S));
}
