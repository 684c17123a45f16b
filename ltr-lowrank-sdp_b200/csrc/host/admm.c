/*
 * admm.c -- phase 2: ADMM on the split X = U V^T, CG solves on the device.
 *
 * Control flow, thresholds and the printed line restate LORADSADMMOptimize / LORADSADMMOptimize_reopt
 * (lorads/src/src_semi/lorads_alg/lorads_admm.c:84-209 and :222-363).  One sweep (admmUpdateVar) is one
 * C-ABI call; objective, dual objective and DIMACS errors are one call each.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lorads_host.h"

#define GPU_TRY(S, call)                                                                   \
    do {                                                                                   \
        if ((call) != 0) {                                                                 \
            fprintf(stderr, "lorads_b200: device error: %s\n", lgpu_last_error((S)->gpu)); \
            return LH_RET_DEVICE;                                                          \
        }                                                                                  \
    } while (0)

static void admm_print(lh_solver *S, const lh_admm_state *st, double t, int64_t cur_rank, int64_t oracle_rank)
{
    lh_log(S,
           "ADMM Iter:%d pObj:%5.5e dObj:%5.5e pInfea(1):%5.5e pInfea(Inf):%5.5e pdGap:%5.5e rho:%3.2f cgIter:%d "
           "CurrRank:%lld OracleRank:%lld Time:%3.2f\n",
           (int)st->iter, st->primal_objective_value, st->dual_objective_value, st->l_1_primal_infeasibility,
           st->l_inf_primal_infeasibility, st->primal_dual_gap, st->rho,
           (int)((double)st->cg_iter / (double)st->nBlks), (long long)cur_rank, (long long)oracle_rank, t);
}

static void admm_record(lh_solver *S, const lh_admm_state *st, double phase_time)
{
    const int64_t cur = lh_sum_rank(S);
    const int64_t orc = lh_oracle_rank(S, 2);
    lh_append_trajectory(S, 2, cur, orc);
    admm_print(S, st, phase_time, cur, orc);
}

/* calObj_admm + LORADSCalDualObj: objective on R = (U+V)/2, dual objective b^T lambda */
static int objs_admm(lh_solver *S)
{
    double p = 0, d = 0;
    GPU_TRY(S, lgpu_cal_obj(S->gpu, 1, &p));
    GPU_TRY(S, lgpu_cal_dual_obj(S->gpu, &d));
    S->pObjVal = p / S->scaleObjHis;
    S->dObjVal = d / S->scaleObjHis;
    return 0;
}
/* updateDimacsADMM: R = (U+V)/2, A(RR^T) from scratch (this also REPLACES constrVal / constrValSum by the
 * values at R, exactly as the reference's primalInfeasibility does), gap from the current objectives */
static int dimacs_admm(lh_solver *S)
{
    double l1 = 0;
    GPU_TRY(S, lgpu_average_uv(S->gpu));
    GPU_TRY(S, lgpu_primal_infeasibility(S->gpu, LGPU_PAIR_RR, &l1));
    S->dimacConstrVio = l1;
    const double gap = S->pObjVal - S->dObjVal;
    S->dimacGap = fabs(gap) / (1 + fabs(S->pObjVal) + fabs(S->dObjVal));
    return 0;
}
static void take_errors(const lh_solver *S, lh_admm_state *st, int with_norms)
{
    st->primal_objective_value = S->pObjVal;
    st->dual_objective_value = S->dObjVal;
    st->primal_dual_gap = S->dimacGap;
    st->l_1_primal_infeasibility = S->dimacConstrVio;
    if (with_norms) {
        st->l_inf_primal_infeasibility = S->dimacConstrVio * (1 + S->bRHSNrm1) / (1 + S->bRHSNrmInf);
        st->l_2_primal_infeasibility = S->dimacConstrVio * (1 + S->bRHSNrm1) / (1 + S->bRHSNrm2);
    }
}

static double buffer_mean(const double *buf, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += fabs(buf[i]);
    return s / (double)n;
}

static int admm_loop(lh_params *p, lh_solver *S, lh_admm_state *st, int64_t iter_celling, double timeSolveStart, int reopt)
{
    if (st->primal_dual_gap <= p->phase2Tol && st->l_1_primal_infeasibility <= p->phase2Tol) return LH_RET_OK;
    double cg_tol;
    const int64_t cg_max = 800;
    st->rho = st->rho < p->rhoMax ? st->rho : p->rhoMax;
    S->cgIter = 0;
    GPU_TRY(S, lgpu_init_constr_val(S->gpu, LGPU_PAIR_UV));
    if (objs_admm(S) || dimacs_admm(S)) return LH_RET_DEVICE;
    take_errors(S, st, 1);
    if (reopt) lh_log(S, "enter admm reopt \n");
    double cur_rho_max = p->rhoMax;
    double old_mean = 1e30;
    double pinf_buf[10];
    memset(pinf_buf, 0, sizeof(pinf_buf));
    int bad_pd = 0;
    const int count = 0; /* never advanced in the reference either: only slot 0 of the window is written */
    const int bad_pd_cap = reopt ? 200 : 800;
    const double t0 = lh_time();
    while (st->iter <= p->maxADMMIter || st->primal_dual_gap >= p->phase2Tol || st->l_1_primal_infeasibility >= p->phase2Tol) {
        if (st->iter >= iter_celling) {
            if (reopt) admm_record(S, st, 0);
            break;
        }
        const double f = reopt ? 1e-4 : 1e-2;
        cg_tol = st->l_1_primal_infeasibility * f < 1e-8 ? st->l_1_primal_infeasibility * f : 1e-8;
        GPU_TRY(S, lgpu_admm_update_var(S->gpu, st->rho, cg_tol, cg_max, &S->cgIter));
        st->cg_iter = S->cgIter;
        if (objs_admm(S) || dimacs_admm(S)) return LH_RET_DEVICE;
        take_errors(S, st, 1);
        admm_record(S, st, lh_time() - t0);
        if (st->l_inf_primal_infeasibility >= 1e10 || st->primal_dual_gap >= 1 - 1e-8) {
            lh_log(S, "Numerical Error!\n");
            return LH_RET_NUM_ERR;
        }
        if (st->primal_dual_gap <= p->phase2Tol * 5) { bad_pd -= 5; if (bad_pd < 0) bad_pd = 0; }
        else if (st->primal_dual_gap <= p->phase2Tol) { bad_pd -= 10; if (bad_pd < 0) bad_pd = 0; }
        if (st->primal_dual_gap >= p->phase1Tol * 1e2) bad_pd += 2;
        if (bad_pd >= bad_pd_cap) {
            if (reopt) lh_log(S, "------\n");
            return LH_RET_OK;
        }
        pinf_buf[count % 10] = st->l_inf_primal_infeasibility;
        if (!reopt) {
            if (st->l_inf_primal_infeasibility <= p->phase2Tol) {
                if (dimacs_admm(S)) return LH_RET_DEVICE;
                take_errors(S, st, 0);
                return LH_RET_OK;
            }
        } else {
            if (st->l_1_primal_infeasibility <= p->phase2Tol) {
                if (dimacs_admm(S)) return LH_RET_DEVICE;
                take_errors(S, st, 0);
                if (st->primal_dual_gap <= p->phase2Tol) return LH_RET_OK;
            }
        }
        GPU_TRY(S, lgpu_update_dual_var(S->gpu, st->rho));
        const int64_t phase = reopt ? st->iter : st->iter + 1;
        if (phase % p->rhoFreq == 0) {
            st->rho *= p->rhoFactor;
            if (st->rho >= cur_rho_max) {
                st->rho = cur_rho_max;
                if (phase % (p->rhoFreq * 100) == 0) {
                    const double mean = buffer_mean(pinf_buf, 10);
                    if (mean / old_mean >= 0.65) {
                        st->rho *= pow(p->rhoFactor, round(log((double)(p->rhoFreq * 100)) / log((double)p->rhoFreq)));
                        cur_rho_max = st->rho;
                    }
                    old_mean = mean;
                }
            }
            if (st->rho >= p->rhoCellingADMM) st->rho = p->rhoCellingADMM;
        }
        if (st->iter % 50 == 0) {
            if (dimacs_admm(S)) return LH_RET_DEVICE;
            take_errors(S, st, 0);
            if (lh_time_is_up(S, p, timeSolveStart, 0)) return LH_RET_TIME_OUT;
        }
        if (st->primal_dual_gap <= p->phase2Tol * 1e-3 && st->l_1_primal_infeasibility <= p->phase2Tol * 1e-3) {
            lh_log(S, "Early Stop When DIMACS Errors Are Well-Satisfied\n");
            return LH_RET_OK;
        }
        st->iter++;
    }
    return LH_RET_OK;
}

int lh_admm_optimize(lh_params *p, lh_solver *S, lh_admm_state *st, int64_t iter_celling, double timeSolveStart)
{
    return admm_loop(p, S, st, iter_celling, timeSolveStart, 0);
}

int lh_admm_optimize_reopt(lh_params *p, lh_solver *S, lh_admm_state *st, int64_t iter_celling, double timeSolveStart)
{
    return admm_loop(p, S, st, iter_celling, timeSolveStart, 1);
}
