/*
 * alm.c -- phase 1: augmented-Lagrangian outer loop, L-BFGS inner loop, exact quartic line search.
 *
 * Control flow and every threshold restate the reference's LORADS_ALMOptimize / LORADS_ALMOptimize_reopt
 * (lorads/src/src_semi/lorads_alg/lorads_alm.c:1220-1484 and :959-1201) so that iteration counts track the
 * reference; the array work of each step is one C-ABI call (include/lorads_b200.h).  The line-search root
 * selection (Shengjin's cubic formulas + the reference's 1e-10 cascade, :191-333) stays on the host in the
 * reference's operation order because it compares against absolute thresholds and exact zeros.
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

enum { DIFF_EASY = 0, DIFF_MEDIUM, DIFF_HARD, DIFF_SUPER };

static double real_root(double base, int n)
{
    if (base < 0 && n % 2 == 0) return NAN;
    return base > 0 ? pow(base, 1.0 / n) : -pow(-base, 1.0 / n);
}

/* real roots of a x^3 + b x^2 + c x + d by Shengjin's discriminants; returns how many entries of res are set */
int lh_cubic_equation(double a, double b, double c, double d, double *res)
{
    const double A = b * b - 3 * a * c;
    const double B = b * c - 9 * a * d;
    const double C = c * c - 3 * b * d;
    const double delta = B * B - 4 * A * C;
    res[0] = res[1] = res[2] = 0.0;
    if (A == 0 && B == 0) {
        const double x = -c / b;
        res[0] = res[0] > x ? res[0] : x;
        return 1;
    }
    if (delta > 0) {
        const double Y1 = A * b + 1.5 * a * (-B + sqrt(delta));
        const double Y2 = A * b + 1.5 * a * (-B - sqrt(delta));
        const double x = (-b - real_root(Y1, 3) - real_root(Y2, 3)) / 3 / a;
        res[0] = res[0] > x ? res[0] : x;
        return 1;
    }
    if (delta == 0 && A != 0 && B != 0) {
        const double K = B / A;
        res[0] = -b / a + K;
        res[1] = -K / 2;
        return 2;
    }
    if (delta < 0) {
        const double sqA = sqrt(A);
        const double T = (A * b - 1.5 * a * B) / (A * sqA);
        const double theta = acos(T);
        const double cs = cos(theta / 3);
        const double sn = sqrt(3) * sin(theta / 3);
        res[0] = (-b - 2 * sqA * cs) / 3 / a;
        res[1] = (-b + sqA * (cs + sn)) / 3 / a;
        res[2] = (-b + sqA * (cs - sn)) / 3 / a;
        return 3;
    }
    return 0;
}

static double quartic(double a, double b, double c, double d, double x)
{
    return a * pow(x, 4) + b * pow(x, 3) + c * pow(x, 2) + d * x;
}

/* terms = { p1, p2, |q2|^2, q1.q2, q0'.q2, |q1|^2, q0'.q1 } from lgpu_alm_linesearch_terms.
 * f(tau) = a tau^4 + b tau^3 + c tau^2 + d tau on [0,1]; candidates 0, 1 and the stationary points. */
int lh_line_search(double rho, const double t[7], double *tau)
{
    const double a = rho * t[2] / 2;
    const double b = rho * t[3];
    const double c = t[1] - rho * t[4] + rho * t[5] / 2;
    const double d = t[0] - rho * t[6];
    double roots[3] = {0.0, 0.0, 0.0};
    const int nroot = lh_cubic_equation(4 * a, 3 * b, 2 * c, d, roots);
    const double f0 = 0.0, f1 = quartic(a, b, c, d, 1.0);
    double fr[3] = {1e+30, 1e+30, 1e+30};
    if (nroot >= 1 && roots[0] > 1e-20 && roots[0] <= 1.0) fr[0] = quartic(a, b, c, d, roots[0]);
    if (nroot >= 2 && roots[1] > 1e-20 && roots[1] <= 1.0) fr[1] = quartic(a, b, c, d, roots[1]);
    if (nroot == 3 && roots[2] > 1e-20 && roots[2] <= 1.0) fr[2] = quartic(a, b, c, d, roots[2]);
    double mn = f0 < f1 ? f0 : f1;
    for (int k = 0; k < 3; ++k) mn = mn < fr[k] ? mn : fr[k];
    /* later candidates win ties within 1e-10, exactly as the reference's cascade of ifs */
    if (fabs(mn - f0) < 1e-10) tau[0] = 0.0;
    if (fabs(mn - f1) < 1e-10) tau[0] = 1.0;
    for (int k = 0; k < 3; ++k)
        if (fabs(mn - fr[k]) < 1e-10) tau[0] = roots[k];
    return nroot;
}

/* stall detector on an exponential moving average of the rho-certificate (LUtilUpdateCheckEma,
 * lorads_utils.c:564-594) */
static int ema_check(double *cur, double *old, double v, double alpha, double thr, int64_t interval, int64_t *counter)
{
    int ok = 1;
    *cur = alpha * v + (1 - alpha) * (*cur);
    if (*counter >= interval) {
        if (*old != 0) {
            const double change = (*cur - *old) / *old;
            ok = (change >= -thr) && (change <= thr);
        }
        *old = *cur;
        *counter = 1;
    } else {
        (*counter)++;
    }
    return ok;
}

static void alm_print(lh_solver *S, const lh_alm_state *st, double t, int64_t cur_rank, int64_t oracle_rank)
{
    lh_log(S,
           "ALM OuterIter:%d InnerIter:%d pObj:%5.5e dObj:%5.5e pInfea(1):%5.5e pInfea(Inf):%5.5e pdGap:%5.5e rho:%3.2f "
           "CurrRank:%lld OracleRank:%lld Time:%3.2f\n",
           (int)st->outerIter, (int)st->innerIter, st->primal_objective_value, st->dual_objective_value,
           st->l_1_primal_infeasibility, st->l_inf_primal_infeasibility, st->primal_dual_gap, st->rho,
           (long long)cur_rank, (long long)oracle_rank, t);
}

static void alm_record(lh_solver *S, const lh_alm_state *st, double phase_time)
{
    const int64_t cur = lh_sum_rank(S);
    const int64_t orc = lh_oracle_rank(S, 1);
    lh_append_trajectory(S, 1, cur, orc);
    alm_print(S, st, phase_time, cur, orc);
}

/* ---- small wrappers: one reference function each --------------------------------------------------------*/
static int cal_grad_cert(lh_solver *S, double rho, double *cert_val)
{
    double lag = 0.0;
    GPU_TRY(S, lgpu_alm_cal_grad(S->gpu, rho, &lag));
    *cert_val = sqrt(lag) / (1 + S->cObjNrmInf);
    return 0;
}
static int cal_objs(lh_solver *S)
{
    double p = 0, d = 0;
    GPU_TRY(S, lgpu_cal_obj(S->gpu, 0, &p));
    GPU_TRY(S, lgpu_cal_dual_obj(S->gpu, &d));
    S->pObjVal = p / S->scaleObjHis;
    S->dObjVal = d / S->scaleObjHis;
    return 0;
}
static int dimacs_alm(lh_solver *S)
{
    double l1 = 0;
    GPU_TRY(S, lgpu_primal_infeasibility(S->gpu, LGPU_PAIR_RR, &l1));
    S->dimacConstrVio = l1;
    const double gap = S->pObjVal - S->dObjVal;
    S->dimacGap = fabs(gap) / (1 + fabs(S->pObjVal) + fabs(S->dObjVal));
    return 0;
}
static double linf_from_l1(const lh_solver *S, double l1) { return l1 * (1 + S->bRHSNrm1) / (1 + S->bRHSNrmInf); }

static double rank_threshold(const lh_params *p, const lh_solver *S)
{
    double thr = 15;
    if (p->dyrankLevel == 0) thr = 1e8;
    else if (p->dyrankLevel == 1) thr = 150;
    else if (p->dyrankLevel == 2) thr = 15;
    else if (p->dyrankLevel == 3) thr = 5;
    /* --nearStallFactor scales the stall threshold when a rank schedule drives the growth (no reference semantics) */
    if (S->scheduleLen > 0 && p->nearStallFactor > 0) thr = ceil(thr * p->nearStallFactor);
    return thr;
}

/* one L-BFGS + line-search step.  Returns 0 ok, 1 rootNum == 0, 2 tau below endTauTol, <0 device error */
static int inner_step(lh_params *p, lh_solver *S, lh_alm_state *st, int64_t clearLBFGS, double *tau_out, double *lag_out)
{
    double terms[7], tau = st->tau;
    if (lgpu_lbfgs_direction(S->gpu, clearLBFGS) != 0) return -1;
    if (lgpu_alm_linesearch_terms(S->gpu, st->rho, terms) != 0) return -1;
    const int nroot = lh_line_search(st->rho, terms, &tau);
    st->tau = tau;
    *tau_out = tau;
    if (nroot == 0) return 1;
    if (fabs(tau) < p->endTauTol) return 2;
    /* step, gradient, L-BFGS pair and the fresh A(RR^T) in one device call; the infeasibility it returns is what
     * updateDimacsALM would compute next (lorads_alm.c:1357) */
    double pinf = 0.0;
    if (lgpu_alm_inner_update(S->gpu, st->rho, tau, lag_out, &pinf) != 0) return -1;
    S->dimacConstrVio = pinf;
    S->dimacGap = fabs(S->pObjVal - S->dObjVal) / (1 + fabs(S->pObjVal) + fabs(S->dObjVal));
    return 0;
}

#define DEV_FAIL(S)                                                                        \
    do {                                                                                   \
        fprintf(stderr, "lorads_b200: device error: %s\n", lgpu_last_error((S)->gpu));     \
        return LH_RET_DEVICE;                                                              \
    } while (0)

int lh_alm_optimize(lh_params *p, lh_solver *S, lh_alm_state *st, double timeSolveStart)
{
    S->maxAlmSubIter = 5000;
    const double t_begin = lh_time();
    int is_rank_max = lh_all_rank_max(S, 1.0);
    int retcode = LH_RET_OK;
    int64_t last_start = 1;
    const double cert = 0.1;
    double cert_tol, cert_val, lag = 0.0, tau = 0.0;
    int restart;
    do {
        restart = 0;
        cert_tol = cert / st->rho;
        GPU_TRY(S, lgpu_init_constr_val(S->gpu, LGPU_PAIR_RR));
        if (cal_grad_cert(S, st->rho, &cert_val)) return LH_RET_DEVICE;
        int difficulty = DIFF_HARD;
        int64_t localIter = 0, clearLBFGS = 0, rank_flag = 0;
        const double rank_update_factor = p->rankUpdateFactor;
        double rho_update_factor = p->ALMRhoFactor;
        int rho_factor_flag = 0;
        const double rank_flag_thres = rank_threshold(p, S);
        const int sub_inc = 10000, sub_ceil = 25000;
        int sub_counter = 0;
        int goto_end = 0, goto_print = 0;
        for (int64_t k = st->outerIter; k <= p->maxALMIter && !goto_end && !goto_print && !restart; k++) {
            double ema_cur = 0.0, ema_old = 0.0;
            int64_t ema_counter = 1, cur_iter_counter = 1;
            int tiny_tau = 0;
            if (sub_counter >= 2) {
                sub_counter = 0;
                S->maxAlmSubIter += sub_inc;
                if (S->maxAlmSubIter > sub_ceil) S->maxAlmSubIter = sub_ceil;
            }
            while (difficulty != DIFF_EASY) {
                localIter = 0;
                const int steady = ema_check(&ema_cur, &ema_old, cert_val, 0.1, 0.005, 5, &ema_counter);
                if (!steady && !p->highAccMode) break;
                if (cur_iter_counter >= S->maxAlmSubIter) { sub_counter += 1; break; }
                if ((double)rank_flag >= rank_flag_thres && !is_rank_max && (k - last_start >= 3)) break;
                if (cert_val <= cert_tol) break;
                while (cert_val - cert_tol > p->endALMSubTol) {
                    if (localIter % 300 == 0) clearLBFGS = 0;
                    const int rc = inner_step(p, S, st, clearLBFGS, &tau, &lag);
                    if (rc < 0) DEV_FAIL(S);
                    if (rc == 1) { retcode = LH_RET_NUM_ERR; goto_end = 1; break; }
                    if (rc == 2) {
                        printf("update rho:%5.8e since tau is too small.\n", tau);
                        st->innerIter++; localIter++; cur_iter_counter++; clearLBFGS++;
                        tiny_tau = 1;
                        break;
                    }
                    st->l_1_primal_infeasibility = S->dimacConstrVio; /* set by inner_step */
                    st->l_inf_primal_infeasibility = linf_from_l1(S, st->l_1_primal_infeasibility);
                    if (st->l_inf_primal_infeasibility <= p->phase1Tol && (st->primal_dual_gap <= p->phase1Tol || !p->highAccMode)) {
                        st->outerIter = k;
                        st->innerIter += 1; localIter += 1; cur_iter_counter += 1; clearLBFGS += 1;
                        goto_end = 1;
                        break;
                    }
                    cert_val = sqrt(lag) / (1 + S->cObjNrmInf);
                    st->innerIter += 1; localIter++; cur_iter_counter++; clearLBFGS++;
                    if (localIter > 800) break;
                }
                if (goto_end || tiny_tau) break;
                GPU_TRY(S, lgpu_update_dual_var(S->gpu, st->rho));
                if (cal_grad_cert(S, st->rho, &cert_val)) return LH_RET_DEVICE;
                if (localIter <= 20) difficulty = DIFF_EASY;
                else if (localIter <= 100) { difficulty = DIFF_MEDIUM; rank_flag += 2; }
                else if (localIter < 400) { difficulty = DIFF_HARD; rank_flag += 3; }
                else { difficulty = DIFF_SUPER; rank_flag += 4; }
                if (difficulty == DIFF_EASY) rank_flag = 0;
            }
            if (goto_end) break;
            /* penalty update: raise rho until the certificate tolerance drops below the certificate */
            do {
                st->rho *= rho_update_factor;
                if (cal_grad_cert(S, st->rho, &cert_val)) return LH_RET_DEVICE;
                cert_tol = cert / st->rho;
            } while (cert_tol >= cert_val);
            if (st->rho >= 5e4 && rho_factor_flag < 4) { rho_update_factor = sqrt(sqrt(rho_update_factor)); rho_factor_flag = 4; }
            else if (st->rho >= 5e6 && rho_factor_flag < 6) { rho_update_factor = sqrt(sqrt(rho_update_factor)); rho_factor_flag = 6; }
            else if (st->rho >= 5e8 && rho_factor_flag < 8) { rho_update_factor = sqrt(sqrt(rho_update_factor)); rho_factor_flag = 8; }
            difficulty = DIFF_HARD;
            clearLBFGS = 0;
            st->outerIter = k;
            if (st->l_inf_primal_infeasibility <= p->phase1Tol && (st->primal_dual_gap <= p->phase1Tol || !p->highAccMode)) {
                goto_end = 1;
                break;
            }
            if (cal_objs(S) || dimacs_alm(S)) return LH_RET_DEVICE;
            st->primal_dual_gap = S->dimacGap;
            st->primal_objective_value = S->pObjVal;
            st->dual_objective_value = S->dObjVal;
            st->l_1_primal_infeasibility = S->dimacConstrVio;
            st->l_inf_primal_infeasibility = linf_from_l1(S, st->l_1_primal_infeasibility);
            st->l_1_dual_infeasibility = 99;
            st->l_inf_dual_infeasibility = 99;
            if (st->primal_dual_gap <= p->phase1Tol * 1e-3 && st->l_1_primal_infeasibility <= p->phase1Tol * 1e-3) {
                goto_print = 1;
                break;
            }
            alm_record(S, st, lh_time() - t_begin);
            if (lh_time_is_up(S, p, timeSolveStart, 0)) { goto_print = 1; break; }
            if ((double)rank_flag >= rank_flag_thres && !is_rank_max) {
                rank_flag = 0;
                if (k - last_start >= 2) {
                    printf("increase the rank, factor:%f.\n", rank_update_factor);
                    const int r = lh_aug_rank(S, rank_update_factor, p);
                    if (r < 0) return LH_RET_DEVICE;
                    is_rank_max = r;
                    st->outerIter = k;
                    last_start = st->outerIter;
                    restart = 1;
                }
            }
        }
        if (restart) continue;
        if (!goto_print) {
            if (cal_objs(S) || dimacs_alm(S)) return LH_RET_DEVICE;
            st->primal_objective_value = S->pObjVal;
            st->dual_objective_value = S->dObjVal;
            st->primal_dual_gap = S->dimacGap;
            st->l_1_primal_infeasibility = S->dimacConstrVio;
            st->l_inf_primal_infeasibility = linf_from_l1(S, st->l_1_primal_infeasibility);
            st->l_1_dual_infeasibility = 99;
            st->l_inf_dual_infeasibility = 99;
        }
    } while (restart);
    lh_log(S, "-----------------------------------------------------------------------\n");
    lh_log(S, "Exit ALM:\n");
    alm_record(S, st, lh_time() - t_begin);
    lh_log(S, "-----------------------------------------------------------------------\n");
    return retcode;
}

/* re-optimisation variant (lorads_alm.c:959-1201): open-ended outer loop with its own exit tests, L-BFGS reset
 * on (localIter-1) % 300, k advanced after every penalty update, rank growth only for <= 10 cones */
int lh_alm_optimize_reopt(lh_params *p, lh_solver *S, lh_alm_state *st, int early_stop, double rho_update_factor,
                          double timeSolveStart)
{
    const double t_begin = lh_time();
    int is_rank_max = lh_all_rank_max(S, 1.0);
    int retcode = LH_RET_OK;
    int64_t last_start = 1;
    const double cert = 0.1;
    double cert_tol, cert_val, lag = 0.0, tau = 0.0;
    int restart;
    do {
        restart = 0;
        cert_tol = cert / st->rho;
        GPU_TRY(S, lgpu_init_constr_val(S->gpu, LGPU_PAIR_RR));
        if (cal_grad_cert(S, st->rho, &cert_val)) return LH_RET_DEVICE;
        int difficulty = DIFF_HARD;
        int64_t localIter = 0, clearLBFGS = 0, rank_flag = 0;
        const double rank_update_factor = p->rankUpdateFactor;
        int64_t k = st->outerIter;
        const int64_t k0 = st->outerIter;
        int rho_factor_flag = 0;
        const double rank_flag_thres = rank_threshold(p, S);
        const int sub_inc = 10000, sub_ceil = 25000;
        int sub_counter = 0;
        int goto_end = 0, goto_print = 0;
        while (!goto_end && !goto_print && !restart) {
            if (k > p->maxALMIter && (st->l_inf_primal_infeasibility <= p->phase1Tol &&
                                      ((st->primal_dual_gap <= (p->phase1Tol > p->phase2Tol * 5 ? p->phase1Tol : p->phase2Tol * 5)) || !p->highAccMode)))
                break;
            double ema_cur = 0.0, ema_old = 0.0;
            int64_t ema_counter = 1, cur_iter_counter = 1;
            int tiny_tau = 0;
            if (sub_counter >= 2) {
                sub_counter = 0;
                S->maxAlmSubIter += sub_inc;
                if (S->maxAlmSubIter > sub_ceil) S->maxAlmSubIter = sub_ceil;
            }
            while (difficulty != DIFF_EASY) {
                localIter = 0;
                const int steady = ema_check(&ema_cur, &ema_old, cert_val, 0.1, 0.005, 5, &ema_counter);
                if (!steady && !p->highAccMode) break;
                if (cur_iter_counter >= S->maxAlmSubIter) { sub_counter += 1; break; }
                if ((double)rank_flag >= rank_flag_thres && !is_rank_max && (k - last_start >= 3)) break;
                if (cert_val <= cert_tol) break;
                while (cert_val - cert_tol > p->endALMSubTol) {
                    if ((localIter - 1) % 300 == 0) clearLBFGS = 0;
                    const int rc = inner_step(p, S, st, clearLBFGS, &tau, &lag);
                    if (rc < 0) DEV_FAIL(S);
                    if (rc == 1) { retcode = LH_RET_NUM_ERR; goto_end = 1; break; }
                    if (rc == 2) {
                        printf("update rho, tau is too small :%5.3e\n", tau);
                        st->innerIter++; localIter++; cur_iter_counter++; clearLBFGS++;
                        tiny_tau = 1;
                        break;
                    }
                    cert_val = sqrt(lag) / (1 + S->cObjNrmInf);
                    st->l_1_primal_infeasibility = S->dimacConstrVio; /* set by inner_step */
                    st->l_inf_primal_infeasibility = linf_from_l1(S, st->l_1_primal_infeasibility);
                    st->innerIter++; localIter++; cur_iter_counter++; clearLBFGS++;
                    if (localIter > 800) break;
                }
                if (goto_end || tiny_tau) break;
                GPU_TRY(S, lgpu_update_dual_var(S->gpu, st->rho));
                if (cal_grad_cert(S, st->rho, &cert_val)) return LH_RET_DEVICE;
                if (localIter <= 20) difficulty = DIFF_EASY;
                else if (localIter <= 100) { difficulty = DIFF_MEDIUM; rank_flag += 2; }
                else { difficulty = DIFF_HARD; rank_flag += 3; } /* the SUPER branch is unreachable here (quirk Q5) */
                if (difficulty == DIFF_EASY) rank_flag = 0;
            }
            if (goto_end) break;
            do {
                st->rho *= rho_update_factor;
                if (cal_grad_cert(S, st->rho, &cert_val)) return LH_RET_DEVICE;
                cert_tol = cert / st->rho;
            } while (cert_tol >= cert_val);
            if (st->rho >= 5e4 && rho_factor_flag < 4) { rho_update_factor = sqrt(sqrt(rho_update_factor)); rho_factor_flag = 4; }
            else if (st->rho >= 5e6 && rho_factor_flag < 6) { rho_update_factor = sqrt(sqrt(rho_update_factor)); rho_factor_flag = 6; }
            else if (st->rho >= 5e8 && rho_factor_flag < 8) { rho_update_factor = sqrt(sqrt(rho_update_factor)); rho_factor_flag = 8; }
            difficulty = DIFF_HARD;
            clearLBFGS = 0;
            k += 1;
            st->outerIter = k;
            if (cal_objs(S) || dimacs_alm(S)) return LH_RET_DEVICE;
            st->primal_dual_gap = S->dimacGap;
            st->primal_objective_value = S->pObjVal;
            st->dual_objective_value = S->dObjVal;
            st->l_1_primal_infeasibility = S->dimacConstrVio;
            st->l_inf_primal_infeasibility = linf_from_l1(S, st->l_1_primal_infeasibility);
            st->l_1_dual_infeasibility = 99;
            st->l_inf_dual_infeasibility = 99;
            if (early_stop) {
                const double gtol = p->phase1Tol > p->phase2Tol * 5 ? p->phase1Tol : p->phase2Tol * 5;
                if (st->l_1_primal_infeasibility <= p->phase1Tol && st->primal_dual_gap <= gtol && (k - k0) > 1) { goto_print = 1; break; }
            } else {
                if (st->primal_dual_gap <= p->phase2Tol && st->l_1_primal_infeasibility <= p->phase2Tol && (k - k0) > 1) { goto_print = 1; break; }
            }
            alm_record(S, st, lh_time() - t_begin);
            if (lh_time_is_up(S, p, timeSolveStart, 0)) { goto_print = 1; break; }
            if ((double)rank_flag >= rank_flag_thres && !is_rank_max && S->nCones <= 10) {
                rank_flag = 0;
                if (k - last_start >= 2) {
                    printf("increase the rank, factor:%f.\n", rank_update_factor);
                    const int r = lh_aug_rank(S, rank_update_factor, p);
                    if (r < 0) return LH_RET_DEVICE;
                    is_rank_max = r;
                    st->outerIter = k;
                    last_start = st->outerIter;
                    restart = 1;
                }
            }
        }
        if (restart) continue;
        if (!goto_print) {
            if (cal_objs(S) || dimacs_alm(S)) return LH_RET_DEVICE;
            st->primal_dual_gap = S->dimacGap;
            /* the reference back-computes l_1 from the stale l_inf here (lorads_alm.c:1191) */
            st->l_1_primal_infeasibility = st->l_inf_primal_infeasibility * (1 + S->bRHSNrmInf) / (1 + S->bRHSNrm1);
            st->l_1_dual_infeasibility = 99;
            st->l_inf_dual_infeasibility = 99;
        }
    } while (restart);
    lh_log(S, "-----------------------------------------------------------------------\n");
    lh_log(S, "Exit ALM:\n");
    alm_record(S, st, lh_time() - t_begin);
    lh_log(S, "-----------------------------------------------------------------------\n");
    return retcode;
}
