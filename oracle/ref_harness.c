/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin flat-array door into the UNMODIFIED reference LoRADS objects.  It is compiled together
 * with the reference sources (where they lie under /root/reference, see oracle/Makefile) into
 * oracle/_ref/liblorads_ref.so and is driven through ctypes by oracle/make_golden.py to
 *   (1) validate the numpy restatement in oracle/lorads_oracle.py, and
 *   (2) generate the golden vectors committed under tests/golden/.
 * Every rh_* function only sequences calls of the reference's OWN functions (named in the
 * comments) exactly as main.c / lorads_alm.c / lorads_admm.c do, and copies arrays in and out.
 * No reference source is copied here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "lorads_file_io.h"
#include "def_lorads_user_data.h"
#include "lorads_user_data.h"
#include "lorads_utils.h"
#include "def_lorads_solver.h"
#include "lorads_solver.h"
#include "lorads_alm.h"
#include "lorads_admm.h"
#include "lorads_alg_common.h"
#include "lorads_logging.h"
#include "lorads_cgs.h"
#include "lorads_sdp_conic.h"

static lorads_solver *S = NULL;
static lorads_params P;
static lorads_alm_state alm_state;
static lorads_admm_state admm_state;
static SDPConst sdpConst;
static lorads_func *F = NULL;
static lorads_int *blkDims = NULL;

extern void initCommandLineArgs(lorads_params *params); /* not in the lib (main.c); defaults set below */

/* main.c:380-418 : read, init solver, cone data, preprocess, rank, ALM/ADMM vars, constants */
int rh_load(const char *fname, double timesLogRank, int fixedRank)
{
    lorads_int nConstrs = 0, nBlks = 0, nCols = 0, nLpCols = 0, nElem = 0;
    double *rowRHS = NULL;
    lorads_int **coneMatBeg = NULL, **coneMatIdx = NULL;
    double **coneMatElem = NULL;
    lorads_int *LpMatBeg = NULL, *LpMatIdx = NULL;
    double *LpMatElem = NULL;
    user_data **SDPDatas = NULL;
    memset(&P, 0, sizeof(P));
    P.fname = (char *)fname;
    P.lbfgsListLength = 2;
    P.initRho = 0.0;
    P.rhoMax = 5000.0;
    P.timesLogRank = timesLogRank;
    P.fixedRank = fixedRank;
    P.initRank = -1;
    P.oracleRankMethod = LORADS_ORACLE_RANK_GRAM;
    if (LReadSDPA((char *)fname, &nConstrs, &nBlks, &blkDims, &rowRHS, &coneMatBeg, &coneMatIdx, &coneMatElem,
                  &nCols, &nLpCols, &LpMatBeg, &LpMatIdx, &LpMatElem, &nElem) != LORADS_RETCODE_OK)
        return 1;
    LORADS_INIT(S, lorads_solver, 1);
    LORADS_INIT(S->var, lorads_variable, 1);
    LORADSInitSolver(S, nConstrs, nBlks, blkDims, nLpCols);
    LORADS_INIT(SDPDatas, user_data *, nBlks);
    LORADSSetDualObjective(S, rowRHS);
    LORADSInitConeData(S, SDPDatas, coneMatElem, coneMatBeg, coneMatIdx, blkDims, nConstrs, nBlks, nLpCols,
                       LpMatBeg, LpMatIdx, LpMatElem);
    LORADSPreprocess(S, blkDims);
    LORADSDetermineRank(S, blkDims, P.timesLogRank, P.fixedRank, P.initRank);
    LORADSInitALMVars(S, S->var->rankElem, blkDims, nBlks, nLpCols, P.lbfgsListLength);
    S->hisRecT = P.lbfgsListLength;
    LORADSInitADMMVars(S, S->var->rankElem, blkDims, nBlks, nLpCols);
    initial_solver_state(&P, S, &alm_state, &admm_state, &sdpConst);
    LORADSInitFuncSet(&F, S->nLpCols);
    return 0;
}

int rh_m(void) { return (int)S->nRows; }
int rh_ncones(void) { return (int)S->nCones; }
int rh_nlp(void) { return (int)S->nLpCols; }
int rh_dim(int c) { return (int)S->var->R[c]->nRows; }
int rh_rank(int c) { return (int)S->var->R[c]->rank; }
int rh_rank_max(int c) { return (int)S->rank_max[c]; }
double rh_rho0(void) { return alm_state.rho; }
int rh_cone_is_sparse_container(int c) { return S->SDPCones[c]->type == LORADS_CONETYPE_SPARSE_SDP; }
int rh_cone_aggregate_is_dense(int c) { return S->SDPCones[c]->sdp_obj_sum->dataType == SDP_COEFF_DENSE; }
void rh_constants(double *o)
{
    o[0] = S->cObjNrm1; o[1] = S->cObjNrm2; o[2] = S->cObjNrmInf;
    o[3] = S->bRHSNrm1; o[4] = S->bRHSNrm2; o[5] = S->bRHSNrmInf;
}
void rh_get_b(double *o) { memcpy(o, S->rowRHS, sizeof(double) * S->nRows); }

static lorads_sdp_dense *pick(int which, int c)
{
    switch (which) {
    case 0: return S->var->R[c];
    case 1: return S->var->U[c];
    case 2: return S->var->V[c];
    case 3: return S->var->Grad[c];
    default: return NULL;
    }
}
void rh_get_factor(int which, int c, double *o)
{
    lorads_sdp_dense *m = pick(which, c);
    memcpy(o, m->matElem, sizeof(double) * m->nRows * m->rank);
}
void rh_set_factor(int which, int c, const double *in)
{
    lorads_sdp_dense *m = pick(which, c);
    memcpy(m->matElem, in, sizeof(double) * m->nRows * m->rank);
}
void rh_get_vec(int which, double *o)
{
    double *src = which == 0 ? S->var->dualVar : which == 1 ? S->var->constrValSum : which == 2 ? S->var->ARDSum
                : which == 3 ? S->var->ADDSum : S->var->M1temp;
    memcpy(o, src, sizeof(double) * S->nRows);
}
void rh_set_vec(int which, const double *in)
{
    double *dst = which == 0 ? S->var->dualVar : which == 1 ? S->var->constrValSum : which == 2 ? S->var->ARDSum
                : which == 3 ? S->var->ADDSum : S->var->M1temp;
    memcpy(dst, in, sizeof(double) * S->nRows);
}
void rh_get_lp(int which, double *o)
{
    lorads_lp_dense *v = which == 0 ? S->var->rLp : which == 1 ? S->var->uLp : which == 2 ? S->var->vLp : S->var->gradLp;
    memcpy(o, v->matElem, sizeof(double) * S->nLpCols);
}

/* per-cone constrVal expanded to a dense m-vector */
static void expand_constr_val(int c, double *o)
{
    memset(o, 0, sizeof(double) * S->nRows);
    double one = 1.0;
    S->var->constrVal[c]->add(&one, S->var->constrVal[c]->data, o);
}

/* LORADSObjConstrValAll (lorads_alg_common.c:169-176): which pair 0:(R,R) 1:(R,U) 2:(U,U) 3:(U,V)
 * out_cv: nCones x m (expanded per-cone A(UV^T)); returns <C, UV^T> accumulated over cones (unscaled) */
double rh_obj_constr_val_all(int pair, double *out_cv)
{
    lorads_sdp_dense **A = (pair == 0 || pair == 1) ? S->var->R : S->var->U;
    lorads_sdp_dense **B = pair == 0 ? S->var->R : pair == 3 ? S->var->V : S->var->U;
    double obj = 0.0;
    LORADSObjConstrValAll(S, A, B, &obj);
    for (int c = 0; c < S->nCones; ++c) expand_constr_val(c, out_cv + (size_t)c * S->nRows);
    return obj;
}

/* InitConstrValAll + InitConstrValSum on (R,R) (lorads_alg_common.c:116-122,221-229) */
void rh_init_constr_val_sum_RR(void)
{
    F->InitConstrValAll(S, S->var->rLp, S->var->rLp, S->var->R, S->var->R);
    F->InitConstrValSum(S);
}

/* ALMCalGrad (lorads_alm.c:32-87) */
double rh_alm_cal_grad(double rho)
{
    double lag = 0.0;
    F->ALMCalGrad(S, S->var->rLp, S->var->gradLp, S->var->R, S->var->Grad, &lag, rho);
    return lag;
}

/* sdp_obj_sum values after zeros + addObjCoeff + sdpDataWSum(w) for cone c (pattern order / packed) */
int rh_wsum(int c, const double *w, int addObj, double *out, int cap)
{
    lorads_sdp_cone *ACone = S->SDPCones[c];
    ACone->sdp_obj_sum->zeros(ACone->sdp_obj_sum->dataMat);
    if (addObj) ACone->addObjCoeff(ACone->coneData, ACone->sdp_obj_sum);
    ACone->sdpDataWSum(ACone->coneData, (double *)w, ACone->sdp_obj_sum);
    if (ACone->sdp_obj_sum->dataType == SDP_COEFF_SPARSE) {
        sdp_coeff_sparse *sp = (sdp_coeff_sparse *)ACone->sdp_obj_sum->dataMat;
        int nn = (int)sp->nTriMatElem;
        if (out) for (int i = 0; i < nn && i < cap; ++i) out[i] = sp->triMatElem[i];
        return nn;
    } else {
        sdp_coeff_dense *ds = (sdp_coeff_dense *)ACone->sdp_obj_sum->dataMat;
        int nn = (int)(ds->nSDPCol * (ds->nSDPCol + 1) / 2);
        if (out) for (int i = 0; i < nn && i < cap; ++i) out[i] = ds->dsMatElem[i];
        return nn;
    }
}
/* pattern (row, col) of the sparse aggregate of cone c; returns nnzP (or -1 if dense aggregate) */
int rh_pattern(int c, int *row, int *col, int cap)
{
    lorads_sdp_cone *ACone = S->SDPCones[c];
    if (ACone->sdp_obj_sum->dataType != SDP_COEFF_SPARSE) return -1;
    sdp_coeff_sparse *sp = (sdp_coeff_sparse *)ACone->sdp_obj_sum->dataMat;
    int nn = (int)sp->nTriMatElem;
    if (row) for (int i = 0; i < nn && i < cap; ++i) { row[i] = (int)sp->triMatRow[i]; col[i] = (int)sp->triMatCol[i]; }
    return nn;
}
/* mul_rk of the aggregate currently held in sdp_obj_sum with factor `which` of cone c */
void rh_mul_rk(int c, int which, double *out)
{
    lorads_sdp_cone *ACone = S->SDPCones[c];
    ACone->sdp_obj_sum->mul_rk(ACone->sdp_obj_sum->dataMat, pick(which, c), out);
}

/* One ALM inner iteration, sequenced exactly as lorads_alm.c:1302-1378 does.
 * out_scalars: [0]=rootNum [1]=tau [2]=p1 [3]=p2 [4]=lagNormSquare [5]=pInf_l1 */
int rh_alm_inner_iter(double rho, int clearLBFGS, double *out_scalars)
{
    lorads_int incx = 1;
    double minusOne = -1.0, tau = 0.0, lag = 0.0;
    F->LBFGSDirection(&P, S, S->lbfgsHis, S->var->gradLp, S->var->uLp, S->var->Grad, S->var->U, clearLBFGS);
    F->LBFGSDirUseGrad(S, S->var->uLp, S->var->gradLp, S->var->U, S->var->Grad);
    double *q0 = S->var->M1temp;
    LORADS_MEMCPY(q0, S->rowRHS, double, S->nRows);
    axpy(&(S->nRows), &minusOne, S->var->constrValSum, &incx, q0, &incx);
    double p12[2];
    F->ALMCalq12p12(S, S->var->rLp, S->var->uLp, S->var->R, S->var->U, S->var->ARDSum, S->var->ADDSum, p12);
    lorads_int rootNum = ALMLineSearch(rho, S->nRows, S->var->dualVar, p12[0], p12[1], q0, S->var->ARDSum,
                                       S->var->ADDSum, &tau);
    out_scalars[0] = (double)rootNum; out_scalars[1] = tau; out_scalars[2] = p12[0]; out_scalars[3] = p12[1];
    F->setAsNegGrad(S, S->var->gradLp, S->var->Grad);
    F->ALMupdateVar(S, S->var->rLp, S->var->uLp, S->var->R, S->var->U, tau);
    double tauSquare = tau * tau;
    axpy(&(S->nRows), &tau, S->var->ARDSum, &incx, S->var->constrValSum, &incx);
    axpy(&(S->nRows), &tauSquare, S->var->ADDSum, &incx, S->var->constrValSum, &incx);
    F->ALMCalGrad(S, S->var->rLp, S->var->gradLp, S->var->R, S->var->Grad, &lag, rho);
    F->setlbfgsHisTwo(S, S->var->gradLp, S->var->uLp, S->var->Grad, S->var->U, tau);
    F->updateDimacsALM(S, S->var->R, S->var->R, S->var->rLp, S->var->rLp);
    out_scalars[4] = lag;
    out_scalars[5] = S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
    return (int)rootNum;
}

/* scalar line search (lorads_alm.c:266-333) on caller-provided vectors (q0 is mutated like the reference) */
int rh_line_search(double rho, int n, double *lambd, double p1, double p2, double *q0, double *q1, double *q2, double *tau)
{
    return (int)ALMLineSearch(rho, n, lambd, p1, p2, q0, q1, q2, tau);
}
int rh_cubic(double a, double b, double c, double d, double *res) { return (int)LORADScubic_equation(a, b, c, d, res); }

void rh_update_dual_var(double rho) { LORADSUpdateDualVar(S, rho); }

/* LORADS_ALMtoADMM copies (lorads_solver.c:1351-1366): V<-R, U<-V */
void rh_alm_to_admm_copy(void)
{
    lorads_params p2 = P;
    p2.heuristicFactor = 1.0; p2.rhoMax = 1e300;
    LORADS_ALMtoADMM(S, &p2, &alm_state, &admm_state);
}
/* one ADMM sweep: LORADSUpdateSDPVar / LORADSUpdateSDPLPVar (lorads_alg_common.c:298-376); constrVal/Sum must be
 * initialised for (U,V) first, as LORADSADMMOptimize does (lorads_admm.c:98-99). returns cgIter accumulated */
int rh_admm_sweep(double rho, double cg_tol, int cg_maxiter, int init_constr)
{
    if (init_constr) {
        LORADSInitConstrValAll(S, S->var->uLp, S->var->vLp, S->var->U, S->var->V);
        LORADSInitConstrValSum(S);
        S->cgIter = 0;
    }
    F->admmUpdateVar(S, rho, cg_tol, cg_maxiter);
    return (int)S->cgIter;
}
double rh_cal_obj_admm(void) { S->scaleObjHis = 1; F->calObj_admm(S); return S->pObjVal; }
double rh_cal_obj_alm(void) { S->scaleObjHis = 1; F->calObj_alm(S); return S->pObjVal; }
double rh_dimacs_admm(void)
{
    F->updateDimacsADMM(S, S->var->U, S->var->V, S->var->uLp, S->var->vLp);
    return S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
}
int rh_oracle_rank(int phase) { return (int)lorads_compute_oracle_rank(S, phase); }
int rh_aug_rank(double factor) { return (int)AUG_RANK(S, S->var->rankElem, S->nCones, factor); }
double rh_dual_infeasibility(void)
{
    S->scaleObjHis = 1;
    calculate_dual_infeasibility_solver(S);
    return S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1];
}
