/*
 * lorads_host.h -- host side (plain C) of the B200-native LoRADS.
 *
 * The host owns everything the reference's main.c / lorads_alm.c / lorads_admm.c own that is NOT
 * arithmetic on big arrays: the CLI, the SDPA reader, the rank heuristic, the seeded initial point,
 * the ALM / ADMM state machines, the scalar line-search root selection, logging and JSON.  All array
 * work is delegated to the C ABI in include/lorads_b200.h.
 */
#ifndef LORADS_HOST_H
#define LORADS_HOST_H

#include <stdint.h>
#include <stdio.h>

#include "lorads_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- command line (reference: lorads_params, lorads.h:131-160; defaults main.c:56-86) ---------*/
typedef struct {
    const char *fname;
    const char *logFile;
    const char *jsonFile;
    double initRho, rhoMax, rhoCellingALM, rhoCellingADMM;
    int64_t maxALMIter, maxADMMIter;
    double timesLogRank;
    int64_t fixedRank, initRank, rhoFreq;
    double rhoFactor, ALMRhoFactor, rankUpdateFactor, phase1Tol, phase2Tol, timeSecLimit, heuristicFactor;
    int64_t lbfgsListLength;
    double endTauTol, endALMSubTol;
    int l2Rescaling;
    int64_t reoptLevel, dyrankLevel;
    int highAccMode;
    int oracleRankMethod;
    /* options benchmark.py passes that the vendored reference does not implement (benchmark.py:245-252) */
    const char *rankScheduleFile;
    double nearStallFactor;
    int disableOracle;
    int jsonFinalMetrics;
    /* B200 additions */
    int device;
    int quiet;
    /* row-block partitioned run: `ranks` processes (one per GPU) forked by the driver, this one is `rank` */
    int ranks, rank;
    unsigned char ncclId[128];
} lh_params;

/* ---- SDPA data exactly as the reference reader hands it on (LReadSDPA, lorads_file_io.c:59) ----*/
typedef struct {
    int64_t m;          /* constraints */
    int64_t nBlks;      /* SDP blocks */
    int64_t *blkDims;
    int64_t nLpCols;
    double *b;
    /* per SDP block: CSC over packed lower indices, m+2 column pointers (col 0 = objective) */
    int64_t **matBeg;
    int64_t **matIdx;
    double **matElem;
    /* LP block: CSC with m+1 columns over LP column ids */
    int64_t *lpBeg;
    int64_t *lpIdx;
    double *lpElem;
    int64_t nElems;
} lh_sdpa;

/* reads SDPA sparse text (.dat-s) or, recognised by its magic, the binary image lh_write_sdpa_binary made of it */
int lh_read_sdpa(const char *fname, lh_sdpa *out, int quiet);
int lh_write_sdpa_binary(const char *fname, const lh_sdpa *d);
void lh_free_sdpa(lh_sdpa *d);
/* features.c: hand-off to the reference's feature extractor (dataset/processor.py:246-345) from the parsed arrays */
int lh_constraint_stats(const lh_sdpa *d, double *out /* m x 7 */, double *out_obj /* 7, or NULL */);
int lh_constraint_rows(const lh_sdpa *d, int64_t *ptr /* m + 1 */, int64_t *rows /* or NULL to size */, int64_t *count);
/* couplings (processor.py:347-366, :497-505, :580-600, :640-643): <A_i, C>, rows shared with C; pairs i < j with a common row */
int lh_constraint_cost_alignment(const lh_sdpa *d, double *inner /* m, or NULL */, int64_t *rows_shared /* m, or NULL */);
int lh_constraint_pairs(const lh_sdpa *d, int64_t *ptr /* m + 1 */, int64_t *col /* or NULL to size */, int64_t *overlap,
                        double *inner, int64_t *count);

/* ---- phase states (reference: lorads_alm_state / lorads_admm_state, def_lorads_solver.h:198-238) */
typedef struct {
    double primal_objective_value, dual_objective_value;
    double l_1_primal_infeasibility, l_inf_primal_infeasibility, l_2_primal_infeasibility;
    double l_1_dual_infeasibility, l_inf_dual_infeasibility, l_2_dual_infeasibility;
    double primal_dual_gap;
    double rho;
    int64_t outerIter, innerIter;
    double tau;
} lh_alm_state;

typedef struct {
    double primal_objective_value, dual_objective_value;
    double l_1_primal_infeasibility, l_inf_primal_infeasibility, l_2_primal_infeasibility;
    double l_1_dual_infeasibility, l_inf_dual_infeasibility, l_2_dual_infeasibility;
    double primal_dual_gap;
    double rho;
    int64_t iter, cg_iter, nBlks;
} lh_admm_state;

enum { LH_STATUS_UNKNOWN = 0, LH_STATUS_PD_OPTIMAL, LH_STATUS_P_OPTIMAL, LH_STATUS_MAXITER, LH_STATUS_TIME_LIMIT };
enum { LH_RET_OK = 0, LH_RET_TIME_OUT = 1, LH_RET_NUM_ERR = 2, LH_RET_BAD_ITER = 4, LH_RET_DEVICE = 64 };

typedef struct {
    lgpu_ctx *gpu;
    int64_t m, nCones, nLpCols;
    int64_t *blkDims;
    int64_t *rank, *rankMax, *nnzRows;
    /* constants (cal_sdp_const) */
    double cObjNrm1, cObjNrm2, cObjNrmInf, bRHSNrm1, bRHSNrm2, bRHSNrmInf;
    double scaleObjHis;
    double pObjVal, dObjVal;
    double dimacConstrVio, dimacDualInf, dimacGap;
    int64_t cgIter;
    int status;
    int maxAlmSubIter; /* the reference's global MAX_ALM_SUB_ITER (lorads_alm.c:20) */
    /* rank schedule (no reference semantics) */
    int64_t *schedule;
    int64_t scheduleLen, schedulePos;
    /* logging */
    FILE *logFp;
    char problemName[4096];
    char inputPath[4096];
    char jsonPath[8192];
    double solveStartTime;
    int oracleMethod;
    int disableOracle;
    int64_t *p1Curr, *p1Oracle, p1Count, p1Cap;
    int64_t *p2Curr, *p2Oracle, p2Count, p2Cap;
} lh_solver;

double lh_time(void);
/* the reference's `time - timeSolveStart >= timeSecLimit` tests, made collective: in a --ranks P run every rank reads
 * its own clock, so the decision is rank 0's, distributed through the communicator (lgpu_agree_flag) */
int lh_time_is_up(lh_solver *S, const lh_params *p, double timeSolveStart, int strict);

/* glibc_rand.c: glibc's srand()/rand() value stream, lock-free */
void lh_srand(unsigned int seed);
int lh_rand(void);
void lh_random_fill(double *a, int64_t n);
void lh_log(lh_solver *S, const char *fmt, ...);

/* setup */
int lh_setup_problem(lh_solver *S, const lh_sdpa *d, const lh_params *p);
void lh_determine_rank(lh_solver *S, const lh_params *p);
int lh_init_variables(lh_solver *S, const lh_params *p);
void lh_initial_state(lh_solver *S, const lh_params *p, lh_alm_state *alm, lh_admm_state *admm);
int lh_all_rank_max(const lh_solver *S, double aug_factor);
int lh_aug_rank(lh_solver *S, double aug_factor, const lh_params *p);
void lh_free_solver(lh_solver *S);

/* scalar line search (reference: LORADScubic_equation / ALMLineSearch, lorads_alm.c:191-333) */
int lh_cubic_equation(double a, double b, double c, double d, double *res);
int lh_line_search(double rho, const double terms[7], double *tau);

/* phases */
int lh_alm_optimize(lh_params *p, lh_solver *S, lh_alm_state *st, double timeSolveStart);
int lh_alm_optimize_reopt(lh_params *p, lh_solver *S, lh_alm_state *st, int early_stop, double rho_update_factor,
                          double timeSolveStart);
void lh_alm_to_admm(lh_solver *S, lh_params *p, lh_alm_state *alm, lh_admm_state *admm);
int lh_admm_optimize(lh_params *p, lh_solver *S, lh_admm_state *st, int64_t iter_celling, double timeSolveStart);
int lh_admm_optimize_reopt(lh_params *p, lh_solver *S, lh_admm_state *st, int64_t iter_celling, double timeSolveStart);
double lh_reopt(lh_params *p, lh_solver *S, lh_alm_state *alm, lh_admm_state *admm, double *reopt_param,
                int64_t *reopt_alm_iter, int64_t *reopt_admm_iter, double timeSolveStart, int *admm_bad_iter_flag,
                int reopt_level);
int lh_dual_infeasibility(lh_solver *S);

/* logging / oracle rank / JSON (reference: lorads_logging.c) */
void lh_logging_init(lh_solver *S, const lh_params *p, double solve_start);
void lh_logging_close(lh_solver *S);
int64_t lh_sum_rank(const lh_solver *S);
int64_t lh_oracle_rank(lh_solver *S, int phase);
void lh_append_trajectory(lh_solver *S, int phase, int64_t cur_rank, int64_t oracle_rank);
void lh_write_json(lh_solver *S, int64_t final_oracle_rank, double pobj, double dobj, double l1, double linf, double gap,
                   double solve_time, double rho_max, double heuristic_factor);
int lh_sym_eigvals(int n, double *a, double *w); /* Jacobi; a is destroyed */
int64_t lh_sym_rank(int n, double *a, double eps);

/* whole program: what `main` of the reference binary does (main.c:256-645); returns the process exit code */
int lorads_b200_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
