/*
 * main.c -- command-line driver of the B200-native LoRADS (drop-in for `LoRADS_v_2_0_1-alpha`).
 *
 * Restates the process-level contract of the reference's main() (lorads/src/src_semi/main.c:256-645):
 * argv[1] is the SDPA file, the 27 long options keep their names, value parsing (atof/atoi) and defaults
 * (main.c:56-86,125-154), unknown options are reported by getopt and ignored, the exit status is 0 unless the
 * device layer fails (there is no CPU fallback), the same banner / parameter echo / phase lines / result table /
 * timing lines are printed, and the same JSON file is written.  On top of that it accepts the three options
 * benchmark.py passes that the vendored reference silently ignores (benchmark.py:245-252):
 *   --rankSchedule <json>   {"rank_schedule":[...], "schedule_length":N}  (benchmark.py:123-133)
 *   --nearStallFactor <f>   scales the ALM stall counter threshold that triggers a rank change
 *   --disableOracle         skip the per-iteration oracle-rank eigen-decomposition (reported as 0)
 * `--device <id>` to choose the GPU, and `--ranks <P>` for a partitioned run on P GPUs of this node (row blocks of one
 * MaxCut-type cone, whole cones otherwise): the driver reads the file once, then forks P processes (one per GPU,
 * devices id .. id+P-1) BEFORE anything touches CUDA, so the ranks share the parsed problem copy-on-write; rank 0
 * creates the NCCL id and hands it to the others through pipes, every rank runs the same state machine on the same
 * reduced scalars, and only rank 0 prints and writes the log / JSON.
 */
#include <ctype.h>
#include <getopt.h>
#include <math.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include "lorads_host.h"

static void init_params(lh_params *p)
{
    memset(p, 0, sizeof(*p));
    p->fname = "NULL";
    p->initRho = 0.0;
    p->rhoMax = 5000.0;
    p->rhoCellingALM = 1e+8;
    p->rhoCellingADMM = p->rhoMax * 200;
    p->maxALMIter = 200;
    p->maxADMMIter = 10000;
    p->timesLogRank = 2.0;
    p->fixedRank = -1;
    p->initRank = -1;
    p->rhoFreq = 5;
    p->rhoFactor = 1.2;
    p->ALMRhoFactor = 2.0;
    p->rankUpdateFactor = 1.5;
    p->phase1Tol = 1e-3;
    p->phase2Tol = 1e-5;
    p->timeSecLimit = 3600.0;
    p->heuristicFactor = 1.0;
    p->lbfgsListLength = 2;
    p->endTauTol = 1e-16;
    p->endALMSubTol = 1e-10;
    p->l2Rescaling = 0;
    p->reoptLevel = 2;
    p->dyrankLevel = 2;
    p->highAccMode = 0;
    p->oracleRankMethod = 0; /* LORADS_ORACLE_RANK_GRAM */
    p->nearStallFactor = 1.0;
}

static struct option long_options[] = {
    {"logfile", required_argument, 0, 1025},
    {"jsonfile", required_argument, 0, 1026},
    {"initRho", required_argument, 0, 1000},
    {"rhoMax", required_argument, 0, 1001},
    {"rhoCellingALM", required_argument, 0, 1002},
    {"rhoCellingADMM", required_argument, 0, 1003},
    {"maxALMIter", required_argument, 0, 1004},
    {"maxADMMIter", required_argument, 0, 1005},
    {"timesLogRank", required_argument, 0, 1006},
    {"fixedRank", required_argument, 0, 1022},
    {"initRank", required_argument, 0, 1023},
    {"rhoFreq", required_argument, 0, 1007},
    {"rhoFactor", required_argument, 0, 1008},
    {"ALMRhoFactor", required_argument, 0, 1009},
    {"rankUpdateFactor", required_argument, 0, 1024},
    {"phase1Tol", required_argument, 0, 1010},
    {"phase2Tol", required_argument, 0, 1011},
    {"timeSecLimit", required_argument, 0, 1012},
    {"heuristicFactor", required_argument, 0, 1013},
    {"lbfgsListLength", required_argument, 0, 1014},
    {"endTauTol", required_argument, 0, 1015},
    {"endALMSubTol", required_argument, 0, 1016},
    {"l2Rescaling", required_argument, 0, 1017},
    {"reoptLevel", required_argument, 0, 1018},
    {"dyrankLevel", required_argument, 0, 1019},
    {"highAccMode", required_argument, 0, 1020},
    {"oracleRankNaive", no_argument, 0, 1021},
    /* accepted on top of the reference's table */
    {"rankSchedule", required_argument, 0, 2000},
    {"nearStallFactor", required_argument, 0, 2001},
    {"disableOracle", no_argument, 0, 2002},
    {"device", required_argument, 0, 2003},
    {"ranks", required_argument, 0, 2004},
    {"jsonFinalMetrics", no_argument, 0, 2005},
    {0, 0, 0, 0}};

static void print_input(const lh_params *p)
{
    printf("Input parameters:\n");
    printf("----------------------------------------------\n");
    printf("fname = %s\n", p->fname);
    printf("initRho = %f\n", p->initRho);
    printf("rhoMax = %f\n", p->rhoMax);
    printf("rhoCellingALM = %f\n", p->rhoCellingALM);
    printf("rhoCellingADMM = %f\n", p->rhoCellingADMM);
    printf("maxALMIter = %lld\n", (long long)p->maxALMIter);
    printf("maxADMMIter = %lld\n", (long long)p->maxADMMIter);
    printf("timesLogRank = %f\n", p->timesLogRank);
    printf("fixedRank = %lld\n", (long long)p->fixedRank);
    printf("initRank = %lld\n", (long long)p->initRank);
    printf("rhoFreq = %lld\n", (long long)p->rhoFreq);
    printf("rhoFactor = %f\n", p->rhoFactor);
    printf("ALMRhoFactor = %f\n", p->ALMRhoFactor);
    printf("rankUpdateFactor = %f\n", p->rankUpdateFactor);
    printf("phase1Tol = %f\n", p->phase1Tol);
    printf("phase2Tol = %f\n", p->phase2Tol);
    printf("timeSecLimit = %f\n", p->timeSecLimit);
    printf("heuristicFactor = %f\n", p->heuristicFactor);
    printf("lbfgsListLength = %lld\n", (long long)p->lbfgsListLength);
    printf("endTauTol = %f\n", p->endTauTol);
    printf("endALMSubTol = %f\n", p->endALMSubTol);
    printf("l2Rescaling = %d\n", p->l2Rescaling);
    printf("reoptLevel = %lld\n", (long long)p->reoptLevel);
    printf("dyrankLevel = %lld\n", (long long)p->dyrankLevel);
    printf("highAccMode = %d\n", p->highAccMode);
    printf("oracleRankMethod = %d\n", p->oracleRankMethod);
    if (p->rankScheduleFile) printf("rankSchedule = %s (nearStallFactor = %f)\n", p->rankScheduleFile, p->nearStallFactor);
    if (p->disableOracle) printf("disableOracle = 1\n");
    printf("----------------------------------------------\n");
}

/* {"rank_schedule":[int,...], "schedule_length":N} as benchmark.py:123-133 writes it.  Only the integer array that
 * follows the "rank_schedule" key is read; anything else in the file is ignored. */
static int read_rank_schedule(const char *path, int64_t **out, int64_t *len)
{
    *out = NULL;
    *len = 0;
    FILE *f = fopen(path, "r");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz <= 0 || sz > (1 << 24)) { fclose(f); return 1; }
    char *buf = (char *)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); return 1; }
    const size_t got = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    buf[got] = '\0';
    char *key = strstr(buf, "\"rank_schedule\"");
    char *s = key ? strchr(key, '[') : NULL;
    if (!s) { free(buf); return 1; }
    int64_t cap = 16, n = 0;
    int64_t *v = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    ++s;
    while (*s && *s != ']') {
        while (*s && (isspace((unsigned char)*s) || *s == ',')) ++s;
        if (*s == ']' || !*s) break;
        char *end = NULL;
        const double x = strtod(s, &end);
        if (end == s) { free(v); free(buf); return 1; }
        if (n == cap) {
            cap *= 2;
            v = (int64_t *)realloc(v, sizeof(int64_t) * (size_t)cap);
        }
        v[n++] = (int64_t)llround(x);
        s = end;
    }
    free(buf);
    if (n == 0) { free(v); return 1; }
    *out = v;
    *len = n;
    return 0;
}

static void on_sigint(int sig)
{
    (void)sig;
    exit(0); /* LUtilStartCtrlCCheck: SIGINT ends the process with status 0 (lorads_utils.c:35-40) */
}

static void print_res(double pObj, double dObj, double constrVio, double dualInfe, double pdgap, double constrVioInf,
                      double dualInfeInf)
{
    printf("-----------------------------------------------------------------------\n");
    printf("Objective function Value are:\n");
    printf("\t 1.Primal Objective:            : %10.6e\n", pObj);
    printf("\t 2.Dual Objective:              : %10.6e\n", dObj);
    printf("Dimacs Error are:\n");
    printf("\t 1.Constraint Violation(1)      : %10.6e\n", constrVio);
    printf("\t 2.Dual Infeasibility(1)        : %10.6e\n", dualInfe);
    printf("\t 3.Primal Dual Gap              : %10.6e\n", pdgap);
    printf("\t 4.Primal Variable Semidefinite : %10.6e\n", 0.0);
    printf("\t 5.Constraint Violation(Inf)    : %10.6e\n", constrVioInf);
    printf("\t 6.Dual Infeasibility(Inf)      : %10.6e\n", dualInfeInf);
    printf("-----------------------------------------------------------------------\n");
}

static void end_program(const lh_solver *S)
{
    printf("final rank: \n"); /* the per-cone ranks are not printed by the Linux build of the reference (quirk Q3) */
    printf("\n");
    printf("-----------------------------------------------------------------------\n");
    if (S->status == LH_STATUS_MAXITER) printf("End Program due to reaching `the maximum number of iterations`:\n");
    else if (S->status == LH_STATUS_PD_OPTIMAL) printf("End Program due to reaching `Official terminate criteria`:\n");
    else if (S->status == LH_STATUS_P_OPTIMAL) printf("End Program due to reaching `final terminate criteria`:\n");
    else if (S->status == LH_STATUS_UNKNOWN) printf("End Program but the status is unknown, please notify the authors\n");
    else if (S->status == LH_STATUS_TIME_LIMIT) printf("End Program since time limit.\n");
    print_res(S->pObjVal, S->dObjVal, S->dimacConstrVio, S->dimacDualInf, S->dimacGap,
              S->dimacConstrVio * (1 + S->bRHSNrm1) / (1 + S->bRHSNrmInf),
              S->dimacDualInf * (1 + S->cObjNrm1) / (1 + S->cObjNrmInf));
}

static void take_final_errors(const lh_solver *S, lh_admm_state *a)
{
    a->l_1_dual_infeasibility = S->dimacDualInf;
    a->l_inf_dual_infeasibility = S->dimacDualInf * (1 + S->cObjNrm1) / (1 + S->cObjNrmInf);
    a->l_2_dual_infeasibility = S->dimacDualInf * (1 + S->cObjNrm1) / (1 + S->cObjNrm2);
    a->primal_dual_gap = S->dimacGap;
    a->l_1_primal_infeasibility = S->dimacConstrVio;
    a->l_inf_primal_infeasibility = S->dimacConstrVio * (1 + S->bRHSNrm1) / (1 + S->bRHSNrmInf);
    a->l_2_primal_infeasibility = S->dimacConstrVio * (1 + S->bRHSNrm1) / (1 + S->bRHSNrm2);
}

int lorads_b200_main(int argc, char **argv)
{
    lh_params params;
    init_params(&params);
    if (argc < 2) {
        fprintf(stderr, "usage: %s <file.dat-s> [--option value ...]\n", argc > 0 ? argv[0] : "lorads_b200");
        return 0;
    }
    int opt, long_index = 0;
    params.fname = argv[1];
    optind = 1;
    while ((opt = getopt_long(argc, argv, "r:", long_options, &long_index)) != -1) {
        switch (opt) {
        case 1025: params.logFile = optarg; break;
        case 1026: params.jsonFile = optarg; break;
        case 1000: params.initRho = atof(optarg); break;
        case 1001: params.rhoMax = atof(optarg); break;
        case 1002: params.rhoCellingALM = atof(optarg); break;
        case 1003: params.rhoCellingADMM = atof(optarg); break;
        case 1004: params.maxALMIter = atoi(optarg); break;
        case 1005: params.maxADMMIter = atoi(optarg); break;
        case 1006: params.timesLogRank = atof(optarg); break;
        case 1022: params.fixedRank = atoi(optarg); break;
        case 1023: params.initRank = atoi(optarg); break;
        case 1007: params.rhoFreq = atoi(optarg); break;
        case 1008: params.rhoFactor = atof(optarg); break;
        case 1009: params.ALMRhoFactor = atof(optarg); break;
        case 1024: params.rankUpdateFactor = atof(optarg); break;
        case 1010: params.phase1Tol = atof(optarg); break;
        case 1011: params.phase2Tol = atof(optarg); break;
        case 1012: params.timeSecLimit = atof(optarg); break;
        case 1013: params.heuristicFactor = atof(optarg); break;
        case 1014: params.lbfgsListLength = atoi(optarg); break;
        case 1015: params.endTauTol = atof(optarg); break;
        case 1016: params.endALMSubTol = atof(optarg); break;
        case 1017: params.l2Rescaling = atoi(optarg); break;
        case 1018: params.reoptLevel = atoi(optarg); break;
        case 1019: params.dyrankLevel = atoi(optarg); break;
        case 1020: params.highAccMode = atoi(optarg); break;
        case 1021: params.oracleRankMethod = 1; break;
        case 2000: params.rankScheduleFile = optarg; break;
        case 2001: params.nearStallFactor = atof(optarg); break;
        case 2002: params.disableOracle = 1; break;
        case 2003: params.device = atoi(optarg); break;
        case 2004: params.ranks = atoi(optarg); break;
        case 2005: params.jsonFinalMetrics = 1; break;
        default: break;
        }
    }
    params.rhoCellingADMM = params.rhoMax * 200;

    printf("-----------------------------------------------------------\n");
    printf("  L         OOO      RRRR       A      DDDD       SSS \n");
    printf("  L        O   O     R   R     A A     D   D     S    \n");
    printf("  L        O   O     RRRR     AAAAA    D   D      SSS \n");
    printf("  L        O   O     R  R     A   A    D   D         S\n");
    printf("  LLLLL     OOO      R   R    A   A    DDDD       SSS \n");
    printf("-----------------------------------------------------------\n");
    print_input(&params);
    signal(SIGINT, on_sigint);

    lh_sdpa data;
    memset(&data, 0, sizeof(data));
    lh_solver SS;
    lh_solver *S = &SS;
    memset(S, 0, sizeof(*S));
    int exit_code = 0;

    const double timeStart = lh_time();
    if (lh_read_sdpa(params.fname, &data, 0) != 0) return 0; /* the reference also exits with status 0 here */
    printf("Reading SDPA file in %f seconds \n", lh_time() - timeStart);
    if (getenv("LORADS_SAVE_BINARY") && params.rank == 0 && lh_write_sdpa_binary(getenv("LORADS_SAVE_BINARY"), &data) != 0)
        fprintf(stderr, "lorads_b200: cannot write the binary image '%s'\n", getenv("LORADS_SAVE_BINARY"));
    printf("nConstrs = %lld, sdp nBlks = %lld, lp Cols = %lld\n", (long long)data.m, (long long)data.nBlks,
           (long long)data.nLpCols);

    /* ---- fork the ranks of a partitioned run ------------------------------------------------------------------
     * The file has been read ONCE, with all the cores; the ranks inherit the parsed problem copy-on-write.  No CUDA call
     * has happened yet and the reader's threads have been joined, so the children start from a single-threaded image. */
    pid_t kids[64];
    int nkids = 0;
    if (params.ranks > 64) params.ranks = 64;
    if (params.ranks > 1) {
        /* the ranks preprocess at the same time: the layout builder divides the cores by this */
        char nr[16];
        snprintf(nr, sizeof(nr), "%d", params.ranks);
        setenv("LORADS_LOCAL_RANKS", nr, 0);
        int fds[64][2];
        for (int r = 1; r < params.ranks; ++r)
            if (pipe(fds[r]) != 0) { perror("pipe"); return 3; }
        fflush(stdout);
        for (int r = 1; r < params.ranks; ++r) {
            const pid_t pid = fork();
            if (pid < 0) { perror("fork"); return 3; }
            if (pid == 0) { /* child: rank r */
                params.rank = r;
                nkids = 0;
                for (int q = 1; q < params.ranks; ++q) {
                    close(fds[q][1]);
                    if (q != r) close(fds[q][0]);
                }
                size_t got = 0;
                while (got < sizeof(params.ncclId)) {
                    const ssize_t k = read(fds[r][0], params.ncclId + got, sizeof(params.ncclId) - got);
                    if (k <= 0) { fprintf(stderr, "lorads_b200: rank %d did not receive the communicator id\n", r); return 3; }
                    got += (size_t)k;
                }
                close(fds[r][0]);
                if (!freopen("/dev/null", "w", stdout)) return 3; /* only rank 0 prints */
                params.logFile = NULL;
                params.jsonFile = NULL;
                break;
            }
            kids[nkids++] = pid;
        }
        if (params.rank == 0) {
            if (lgpu_nccl_unique_id(params.ncclId) != 0) { fprintf(stderr, "lorads_b200: cannot create the NCCL id\n"); return 3; }
            for (int r = 1; r < params.ranks; ++r) {
                close(fds[r][0]);
                if (write(fds[r][1], params.ncclId, sizeof(params.ncclId)) != (ssize_t)sizeof(params.ncclId)) return 3;
                close(fds[r][1]);
            }
        }
        params.device += params.rank;
    }

    const double timeSolveStart = lh_time();
    if (params.rankScheduleFile) {
        if (read_rank_schedule(params.rankScheduleFile, &S->schedule, &S->scheduleLen) != 0) {
            fprintf(stderr, "lorads_b200: cannot read rank schedule '%s'; using the default rank rule\n", params.rankScheduleFile);
            S->schedule = NULL;
            S->scheduleLen = 0;
        }
    }
    printf("Pre-solver starts \n");       /* LORADSPreprocess prints these three lines (lorads_solver.c:284-290) */
    printf("  Processing the cones \n");
    if (lh_setup_problem(S, &data, &params) != LH_RET_OK) { exit_code = 3; goto cleanup; }
    printf("  End preprocess \n");
    const int want_profile = getenv("LORADS_PROFILE") != NULL; /* per-kernel-class device time, printed at the end */
    if (want_profile) lgpu_profile_enable(S->gpu, 1);
    lh_determine_rank(S, &params);
    if (lh_init_variables(S, &params) != LH_RET_OK) { exit_code = 3; goto cleanup; }

    lh_alm_state alm;
    lh_admm_state admm;
    lh_initial_state(S, &params, &alm, &admm);
    lh_logging_init(S, &params, timeSolveStart);

    double reopt_param = 5;
    int64_t reopt_alm_iter = 3, reopt_admm_iter = 50;
    int64_t alm_reopt_min_iter = 3, admm_reopt_min_iter = params.highAccMode ? 1000 : 50;
    double initial_solving_time = 0.0, all_time = 0.0, all_dual_infea = 0.0;
    int admm_bad_iter_flag = 0;
    int rc;

    S->status = LH_STATUS_UNKNOWN;
    printf("-----------------------------------------------------------------------\n");
    printf("Start solving by ALM and ADMM\n");
    printf("-----------------------------------------------------------------------\n");
    const double all_time_start = lh_time();
    double time_start = lh_time(), time_end;
    rc = lh_alm_optimize(&params, S, &alm, timeSolveStart);
    if (rc == LH_RET_DEVICE) { exit_code = 3; goto close_log; }
    if (lh_time_is_up(S, &params, timeSolveStart, 1)) {
        printf("Time limit reached\n");
        S->status = LH_STATUS_TIME_LIMIT;
        goto end_solving;
    }
    lh_alm_to_admm(S, &params, &alm, &admm);
    rc = lh_admm_optimize(&params, S, &admm, params.maxADMMIter, timeSolveStart);
    if (rc == LH_RET_DEVICE) { exit_code = 3; goto close_log; }
    if (rc == LH_RET_BAD_ITER) admm_bad_iter_flag = 1;
    time_end = lh_time();
    initial_solving_time = time_end - time_start;
    all_time += initial_solving_time;

    if (params.reoptLevel >= 1) {
        int cnt = 0;
        while ((alm.primal_dual_gap > params.phase2Tol || alm.l_1_primal_infeasibility > params.phase2Tol) &&
               (admm.primal_dual_gap > params.phase2Tol || admm.l_1_primal_infeasibility > params.phase2Tol)) {
            if (cnt >= 1) break;
            printf("******  reopt parameter:%.3f\n", reopt_param);
            time_start = lh_time();
            lh_reopt(&params, S, &alm, &admm, &reopt_param, &alm_reopt_min_iter, &admm_reopt_min_iter, timeSolveStart,
                     &admm_bad_iter_flag, 1);
            time_end = lh_time();
            all_time += (time_end - time_start);
            cnt += 1;
            if (lh_time_is_up(S, &params, timeSolveStart, 1)) {
                printf("Time limit reached\n");
                S->status = LH_STATUS_TIME_LIMIT;
                goto end_solving;
            }
        }
    }
    time_start = lh_time();
    if (lh_dual_infeasibility(S) != LH_RET_OK) { exit_code = 3; goto close_log; }
    time_end = lh_time();
    all_dual_infea += (time_end - time_start);
    all_time += (time_end - time_start);
    take_final_errors(S, &admm);
    printf("-----------------------------------------------------------------------\n");
    printf("Dual infeasibility: l_1 = %f, l_inf = %f, l_2 = %f\n", admm.l_1_dual_infeasibility,
           admm.l_inf_dual_infeasibility, admm.l_2_dual_infeasibility);
    printf("-----------------------------------------------------------------------\n");
    if (params.reoptLevel >= 2) {
        int dual_cnt = 0;
        while (admm.l_1_dual_infeasibility > params.phase2Tol || admm.primal_dual_gap > params.phase2Tol ||
               admm.l_1_primal_infeasibility > params.phase2Tol) {
            if (dual_cnt >= 2) break;
            if (!params.highAccMode && admm.l_1_dual_infeasibility <= 5 * params.phase2Tol &&
                admm.primal_dual_gap <= 5 * params.phase2Tol && admm.l_1_primal_infeasibility <= 1 * params.phase2Tol)
                break;
            printf("******  reopt parameter:%.3f\n", reopt_param);
            time_start = lh_time();
            lh_reopt(&params, S, &alm, &admm, &reopt_param, &reopt_alm_iter, &reopt_admm_iter, timeSolveStart,
                     &admm_bad_iter_flag, 2);
            /* R = (U+V)/2 ; V = R (main.c:545-556) */
            if (lgpu_average_uv(S->gpu) != 0 || lgpu_copy_r_to_v(S->gpu) != 0) { exit_code = 3; goto close_log; }
            time_end = lh_time();
            all_time += (time_end - time_start);
            time_start = lh_time();
            if (lh_dual_infeasibility(S) != LH_RET_OK) { exit_code = 3; goto close_log; }
            time_end = lh_time();
            all_dual_infea += (time_end - time_start);
            all_time += (time_end - time_start);
            take_final_errors(S, &admm);
            printf("-----------------------------------------------------------------------\n");
            printf("reopt %d:Dual infeasibility: l_1 = %f, l_inf = %f, l_2 = %f\n", dual_cnt, admm.l_1_dual_infeasibility,
                   admm.l_inf_dual_infeasibility, admm.l_2_dual_infeasibility);
            printf("-----------------------------------------------------------------------\n");
            dual_cnt += 1;
            if (lh_time_is_up(S, &params, timeSolveStart, 1)) {
                printf("Time limit reached\n");
                S->status = LH_STATUS_TIME_LIMIT;
                goto end_solving;
            }
        }
    }
    if (admm.l_1_dual_infeasibility <= 5 * params.phase2Tol && admm.primal_dual_gap <= 5 * params.phase2Tol &&
        admm.l_1_primal_infeasibility <= 1 * params.phase2Tol)
        S->status = LH_STATUS_PD_OPTIMAL;
    else if (admm.primal_dual_gap <= 5 * params.phase2Tol && admm.l_1_primal_infeasibility <= 1 * params.phase2Tol)
        S->status = LH_STATUS_P_OPTIMAL;
    else
        S->status = LH_STATUS_MAXITER;

end_solving: {
    int64_t final_oracle_rank = lh_oracle_rank(S, 2);
    if (final_oracle_rank < 0) final_oracle_rank = 0;
    /* Quirk Q1 of the reference is kept by default: metrics.primal_obj / dual_obj come from the ADMM state, which stays at
     * 1e30 when ADMM returned at its first line (main.c:610, lorads_admm.c:86-88).  --jsonFinalMetrics writes the values of
     * the result table instead, so that benchmark.py:274 always reads a meaningful number; the flags only benchmark.py
     * passes (--rankSchedule / --nearStallFactor / --disableOracle, benchmark.py:245-252 -- the vendored C ignores them, so no
     * reference output exists under them) imply it.  LORADS_JSON_FINAL_OBJ=1/0 forces it on / off. */
    int json_final = params.jsonFinalMetrics || params.rankScheduleFile != NULL || params.disableOracle;
    if (getenv("LORADS_JSON_FINAL_OBJ") != NULL) json_final = atoi(getenv("LORADS_JSON_FINAL_OBJ")) != 0;
    const int admm_ran = admm.primal_objective_value < 1e29;
    if (json_final)
        lh_write_json(S, final_oracle_rank, S->pObjVal, S->dObjVal,
                      admm_ran ? admm.l_1_primal_infeasibility : alm.l_1_primal_infeasibility,
                      admm_ran ? admm.l_inf_primal_infeasibility : alm.l_inf_primal_infeasibility,
                      admm_ran ? admm.primal_dual_gap : alm.primal_dual_gap, all_time, params.rhoMax, params.heuristicFactor);
    else
        lh_write_json(S, final_oracle_rank, admm.primal_objective_value, admm.dual_objective_value, admm.l_1_primal_infeasibility,
                      admm.l_inf_primal_infeasibility, admm.primal_dual_gap, all_time, params.rhoMax, params.heuristicFactor);
    lh_logging_close(S);
    end_program(S);
    all_time = lh_time() - all_time_start;
    if (S->status == LH_STATUS_TIME_LIMIT) printf("Time limit reached :%f\n", params.timeSecLimit);
    printf("initial solving: %f\n", initial_solving_time);
    printf("all_time - all_dual_infea: %f\n", all_time - all_dual_infea);
    printf("all_dual_infea: %f\n", all_dual_infea);
    printf("all_time: %f\n", all_time);
    fprintf(stderr, "lorads_b200: %lld kernel launches\n", (long long)lgpu_launch_count(S->gpu)); /* stdout stays the reference's */
    if (want_profile) {
        double ms[32];
        int64_t cnt[32];
        const int nc = lgpu_profile_num_classes();
        if (nc <= 32 && lgpu_profile_read(S->gpu, nc, ms, cnt) == 0)
            for (int k = 0; k < nc; ++k)
                if (cnt[k] > 0)
                    printf("profile %-14s launches %9lld  device ms %12.3f\n", lgpu_profile_class_name(k), (long long)cnt[k], ms[k]);
    }
    goto cleanup;
}
close_log:
    lh_logging_close(S);
cleanup:
    lh_free_solver(S);
    lh_free_sdpa(&data);
    for (int k = 0; k < nkids; ++k) { /* rank 0 reaps the other ranks; a failed rank fails the run */
        int st = 0;
        if (waitpid(kids[k], &st, 0) > 0 && (!WIFEXITED(st) || WEXITSTATUS(st) != 0) && exit_code == 0) exit_code = 3;
    }
    return exit_code;
}

#ifndef LORADS_B200_NO_MAIN
int main(int argc, char **argv) { return lorads_b200_main(argc, argv); }
#endif
