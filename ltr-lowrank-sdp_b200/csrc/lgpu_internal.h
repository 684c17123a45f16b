/*
 * lgpu_internal.h -- internal structures of liblorads_b200.so (not part of the C ABI).
 *
 * Data layout in HBM (DESIGN.md "Data layout"):
 *   factors   row-major n x ld doubles, ld = r rounded up to a multiple of 4 (32-byte rows), padding
 *             columns are identically zero in every vector so flat BLAS-1 style kernels can run over
 *             the padded storage.  All cones (and the LP scalars behind them) live in ONE flat buffer
 *             per variable so L-BFGS / vector kernels are single launches over N = sum n_c ld_c + nLp.
 *   cone data uploaded once: aggregated lower pattern (row,col), C on the pattern, constraints as
 *             CSR-by-constraint over pattern slots, the slot-transposed CSR (so A*(w) is a gather, no
 *             atomics), and the full symmetric CSR (row -> (col, slot)) for the SpMM.
 */
#ifndef LGPU_INTERNAL_H
#define LGPU_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#define LGPU_NSCALAR 256
#define LGPU_MAX_WORLD 16   /* ranks of one node */
#define LGPU_PEER_RED 24    /* widest scalar pack of the one-shot peer all-reduce */
#define LGPU_MAX_PARTIAL_BLOCKS 1184 /* 148 SMs x 8 resident CTAs of 256 threads */
#define LGPU_MAX_REDUCE 16           /* widest fused reduction (k_reduce / grid_reduce_finish) */

/* device scalar slots (ctx->dsc) */
enum {
    SC_TMP = 0,
    SC_TMP2,
    SC_NEG,
    SC_LAG,    /* sum |Grad|^2 */
    SC_YS,     /* <y, s> */
    SC_PINF,   /* |b - A|^2        (SC_LAG .. SC_YNYN stay contiguous: one fetch / one all-reduce) */
    SC_GSN,    /* carried L-BFGS inner products (history length 2): <g, s_new> */
    SC_GYN,    /* <g, y_new> */
    SC_GSO,    /* <g, s_old> */
    SC_GYO,    /* <g, y_old> */
    SC_SOYN,   /* <s_old, y_new> */
    SC_YOYN,   /* <y_old, y_new> */
    SC_YNYN,   /* <y_new, y_new> */
    SC_YOYO,   /* <y_old, y_old> (refresh pass only) */
    SC_BN,     /* refresh pass only: <y_new, s_new>, <y_old, s_old> */
    SC_BO,
    SC_CRG,    /* carried <C R, .> with g, s_new, y_new, s_old, y_old (bulk step pass): <C R, D> without reading C R again */
    SC_CRSN,
    SC_CRYN,
    SC_CRSO,
    SC_CRYO,   /* SC_LAG .. SC_CRYO stay contiguous: one fetch / one all-reduce of the step's pack */
    SC_DG,     /* <D, Grad> */
    SC_P1,     /* <C, R D^T> accumulated over cones (not yet doubled) */
    SC_P2,     /* <C, D D^T> */
    SC_LS0,    /* |q2|^2 */
    SC_LS1,    /* q1.q2 */
    SC_LS2,    /* q0'.q2 */
    SC_LS3,    /* |q1|^2 */
    SC_LS4,    /* q0'.q1 */
    SC_OBJ,    /* <C, RR^T> */
    SC_DOBJ,   /* b^T lambda */
    SC_CG_RR,  /* r.r */
    SC_CG_PQ,  /* p.Q */
    SC_CG_ALPHA,
    SC_CG_NEGALPHA,
    SC_CG_RES, /* |r|^2 after update */
    SC_CG_B1,  /* |b|_1 */
    SC_CG_BETA,
    SC_ALPHA0, /* L-BFGS alpha[node], LGPU_MAX_HIST slots */
    SC_BETA0 = SC_ALPHA0 + 16, /* L-BFGS beta[node] */
    SC_LANCZOS = SC_BETA0 + 16,
    SC_END = SC_LANCZOS + 8
};

/* kernel classes for the live per-launch timing (lgpu_profile_*), bench.py's roofline source */
enum {
    KC_UVT = 0,     /* k_uvt: pattern samples of sym(U V^T) */
    KC_GATHER,      /* k_con_gather / objective gather / k_mc_rowdot / k_mc_epi */
    KC_WSUM,        /* k_wsum: A^*(w) (+C) on the pattern */
    KC_SPMM,        /* k_spmm: S X */
    KC_VEC,         /* flat elementwise kernels */
    KC_REDUCE,      /* flat fused elementwise + reduction kernels */
    KC_SCALAR,      /* one-thread scalar kernels */
    KC_LAYOUT,      /* transposes / re-stride / gram */
    KC_MC_SPMM,     /* MaxCut-type fused: T = C D with q1/q2/p1/p2 epilogue */
    KC_MC_STEP,     /* MaxCut-type fused: step + gradient + L-BFGS pair + A(RR^T) */
    KC_MC_DIR,      /* fused two-loop passes */
    KC_DENSE,       /* dense-aggregate cones: DMMA SYR2K / SYMM */
    KC_EXCH,        /* partitioned runs: peer PUT of the halo rows + wait for the sources */
    KC_COUNT
};

struct ProfRec {
    int cls;
    cudaEvent_t a, b;
};

/* one per rank, IPC-shared: written by the peers, read by the owner */
struct PeerBlock {
    unsigned long long xflag[LGPU_MAX_WORLD];                 /* [src] = last exchange sequence number src completed */
    unsigned long long aflag[LGPU_MAX_WORLD];                 /* [src] = last all-reduce sequence number src posted */
    double inbox[2][LGPU_MAX_WORLD][LGPU_PEER_RED];           /* [seq & 1][src][k] */
};

struct DevCone {
    /* sizes */
    int64_t n = 0;       /* block dimension; in a row-block partitioned run: number of rows THIS rank owns */
    int64_t n_glob = 0;  /* block dimension of the whole problem */
    int64_t row_lo = 0;  /* first owned row */
    int64_t n_alloc = 0; /* rows allocated per rank (= n, or ceil(n_glob / world) so that all-gather counts are equal) */
    int64_t m_loc = 0;   /* constraints owned by this rank (partitioned) */
    int64_t mA = 0;      /* number of non-zero constraints in this block */
    int64_t nnzP = 0;    /* aggregated lower pattern */
    int64_t nnzA = 0;    /* sum_i nnz(A_i) */
    int64_t nnzC = 0;
    int64_t nnzF = 0;    /* full symmetric pattern = 2 nnzP - #diag */
    /* reference storage classes (for parity of the rules; the device layout is uniform) */
    int obj_type = 0;
    bool dense_aggregate = false;
    bool sparse_container = false;
    bool diag_only = false; /* every non-zero A_i is one diagonal entry (MaxCut-type) */
    int64_t max_con_len = 0;
    int64_t max_slot_len = 0;
    /* C norms */
    double c_nrm1 = 0, c_nrm2sq = 0, c_nrminf = 0;
    /* host copies kept for lgpu_cone_pattern */
    std::vector<int32_t> h_pat_row, h_pat_col;
    /* row relabelling for gather locality (fused layout): device row perm[i] = caller's row i; iperm = the inverse */
    bool reordered = false;
    double window_before = 0.0, window_after = 0.0;
    int32_t *perm = nullptr, *iperm = nullptr;
    std::vector<int32_t> h_iperm;
    /* device arrays */
    int32_t *pat_row = nullptr, *pat_col = nullptr;
    double *cval = nullptr;                      /* [nnzP] C on the pattern */
    int32_t *c_slot = nullptr; double *c_coef = nullptr; /* [nnzC] (2-delta) c */
    int32_t *a_ptr = nullptr, *a_slot = nullptr; double *a_coef = nullptr; /* CSR by constraint, coef=(2-delta)a */
    int32_t *con_gid = nullptr;                  /* [mA] global constraint index */
    int32_t *long_con = nullptr; int64_t n_long_con = 0; int con_group = 1; /* constraints with outlier list lengths */
    int32_t *t_ptr = nullptr, *t_loc = nullptr, *t_gid = nullptr; double *t_val = nullptr; /* by slot */
    int32_t *f_ptr = nullptr, *f_col = nullptr, *f_slot = nullptr; /* full CSR */
    /* rows with > LGPU_LONG_ROW entries, cut into chunks: work item w = entries [lw_beg, lw_end) of row lw_row */
    int32_t *long_rows = nullptr, *long_first = nullptr; int64_t n_long = 0;
    int32_t *lw_row = nullptr, *lw_beg = nullptr, *lw_end = nullptr; int64_t n_lwork = 0;
    double *long_scratch = nullptr; int64_t long_scratch_ld = 0;
    double *symm_scratch = nullptr; int64_t symm_scratch_len = 0; /* split-k partials of the dense SYMM */
    int32_t *d_row = nullptr; double *d_val = nullptr; /* diag_only: row and value per constraint */
    /* diag_only fused path: C per full-CSR entry, and row -> constraints (global id, a_k) */
    double *mc_val = nullptr;
    int32_t *rc_ptr = nullptr, *rc_gid = nullptr; double *rc_a = nullptr;
    /* scratch */
    double *uvt = nullptr;  /* [nnzP] */
    double *S = nullptr;    /* [nnzP] aggregate values (sdp_obj_sum / sdp_coeff_w_sum / slack) */
    double *cv = nullptr;   /* [mA] constrVal of this cone (compact) */
    double *wtmp = nullptr; /* [mA] CG weight vector (compact) */
    /* variables */
    int64_t r = 0, ld = 0;
    int64_t off = 0;        /* offset of this cone inside the flat factor buffers */
};

struct DevLp {
    int64_t n = 0;          /* number of LP columns */
    int64_t nnz = 0;
    double *obj = nullptr;  /* [n] */
    /* by constraint (AUV): y_i += sum a_ij u_j v_j */
    int32_t *r_ptr = nullptr, *r_col = nullptr; double *r_val = nullptr;
    /* by column (WSum) */
    int32_t *c_ptr = nullptr, *c_row = nullptr; double *c_val = nullptr;
    double *nrm2sq = nullptr; /* [n] |a_j|^2 (host copy below) */
    std::vector<double> h_obj, h_nrm2sq;
    std::vector<int32_t> h_c_ptr, h_c_row; std::vector<double> h_c_val;
    double nrm1 = 0, nrminf_q = 0;
    int64_t off = 0;        /* offset inside the flat buffers */
};

struct lgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool failed = false; /* sticky: set by every LGPU_FAIL, tested by CHECK_LAUNCH */
    int64_t launches = 0;
    int num_sms = 148;

    int64_t m = 0;
    int ncones = 0;
    std::vector<DevCone> cones;
    DevLp lp;
    std::vector<double> h_b;
    std::vector<int32_t> h_cperm; /* non-empty: device m-vectors are in renumbered constraint order, cperm[k] = device id of k */
    double b_nrm1 = 0, b_nrm2 = 0, b_nrminf_q = 0;

    /* m-vectors */
    double *b = nullptr, *lam = nullptr, *cvs = nullptr, *q1 = nullptr, *q2 = nullptr, *M1 = nullptr,
           *mtmp = nullptr;
    /* flat N-vectors */
    int64_t N = 0;
    bool vars_ready = false;
    double *R = nullptr, *U = nullptr, *V = nullptr, *G = nullptr, *M2 = nullptr, *bLin = nullptr,
           *cg_r = nullptr, *cg_p = nullptr, *cg_Q = nullptr, *stage = nullptr;
    /* row-block partition over `world` ranks (NCCL); world == 1: single GPU */
    int rank = 0, world = 1;
    void *comm = nullptr;     /* ncclComm_t */
    /* by-cone partition (world > 1 and the problem is not the single MaxCut-type cone): everything replicated, the operator
     * work of cone c done by rank cone_owner[c] (lgpu_api.cu, "by-cone partition") */
    bool cone_par = false;
    std::vector<int> cone_owner;
    double *gfull = nullptr;  /* all-gather mode: [world * n_alloc * ld] factor rows of every rank, global row order */
    /* halo mode: only the remote rows this rank's CSR rows reference are exchanged (ncclSend/ncclRecv), into
     * `halo` (rows grouped by owner, ascending); CSR columns are remapped: local row j - lo, halo row n_alloc + k */
    bool use_halo = false;
    int64_t halo_rows = 0, send_rows = 0;
    std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt; /* per peer, in rows */
    int32_t *send_idx = nullptr; /* [send_rows] local row to pack, grouped by destination */
    int32_t *halo_gid = nullptr; /* [halo_rows] global row of each halo row */
    double *sendbuf = nullptr;   /* [send_rows * ld] */
    double *halo = nullptr;      /* [halo_rows * ld]  (peer mode: [2][halo_rows * ld], alternating per exchange) */
    /* Peer-memory exchange (DESIGN.md "Multi-GPU"): every rank maps the other ranks' halo buffers and a small
     * communication block through CUDA IPC; the halo rows are PUT straight into the consumers' buffers by k_put_rows
     * (NVLink stores), completion is a per-source sequence number in the consumer's block, and the scalar packs are
     * all-reduced by one small kernel through the peers' inboxes.  Falls back to NCCL when the mapping is unavailable. */
    bool peer = false;
    struct PeerBlock *blk = nullptr;              /* this rank's block (device memory, IPC-exported) */
    struct PeerBlock *peer_blk[LGPU_MAX_WORLD] = {nullptr}; /* [q] = rank q's block as mapped here ([rank] = blk) */
    double *peer_halo[LGPU_MAX_WORLD] = {nullptr};          /* [q] = rank q's halo allocation as mapped here */
    std::vector<int64_t> dst_off;                 /* [q] = first row of MY block inside rank q's halo */
    int64_t peer_halo_rows[LGPU_MAX_WORLD] = {0}; /* [q] = rank q's halo size in rows (its two buffers are that far apart) */
    unsigned long long xseq = 0, aseq = 0;        /* exchange / all-reduce sequence numbers (same on all ranks) */
    const double *halo_cur = nullptr;             /* halo buffer the current product reads */
    unsigned int *put_counter = nullptr;
    bool fuse_put = true;                         /* LORADS_FUSE_PUT=0: separate k_put_rows after the direction pass (A/B) */
    int32_t *put_dest = nullptr;                  /* [n_loc * (world - 1)] row of each own row in each peer's halo, or -1 */
    const double *pending_put_x = nullptr;        /* vector whose halo rows the direction pass already PUT ... */
    unsigned long long pending_put_seq = 0;       /* ... under this exchange sequence number */
    /* fused MaxCut-type path (single diag_only cone, no LP): CR = C R carried across iterations, CD = C D */
    bool dense_dmma = true;   /* dense-aggregate cones: SYR2K / SYMM on the FP64 tensor pipe */
    bool fast_enabled = true;
    bool mc = false;
    double *CR = nullptr, *CD = nullptr;
    bool cr_valid = false, cd_valid = false;
    int cr_updates = 0;
    /* carried inner products of the two L-BFGS pairs and the current gradient (fused path, history length 2) */
    int step_variant = 1; /* min CTAs/SM of k_mc_step: 0 -> 2, 1 -> 3 (default; measured best: 80 registers, no spills), 2 -> 4, 3 -> 5 */
    bool gram_enabled = true;
    int spmm_dot = 3;       /* <D, C D> in the sparse product's epilogue: 0 separate pass; 1..4 kernel variants (A/B) */
    int step_bulk = 1;      /* k_mc_step through the bulk-copy pipeline (LORADS_STEP_BULK=0: register-staged kernel) */
    int step_tile_rows = 0, step_stages = 0; /* 0: chosen from ld (LORADS_STEP_TILE / LORADS_STEP_STAGES override) */
    bool gram_valid = false;
    bool gram_pair_ok[2] = {false, false};
    bool defer_allreduce = false; /* partitioned: keep local sums, a later call all-reduces the whole pack */
    struct {
        double gg, sg[2], yg[2], yy[2], beta[2], so_yn, yo_yn; /* indexed by ring slot; cross terms: (older, newer) */
        double cr_g, cr_s[2], cr_y[2];                          /* <C R, g>, <C R, s_j>, <C R, y_j> */
    } gram;
    /* per-row inner products <R_i, g_i>, <R_i, s_new_i>, <R_i, y_new_i>, <R_i, s_old_i>, <R_i, y_old_i> written by the bulk
     * step pass ([n_alloc][5]); with them and the carried <C R, .> the direction pass needs neither R nor C R */
    double *rowdots = nullptr;
    bool rowdots_valid = false, rowdots_enabled = true, rowdots_force = false;
    double p1_host = 0.0; /* <C R, D> formed on the host from the carried products (added to dsc[SC_P1], which is then zero) */
    bool epi_done = false; /* q1, q2 and <C, R D^T> already produced by the direction pass for the current D */
    int h = 0, head = 0;
    std::vector<double *> s, y;
    std::vector<int64_t> cg_last_iter; /* per cone: cg->iter persists across calls (reference quirk) */

    /* scalars and reduction scratch */
    double *dsc = nullptr;      /* device scalars */
    double *hsc = nullptr;      /* pinned host mirror (device-mapped) */
    double *hsc_dev = nullptr;  /* the mirror's device address */
    unsigned long long *hflag = nullptr, *hflag_dev = nullptr; /* read-back sequence number, host / device address */
    unsigned long long fetch_seq = 0;
    bool fast_fetch = true;
    double *partials = nullptr; /* [LGPU_MAX_REDUCE * LGPU_MAX_PARTIAL_BLOCKS] */
    unsigned int *counter = nullptr;
    /* live per-launch timing */
    bool prof = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[KC_COUNT] = {0};
    int64_t prof_cnt[KC_COUNT] = {0};
    cudaEvent_t timers[8] = {nullptr};
    /* generic staging for host<->device operator calls */
    void *dstage = nullptr; size_t dstage_bytes = 0;
};

#endif
