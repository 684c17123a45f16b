/*
 * lorads_b200.h -- C ABI of the B200-native LoRADS inner loop (liblorads_b200.so).
 *
 * The reference solver (muhd-umer/ltr-lowrank-sdp, lorads/src/src_semi) has no FFI: its drop-in
 * boundary is the process (argv + .dat-s in, log/JSON out; main.c:256-645).  Inside that process
 * every heavy operation goes through three internal vtables, and THOSE are what this ABI replaces:
 *
 *   cone vtable      coneAUV / objAUV / sdpDataWSum / addObjCoeff      def_lorads_sdp_conic.h:101-124
 *   sdp_coeff vtable mul_rk / mv / mul_inner_rk_double / add_sdp_coeff  def_lorads_sdp_data.h:66-85
 *   lorads_func      ALMCalGrad, ALMCalq12p12, LBFGSDirection, LBFGSDirUseGrad, setAsNegGrad,
 *                    setlbfgsHisTwo, ALMupdateVar, updateDimacsALM/ADMM, calObj_alm/admm,
 *                    admmUpdateVar, copyRtoV, InitConstrValAll/Sum       def_lorads_solver.h:169-187
 *
 * Conventions
 *   - plain C: opaque handle, plain pointers and sizes, int return (0 = ok, nonzero = error; the
 *     message is available from lgpu_last_error()).  No exceptions cross the boundary.
 *   - every `double *` / index pointer in a signature is a HOST buffer owned by the caller unless
 *     the name ends in `_dev`.  Device memory is owned by the library.
 *   - factors are exchanged in the reference's layout: column-major n x r doubles
 *     (lorads_sdp_dense.matElem, def_lorads_elements.h:60-64).  On the device they are row-major
 *     with padded leading dimension; the transposition happens inside lgpu_set/get_factor.
 *   - there is NO CPU fallback: every entry point fails (nonzero) if no CUDA device is usable.
 */
#ifndef LORADS_B200_H
#define LORADS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lgpu_ctx lgpu_ctx;

/* which factor / vector */
enum { LGPU_R = 0, LGPU_U = 1, LGPU_V = 2, LGPU_GRAD = 3 };
enum { LGPU_VEC_DUAL = 0, LGPU_VEC_CONSTR_SUM = 1, LGPU_VEC_ARD = 2, LGPU_VEC_ADD = 3, LGPU_VEC_M1 = 4,
       LGPU_VEC_B = 5 };
/* factor pairs for the A(UV^T) entry points */
enum { LGPU_PAIR_RR = 0, LGPU_PAIR_RU = 1, LGPU_PAIR_UU = 2, LGPU_PAIR_UV = 3 };

/* ---- lifecycle -------------------------------------------------------------------------------*/
int lgpu_create(lgpu_ctx **ctx, int device);
void lgpu_destroy(lgpu_ctx *ctx);
const char *lgpu_last_error(const lgpu_ctx *ctx);
const char *lgpu_version(void);
/* number of kernels this context has launched since creation (bench.py "gpu_launches") */
int64_t lgpu_launch_count(const lgpu_ctx *ctx);

/* ---- timing hooks (no reference counterpart; bench.py times on the library's own stream with these) ----*/
int lgpu_sync(lgpu_ctx *ctx);
/* CUDA events on the launching stream: record into slot 0..7, elapsed between two recorded slots */
int lgpu_timer_record(lgpu_ctx *ctx, int slot);
int lgpu_timer_elapsed_ms(lgpu_ctx *ctx, int slot_a, int slot_b, double *ms);
/* per-launch device time by kernel class (CUDA events around every launch while enabled) */
int lgpu_profile_enable(lgpu_ctx *ctx, int on);
int lgpu_profile_read(lgpu_ctx *ctx, int ncls, double *ms, int64_t *count);
int lgpu_profile_num_classes(void);
const char *lgpu_profile_class_name(int cls);

/* ---- problem upload (once) -------------------------------------------------------------------
 * Replaces LORADSInitSolver/LORADSSetDualObjective/LORADSInitConeData/LORADSPreprocess
 * (main.c:395-403; AConeProcData + AConePresolveData, lorads_sdp_conic.c:1046,1185-1393).
 * A cone is given exactly as the reference reader produces it (LReadSDPA, lorads_file_io.c:59):
 * CSC over PACKED lower-triangular indices, column 0 = objective (already negated), columns
 * 1..m = A_1..A_m; mat_beg has m+2 entries.  Indices are 64-bit (the reference's INT32 build
 * overflows PACK_IDX for n > 46340).  The library applies the reference's storage rules
 * (zero/sparse/dense coefficient, sparse/dense aggregate, sparse/dense container) and builds the
 * device layout: aggregated pattern, CSR-by-constraint, slot-transposed CSR, full symmetric CSR. */
int lgpu_set_problem(lgpu_ctx *ctx, int64_t m, const double *b, int n_cones, const int64_t *blk_dims,
                     int64_t n_lp_cols);
int lgpu_cone_upload(lgpu_ctx *ctx, int cone, const int64_t *mat_beg, const int64_t *mat_idx,
                     const double *mat_elem);
/* LP block (LORADSSetLpCone, lorads_lp_conic.c:347-365): CSC with m+1 columns over LP column ids. */
int lgpu_lp_upload(lgpu_ctx *ctx, const int64_t *lp_beg, const int64_t *lp_idx, const double *lp_elem);
/* per-cone facts the host rank heuristic needs (LORADSDetermineRank, lorads_solver.c:406-459):
 * out[0] = #nonzero A_i, out[1] = aggregate is dense, out[2] = SPARSE_CONE container, out[3] = nnzP,
 * out[4] = all constraints are single diagonal entries (MaxCut-type fast path), out[5] = nnzA */
int lgpu_cone_info(const lgpu_ctx *ctx, int cone, int64_t out[6]);
/* the same six facts from the reader's arrays alone: pure host code, needs neither a context nor a GPU */
int lgpu_cone_classify(int64_t n, int64_t m, const int64_t *mat_beg, const int64_t *mat_idx, int64_t out[6]);
/* The device layout of one cone as lgpu_cone_upload would build it for rank `rank` of `world` (world = 1: the
 * single-GPU layout), again pure host code: build once, read arrays by name, free.  `name` is the array
 * (csrc/lgpu_layout.h: pat_row, pat_col, cval, c_slot, c_coef, a_ptr, a_slot, a_coef, con_gid, t_ptr, t_loc, t_gid,
 * t_val, f_ptr, f_col, f_slot, d_row, d_val, mc_val, rc_ptr, rc_gid, rc_a, and for world > 1 lf_ptr, lf_col, lmc_val,
 * lrc_ptr, lrc_gid, lrc_a, send_idx, halo_gid, send_off, send_cnt, recv_off, recv_cnt) or "scalars" (16 doubles: mA
 * nnzP nnzA nnzC nnzF max_con_len max_slot_len |C|_1 |C|_2^2 |C|_inf dense sparse_container diag_only use_halo
 * halo_rows send_rows).  *count receives the number of elements, *elem_bytes 4 (int32), 8 (double) or -8 (int64); at
 * most cap_bytes are written to out.  An inspection/test entry: the replacement for walking the reference's
 * sdp_coeff_sparse / lorads_cone_sdp_* structs (def_lorads_sdp_data.h:100-117, def_lorads_sdp_conic.h:131-155). */
typedef struct lgpu_layout lgpu_layout;
int lgpu_cone_layout_build(lgpu_layout **out, int64_t n, int64_t m, const int64_t *mat_beg, const int64_t *mat_idx,
                           const double *mat_elem, int world, int rank);
int lgpu_cone_layout_get(const lgpu_layout *layout, const char *name, int64_t cap_bytes, void *out, int64_t *count,
                         int *elem_bytes);
void lgpu_cone_layout_free(lgpu_layout *layout);
/* Row relabelling for gather locality: when the fused MaxCut-type layout is used and a breadth-first relabelling of the
 * graph of C brings the entries of the symmetric CSR near the diagonal (the factor rows a row gathers then sit in L2, the
 * halos of a row-block partition become thin), the library stores the factors in that order.  Nothing changes at this
 * interface: factors are passed in and returned in the caller's row order.  out = {applied, share of CSR entries within
 * +-65536 rows of the diagonal before, after}.  LORADS_REORDER=0/1 forbids / forces it.  The reference has no counterpart
 * (it walks its entries in file order, lorads_sdp_data.c:750-763). */
int lgpu_cone_reorder_info(const lgpu_ctx *ctx, int cone, double out[3]);

/* aggregated lower pattern of a cone in the reference's order (sorted by (col,row)) */
int lgpu_cone_pattern(const lgpu_ctx *ctx, int cone, int64_t cap, int32_t *row, int32_t *col);
/* cal_sdp_const (lorads_solver.c:1457-1485): out = {|C|_1, |C|_2, |C|_inf, |b|_1, |b|_2, |b|_inf(Q2 quirk)} */
int lgpu_constants(const lgpu_ctx *ctx, double out[6]);
/* objScale_dualvar (lorads_solver.c:1438-1450): C *= s (all cones + LP), lambda *= s */
int lgpu_obj_scale(lgpu_ctx *ctx, double s);

/* ---- variables -------------------------------------------------------------------------------
 * LORADSInitALMVars / LORADSInitADMMVars storage (lorads_solver.c:616-708,851-946): R, Grad, U(=D), V,
 * M2temp, bLinSys, CG work vectors, L-BFGS ring of `lbfgs_len` (s,y) pairs, all m-vectors.  Values
 * are zero until set. */
int lgpu_alloc_vars(lgpu_ctx *ctx, const int64_t *rank, int lbfgs_len);
/* When the problem is one SDP block whose constraints are all single diagonal entries (MaxCut-type: diag(X) = b)
 * and there is no LP block, the library runs a fused path: A(UV^T) is row-local, <C, UV^T> = <U, C V>, and
 * (C + A^*(w)) X = C X + Diag(.) X, so an ALM inner iteration costs ONE sparse product.  Results agree with the
 * general path to rounding.  This switch (default 1) exists for A/B parity tests; it takes effect at the next
 * lgpu_alloc_vars / lgpu_aug_rank.  lgpu_uses_fused_path reports what the current variables use. */
int lgpu_set_fused_path(lgpu_ctx *ctx, int on);
int lgpu_uses_fused_path(const lgpu_ctx *ctx);
/* On the fused path with the default history length 2 the L-BFGS scalars (alpha, beta <y,q>, <D,Grad>) are formed
 * from inner products carried from the step pass instead of by the recursion's own four passes; the direction then
 * costs one pass.  Same operations on the same numbers up to the rounding of those inner products.  Switch
 * (default 1) for A/B parity tests. */
int lgpu_set_carried_dots(lgpu_ctx *ctx, int on);
/* Cones whose aggregate is dense (the reference's SDP_COEFF_DENSE aggregate: n < 20, a dense member, or >= 10 % fill,
 * lorads_sdp_conic.c:1185-1393) run LORADSUVt's dsyr2k and mul_rk's dsymm (lorads_alg_common.c:72-89,
 * lorads_sdp_data.c:948-973) as DMMA (FP64 tensor core) kernels over the packed triangle.  Switch (default 1) for A/B
 * tests against the pattern-gather kernels. */
int lgpu_set_dense_tensor_path(lgpu_ctx *ctx, int on);
int lgpu_set_factor(lgpu_ctx *ctx, int which, int cone, const double *colmajor);
int lgpu_get_factor(lgpu_ctx *ctx, int which, int cone, double *colmajor);
int lgpu_set_lp(lgpu_ctx *ctx, int which, const double *v);
int lgpu_get_lp(lgpu_ctx *ctx, int which, double *v);
int lgpu_set_vec(lgpu_ctx *ctx, int which, const double *v);
int lgpu_get_vec(lgpu_ctx *ctx, int which, double *v);
int lgpu_get_rank(const lgpu_ctx *ctx, int cone, int64_t *rank);
/* device-side pseudo-random initial point for synthetic benchmarks (not the reference's rand()) */
int lgpu_fill_factor_random(lgpu_ctx *ctx, int which, uint64_t seed);
/* AUG_RANK (lorads_solver.c:1154-1254): grow every factor to new_rank[c] columns; new column j of
 * R, U, V, Grad gets 1/sqrt(delta r) at row j; M2/bLinSys/CG/L-BFGS storage re-created zeroed. */
int lgpu_aug_rank(lgpu_ctx *ctx, const int64_t *new_rank);

/* ---- lorads_func mirror on device-resident state ----------------------------------------------*/
/* InitConstrValAll(+LP) then InitConstrValSum (lorads_alg_common.c:116-122,221-247) on a factor pair */
int lgpu_init_constr_val(lgpu_ctx *ctx, int pair);
/* ALMCalGrad[LP] (lorads_alm.c:32-130): M1 = -lambda - rho b + rho constrValSum; Grad_c = 2 (C + A*(M1)) R_c */
int lgpu_alm_cal_grad(lgpu_ctx *ctx, double rho, double *lag_norm_square);
/* LBFGSDirection[LP] + LBFGSDirUseGrad[LP] (lorads_alm.c:347-705): D (stored in U) from Grad and the ring */
int lgpu_lbfgs_direction(lgpu_ctx *ctx, int64_t inner_iter);
/* q0 = b - constrValSum; ALMCalq12p12[LP] (lorads_alm.c:714-761); then the five reductions of
 * ALMLineSearch (lorads_alm.c:266-279) with q0' = q0 + lambda/rho.
 * out = { p1, p2, |q2|^2, q1.q2, q0'.q2, |q1|^2, q0'.q1 } */
int lgpu_alm_linesearch_terms(lgpu_ctx *ctx, double rho, double out[7]);
/* setAsNegGrad; ALMupdateVar (R += tau D); constrValSum += tau q1 + tau^2 q2 (lorads_alm.c:1342-1353) */
int lgpu_alm_step(lgpu_ctx *ctx, double tau);
/* the four calls above + below in one: setAsNegGrad, ALMupdateVar, constrValSum update, ALMCalGrad, setlbfgsHisTwo,
 * updateDimacsALM (lorads_alm.c:1342-1357) -- what the host loop calls after the line search.  Returns
 * sum |Grad|^2 and |b - A(RR^T)|_2 / (1 + |b|_1). */
int lgpu_alm_inner_update(lgpu_ctx *ctx, double rho, double tau, double *lag_norm_square, double *pinf_l1);
/* setlbfgsHisTwo (lorads_alm.c:842-890): s = tau D, y += Grad, beta = 1/<y,s>, advance the ring */
int lgpu_lbfgs_push(lgpu_ctx *ctx, double tau);
/* updateDimacsALM -> primalInfeasibility[LP] (lorads_alg_common.c:386-407): recompute A(RR^T) and
 * constrValSum from scratch, return |b - A|_2 / (1 + |b|_1) */
int lgpu_primal_infeasibility(lgpu_ctx *ctx, int pair, double *pinf_l1);
/* LORADSUpdateDualVar (lorads_alg_common.c:511-524) */
int lgpu_update_dual_var(lgpu_ctx *ctx, double rho);
/* calObj_alm / calObj_admm (lorads_alm.c:1488-1510, lorads_admm.c:398-430): <C, RR^T> (admm: R=(U+V)/2 first),
 * NOT divided by scaleObjHis; LORADSCalDualObj (lorads_alg_common.c:531-537): b^T lambda, same */
int lgpu_cal_obj(lgpu_ctx *ctx, int admm, double *pobj);
int lgpu_cal_dual_obj(lgpu_ctx *ctx, double *dobj);
/* LORADS_ALMtoADMM copies (lorads_solver.c:1351-1366): V <- R, U <- V ; averageUV (lorads_admm.c:372-377);
 * copyRtoV (lorads_alg_common.c:260-268) */
int lgpu_alm_to_admm(lgpu_ctx *ctx);
int lgpu_average_uv(lgpu_ctx *ctx);
int lgpu_copy_r_to_v(lgpu_ctx *ctx);
/* admmUpdateVar = LORADSUpdateSDPVar[+LP] (lorads_alg_common.c:298-376): Gauss-Seidel sweep, per cone U then V
 * by CGSolve (lorads_cgs.c:128-287) on x -> x + (sum_i <A_i, sym(x V^T)> A_i) V.  cg_iter_total is the
 * running ASolver->cgIter counter (in/out). */
int lgpu_admm_update_var(lgpu_ctx *ctx, double rho, double cg_tol, int64_t cg_max_iter, int64_t *cg_iter_total);
/* r x r Gram matrices for the oracle rank (lorads_logging.c:216-270): phase 1 -> R^T R, phase 2 ->
 * ((U+V)/2)^T (U+V)/2, row-major r x r into gram */
int lgpu_gram(lgpu_ctx *ctx, int phase, int cone, double *gram);
/* calculate_dual_infeasibility_solver (lorads_solver.c:1396-1426): sum over cones of |min(lambda_min(C - A*(lambda)),0)|
 * (+ LP part), NOT yet divided by scaleObjHis and (1+|C|_1).  Lanczos on the device replaces ARPACK. */
int lgpu_dual_infeasibility(lgpu_ctx *ctx, double *sum_neg_eig);

/* ---- operator-level entry points on HOST buffers (cone / sdp_coeff vtable drop-ins) ------------
 * These copy their inputs to the device, run the same kernels as above, and copy the result back;
 * parity tests and bench.py's e2e leg call them. */
/* LORADSUVt + coneAUV + objAUV (lorads_alg_common.c:43-90,153-158): constr_val is the dense m-vector */
int lgpu_op_auv(lgpu_ctx *ctx, int cone, int64_t r, const double *U, const double *V, double *constr_val,
                double *obj);
/* LORADSUVt alone: pattern samples of (UV^T + VU^T)/2, nnzP values */
int lgpu_op_uvt(lgpu_ctx *ctx, int cone, int64_t r, const double *U, const double *V, double *uvt);
/* zeros + [addObjCoeff] + sdpDataWSum (lorads_sdp_conic.c:448-460,608-616): nnzP values */
int lgpu_op_wsum(lgpu_ctx *ctx, int cone, const double *w, int add_obj, double *S);
/* the same followed by mul_rk (lorads_sdp_data.c:750-763,948-973): Y = (C? + A*(w)) X */
int lgpu_op_wsum_mulrk(lgpu_ctx *ctx, int cone, int64_t r, const double *w, int add_obj, const double *X,
                       double *Y);

/* ---- multi-GPU: row-block partition over the GPUs of one node, one process per GPU -----------
 * Rank p owns rows [lo, hi) of every factor-shaped array, the CSR rows of C and the constraints attached to those
 * rows.  Exchange steps: one NCCL all-gather of the direction's rows before the sparse product, and NCCL all-reduces
 * of the scalar packs (L-BFGS dots, the seven line-search terms, |Grad|^2 / <y,s> / |b - A|^2, CG dots).  This build
 * partitions the fused MaxCut-type layout (one SDP block, single-diagonal-entry constraints).
 * Order of calls: lgpu_create, lgpu_comm_init, lgpu_set_problem, lgpu_cone_upload (every rank passes the WHOLE
 * problem and keeps its slice), lgpu_alloc_vars, ... ; host factors/vectors passed IN are always whole (a rank reads
 * only its rows from them); lgpu_get_factor writes only the rows the rank owns into the caller's whole-size array;
 * lgpu_get_vec returns the whole m-vector on every rank. */
int lgpu_partition_rows(int64_t n, int world, int rank, int64_t *lo, int64_t *hi, int64_t *rows_per_rank);
/* ncclUniqueId is 128 bytes; rank 0 calls lgpu_nccl_unique_id and the host side distributes it to the other ranks */
int lgpu_nccl_unique_id(unsigned char id[128]);
int lgpu_comm_init(lgpu_ctx *ctx, const unsigned char id[128], int rank, int world);
/* Collective yes/no decision for anything a rank derives from its OWN clock or environment (the reference's
 * wall-clock tests against timeSecLimit, lorads_alm.c:1441, lorads_admm.c:170-173, main.c:471,519,583): on return
 * *flag is rank 0's value on every rank (rank 0 is authoritative), so all ranks leave a loop in the same iteration
 * and nobody is left waiting inside the next collective.  One GPU: *flag is returned unchanged. */
int lgpu_agree_flag(lgpu_ctx *ctx, int *flag);
/* 1 when the ranks exchange halo rows and scalar packs through peer-mapped memory (NVLink stores from our own kernels,
 * CUDA IPC), 0 when they use ncclSend/ncclRecv/ncclAllReduce (LORADS_PEER=0, no peer access, or two ranks on one GPU) */
int lgpu_uses_peer_exchange(const lgpu_ctx *ctx);
/* By-cone partition (several cones / LP block / general constraints on several GPUs): cones couple only through the
 * length-m constraint vector and scalars (LORADSInitConstrValSum, lorads_alg_common.c:221-229), so each cone's operator
 * work runs on one owner.  The map every rank derives: largest cost first onto the least loaded rank.  Pure function. */
int lgpu_cone_owner_map(int ncones, const double *cost, int world, int *owner);

#ifdef __cplusplus
}
#endif
#endif /* LORADS_B200_H */
