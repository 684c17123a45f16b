/*
 * lgpu_kernels.cuh -- hand-written sm_100a kernels of the LoRADS inner loop.
 *
 * Every kernel here is HBM-bandwidth bound FP64 work (<= 0.25 flop/byte): gathers over row-major
 * factor rows, CSR walks, flat vector algebra and deterministic reductions.  The rules that matter
 * are coalescing (a sub-warp group of G lanes owns one factor row and reads it as consecutive
 * 16-byte double2 words), enough loads in flight (grid-stride loops, grids sized in multiples of the
 * 148 SMs) and no atomics on the data path (reductions are two-stage with a fixed summation order,
 * so results are bit-reproducible run to run).
 *
 * Reference functions restated by each kernel are cited at the kernel.
 */
#ifndef LGPU_KERNELS_CUH
#define LGPU_KERNELS_CUH

#include "lgpu_internal.h"

#define LGPU_TPB 256
#define LGPU_LONG_ROW 96    /* CSR rows longer than this leave the row-per-group kernels ...                  */
#define LGPU_LONG_CHUNK 512 /* ... and are processed in chunks of this many entries, one CTA per chunk        */

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* streamed-once data: evict-first loads / stores keep the L2 for what is actually reused (the gathered factor rows) */
template <bool CS>
__device__ __forceinline__ double2 ldv(const double *p, size_t w)
{
    return CS ? __ldcs(reinterpret_cast<const double2 *>(p) + w) : reinterpret_cast<const double2 *>(p)[w];
}
template <bool CS>
__device__ __forceinline__ void stv(double *p, size_t w, double2 v)
{
    if (CS) __stcs(reinterpret_cast<double2 *>(p) + w, v);
    else reinterpret_cast<double2 *>(p)[w] = v;
}

template <int K>
struct SlotSpec {
    int slot[K];
    int accumulate; /* 1: dsc[slot] += result, 0: dsc[slot] = result */
};

/* Deterministic grid reduction tail: block-sum K values, publish per-block partials, the LAST block
 * to arrive sums the partials in a fixed (thread-strided, then tree) order and writes dsc[slot]. */
struct NoPost {
    __device__ void operator()(double *) const {}
};

/* `post(dsc)` runs once, in the finishing thread, after the slots are written: derived scalars (an L-BFGS alpha,
 * 1/<y,s>, ...) are produced without a separate one-thread launch. */
/* SYS: the CTA's earlier stores include stores to PEER memory (NVLink) that `post` is about to publish with a flag:
 * the fence ahead of the arrival counter is then system-wide. */
template <int K, class P = NoPost, bool SYS = false>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[K], double *partials, unsigned int *counter,
                                                   double *dsc, const SlotSpec<K> &spec, P post = P())
{
    __shared__ double sh[K][LGPU_TPB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double w = warp_sum(v[k]);
        if (lane == 0) sh[k][wid] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = 0.0;
            for (int w = 0; w < LGPU_TPB / 32; ++w) t += sh[k][w];
            partials[(size_t)k * gridDim.x + blockIdx.x] = t;
        }
        if (SYS) __threadfence_system();
        else __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += LGPU_TPB)
            t += ((volatile double *)partials)[(size_t)k * gridDim.x + i];
        t = warp_sum(t);
        __syncthreads();
        if (lane == 0) sh[k][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = 0.0;
            for (int w = 0; w < LGPU_TPB / 32; ++w) t += sh[k][w];
            if (spec.accumulate) dsc[spec.slot[k]] += t;
            else dsc[spec.slot[k]] = t;
        }
        post(dsc);
        *counter = 0u;
    }
}

/* generic flat kernels: f is a __device__ lambda */
template <class F>
__global__ void __launch_bounds__(LGPU_TPB) k_map(int64_t n, F f)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <int K, class F, class P>
__global__ void __launch_bounds__(LGPU_TPB) k_reduce(int64_t n, F f, double *partials, unsigned int *counter,
                                                     double *dsc, SlotSpec<K> spec, P post)
{
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, acc);
    grid_reduce_finish<K, P>(acc, partials, counter, dsc, spec, post);
}

/* D = -G where the device scalar dsc[slot] >= 0 (LBFGSDirectionUseGrad, lorads_alm.c:607-627); the test is made once
 * per thread, so the common case (descent direction found) touches no memory */
__global__ void __launch_bounds__(LGPU_TPB) k_neg_if_nonneg(int64_t n, const double *__restrict__ dsc, int slot,
                                                            const double *__restrict__ G, double *__restrict__ D)
{
    if (!(dsc[slot] >= 0.0)) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) D[i] = -G[i];
}

template <class F>
__global__ void k_scalar(F f)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) f();
}

/* ------------------------------------------------------------------------------------------------
 * K1  pattern samples of (U V^T + V U^T)/2          reference: LORADSUVt, lorads_alg_common.c:43-90
 * One group of G lanes per pattern entry (row i >= col j); each lane reads double2 words of the two
 * (or four) factor rows, group-shuffle reduction, one 8-byte store per entry.
 * ------------------------------------------------------------------------------------------------*/
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_uvt(int64_t nnzP, const int32_t *__restrict__ prow,
                                                  const int32_t *__restrict__ pcol, const double *__restrict__ U,
                                                  const double *__restrict__ V, int ld, int same,
                                                  double *__restrict__ out)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    /* all lanes of a warp run the same number of iterations so the shuffles stay convergent */
    const int64_t iters = (nnzP + groups - 1) / groups;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t k = g0 + it * groups;
        const bool live = k < nnzP;
        double a1 = 0.0, a2 = 0.0;
        int i = 0, j = 0;
        if (live) {
            i = prow[k];
            j = pcol[k];
            const double2 *Ui = reinterpret_cast<const double2 *>(U + (size_t)i * ld);
            const double2 *Vj = reinterpret_cast<const double2 *>(V + (size_t)j * ld);
            if (i == j || same) {
                for (int c = lane; c < ld2; c += G) {
                    double2 u = Ui[c], v = Vj[c];
                    a1 = fma(u.x, v.x, a1);
                    a1 = fma(u.y, v.y, a1);
                }
                a2 = a1;
            } else {
                const double2 *Uj = reinterpret_cast<const double2 *>(U + (size_t)j * ld);
                const double2 *Vi = reinterpret_cast<const double2 *>(V + (size_t)i * ld);
                for (int c = lane; c < ld2; c += G) {
                    double2 u = Ui[c], v = Vj[c], uu = Uj[c], vv = Vi[c];
                    a1 = fma(u.x, v.x, a1);
                    a1 = fma(u.y, v.y, a1);
                    a2 = fma(uu.x, vv.x, a2);
                    a2 = fma(uu.y, vv.y, a2);
                }
            }
        }
        a1 = group_sum<G>(a1);
        a2 = group_sum<G>(a2);
        if (live && lane == 0) out[k] = (i == j) ? a1 : (0.5 * a1 + 0.5 * a2);
    }
}

/* ================================================================================================
 * Dense-aggregate cones (theta / matrix-completion type: the pattern is the whole lower triangle, slot = packed
 * column-major index).  Here the two heavy operators ARE matrix products, so they run on the FP64 tensor pipe:
 * DMMA m8n8k4 (mma.sync ... f64), fragments straight from the row-major factors.
 * fragment layout (PTX ISA, m8n8k4 .f64): g = lane >> 2, t = lane & 3
 *   A (8x4, row):  a  = A[g][t]          B (4x8, col):  b = B[t][g]          C (8x8):  c0 = C[g][2t], c1 = C[g][2t+1]
 * ================================================================================================*/
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

/* K1d  W = (U V^T + V U^T)/2 on the packed lower triangle         reference: LORADSUVt dense branch -> fds_syr2k
 *      (dsyr2k, alpha = 1/2), lorads_alg_common.c:72-89, lorads_dense_opts.c:773-783
 * One warp per 16 x 16 tile (I >= J) of W, 2 x 2 DMMA tiles, k in steps of 4 over the padded rank. */
__global__ void __launch_bounds__(128) k_dense_uvt(int64_t n, int ld, const double *__restrict__ U,
                                                   const double *__restrict__ V, int same, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t nt = (n + 15) / 16;
    const int64_t ntiles = nt * (nt + 1) / 2;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < ntiles; w += warps) {
        /* w -> (I, J), I >= J, tiles of row I start at I (I + 1) / 2 */
        int64_t I = (int64_t)((sqrt(8.0 * (double)w + 1.0) - 1.0) * 0.5);
        while (I * (I + 1) / 2 > w) --I;
        while ((I + 1) * (I + 2) / 2 <= w) ++I;
        const int64_t J = w - I * (I + 1) / 2;
        const int64_t i0 = I * 16, j0 = J * 16;
        double c[2][2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) c[a][b][0] = c[a][b][1] = 0.0;
        for (int k0 = 0; k0 < ld; k0 += 4) {
            double ui[2], vi[2], uj[2], vj[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int64_t ri = i0 + a * 8 + g, rj = j0 + a * 8 + g;
                ui[a] = ri < n ? U[(size_t)ri * ld + k0 + t] : 0.0;
                uj[a] = rj < n ? U[(size_t)rj * ld + k0 + t] : 0.0;
                vi[a] = same ? ui[a] : (ri < n ? V[(size_t)ri * ld + k0 + t] : 0.0);
                vj[a] = same ? uj[a] : (rj < n ? V[(size_t)rj * ld + k0 + t] : 0.0);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    dmma_m8n8k4(c[a][b][0], c[a][b][1], ui[a], vj[b]); /* U_i V_j^T */
                    dmma_m8n8k4(c[a][b][0], c[a][b][1], vi[a], uj[b]); /* V_i U_j^T */
                }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t i = i0 + a * 8 + g, j = j0 + b * 8 + 2 * t + e;
                    if (i < n && j <= i) out[(2 * n - j - 1) * j / 2 + i] = 0.5 * c[a][b][e];
                }
    }
}

/* K4d  Y = alpha S X + beta Z, S dense symmetric given on the packed lower triangle (never unpacked to n x n)
 *      reference: dataMatDenseMultiRkMat (unpack + dsymm on every call), lorads_sdp_data.c:948-973
 * One warp per (8-row tile, 8-column tile, k-range): the block dimension n is small (hundreds to a few thousand), so
 * the parallelism has to come from the column tiles and from splitting the k loop -- the first version (one warp per
 * 8 rows x 64 columns, whole k loop) ran 94 CTAs at 6 % warps active on n = 3000.  Column tiles vary fastest, so the
 * warps of a CTA share the S fragment through L1.  ksplit > 1: raw partial sums go to scratch[ks][n][ld] and
 * k_dense_symm_finish adds them in k order (fixed order) and applies alpha / beta. */
__global__ void __launch_bounds__(128) k_dense_symm(int64_t n, int ld, int ksplit, int64_t klen, const double *__restrict__ S,
                                                    const double *__restrict__ X, double alpha, double beta,
                                                    const double *__restrict__ Z, double *__restrict__ Y,
                                                    double *__restrict__ scratch)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t rt = (n + 7) / 8;
    const int ct = (ld + 7) / 8;
    const int64_t ntiles = rt * ct * ksplit;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < ntiles; w += warps) {
        const int c0 = (int)(w % ct) * 8;
        const int64_t rest = w / ct;
        const int64_t i0 = (rest % rt) * 8;
        const int ks = (int)(rest / rt);
        const int64_t kbeg = (int64_t)ks * klen, kend = (kbeg + klen < n) ? kbeg + klen : n;
        const int64_t i = i0 + g;
        const int colb = c0 + g;
        double acc0 = 0.0, acc1 = 0.0;
        for (int64_t k0 = kbeg; k0 < kend; k0 += 4) {
            const int64_t k = k0 + t;
            double a = 0.0, b = 0.0;
            if (k < kend) {
                if (i < n) a = (i >= k) ? S[(2 * n - k - 1) * k / 2 + i] : S[(2 * n - i - 1) * i / 2 + k];
                if (colb < ld) b = X[(size_t)k * ld + colb];
            }
            dmma_m8n8k4(acc0, acc1, a, b);
        }
        if (i < n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = c0 + 2 * t + e;
                if (col < ld) {
                    const double v = e ? acc1 : acc0;
                    if (ksplit > 1) {
                        scratch[((size_t)ks * n + i) * ld + col] = v;
                    } else {
                        double o = alpha * v;
                        if (Z != nullptr) o = fma(beta, Z[(size_t)i * ld + col], o);
                        Y[(size_t)i * ld + col] = o;
                    }
                }
            }
        }
    }
}
__global__ void __launch_bounds__(LGPU_TPB) k_dense_symm_finish(int64_t total, int64_t nld, int ksplit,
                                                                const double *__restrict__ scratch, double alpha, double beta,
                                                                const double *__restrict__ Z, double *__restrict__ Y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        double tsum = 0.0;
        for (int ks = 0; ks < ksplit; ++ks) tsum += scratch[(size_t)ks * nld + q];
        double o = alpha * tsum;
        if (Z != nullptr) o = fma(beta, Z[q], o);
        Y[q] = o;
    }
}

/* ------------------------------------------------------------------------------------------------
 * K2  per-constraint gather  cv[t] = sum_e coef[e] * uvt[slot[e]]
 *     reference: sdpDenseConeAUVImpl / sdpSparseConeAUVImpl -> sparseAUV / denseAUV,
 *     lorads_sdp_conic.c:378-385,681-688 ; lorads_sdp_data.c:803-876,1017-1034
 *     coef = 2a off the diagonal, a on it (the reference's "2 a UVt, halved on the diagonal").
 * ------------------------------------------------------------------------------------------------*/
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_con_gather(int64_t mA, const int32_t *__restrict__ aptr,
                                                         const int32_t *__restrict__ aslot,
                                                         const double *__restrict__ acoef,
                                                         const double *__restrict__ uvt, double *__restrict__ cv)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t iters = (mA + groups - 1) / groups;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t t = g0 + it * groups;
        const bool live = t < mA;
        double a = 0.0;
        bool mine = live;
        if (live) {
            const int e0 = aptr[t], e1 = aptr[t + 1];
            if (e1 - e0 > LGPU_LONG_ROW * G) mine = false; /* a long list among short ones: k_con_gather_long */
            else
                for (int e = e0 + lane; e < e1; e += G) a = fma(acoef[e], uvt[aslot[e]], a);
        }
        a = group_sum<G>(a);
        if (mine && lane == 0) cv[t] = a;
    }
}

/* constraints whose entry list is much longer than the rest (e.g. the trace constraint A = I of a theta problem among
 * 37 466 single-entry ones): one CTA per constraint, fixed-order block sum.  A single thread walking such a list cost
 * 109 us per launch on theta102. */
__global__ void __launch_bounds__(LGPU_TPB) k_con_gather_long(int64_t nlong, const int32_t *__restrict__ which,
                                                              const int32_t *__restrict__ aptr, const int32_t *__restrict__ aslot,
                                                              const double *__restrict__ acoef, const double *__restrict__ uvt,
                                                              double *__restrict__ cv)
{
    __shared__ double sh[LGPU_TPB / 32];
    for (int64_t k = blockIdx.x; k < nlong; k += gridDim.x) {
        const int t = which[k];
        double a = 0.0;
        for (int e = aptr[t] + threadIdx.x; e < aptr[t + 1]; e += LGPU_TPB) a = fma(acoef[e], uvt[aslot[e]], a);
        a = warp_sum(a);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < LGPU_TPB / 32; ++w) s += sh[w];
            cv[t] = s;
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------------
 * K3  S = [C] + sum_i w_i A_i on the pattern, as a GATHER over the slot-transposed CSR
 *     reference: zeros + addObjCoeff + sdp*DataWeightSumImpl, lorads_sdp_conic.c:448-460,608-616,894-902
 *     w is indexed by global constraint id (use_gid) or by the cone's compact id.
 * ------------------------------------------------------------------------------------------------*/
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_wsum(int64_t nnzP, const int32_t *__restrict__ tptr,
                                                   const int32_t *__restrict__ tidx, const double *__restrict__ tval,
                                                   const double *__restrict__ w, const double *__restrict__ cval,
                                                   int add_obj, double wscale, double *__restrict__ S)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t iters = (nnzP + groups - 1) / groups;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t k = g0 + it * groups;
        const bool live = k < nnzP;
        double a = 0.0;
        if (live) {
            const int e0 = tptr[k], e1 = tptr[k + 1];
            for (int e = e0 + lane; e < e1; e += G) a = fma(wscale * w[tidx[e]], tval[e], a);
        }
        a = group_sum<G>(a);
        if (live && lane == 0) S[k] = add_obj ? (cval[k] + a) : a;
    }
}

/* ------------------------------------------------------------------------------------------------
 * K4  Y = alpha * S X + beta * Z   over the full symmetric CSR, S given on pattern slots
 *     reference: dataMatSparseMultiRkMat / dataMatDenseMultiRkMat, lorads_sdp_data.c:750-763,948-973
 *     One group of G lanes per row; each lane owns double2 column words c = lane + G*t.  Rows of X
 *     are read as contiguous 16-byte words (row-major, ld multiple of 4).  With diag != nullptr the
 *     operator is (S + Diag(diag)) X  (MaxCut-type A*(w) folded in without touching S).
 * ------------------------------------------------------------------------------------------------*/
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_spmm(int64_t n, const int32_t *__restrict__ fptr,
                                                   const int32_t *__restrict__ fcol, const int32_t *__restrict__ fslot,
                                                   const double *__restrict__ S, const double *__restrict__ X, int ld,
                                                   double alpha, double beta, const double *__restrict__ Z,
                                                   double *__restrict__ Y)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    for (int64_t i = g0; i < n; i += groups) {
        const int e0 = fptr[i], e1 = fptr[i + 1];
        if (e1 - e0 > LGPU_LONG_ROW) continue; /* -> k_spmm_long_chunks */
        for (int cb = 0; cb < ld2; cb += G) {
            const int c = cb + lane;
            double2 acc = make_double2(0.0, 0.0);
            if (c < ld2) {
                int e = e0;
                /* two entries per trip: more independent loads in flight */
                for (; e + 1 < e1; e += 2) {
                    const double s0 = S[fslot[e]], s1 = S[fslot[e + 1]];
                    const double2 x0 = reinterpret_cast<const double2 *>(X + (size_t)fcol[e] * ld)[c];
                    const double2 x1 = reinterpret_cast<const double2 *>(X + (size_t)fcol[e + 1] * ld)[c];
                    acc.x = fma(s0, x0.x, acc.x);
                    acc.y = fma(s0, x0.y, acc.y);
                    acc.x = fma(s1, x1.x, acc.x);
                    acc.y = fma(s1, x1.y, acc.y);
                }
                if (e < e1) {
                    const double s0 = S[fslot[e]];
                    const double2 x0 = reinterpret_cast<const double2 *>(X + (size_t)fcol[e] * ld)[c];
                    acc.x = fma(s0, x0.x, acc.x);
                    acc.y = fma(s0, x0.y, acc.y);
                }
                double2 o = make_double2(alpha * acc.x, alpha * acc.y);
                if (Z != nullptr) {
                    const double2 z = reinterpret_cast<const double2 *>(Z + (size_t)i * ld)[c];
                    o.x = fma(beta, z.x, o.x);
                    o.y = fma(beta, z.y, o.y);
                }
                reinterpret_cast<double2 *>(Y + (size_t)i * ld)[c] = o;
            }
        }
    }
}

/* host column-major (cm_ld rows, this rank's rows start at cm_row0)  <->  device row-major n x ld (padding columns
 * zeroed) */
/* perm != nullptr (row relabelling for gather locality, lgpu_layout.h): the caller's row `row` lives in device row perm[row] */
__global__ void __launch_bounds__(LGPU_TPB) k_col2row(int64_t n, int r, int ld, const double *__restrict__ cm,
                                                      int64_t cm_ld, int64_t cm_row0, double *__restrict__ rm,
                                                      const int32_t *__restrict__ perm)
{
    __shared__ double tile[32][33];
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    for (int c0 = 0; c0 < ld; c0 += 32) {
        /* read: threads along rows (contiguous in column-major) */
        for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
            const int64_t row = row0 + threadIdx.x;
            const int col = c0 + cc;
            tile[cc][threadIdx.x] = (row < n && col < r) ? cm[(size_t)col * cm_ld + cm_row0 + row] : 0.0;
        }
        __syncthreads();
        for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
            const int64_t row = row0 + rr;
            const int col = c0 + threadIdx.x;
            if (row < n && col < ld) rm[(size_t)(perm ? perm[row] : row) * ld + col] = tile[threadIdx.x][rr];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(LGPU_TPB) k_row2col(int64_t n, int r, int ld, const double *__restrict__ rm,
                                                      double *__restrict__ cm, int64_t cm_ld, int64_t cm_row0,
                                                      const int32_t *__restrict__ perm)
{
    __shared__ double tile[32][33];
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    for (int c0 = 0; c0 < r; c0 += 32) {
        for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
            const int64_t row = row0 + rr;
            const int col = c0 + threadIdx.x;
            tile[rr][threadIdx.x] = (row < n && col < r) ? rm[(size_t)(perm ? perm[row] : row) * ld + col] : 0.0;
        }
        __syncthreads();
        for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
            const int64_t row = row0 + threadIdx.x;
            const int col = c0 + cc;
            if (row < n && col < r) cm[(size_t)col * cm_ld + cm_row0 + row] = tile[threadIdx.x][cc];
        }
        __syncthreads();
    }
}

/* partitioned run with relabelled rows: device row t of this rank is the caller's row src[t] of the whole column-major factor */
__global__ void __launch_bounds__(LGPU_TPB) k_gather_rows_cm(int64_t nloc, int r, int ld, const double *__restrict__ cm,
                                                             int64_t cm_ld, const int32_t *__restrict__ src, double *__restrict__ rm)
{
    const int64_t total = nloc * (int64_t)ld;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t t = i / ld;
        const int col = (int)(i - t * ld);
        rm[i] = col < r ? cm[(size_t)col * cm_ld + src[t]] : 0.0;
    }
}

/* re-stride a row-major factor from (r_old, ld_old) to ld_new and plant the AUG_RANK seed:
 * new column r_old + j gets 1/sqrt(dr) at row j (j < min(n, dr))      lorads_solver.c:1096-1106,1180-1214 */
__global__ void __launch_bounds__(LGPU_TPB) k_restride_aug(int64_t n, int r_old, int ld_old, int r_new, int ld_new,
                                                           const double *__restrict__ src, double *__restrict__ dst,
                                                           int plant, int64_t n_glob, int64_t row_lo,
                                                           const int32_t *__restrict__ iperm)
{
    /* n rows of this rank, global rows row_lo .. row_lo + n of a cone of dimension n_glob */
    const int64_t total = n * (int64_t)ld_new;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int dr = r_new - r_old;
    const double seed = dr > 0 ? 1.0 / sqrt((double)(n_glob < dr ? n_glob : dr)) : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t row = i / ld_new;
        const int col = (int)(i - row * ld_new);
        double v = 0.0;
        if (col < r_old) v = src[(size_t)row * ld_old + col];
        else if (plant && col < r_new && (int64_t)(col - r_old) == (iperm ? (int64_t)iperm[row + row_lo] : row + row_lo)) v = seed;
        dst[i] = v;
    }
}

/* ------------------------------------------------------------------------------------------------
 * K12 r x r Gram of a tall factor (oracle rank)      reference: build_gram_from_factor/_average,
 *     lorads_logging.c:216-270.  Tile (ti,tj) of 16x16 columns per block column, row chunks per block
 *     row; partial tiles go to `part` and are summed in fixed order by k_gram_finish.
 * ------------------------------------------------------------------------------------------------*/
__global__ void __launch_bounds__(256) k_gram_partial(int64_t n, int r, int ld, const double *__restrict__ A,
                                                      const double *__restrict__ B, int average, int64_t rows_per_chunk,
                                                      double *__restrict__ part)
{
    __shared__ double sa[64][17], sb[64][17];
    const int nt = (r + 15) / 16;
    const int ti = blockIdx.y / nt, tj = blockIdx.y % nt;
    const int a = threadIdx.x / 16, b2 = threadIdx.x % 16;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
    const int64_t r1 = (r0 + rows_per_chunk < n) ? r0 + rows_per_chunk : n;
    double acc = 0.0;
    for (int64_t base = r0; base < r1; base += 64) {
        for (int q = threadIdx.x; q < 64 * 16; q += 256) {
            const int rr = q / 16, cc = q % 16;
            const int64_t row = base + rr;
            const int ca = ti * 16 + cc, cb = tj * 16 + cc;
            double va = 0.0, vb = 0.0;
            if (row < r1) {
                if (ca < r) va = average ? 0.5 * (A[(size_t)row * ld + ca] + B[(size_t)row * ld + ca]) : A[(size_t)row * ld + ca];
                if (cb < r) vb = average ? 0.5 * (A[(size_t)row * ld + cb] + B[(size_t)row * ld + cb]) : A[(size_t)row * ld + cb];
            }
            sa[rr][cc] = va;
            sb[rr][cc] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < 64; ++rr) acc = fma(sa[rr][a], sb[rr][b2], acc);
        __syncthreads();
    }
    part[((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 256 + threadIdx.x] = acc;
}

/* K12 on the FP64 tensor pipe: one warp per (row chunk, 16 x 16 block of the Gram), 2 x 2 DMMA tiles, four factor rows per
 * step.  With M the factor (or (A + B) / 2) the fragments are A[i][k] = M[row0 + k][i0 + i], B[k][j] = M[row0 + k][j0 + j],
 * loaded straight from the row-major rows (8 consecutive doubles per row and fragment).  The warps of one chunk are
 * neighbours in the grid, so the chunk's rows come from DRAM once and from L2 afterwards.  Partial blocks go to `part` in
 * k_gram_partial's layout and k_gram_finish adds them in chunk order. */
__global__ void __launch_bounds__(128) k_gram_dmma(int64_t n, int r, int ld, const double *__restrict__ A,
                                                   const double *__restrict__ B, int average, int64_t rows_per_chunk,
                                                   int nchunks, double *__restrict__ part)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nt = (r + 15) / 16, nt2 = nt * nt;
    const int64_t units = (int64_t)nchunks * nt2;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < units; w += warps) {
        const int tile = (int)(w % nt2);
        const int64_t chunk = w / nt2;
        const int ti = tile / nt, tj = tile % nt;
        const int64_t r0 = chunk * rows_per_chunk;
        const int64_t r1 = (r0 + rows_per_chunk < n) ? r0 + rows_per_chunk : n;
        double c[2][2][2];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y) c[x][y][0] = c[x][y][1] = 0.0;
        for (int64_t k0 = r0; k0 < r1; k0 += 4) {
            const int64_t row = k0 + t;
            double a[2], b[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int ca = ti * 16 + x * 8 + g, cb = tj * 16 + x * 8 + g;
                a[x] = b[x] = 0.0;
                if (row < r1) {
                    if (ca < ld) a[x] = average ? 0.5 * (A[(size_t)row * ld + ca] + B[(size_t)row * ld + ca]) : A[(size_t)row * ld + ca];
                    if (cb < ld) b[x] = average ? 0.5 * (A[(size_t)row * ld + cb] + B[(size_t)row * ld + cb]) : A[(size_t)row * ld + cb];
                }
            }
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int y = 0; y < 2; ++y) dmma_m8n8k4(c[x][y][0], c[x][y][1], a[x], b[y]);
        }
        double *dst = part + ((size_t)chunk * nt2 + tile) * 256;
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y)
#pragma unroll
                for (int e = 0; e < 2; ++e) dst[(x * 8 + g) * 16 + y * 8 + 2 * t + e] = c[x][y][e];
    }
}

__global__ void __launch_bounds__(256) k_gram_finish(int nchunks, int ntile2, int r, const double *__restrict__ part,
                                                     double *__restrict__ gram)
{
    const int nt = (r + 15) / 16;
    const int tile = blockIdx.x;
    const int ti = tile / nt, tj = tile % nt;
    const int a = threadIdx.x / 16, b2 = threadIdx.x % 16;
    double t = 0.0;
    for (int c = 0; c < nchunks; ++c) t += part[((size_t)c * ntile2 + tile) * 256 + threadIdx.x];
    const int ca = ti * 16 + a, cb = tj * 16 + b2;
    if (ca < r && cb < r) gram[(size_t)ca * r + cb] = t;
}

/* ================================================================================================
 * Fused path for cones whose constraints are all single diagonal entries (MaxCut-type: A_k = a_k e_d e_d^T).
 * There A(sym(U V^T))_k = a_k <U_d, V_d> needs no off-diagonal sample, <C, U V^T> = <U, C V>, and
 * (C + A^*(w)) X = C X + Diag(sum_k w_k a_k) X, so one ALM inner iteration is ONE sparse product T = C D
 * plus row-local streaming work.  Row -> constraints is a CSR (rcptr, rcgid, rca) so several (or no)
 * constraints per row are handled; MaxCut has exactly one.
 * ================================================================================================*/

/* T = C X over the full symmetric CSR with C's values stored per CSR entry; X is indexed by GLOBAL row (the
 * all-gathered factor in a partitioned run), T by this rank's rows.       reference: mul_rk, lorads_sdp_data.c:750-763 */
/* Measured on B200 (profiles/r1_spmm_variants.md): this gather is bound by DRAM, and what keeps DRAM busy is resident
 * warps, not per-thread tricks -- the same product with a software-pipelined (column, value) prefetch and 8 loads in
 * flight per lane needed 48-115 registers and ran 1.8-3.9x slower than this 32-register form at 8 CTAs per SM.
 * A group of G lanes owns a row; lane c holds 16-byte column word c; U entries are fetched per trip. */
/* DOT: the product's epilogue also accumulates sum_i <X_i, T_i> = <C, X X^T> (the line search's p2 with X = D,
 * lorads_alm.c:714-734).  The layout puts a row's diagonal entry last, so the last factor row the walk gathers IS row
 * i's own: it serves both the product and the dot, and the separate 2 F pass over (D, T) disappears at no extra
 * traffic.  (Measured alternatives: re-reading the own row after the walk cost as much as the pass it replaced, 0.7 ms
 * at C5; capturing it inside the walk by a column compare spilled 120 bytes at the 32 registers 8 CTAs per SM allow.) */
/* DOT 0: product only.  DOT 1: the diagonal entry is peeled off the walk (one look at the row's last column id first).
 * DOT 2: plain walk over all entries that remembers the last gathered words and their column; the own row is that one
 * when the column matches, else one more load.  MINB = CTAs per SM the register budget is cut for (8 -> 32 registers). */
template <int G, int U, bool HALO, int DOT, int MINB>
__global__ void __launch_bounds__(LGPU_TPB, MINB) k_mc_spmm(int64_t n, const int32_t *__restrict__ fptr,
                                                            const int32_t *__restrict__ fcol, const double *__restrict__ fval,
                                                            const double *__restrict__ Xin, const double *__restrict__ Xhalo,
                                                            int nsplit, int ld, double *__restrict__ T, int64_t self_off,
                                                            double *partials, unsigned int *counter, double *dsc, SlotSpec<1> spec)
{
    /* HALO: column ids < nsplit address this rank's own rows (Xin), the others the received halo rows; Xhalo is
     * passed pre-offset by -nsplit rows so both cases index with the column id itself */
#define X_ROW(col) ((HALO && (col) >= nsplit ? Xhalo : Xin) + (size_t)(col) * ld)
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    double dot = 0.0;
    for (int64_t i = g0; i < n; i += groups) {
        const int e0 = fptr[i], e1 = fptr[i + 1];
        if (e1 - e0 > LGPU_LONG_ROW) continue; /* hub rows go to k_spmm_long_chunks */
        /* DOT 1: the row's diagonal entry, if it has one, is its LAST entry (lgpu_layout.h) */
        const int ew = (DOT == 1 && e1 > e0 && fcol[e1 - 1] == (int)i + (int)self_off) ? e1 - 1 : e1;
        for (int c = lane; c < ld2; c += G) {
            double2 acc = make_double2(0.0, 0.0);
            double2 xl = make_double2(0.0, 0.0);
            int cl = -1;
            int e = e0;
            for (; e + U <= ew; e += U) {
                double2 x[U];
                double v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    v[u] = fval[e + u];
                    const int col = fcol[e + u];
                    x[u] = reinterpret_cast<const double2 *>(X_ROW(col))[c];
                    if (DOT == 2 && u == U - 1) cl = col;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { acc.x = fma(v[u], x[u].x, acc.x); acc.y = fma(v[u], x[u].y, acc.y); }
                if (DOT == 2) xl = x[U - 1];
            }
            for (; e < ew; ++e) {
                const double v = fval[e];
                const int col = fcol[e];
                const double2 x = reinterpret_cast<const double2 *>(X_ROW(col))[c];
                acc.x = fma(v, x.x, acc.x); acc.y = fma(v, x.y, acc.y);
                if (DOT == 2) { xl = x; cl = col; }
            }
            if (DOT == 1) {
                /* the row's own words: the last gathered row when the diagonal is present, else one more load */
                const double2 xd = reinterpret_cast<const double2 *>(Xin + (size_t)((int)i + (int)self_off) * ld)[c];
                if (ew < e1) { const double v = fval[ew]; acc.x = fma(v, xd.x, acc.x); acc.y = fma(v, xd.y, acc.y); }
                dot = fma(xd.x, acc.x, dot); dot = fma(xd.y, acc.y, dot);
            }
            if (DOT == 2) {
                if (cl != (int)i + (int)self_off) xl = reinterpret_cast<const double2 *>(Xin + (size_t)((int)i + (int)self_off) * ld)[c];
                dot = fma(xl.x, acc.x, dot); dot = fma(xl.y, acc.y, dot);
            }
            reinterpret_cast<double2 *>(T + (size_t)i * ld)[c] = acc;
        }
    }
#undef X_ROW
    if (DOT) {
        double red[1] = {dot};
        grid_reduce_finish<1>(red, partials, counter, dsc, spec);
    }
}

/* Rows of the symmetric CSR with more than LGPU_LONG_ROW entries (hub vertices, arrow-shaped patterns) are cut into
 * chunks of LGPU_LONG_CHUNK entries; one CTA per chunk (its 256/G groups stride over the chunk, partial sums meet in
 * shared memory in group order), partial rows go to scratch, and a second small kernel adds a row's chunks in chunk
 * order and applies alpha/beta.  Fixed summation order: bit-reproducible.  Without this a single group walks the
 * whole row while the rest of the GPU idles: ice_2.0 (n = 8113, one row of 8112 entries) spent 1.1 ms per product there.
 * value of entry e: vals[e] (slots == nullptr) or vals[slots[e]] */
template <int G, bool HALO>
__global__ void __launch_bounds__(LGPU_TPB) k_spmm_long_chunks(int64_t nwork, const int32_t *__restrict__ wrow,
                                                               const int32_t *__restrict__ wbeg, const int32_t *__restrict__ wend,
                                                               const int32_t *__restrict__ fcol, const int32_t *__restrict__ slots,
                                                               const double *__restrict__ vals, const double *__restrict__ Xin,
                                                               const double *__restrict__ Xhalo, int nsplit, int ld,
                                                               double *__restrict__ scratch)
{
    constexpr int NG = LGPU_TPB / G;
    __shared__ double2 part[NG][G];
    const int lane = threadIdx.x % G, grp = threadIdx.x / G;
    const int ld2 = ld >> 1;
    for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int e0 = wbeg[w], e1 = wend[w];
        for (int cb = 0; cb < ld2; cb += G) {
            const int c = cb + lane;
            double2 acc = make_double2(0.0, 0.0);
            if (c < ld2)
                for (int e = e0 + grp; e < e1; e += NG) {
                    const int col = fcol[e];
                    const double v = slots ? vals[slots[e]] : vals[e];
                    const double *xr = ((HALO && col >= nsplit) ? Xhalo : Xin) + (size_t)col * ld;
                    const double2 x = reinterpret_cast<const double2 *>(xr)[c];
                    acc.x = fma(v, x.x, acc.x); acc.y = fma(v, x.y, acc.y);
                }
            part[grp][lane] = acc;
            __syncthreads();
            if (grp == 0 && c < ld2) {
                double2 t = part[0][lane];
                for (int g = 1; g < NG; ++g) { t.x += part[g][lane].x; t.y += part[g][lane].y; }
                reinterpret_cast<double2 *>(scratch + (size_t)w * ld)[c] = t;
            }
            __syncthreads();
        }
    }
    (void)wrow;
}
__global__ void __launch_bounds__(LGPU_TPB) k_spmm_long_finish(int64_t nlong, const int32_t *__restrict__ rows,
                                                               const int32_t *__restrict__ first, int ld,
                                                               const double *__restrict__ scratch, double alpha, double beta,
                                                               const double *__restrict__ Z, double *__restrict__ T)
{
    const int64_t total = nlong * (int64_t)ld;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int64_t k = q / ld;
        const int c = (int)(q - k * ld);
        double t = 0.0;
        for (int w = first[k]; w < first[k + 1]; ++w) t += scratch[(size_t)w * ld + c];
        const size_t o = (size_t)rows[k] * ld + c;
        double v = alpha * t;
        if (Z != nullptr) v = fma(beta, Z[o], v);
        T[o] = v;
    }
}

/* One CG iteration's first half for diagonal constraints, fused (lorads_cgs.c:209-236 with linSysProduct,
 * lorads_admm.c:471-486): [p = beta p + r ;]  Q = p + Diag(c_i <p_i, V_i>) V ;  sums <r, r> and <p, Q> ; the finishing
 * thread forms alpha = <r,r> / <p,Q>.  beta is read from dsc[beta_slot] (beta_slot < 0: p is used as it is). */
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_mc_cg_first(int64_t n, int ld, int beta_slot, double *__restrict__ P,
                                                          const double *__restrict__ Rr, const double *__restrict__ Vf,
                                                          const int32_t *__restrict__ rcptr, const double *__restrict__ rca,
                                                          double *__restrict__ Q, double *partials, unsigned int *counter,
                                                          double *dsc, SlotSpec<2> spec, int alpha_slot)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    const int64_t iters = (n + groups - 1) / groups;
    const double beta = beta_slot >= 0 ? dsc[beta_slot] : 0.0;
    double red[2] = {0.0, 0.0};
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        double pv = 0.0;
        if (live)
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                double2 p = reinterpret_cast<double2 *>(P)[w];
                const double2 r = reinterpret_cast<const double2 *>(Rr)[w];
                if (beta_slot >= 0) {
                    p.x = fma(beta, p.x, r.x); p.y = fma(beta, p.y, r.y);
                    reinterpret_cast<double2 *>(P)[w] = p;
                }
                const double2 v = reinterpret_cast<const double2 *>(Vf)[w];
                pv = fma(p.x, v.x, pv); pv = fma(p.y, v.y, pv);
                red[0] = fma(r.x, r.x, red[0]); red[0] = fma(r.y, r.y, red[0]);
            }
        pv = group_sum<G>(pv);
        if (live) {
            double dg = 0.0;
            for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) dg = fma(rca[t] * pv, rca[t], dg);
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                const double2 p = reinterpret_cast<const double2 *>(P)[w];
                const double2 v = reinterpret_cast<const double2 *>(Vf)[w];
                const double2 q = make_double2(fma(dg, v.x, p.x), fma(dg, v.y, p.y));
                reinterpret_cast<double2 *>(Q)[w] = q;
                red[1] = fma(p.x, q.x, red[1]); red[1] = fma(p.y, q.y, red[1]);
            }
        }
    }
    const int rr = spec.slot[0], pq = spec.slot[1];
    grid_reduce_finish<2>(red, partials, counter, dsc, spec, [=] __device__(double *sc) { sc[alpha_slot] = sc[rr] / sc[pq]; });
}

/* ADMM update of the LP block: LORADSUpdateLPVarOne over the columns, u_j then v_j, constrValSum refreshed after every
 * single update (lorads_admm.c:759-792, lorads_alg_common.c:356-374).  The sweep is a Gauss-Seidel recurrence through
 * constrValSum, sequential by construction, so it runs as ONE warp that walks the columns in order; the lanes share a
 * column's entries (a column touches distinct constraints, so its constrValSum updates do not collide). */
__global__ void __launch_bounds__(32) k_lp_admm_sweep(int64_t nlp, double rho, const int32_t *__restrict__ cptr,
                                                      const int32_t *__restrict__ crow, const double *__restrict__ cval,
                                                      const double *__restrict__ obj, const double *__restrict__ nrm2sq,
                                                      const double *__restrict__ b, const double *__restrict__ lam,
                                                      double *cvs, double *u, double *v)
{
    const int lane = threadIdx.x;
    for (int64_t j = 0; j < nlp; ++j) {
        const int e0 = cptr[j], e1 = cptr[j + 1];
        for (int pass = 0; pass < 2; ++pass) {
            const double uj = ((volatile double *)u)[j], vj = ((volatile double *)v)[j];
            const double fixed = pass == 0 ? vj : uj;
            const double uv_old = uj * vj;
            double part = 0.0;
            for (int e = e0 + lane; e < e1; e += 32) {
                const int i = crow[e];
                const double a = cval[e];
                const double m1 = rho * (-b[i] + ((volatile double *)cvs)[i] - a * uv_old) - lam[i];
                part = fma(m1, a, part);
            }
            const double w = obj[j] + warp_sum(part);
            double M2 = w * fixed;
            M2 = M2 - rho * fixed;
            const double blin = -1.0 * M2 / rho;
            const double nv = blin / (1 + nrm2sq[j] * fixed * fixed);
            const double uv_new = pass == 0 ? nv * vj : uj * nv;
            if (lane == 0) {
                if (pass == 0) u[j] = nv;
                else v[j] = nv;
            }
            for (int e = e0 + lane; e < e1; e += 32) {
                const int i = crow[e];
                const double a = cval[e];
                double c = ((volatile double *)cvs)[i];
                c -= a * uv_old;
                c += a * uv_new;
                cvs[i] = c;
            }
            __threadfence_block();
            __syncwarp();
        }
    }
}

/* pack the rows other ranks need into the send buffer (grouped by destination): out[k] = X[idx[k]] */
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_pack_rows(int64_t nrows, int ld, const int32_t *__restrict__ idx,
                                                        const double *__restrict__ X, double *__restrict__ out)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    for (int64_t k = g0; k < nrows; k += groups) {
        const size_t src = (size_t)idx[k] * ld2, dst = (size_t)k * ld2;
        for (int c = lane; c < ld2; c += G)
            reinterpret_cast<double2 *>(out)[dst + c] = reinterpret_cast<const double2 *>(X)[src + c];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Peer-memory exchange over NVLink (one process per GPU, buffers mapped through CUDA IPC).
 * ------------------------------------------------------------------------------------------------*/
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
/* wait until *flag >= seq; a peer that never arrives is a failed run, not a hung GPU: trap after 30 s */
__device__ __forceinline__ void wait_flag_ge(const unsigned long long *flag, unsigned long long seq)
{
    const unsigned long long t0 = global_timer_ns();
    while (ld_volatile_u64(flag) < seq) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > 30000000000ull) __trap();
    }
}

/* PUT: row idx[k] of X goes to row (k - send_off[q]) of dst[q], q = the destination whose segment holds k -- the pack
 * kernel and the send/receive pair in one pass of NVLink stores; the last CTA publishes `seq` in every peer's block.
 * reference: none (the reference is single-process); replaces k_pack_rows + ncclSend/ncclRecv */
struct PeerPut {
    int world, rank;
    long long send_off[LGPU_MAX_WORLD + 1];
    double *dst[LGPU_MAX_WORLD];
    unsigned long long *flag[LGPU_MAX_WORLD]; /* &peer_blk[q]->xflag[rank] */
};
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_put_rows(int64_t nrows, int ld, const int32_t *__restrict__ idx,
                                                       const double *__restrict__ X, PeerPut pp, unsigned long long seq,
                                                       unsigned int *counter)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    for (int64_t k = g0; k < nrows; k += groups) {
        int q = 0;
        while (k >= pp.send_off[q + 1]) ++q;
        const double2 *src = reinterpret_cast<const double2 *>(X) + (size_t)idx[k] * ld2;
        double2 *dst = reinterpret_cast<double2 *>(pp.dst[q]) + (size_t)(k - pp.send_off[q]) * ld2;
        for (int c = lane; c < ld2; c += G) dst[c] = src[c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool last = atomicAdd(counter, 1u) == gridDim.x - 1;
        if (last) {
            *counter = 0u;
            __threadfence_system();
            for (int q = 0; q < pp.world; ++q)
                if (q != pp.rank) st_release_sys_u64(pp.flag[q], seq);
        }
    }
}
/* the consumer side: one lane per source spins on this rank's own block until every source has published `seq` */
__global__ void __launch_bounds__(32) k_wait_sources(const unsigned long long *xflag, int world, int rank, unsigned long long seq)
{
    const int q = threadIdx.x;
    if (q < world && q != rank) wait_flag_ge(xflag + q, seq);
    __threadfence_system();
}

/* One-shot all-reduce of `count` (<= LGPU_PEER_RED) device scalars: every rank stores its values into every rank's inbox
 * (its own included), publishes `seq`, waits for the others and adds the world's values in RANK ORDER -- the same
 * bits on every rank.  Replaces a latency-bound ncclAllReduce of a few doubles. */
struct PeerRed {
    int world, rank;
    PeerBlock *blk[LGPU_MAX_WORLD]; /* [q] = rank q's block as mapped here */
};
__global__ void __launch_bounds__(LGPU_MAX_WORLD * LGPU_PEER_RED) k_peer_allreduce(double *dsc, int first, int count, PeerRed pr,
                                                                                   unsigned long long seq)
{
    const int t = threadIdx.x, q = t / LGPU_PEER_RED, k = t % LGPU_PEER_RED;
    const int par = (int)(seq & 1ull);
    if (q < pr.world && k < count) pr.blk[q]->inbox[par][pr.rank][k] = dsc[first + k];
    __threadfence_system();
    __syncthreads();
    if (t < pr.world && t != pr.rank) {
        st_release_sys_u64(&pr.blk[t]->aflag[pr.rank], seq);
        wait_flag_ge(&pr.blk[pr.rank]->aflag[t], seq);
    }
    __threadfence_system();
    __syncthreads();
    if (t < count) {
        double s = 0.0;
        const volatile double *in = &pr.blk[pr.rank]->inbox[par][0][0];
        for (int r = 0; r < pr.world; ++r) s += in[r * LGPU_PEER_RED + t];
        dsc[first + t] = s;
    }
}

/* Row-local pass after T = C D: q1_k = 2 a_k <R_i, D_i>, q2_k = a_k <D_i, D_i> for the constraints k of row i, and the
 * objective terms sum_i <R_i, T_i> (p1 / 2) and sum_i <D_i, T_i> (p2)
 *     reference: ALMCalq12p12 -> LORADSObjConstrValAll -> LORADSUVt/objAUV/coneAUV, lorads_alm.c:714-734,
 *     lorads_alg_common.c:43-90,153-176 */
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_mc_epi(int64_t n, int ld, const double *__restrict__ Rm,
                                                     const double *__restrict__ D, const double *__restrict__ T,
                                                     const int32_t *__restrict__ rcptr, const int32_t *__restrict__ rcgid,
                                                     const double *__restrict__ rca, double *__restrict__ q1,
                                                     double *__restrict__ q2, double *partials, unsigned int *counter,
                                                     double *dsc, SlotSpec<2> spec)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    const int64_t iters = (n + groups - 1) / groups;
    double red[2] = {0.0, 0.0};
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        double rd = 0.0, dd = 0.0;
        if (live)
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                const double2 r = reinterpret_cast<const double2 *>(Rm)[w];
                const double2 d = reinterpret_cast<const double2 *>(D)[w];
                const double2 t = reinterpret_cast<const double2 *>(T)[w];
                rd = fma(r.x, d.x, rd); rd = fma(r.y, d.y, rd);
                dd = fma(d.x, d.x, dd); dd = fma(d.y, d.y, dd);
                red[0] = fma(r.x, t.x, red[0]); red[0] = fma(r.y, t.y, red[0]);
                red[1] = fma(d.x, t.x, red[1]); red[1] = fma(d.y, t.y, red[1]);
            }
        rd = group_sum<G>(rd);
        dd = group_sum<G>(dd);
        if (live && lane == 0)
            for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) {
                const double a = rca[t];
                q1[rcgid[t]] = 2.0 * (a * rd);
                q2[rcgid[t]] = a * dd;
            }
    }
    grid_reduce_finish<2>(red, partials, counter, dsc, spec);
}

/* L-BFGS direction from precomputed coefficients (history length 2), one pass, with the constraint-operator
 * epilogue:  q = G ; q -= a1 y1 ; [q -= a0 y0 ; q += w0 s0 ;] q += w1 s1 ; D = -q   -- the reference's axpy sequence
 * (lorads_alm.c:468-505) with alpha / beta scalars obtained from carried inner products instead of four extra
 * passes.  nn = 0: D = -G.  Also q1_k = 2 a_k <R_i, D_i>, q2_k = a_k <D_i, D_i> and sum_i <(C R)_i, D_i> = <C, R D^T>. */
struct DirCoef {
    double a1, a0, w0, w1;
    int nn;
};
/* PUT (partitioned runs, peer memory): the row of D just formed is also stored into the halo buffer of every peer whose
 * CSR rows reference it -- dest[i * npeer + j] is its row there (or -1) -- so the halo exchange rides inside the
 * direction pass: no pack kernel, no second read of D, and the NVLink stores overlap the pass's own HBM streams.  The
 * finishing thread publishes the exchange's sequence number in every peer's block (fused compute + transfer over
 * peer-mapped memory; the consumers wait in k_wait_sources). */
struct PeerDirect {
    int npeer;                                /* world - 1 */
    double *dst[LGPU_MAX_WORLD];              /* [j] halo base of the j-th peer (this exchange's buffer) */
    unsigned long long *flag[LGPU_MAX_WORLD]; /* [j] &peer_blk->xflag[my rank] */
    unsigned long long seq;
};
struct PublishSeq { /* runs in the finishing thread of the grid, after every CTA's system-wide fence */
    PeerDirect pd;
    __device__ void operator()(double *) const
    {
        __threadfence_system();
        for (int j = 0; j < pd.npeer; ++j) st_release_sys_u64(pd.flag[j], pd.seq);
    }
};
/* ROWDOTS: <R_i, D_i> comes from the per-row products the bulk step pass left (rowdots[i][0..4] against g, s1, y1, s0, y0)
 * combined with the direction's coefficients, and <C R, D> from the carried global products on the host: the pass then reads
 * neither R nor C R (6 F instead of 8 F; rowdots adds 40 bytes per row) */
template <int G, bool PUT, bool ROWDOTS>
__global__ void __launch_bounds__(LGPU_TPB) k_mc_combine(int64_t n, int ld, DirCoef cf, const double *__restrict__ Gd,
                                                         const double *__restrict__ s1, const double *__restrict__ y1,
                                                         const double *__restrict__ s0, const double *__restrict__ y0,
                                                         const double *__restrict__ Rm, const double *__restrict__ CR,
                                                         double *__restrict__ D, const int32_t *__restrict__ rcptr,
                                                         const int32_t *__restrict__ rcgid, const double *__restrict__ rca,
                                                         double *__restrict__ q1, double *__restrict__ q2, double *partials,
                                                         unsigned int *counter, double *dsc, SlotSpec<1> spec,
                                                         const int32_t *__restrict__ dest, PeerDirect pd,
                                                         const double *__restrict__ rowdots)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    const int64_t iters = (n + groups - 1) / groups;
    double red[1] = {0.0};
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        double rd = 0.0, dd = 0.0;
        if (live)
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                double2 q = reinterpret_cast<const double2 *>(Gd)[w];
                if (cf.nn >= 1) {
                    const double2 yb = reinterpret_cast<const double2 *>(y1)[w];
                    q.x = fma(-cf.a1, yb.x, q.x); q.y = fma(-cf.a1, yb.y, q.y);
                    if (cf.nn >= 2) {
                        const double2 ya = reinterpret_cast<const double2 *>(y0)[w];
                        const double2 sa = reinterpret_cast<const double2 *>(s0)[w];
                        q.x = fma(-cf.a0, ya.x, q.x); q.y = fma(-cf.a0, ya.y, q.y);
                        q.x = fma(cf.w0, sa.x, q.x); q.y = fma(cf.w0, sa.y, q.y);
                    }
                    const double2 sb = reinterpret_cast<const double2 *>(s1)[w];
                    q.x = fma(cf.w1, sb.x, q.x); q.y = fma(cf.w1, sb.y, q.y);
                }
                const double2 d = make_double2(-q.x, -q.y);
                reinterpret_cast<double2 *>(D)[w] = d;
                if (PUT) {
                    for (int j = 0; j < pd.npeer; ++j) {
                        const int slot = dest[(size_t)i * pd.npeer + j];
                        if (slot >= 0) reinterpret_cast<double2 *>(pd.dst[j])[(size_t)slot * ld2 + c] = d;
                    }
                }
                if (!ROWDOTS) {
                    const double2 r = reinterpret_cast<const double2 *>(Rm)[w];
                    const double2 cr = reinterpret_cast<const double2 *>(CR)[w];
                    rd = fma(r.x, d.x, rd); rd = fma(r.y, d.y, rd);
                    red[0] = fma(cr.x, d.x, red[0]); red[0] = fma(cr.y, d.y, red[0]);
                }
                dd = fma(d.x, d.x, dd); dd = fma(d.y, d.y, dd);
            }
        if (!ROWDOTS) rd = group_sum<G>(rd);
        dd = group_sum<G>(dd);
        if (ROWDOTS && live && lane == 0) {
            /* <R_i, D_i> with D = -g + a1 y1 + a0 y0 - w0 s0 - w1 s1 (the axpy sequence above, applied to the row products) */
            const double *o = rowdots + (size_t)i * 5;
            rd = -o[0];
            if (cf.nn >= 1) {
                rd = fma(cf.a1, o[2], rd);
                if (cf.nn >= 2) { rd = fma(cf.a0, o[4], rd); rd = fma(-cf.w0, o[3], rd); }
                rd = fma(-cf.w1, o[1], rd);
            }
        }
        if (live && lane == 0)
            for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) {
                const double a = rca[t];
                q1[rcgid[t]] = 2.0 * (a * rd);
                q2[rcgid[t]] = a * dd;
            }
    }
    if (PUT) {
        __threadfence_system(); /* every thread's peer stores are performed before its CTA reports in */
        grid_reduce_finish<1, PublishSeq, true>(red, partials, counter, dsc, spec, PublishSeq{pd});
    } else {
        grid_reduce_finish<1>(red, partials, counter, dsc, spec);
    }
}

/* One fused streaming pass for everything that follows the line search (tau known):
 *   setAsNegGrad + ALMupdateVar + constrValSum += tau q1 + tau^2 q2      lorads_alm.c:1342-1353
 *   ALMCalGrad: M1 = -lambda - rho b + rho constrValSum ; Grad = 2 (C R + Diag(A^*(M1)) R)   lorads_alm.c:32-87
 *   setlbfgsHisTwo: s = tau D, y = Grad_new - Grad_old, <y,s>             lorads_alm.c:842-863
 *   updateDimacsALM: constrValSum <- A(R R^T) from scratch, |b - A|^2      lorads_alg_common.c:386-394,424-428
 * C R is carried as CR <- CR + tau (C D) with C D = T from k_mc_spmm.
 * reductions: [0] sum Grad^2, [1] <y,s>, [2] sum (b - A(RR^T))^2 */
template <int G, bool GRAM, int MINB, bool CS>
__global__ void __launch_bounds__(LGPU_TPB, MINB) k_mc_step(int64_t n, int ld, double tau, double rho, double *__restrict__ Rm,
                                                      const double *__restrict__ D, double *__restrict__ CR,
                                                      const double *__restrict__ T, double *__restrict__ Gd,
                                                      double *__restrict__ sh, double *__restrict__ yh,
                                                      const int32_t *__restrict__ rcptr, const int32_t *__restrict__ rcgid,
                                                      const double *__restrict__ rca, const double *__restrict__ lam,
                                                      const double *__restrict__ b, double *__restrict__ cvs,
                                                      const double *__restrict__ q1, const double *__restrict__ q2,
                                                      double *__restrict__ M1, const double *__restrict__ so,
                                                      const double *__restrict__ yo, double *partials, unsigned int *counter,
                                                      double *dsc, SlotSpec<GRAM ? 10 : 3> spec, int beta_slot)
{
    /* GRAM: also the inner products the next direction needs, with (sn, yn) the pair formed here and (so, yo) the
     * pair that stays in the history: [3] <g,sn> [4] <g,yn> [5] <g,so> [6] <g,yo> [7] <so,yn> [8] <yo,yn> [9] <yn,yn>
     * (g = the new gradient) */
    constexpr int NR = GRAM ? 10 : 3;
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    const int64_t iters = (n + groups - 1) / groups;
    const double t2 = tau * tau;
    double red[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) red[k] = 0.0;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        double rr = 0.0;
        int k0 = 0, k1 = 0;
        if (live) {
            k0 = rcptr[i];
            k1 = rcptr[i + 1];
            double coef = 0.0;
            for (int t = k0; t < k1; ++t) {
                const int k = rcgid[t];
                const double cv = fma(t2, q2[k], fma(tau, q1[k], cvs[k]));
                const double m1 = -lam[k] - rho * b[k] + rho * cv;
                if (lane == 0) M1[k] = m1;
                coef = fma(m1, rca[t], coef);
            }
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                double2 r = ldv<CS>(Rm, w);
                const double2 d = ldv<CS>(D, w);
                double2 cr = ldv<CS>(CR, w);
                const double2 t = ldv<CS>(T, w);
                const double2 go = ldv<CS>(Gd, w);
                r.x = fma(tau, d.x, r.x); r.y = fma(tau, d.y, r.y);
                cr.x = fma(tau, t.x, cr.x); cr.y = fma(tau, t.y, cr.y);
                double2 gn;
                gn.x = 2.0 * fma(coef, r.x, cr.x);
                gn.y = 2.0 * fma(coef, r.y, cr.y);
                const double2 sv = make_double2(tau * d.x, tau * d.y);
                const double2 yv = make_double2(-go.x + gn.x, -go.y + gn.y);
                stv<CS>(Rm, w, r);
                stv<CS>(CR, w, cr);
                stv<CS>(Gd, w, gn);
                stv<CS>(sh, w, sv);
                stv<CS>(yh, w, yv);
                red[0] = fma(gn.x, gn.x, red[0]); red[0] = fma(gn.y, gn.y, red[0]);
                red[1] = fma(yv.x, sv.x, red[1]); red[1] = fma(yv.y, sv.y, red[1]);
                if (GRAM) {
                    const double2 os = ldv<CS>(so, w);
                    const double2 oy = ldv<CS>(yo, w);
                    red[3] = fma(gn.x, sv.x, red[3]); red[3] = fma(gn.y, sv.y, red[3]);
                    red[4] = fma(gn.x, yv.x, red[4]); red[4] = fma(gn.y, yv.y, red[4]);
                    red[5] = fma(gn.x, os.x, red[5]); red[5] = fma(gn.y, os.y, red[5]);
                    red[6] = fma(gn.x, oy.x, red[6]); red[6] = fma(gn.y, oy.y, red[6]);
                    red[7] = fma(os.x, yv.x, red[7]); red[7] = fma(os.y, yv.y, red[7]);
                    red[8] = fma(oy.x, yv.x, red[8]); red[8] = fma(oy.y, yv.y, red[8]);
                    red[9] = fma(yv.x, yv.x, red[9]); red[9] = fma(yv.y, yv.y, red[9]);
                }
                rr = fma(r.x, r.x, rr); rr = fma(r.y, r.y, rr);
            }
        }
        rr = group_sum<G>(rr);
        if (live && lane == 0)
            for (int t = k0; t < k1; ++t) {
                const int k = rcgid[t];
                const double cv = rca[t] * rr;
                cvs[k] = cv;
                const double df = b[k] - cv;
                red[2] = fma(df, df, red[2]);
            }
    }
    const int ys_slot = spec.slot[1];
    grid_reduce_finish<NR>(red, partials, counter, dsc, spec,
                           [=] __device__(double *sc) { sc[beta_slot] = 1.0 / sc[ys_slot]; });
}

/* ------------------------------------------------------------------------------------------------
 * The same pass as k_mc_step, staged through shared memory by the bulk-copy engine (TMA, cp.async.bulk + mbarrier).
 *
 * Why: the register-staged kernel above needs 80 registers for its seven 16-byte loads and ten accumulators (3 CTAs per
 * SM = 37 % of the warps) and every row starts with a dependent chain row -> constraint -> (q1, q2, cvs, lambda, b) ->
 * coefficient that must resolve before the row's arithmetic can retire, so it ran at 0.82 of the copy bandwidth.  Here
 * ONE persistent CTA per SM owns a ring of `nstage` shared-memory stages.  Producer warp p owns stage p: lane 0 arms
 * the stage's mbarrier with the byte count and issues one 1-D bulk copy per input stream (a tile of `tr` consecutive
 * rows of a row-major factor is one contiguous piece), while all its lanes walk the row -> constraint chain of the
 * tile's rows (M1, the row coefficient, and for single-constraint rows the constraint id, a_k and b_k) into the
 * stage.  Several tiles' chains are in flight at once because every producer warp runs its own.  The 8 consumer
 * warps read the stage with conflict-free 16-byte shared loads, do exactly k_mc_step's arithmetic, store the five
 * outputs straight to global memory and hand the stage back through a second mbarrier.  Bytes in flight per SM are set by
 * nstage x tile size (>= 100 KB), not by registers.
 * Waits are bounded: a protocol error traps (reported as a launch failure) instead of hanging the GPU.
 * ------------------------------------------------------------------------------------------------*/
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0, spins = 0;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok) return;
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#define LGPU_STEP_MAX_STAGES 4
__host__ __device__ inline size_t step_bulk_stage_bytes(int nstream, int tr, int ld)
{
    /* nstream factor tiles + per-row {coef, a, b} doubles + {constraint id, count} ints, rounded to 128 bytes */
    size_t b = (size_t)nstream * tr * ld * 8 + (size_t)tr * (3 * 8 + 2 * 4);
    return (b + 127) / 128 * 128;
}

template <int G, bool GRAM, bool ROWD>
__global__ void __launch_bounds__(LGPU_TPB + 32 * LGPU_STEP_MAX_STAGES, 1)
k_mc_step_bulk(int64_t n, int ld, int tr, int nstage, double tau, double rho, double *__restrict__ Rm, const double *__restrict__ D,
               double *__restrict__ CR, const double *__restrict__ T, double *__restrict__ Gd, double *__restrict__ sh,
               double *__restrict__ yh, const int32_t *__restrict__ rcptr, const int32_t *__restrict__ rcgid,
               const double *__restrict__ rca, const double *__restrict__ lam, const double *__restrict__ b,
               double *__restrict__ cvs, const double *__restrict__ q1, const double *__restrict__ q2, double *__restrict__ M1,
               const double *__restrict__ so, const double *__restrict__ yo, double *partials, unsigned int *counter, double *dsc,
               SlotSpec<GRAM ? (ROWD ? 15 : 10) : 3> spec, int beta_slot, double *__restrict__ rowdots)
{
    /* ROWD (with GRAM): reductions [10..14] = <CR, g>, <CR, s_new>, <CR, y_new>, <CR, s_old>, <CR, y_old> and, per row, the
     * same five partners against R_i into rowdots[i][0..4]: the next direction pass gets <R_i, D_i> and <C R, D> from them */
    constexpr int NSTREAM = GRAM ? 7 : 5;
    constexpr int NR = GRAM ? (ROWD ? 15 : 10) : 3;
    constexpr int NG = LGPU_TPB / G;
    extern __shared__ __align__(128) unsigned char smem[];
    const int ld2 = ld >> 1;
    const uint32_t tile_bytes = (uint32_t)tr * (uint32_t)ld * 8u;
    const size_t stage_bytes = step_bulk_stage_bytes(NSTREAM, tr, ld);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)nstage * stage_bytes);
    uint64_t *empty = full + LGPU_STEP_MAX_STAGES;
    const int64_t ntiles = (n + tr - 1) / tr;
    const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1 + 32); /* lane 0's arrive.expect_tx + one arrive per producer lane */
            mbar_init(&empty[s], LGPU_TPB / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= LGPU_TPB / 32) {
        /* ---------------- producer warp p: stage p, tiles blockIdx.x + (p + j nstage) gridDim.x ---------------- */
        const int p = warp - LGPU_TPB / 32;
        if (p >= nstage) return;
        unsigned char *base = smem + (size_t)p * stage_bytes;
        double *scoef = reinterpret_cast<double *>(base + (size_t)NSTREAM * tile_bytes);
        double *sra = scoef + tr, *srb = sra + tr;
        int *srk = reinterpret_cast<int *>(srb + tr), *scnt = srk + tr;
        const double t2 = tau * tau;
        int it = 0;
        for (int64_t s = p;; s += nstage, ++it) {
            const int64_t tile = (int64_t)blockIdx.x + s * gridDim.x;
            if (tile >= ntiles) break;
            if (it > 0) mbar_wait(&empty[p], (uint32_t)((it - 1) & 1));
            const int64_t row0 = tile * tr;
            const int rows = (int)((n - row0 < tr) ? (n - row0) : tr);
            if (lane32 == 0) {
                const uint32_t bytes = (uint32_t)rows * (uint32_t)ld * 8u;
                mbar_arrive_expect_tx(&full[p], NSTREAM * bytes);
                const size_t off = (size_t)row0 * ld;
                bulk_g2s(base + 0 * (size_t)tile_bytes, Rm + off, bytes, &full[p]);
                bulk_g2s(base + 1 * (size_t)tile_bytes, D + off, bytes, &full[p]);
                bulk_g2s(base + 2 * (size_t)tile_bytes, CR + off, bytes, &full[p]);
                bulk_g2s(base + 3 * (size_t)tile_bytes, T + off, bytes, &full[p]);
                bulk_g2s(base + 4 * (size_t)tile_bytes, Gd + off, bytes, &full[p]);
                if (GRAM) {
                    bulk_g2s(base + 5 * (size_t)tile_bytes, so + off, bytes, &full[p]);
                    bulk_g2s(base + 6 * (size_t)tile_bytes, yo + off, bytes, &full[p]);
                }
            }
            for (int rr = lane32; rr < rows; rr += 32) {
                const int64_t i = row0 + rr;
                const int k0 = rcptr[i], k1 = rcptr[i + 1];
                double coef = 0.0;
                for (int t = k0; t < k1; ++t) {
                    const int k = rcgid[t];
                    const double cv = fma(t2, q2[k], fma(tau, q1[k], cvs[k]));
                    const double m1 = -lam[k] - rho * b[k] + rho * cv;
                    M1[k] = m1;
                    coef = fma(m1, rca[t], coef);
                }
                scoef[rr] = coef;
                scnt[rr] = k1 - k0;
                if (k1 - k0 == 1) {
                    const int k = rcgid[k0];
                    srk[rr] = k;
                    sra[rr] = rca[k0];
                    srb[rr] = b[k];
                } else {
                    srk[rr] = k0;
                }
            }
            mbar_arrive(&full[p]);
        }
        return;
    }

    /* ---------------- consumers: 8 warps, a group of G lanes per row ---------------- */
    const int lane = threadIdx.x % G, grp = threadIdx.x / G;
    double red[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) red[k] = 0.0;
    int st = 0, it = 0;
    for (int64_t s = 0;; ++s) {
        const int64_t tile = (int64_t)blockIdx.x + s * gridDim.x;
        if (tile >= ntiles) break;
        mbar_wait(&full[st], (uint32_t)(it & 1));
        const unsigned char *base = smem + (size_t)st * stage_bytes;
        const double2 *sR = reinterpret_cast<const double2 *>(base);
        const double2 *sD = reinterpret_cast<const double2 *>(base + 1 * (size_t)tile_bytes);
        const double2 *sC = reinterpret_cast<const double2 *>(base + 2 * (size_t)tile_bytes);
        const double2 *sT = reinterpret_cast<const double2 *>(base + 3 * (size_t)tile_bytes);
        const double2 *sG = reinterpret_cast<const double2 *>(base + 4 * (size_t)tile_bytes);
        const double2 *sSo = reinterpret_cast<const double2 *>(base + 5 * (size_t)tile_bytes);
        const double2 *sYo = reinterpret_cast<const double2 *>(base + 6 * (size_t)tile_bytes);
        const double *scoef = reinterpret_cast<const double *>(base + (size_t)NSTREAM * tile_bytes);
        const double *sra = scoef + tr, *srb = sra + tr;
        const int *srk = reinterpret_cast<const int *>(srb + tr), *scnt = srk + tr;
        const int64_t row0 = tile * tr;
        const int rows = (int)((n - row0 < tr) ? (n - row0) : tr);
        for (int rb = 0; rb < tr; rb += NG) { /* every group runs the same trip count: the shuffles stay convergent */
            const int rr = rb + grp;
            const bool live = rr < rows;
            double rsq = 0.0;
            double rdot[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            if (live) {
                const double coef = scoef[rr];
                for (int c = lane; c < ld2; c += G) {
                    const int ws = rr * ld2 + c;
                    const size_t w = (size_t)(row0 + rr) * ld2 + c;
                    double2 r = sR[ws];
                    const double2 d = sD[ws];
                    double2 cr = sC[ws];
                    const double2 t = sT[ws];
                    const double2 go = sG[ws];
                    r.x = fma(tau, d.x, r.x); r.y = fma(tau, d.y, r.y);
                    cr.x = fma(tau, t.x, cr.x); cr.y = fma(tau, t.y, cr.y);
                    double2 gn;
                    gn.x = 2.0 * fma(coef, r.x, cr.x);
                    gn.y = 2.0 * fma(coef, r.y, cr.y);
                    const double2 sv = make_double2(tau * d.x, tau * d.y);
                    const double2 yv = make_double2(-go.x + gn.x, -go.y + gn.y);
                    reinterpret_cast<double2 *>(Rm)[w] = r;
                    reinterpret_cast<double2 *>(CR)[w] = cr;
                    reinterpret_cast<double2 *>(Gd)[w] = gn;
                    reinterpret_cast<double2 *>(sh)[w] = sv;
                    reinterpret_cast<double2 *>(yh)[w] = yv;
                    red[0] = fma(gn.x, gn.x, red[0]); red[0] = fma(gn.y, gn.y, red[0]);
                    red[1] = fma(yv.x, sv.x, red[1]); red[1] = fma(yv.y, sv.y, red[1]);
                    if (GRAM) {
                        const double2 os = sSo[ws];
                        const double2 oy = sYo[ws];
                        red[3] = fma(gn.x, sv.x, red[3]); red[3] = fma(gn.y, sv.y, red[3]);
                        red[4] = fma(gn.x, yv.x, red[4]); red[4] = fma(gn.y, yv.y, red[4]);
                        red[5] = fma(gn.x, os.x, red[5]); red[5] = fma(gn.y, os.y, red[5]);
                        red[6] = fma(gn.x, oy.x, red[6]); red[6] = fma(gn.y, oy.y, red[6]);
                        red[7] = fma(os.x, yv.x, red[7]); red[7] = fma(os.y, yv.y, red[7]);
                        red[8] = fma(oy.x, yv.x, red[8]); red[8] = fma(oy.y, yv.y, red[8]);
                        red[9] = fma(yv.x, yv.x, red[9]); red[9] = fma(yv.y, yv.y, red[9]);
                    }
                    if (GRAM && ROWD) {
                        const double2 os = sSo[ws];
                        const double2 oy = sYo[ws];
                        red[10] = fma(cr.x, gn.x, red[10]); red[10] = fma(cr.y, gn.y, red[10]);
                        red[11] = fma(cr.x, sv.x, red[11]); red[11] = fma(cr.y, sv.y, red[11]);
                        red[12] = fma(cr.x, yv.x, red[12]); red[12] = fma(cr.y, yv.y, red[12]);
                        red[13] = fma(cr.x, os.x, red[13]); red[13] = fma(cr.y, os.y, red[13]);
                        red[14] = fma(cr.x, oy.x, red[14]); red[14] = fma(cr.y, oy.y, red[14]);
                        rdot[0] = fma(r.x, gn.x, rdot[0]); rdot[0] = fma(r.y, gn.y, rdot[0]);
                        rdot[1] = fma(r.x, sv.x, rdot[1]); rdot[1] = fma(r.y, sv.y, rdot[1]);
                        rdot[2] = fma(r.x, yv.x, rdot[2]); rdot[2] = fma(r.y, yv.y, rdot[2]);
                        rdot[3] = fma(r.x, os.x, rdot[3]); rdot[3] = fma(r.y, os.y, rdot[3]);
                        rdot[4] = fma(r.x, oy.x, rdot[4]); rdot[4] = fma(r.y, oy.y, rdot[4]);
                    }
                    rsq = fma(r.x, r.x, rsq); rsq = fma(r.y, r.y, rsq);
                }
            }
            rsq = group_sum<G>(rsq);
            if (GRAM && ROWD) {
#pragma unroll
                for (int q = 0; q < 5; ++q) rdot[q] = group_sum<G>(rdot[q]);
                if (live && lane == 0) {
                    double *o = rowdots + (size_t)(row0 + rr) * 5;
#pragma unroll
                    for (int q = 0; q < 5; ++q) o[q] = rdot[q];
                }
            }
            if (live && lane == 0) {
                const int cnt = scnt[rr];
                if (cnt == 1) {
                    const double cv = sra[rr] * rsq;
                    cvs[srk[rr]] = cv;
                    const double df = srb[rr] - cv;
                    red[2] = fma(df, df, red[2]);
                } else {
                    const int k0 = srk[rr];
                    for (int t = k0; t < k0 + cnt; ++t) {
                        const int k = rcgid[t];
                        const double cv = rca[t] * rsq;
                        cvs[k] = cv;
                        const double df = b[k] - cv;
                        red[2] = fma(df, df, red[2]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane32 == 0) mbar_arrive(&empty[st]);
        if (++st == nstage) { st = 0; ++it; }
    }
    const int ys_slot = spec.slot[1];
    grid_reduce_finish<NR>(red, partials, counter, dsc, spec, [=] __device__(double *sc) { sc[beta_slot] = 1.0 / sc[ys_slot]; });
}

/* Grad = 2 (CR + Diag(sum_k M1_k a_k) R) and sum Grad^2, with CR = C R already formed    lorads_alm.c:32-87 */
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_mc_grad(int64_t n, int ld, const double *__restrict__ Rm,
                                                      const double *__restrict__ CR, double *__restrict__ Gd,
                                                      const int32_t *__restrict__ rcptr, const int32_t *__restrict__ rcgid,
                                                      const double *__restrict__ rca, const double *__restrict__ M1,
                                                      double *partials, unsigned int *counter, double *dsc, SlotSpec<1> spec)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    double red[1] = {0.0};
    for (int64_t i = g0; i < n; i += groups) {
        double coef = 0.0;
        for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) coef = fma(M1[rcgid[t]], rca[t], coef);
        for (int c = lane; c < ld2; c += G) {
            const size_t w = (size_t)i * ld2 + c;
            const double2 r = reinterpret_cast<const double2 *>(Rm)[w];
            const double2 cr = reinterpret_cast<const double2 *>(CR)[w];
            double2 gn;
            gn.x = 2.0 * fma(coef, r.x, cr.x);
            gn.y = 2.0 * fma(coef, r.y, cr.y);
            reinterpret_cast<double2 *>(Gd)[w] = gn;
            red[0] = fma(gn.x, gn.x, red[0]); red[0] = fma(gn.y, gn.y, red[0]);
        }
    }
    grid_reduce_finish<1>(red, partials, counter, dsc, spec);
}

/* cvs_k = a_k <A_i, B_i> for the constraints k of row i (A(sym(A B^T)) for diagonal constraints); optional
 * sum (b - cvs)^2          reference: LORADSInitConstrValAll/Sum + primalInfeasibility, lorads_alg_common.c:116-122,221-229,386-394 */
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_mc_rowdot(int64_t n, int ld, const double *__restrict__ A,
                                                        const double *__restrict__ B, const int32_t *__restrict__ rcptr,
                                                        const int32_t *__restrict__ rcgid, const double *__restrict__ rca,
                                                        double scale, double *__restrict__ out, const double *__restrict__ b,
                                                        double *partials, unsigned int *counter, double *dsc, SlotSpec<1> spec)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    const int64_t iters = (n + groups - 1) / groups;
    double red[1] = {0.0};
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        double ab = 0.0;
        if (live)
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                const double2 x = reinterpret_cast<const double2 *>(A)[w];
                const double2 y = reinterpret_cast<const double2 *>(B)[w];
                ab = fma(x.x, y.x, ab); ab = fma(x.y, y.y, ab);
            }
        ab = group_sum<G>(ab);
        if (live && lane == 0)
            for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) {
                const int k = rcgid[t];
                const double cv = scale * (rca[t] * ab);
                out[k] = cv;
                if (b != nullptr) {
                    const double df = b[k] - cv;
                    red[0] = fma(df, df, red[0]);
                }
            }
    }
    if (b != nullptr) grid_reduce_finish<1>(red, partials, counter, dsc, spec);
}

/* CG operator for diagonal constraints: out = x + Diag(c_i <x_i, V_i>) V with c_i = sum_k a_k^2, and <p, out>
 *     reference: linSysProduct / LORADSUpdateConstrValCG, lorads_admm.c:442-486 */
template <int G>
__global__ void __launch_bounds__(LGPU_TPB) k_mc_cg_mvec(int64_t n, int ld, const double *__restrict__ X,
                                                         const double *__restrict__ Vf, const int32_t *__restrict__ rcptr,
                                                         const double *__restrict__ rca, double *__restrict__ out)
{
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ld2 = ld >> 1;
    const int64_t iters = (n + groups - 1) / groups;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        double xv = 0.0;
        if (live)
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                const double2 x = reinterpret_cast<const double2 *>(X)[w];
                const double2 v = reinterpret_cast<const double2 *>(Vf)[w];
                xv = fma(x.x, v.x, xv); xv = fma(x.y, v.y, xv);
            }
        xv = group_sum<G>(xv);
        if (live) {
            /* weight_k = a_k <x_i, V_i> ; the aggregate's diagonal entry is sum_k weight_k a_k (constraint order) */
            double dg = 0.0;
            for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) dg = fma(rca[t] * xv, rca[t], dg);
            for (int c = lane; c < ld2; c += G) {
                const size_t w = (size_t)i * ld2 + c;
                const double2 x = reinterpret_cast<const double2 *>(X)[w];
                const double2 v = reinterpret_cast<const double2 *>(Vf)[w];
                reinterpret_cast<double2 *>(out)[w] = make_double2(fma(dg, v.x, x.x), fma(dg, v.y, x.y));
            }
        }
    }
}

#endif
