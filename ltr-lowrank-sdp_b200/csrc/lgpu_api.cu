/*
 * lgpu_api.cu -- C ABI implementation of liblorads_b200.so (declared in include/lorads_b200.h).
 *
 * Host side of the device layer: problem preprocessing (the reference's AConeProcData /
 * AConePresolveData rules re-derived for a uniform device layout), device memory, kernel
 * sequencing per reference function.  Control flow of the solver (ALM/ADMM state machines,
 * line-search root selection, printing) stays in the C host driver (csrc/host/).
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <type_traits>
#include <unordered_map>

#include "../../include/lorads_b200.h"
#include "lgpu_kernels.cuh"
#include "lgpu_layout.h"

/* ------------------------------------------------------------------------------------------------
 * error handling
 * ------------------------------------------------------------------------------------------------*/
#define LGPU_FAIL(ctx, ...)                                   \
    do {                                                      \
        char _buf[512];                                       \
        snprintf(_buf, sizeof(_buf), __VA_ARGS__);            \
        (ctx)->err = _buf;                                    \
        (ctx)->failed = true;                                 \
        return 1;                                             \
    } while (0)

#define CU(ctx, call)                                                                                      \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) LGPU_FAIL(ctx, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

/* also trips on a sticky earlier failure (an NCCL call inside a void helper): the solver must not continue on
 * per-rank partial sums or a stale product */
#define CHECK_LAUNCH(ctx)                                                                                  \
    do {                                                                                                   \
        if ((ctx)->failed) return 1;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                               \
        if (_e != cudaSuccess) LGPU_FAIL(ctx, "%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)

#define TRY(x)              \
    do {                    \
        int _rc = (x);      \
        if (_rc) return _rc; \
    } while (0)

static thread_local std::string g_create_err;

template <class T>
static int dev_alloc(lgpu_ctx *ctx, T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    CU(ctx, cudaMalloc((void **)p, count * sizeof(T)));
    return 0;
}
template <class T>
static int dev_upload(lgpu_ctx *ctx, T **p, const std::vector<T> &h)
{
    TRY(dev_alloc(ctx, p, h.size()));
    if (!h.empty()) CU(ctx, cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
template <class T>
static void dev_free(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

/* ------------------------------------------------------------------------------------------------
 * launch helpers
 * ------------------------------------------------------------------------------------------------*/
/* grid = min(blocks needed, SMs x resident CTAs of THIS kernel): every launch is at most one full wave, grid-stride
 * loops cover the rest, so there is no partial last wave (148 SMs; occupancy queried once per kernel and cached) */
static std::unordered_map<const void *, int> g_occ;
static inline int grid_for(const lgpu_ctx *ctx, int64_t threads_needed, const void *kernel = nullptr)
{
    int64_t blocks = (threads_needed + LGPU_TPB - 1) / LGPU_TPB;
    int per_sm = 8;
    if (kernel != nullptr) {
        auto it = g_occ.find(kernel);
        if (it == g_occ.end()) {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, LGPU_TPB, 0) != cudaSuccess || occ < 1) occ = 4;
            if (occ > 8) occ = 8;
            it = g_occ.emplace(kernel, occ).first;
        }
        per_sm = it->second;
    }
    int64_t cap = (int64_t)ctx->num_sms * per_sm;
    if (cap > LGPU_MAX_PARTIAL_BLOCKS) cap = LGPU_MAX_PARTIAL_BLOCKS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

/* live timing of one launch: CUDA events on the launching stream around the kernel (only when enabled) */
static void prof_flush(lgpu_ctx *ctx)
{
    if (ctx->prof_recs.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &r : ctx->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ctx->prof_ms[r.cls] += (double)ms;
            ctx->prof_cnt[r.cls] += 1;
        }
        ctx->prof_free.push_back(r.a);
        ctx->prof_free.push_back(r.b);
    }
    ctx->prof_recs.clear();
}
struct Prof {
    lgpu_ctx *ctx;
    cudaEvent_t b = nullptr;
    Prof(lgpu_ctx *c, int cls) : ctx(c)
    {
        ctx->launches++;
        if (!ctx->prof) return;
        if (ctx->prof_recs.size() >= 16384) prof_flush(ctx);
        cudaEvent_t ev[2];
        for (int k = 0; k < 2; ++k) {
            if (!ctx->prof_free.empty()) { ev[k] = ctx->prof_free.back(); ctx->prof_free.pop_back(); }
            else cudaEventCreate(&ev[k]);
        }
        cudaEventRecord(ev[0], ctx->stream);
        b = ev[1];
        ctx->prof_recs.push_back({cls, ev[0], ev[1]});
    }
    ~Prof() { if (b) cudaEventRecord(b, ctx->stream); }
};

template <class F>
static void launch_map(lgpu_ctx *ctx, int64_t n, F f)
{
    if (n <= 0) return;
    Prof pr(ctx, KC_VEC);
    k_map<<<grid_for(ctx, n, (const void *)k_map<F>), LGPU_TPB, 0, ctx->stream>>>(n, f);
}
template <int K> static int allreduce_spec(lgpu_ctx *ctx, const SlotSpec<K> &sp);
template <class F> static void launch_scalar(lgpu_ctx *ctx, F f);
/* fused elementwise + reduction pass.  One GPU: the finishing thread also forms the derived scalar (`post`).
 * Partitioned: local sums, NCCL all-reduce of the slots in stream, then `post` as a one-thread kernel. */
template <int K, class F, class P>
static void launch_reduce_post(lgpu_ctx *ctx, int64_t n, F f, SlotSpec<K> spec, P post, int cls = KC_REDUCE)
{
    if (ctx->world > 1 && !ctx->defer_allreduce && !ctx->cone_par) { /* by-cone mode: flat vectors are replicated */
        {
            Prof pr(ctx, cls);
            k_reduce<K, F, NoPost><<<grid_for(ctx, n > 0 ? n : 1, (const void *)k_reduce<K, F, NoPost>), LGPU_TPB, 0, ctx->stream>>>(
                n, f, ctx->partials, ctx->counter, ctx->dsc, spec, NoPost());
        }
        if (allreduce_spec<K>(ctx, spec) != 0) return;
        if (!std::is_same<P, NoPost>::value) {
            double *dsc = ctx->dsc;
            launch_scalar(ctx, [=] __device__() { post(dsc); });
        }
        return;
    }
    Prof pr(ctx, cls);
    k_reduce<K, F, P><<<grid_for(ctx, n > 0 ? n : 1, (const void *)k_reduce<K, F, P>), LGPU_TPB, 0, ctx->stream>>>(
        n, f, ctx->partials, ctx->counter, ctx->dsc, spec, post);
}
template <int K, class F>
static void launch_reduce(lgpu_ctx *ctx, int64_t n, F f, SlotSpec<K> spec)
{
    launch_reduce_post<K>(ctx, n, f, spec, NoPost());
}
template <class F>
static void launch_scalar(lgpu_ctx *ctx, F f)
{
    Prof pr(ctx, KC_SCALAR);
    k_scalar<<<1, 32, 0, ctx->stream>>>(f);
}
static SlotSpec<1> slot1(int s, int acc = 0);
static SlotSpec<1> slot1(int s, int acc)
{
    SlotSpec<1> sp;
    sp.slot[0] = s;
    sp.accumulate = acc;
    return sp;
}

/* Scalars back to the host.  The plain way -- a device-to-host copy through the copy engine plus a stream synchronise --
 * costs 15-20 us per read-back, and an ALM inner iteration has two of them; on the latency-bound instances (n <= 2e4) that
 * was a quarter of the iteration.  Fast path: a one-warp kernel at the end of the stream stores the values into the pinned,
 * device-mapped mirror (ctx->hsc) and then a sequence number (system-wide fence in between); the host spins on the number.
 * A stream error is noticed by polling cudaStreamQuery now and then.  LORADS_FAST_FETCH=0 selects the plain way. */
__global__ void __launch_bounds__(32) k_publish(const double *__restrict__ src, double *dst_host, int count,
                                                unsigned long long *flag_host, unsigned long long seq)
{
    for (int k = threadIdx.x; k < count; k += 32) dst_host[k] = src[k];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *(volatile unsigned long long *)flag_host = seq;
}
static int fetch_scalars(lgpu_ctx *ctx, int first, int count)
{
    if (ctx->fast_fetch && ctx->hsc_dev != nullptr) {
        const unsigned long long seq = ++ctx->fetch_seq;
        k_publish<<<1, 32, 0, ctx->stream>>>(ctx->dsc + first, ctx->hsc_dev + first, count, ctx->hflag_dev, seq);
        volatile unsigned long long *flag = ctx->hflag;
        for (unsigned long long spins = 1;; ++spins) {
            if (*flag == seq) break;
            if ((spins & 0x3fff) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q == cudaSuccess) { if (*flag == seq) break; }
                else if (q != cudaErrorNotReady) LGPU_FAIL(ctx, "%s:%d stream failed while waiting for scalars: %s", __FILE__, __LINE__, cudaGetErrorString(q));
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE); /* the values were stored before the number: read them after it */
        return 0;
    }
    CU(ctx, cudaMemcpyAsync(ctx->hsc + first, ctx->dsc + first, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * NCCL (row-block partitioned runs).  The library is bound at run time so that single-GPU use has no NCCL
 * dependency; only the few entry points used here are declared.
 * ------------------------------------------------------------------------------------------------*/
#include <dlfcn.h>
typedef struct { char internal[128]; } lg_ncclUniqueId;
typedef void *lg_ncclComm_t;
enum { LG_NCCL_FLOAT64 = 8, LG_NCCL_SUM = 0 };
static struct {
    bool ready = false;
    int (*GetUniqueId)(lg_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(lg_ncclComm_t *, int, lg_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(lg_ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, lg_ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, lg_ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, lg_ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, lg_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} g_nccl;

static bool nccl_load(std::string &err)
{
    if (g_nccl.ready) return true;
    const char *cands[] = {getenv("LORADS_NCCL_LIB"), "libnccl.so.2",
#ifdef LORADS_NCCL_PATH
                           LORADS_NCCL_PATH,
#endif
                           "libnccl.so"};
    void *h = nullptr;
    for (const char *cnd : cands) {
        if (!cnd) continue;
        h = dlopen(cnd, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { err = "cannot load libnccl.so.2 (set LORADS_NCCL_LIB)"; return false; }
    *(void **)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void **)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void **)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void **)&g_nccl.AllReduce = dlsym(h, "ncclAllReduce");
    *(void **)&g_nccl.AllGather = dlsym(h, "ncclAllGather");
    *(void **)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
    *(void **)&g_nccl.Send = dlsym(h, "ncclSend");
    *(void **)&g_nccl.Recv = dlsym(h, "ncclRecv");
    *(void **)&g_nccl.GroupStart = dlsym(h, "ncclGroupStart");
    *(void **)&g_nccl.GroupEnd = dlsym(h, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.Send ||
        !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        err = "libnccl lacks a required symbol";
        return false;
    }
    g_nccl.ready = true;
    return true;
}
#define NC(ctx, call)                                                                                   \
    do {                                                                                                \
        int _r = (call);                                                                                \
        if (_r != 0) LGPU_FAIL(ctx, "%s:%d NCCL error %s: %s", __FILE__, __LINE__, #call,               \
                               g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                \
    } while (0)

/* sum dsc[first .. first+count) over the ranks, in stream (no-op on one GPU) */
static int allreduce_scalars(lgpu_ctx *ctx, int first, int count)
{
    if (ctx->world <= 1) return 0;
    if (ctx->peer && count <= LGPU_PEER_RED) {
        PeerRed pr;
        pr.world = ctx->world;
        pr.rank = ctx->rank;
        for (int q = 0; q < ctx->world; ++q) pr.blk[q] = ctx->peer_blk[q];
        Prof pf(ctx, KC_SCALAR);
        k_peer_allreduce<<<1, LGPU_MAX_WORLD * LGPU_PEER_RED, 0, ctx->stream>>>(ctx->dsc, first, count, pr, ++ctx->aseq);
        return 0;
    }
    NC(ctx, g_nccl.AllReduce(ctx->dsc + first, ctx->dsc + first, (size_t)count, LG_NCCL_FLOAT64, LG_NCCL_SUM,
                             (lg_ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}
template <int K>
static int allreduce_spec(lgpu_ctx *ctx, const SlotSpec<K> &sp)
{
    if (ctx->world <= 1) return 0;
    int lo = sp.slot[0], hi = sp.slot[0];
    for (int k = 1; k < K; ++k) { lo = std::min(lo, sp.slot[k]); hi = std::max(hi, sp.slot[k]); }
    if (hi - lo + 1 == K) return allreduce_scalars(ctx, lo, K);
    for (int k = 0; k < K; ++k) TRY(allreduce_scalars(ctx, sp.slot[k], 1));
    return 0;
}

/* ---- peer-memory mapping (CUDA IPC) ------------------------------------------------------------------------
 * A record per rank (IPC handle of an allocation + the PCI bus id of its GPU) is all-gathered through the NCCL
 * communicator; every rank opens the others' handles.  Peer mode needs: one GPU per rank (distinct bus ids), peer
 * access between all pairs, and every handle opening -- the decision is all-reduced so that either every rank uses the
 * peer path or none does.  LORADS_PEER=0 keeps the NCCL path. */
struct PeerRecord {
    cudaIpcMemHandle_t handle; /* 64 bytes */
    char bus[32];
    int ok;
    int pad;
    long long rows; /* halo rows of the exporting rank */
    long long pad2[2];
};
static int peer_allgather_records(lgpu_ctx *ctx, const PeerRecord &mine, std::vector<PeerRecord> &all)
{
    const int P = ctx->world;
    all.assign(P, PeerRecord());
    PeerRecord *d = nullptr;
    CU(ctx, cudaMalloc((void **)&d, sizeof(PeerRecord) * P));
    CU(ctx, cudaMemcpyAsync(d + ctx->rank, &mine, sizeof(PeerRecord), cudaMemcpyHostToDevice, ctx->stream));
    NC(ctx, g_nccl.AllGather(d + ctx->rank, d, sizeof(PeerRecord), /* ncclInt8 */ 0, (lg_ncclComm_t)ctx->comm, ctx->stream));
    CU(ctx, cudaMemcpyAsync(all.data(), d, sizeof(PeerRecord) * P, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return 0;
}
/* every rank passes its local verdict; returns the conjunction (NCCL sum of the refusals) */
static int peer_all_agree(lgpu_ctx *ctx, bool mine, bool *all)
{
    ctx->hsc[SC_TMP] = mine ? 0.0 : 1.0;
    CU(ctx, cudaMemcpyAsync(ctx->dsc + SC_TMP, ctx->hsc + SC_TMP, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NC(ctx, g_nccl.AllReduce(ctx->dsc + SC_TMP, ctx->dsc + SC_TMP, 1, LG_NCCL_FLOAT64, LG_NCCL_SUM, (lg_ncclComm_t)ctx->comm,
                             ctx->stream));
    TRY(fetch_scalars(ctx, SC_TMP, 1));
    *all = ctx->hsc[SC_TMP] < 0.5;
    return 0;
}
static int peer_setup(lgpu_ctx *ctx)
{
    ctx->peer = false;
    const int P = ctx->world;
    bool want = P > 1 && P <= LGPU_MAX_WORLD;
    if (const char *v = getenv("LORADS_PEER")) want = want && atoi(v) != 0;
    PeerRecord mine;
    memset(&mine, 0, sizeof(mine));
    if (want) {
        if (cudaMalloc((void **)&ctx->blk, sizeof(PeerBlock)) != cudaSuccess) { cudaGetLastError(); want = false; ctx->blk = nullptr; }
        else {
            cudaMemset(ctx->blk, 0, sizeof(PeerBlock));
            if (cudaIpcGetMemHandle(&mine.handle, ctx->blk) != cudaSuccess) { cudaGetLastError(); want = false; }
        }
        if (cudaDeviceGetPCIBusId(mine.bus, sizeof(mine.bus), ctx->device) != cudaSuccess) { cudaGetLastError(); want = false; }
    }
    mine.ok = want ? 1 : 0;
    std::vector<PeerRecord> all;
    TRY(peer_allgather_records(ctx, mine, all));
    bool ok = want;
    for (int q = 0; q < P && ok; ++q) {
        if (!all[q].ok) ok = false;
        for (int s = 0; s < q && ok; ++s)
            if (strncmp(all[q].bus, all[s].bus, sizeof(mine.bus)) == 0) ok = false; /* two ranks on one GPU: spinning kernels may not co-run */
    }
    for (int q = 0; q < P && ok; ++q) {
        if (q == ctx->rank) { ctx->peer_blk[q] = ctx->blk; continue; }
        int dev = -1, can = 0;
        if (cudaDeviceGetByPCIBusId(&dev, all[q].bus) != cudaSuccess || cudaDeviceCanAccessPeer(&can, ctx->device, dev) != cudaSuccess || !can) {
            cudaGetLastError();
            ok = false;
            break;
        }
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        ctx->peer_blk[q] = (PeerBlock *)ptr;
    }
    bool every = false;
    TRY(peer_all_agree(ctx, ok, &every));
    ctx->peer = every;
    if (every && ctx->put_counter == nullptr) {
        CU(ctx, cudaMalloc((void **)&ctx->put_counter, sizeof(unsigned int)));
        CU(ctx, cudaMemset(ctx->put_counter, 0, sizeof(unsigned int)));
    }
    if (getenv("LORADS_PEER_VERBOSE") && ctx->rank == 0)
        fprintf(stderr, "lorads_b200: peer-memory exchange %s (%d ranks)\n", every ? "enabled" : "unavailable, using NCCL", P);
    return 0;
}
static int peer_unmap_halo(lgpu_ctx *ctx, bool barrier)
{
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int q = 0; q < ctx->world; ++q) {
        if (q != ctx->rank && ctx->peer_halo[q]) cudaIpcCloseMemHandle(ctx->peer_halo[q]);
        ctx->peer_halo[q] = nullptr;
    }
    if (!barrier) return 0;
    bool all = false;
    return peer_all_agree(ctx, true, &all); /* barrier: nobody frees a buffer a peer still maps */
}
/* (re)map the halo allocations of all ranks: called by every rank whenever the halo is (re)allocated */
static int peer_map_halo(lgpu_ctx *ctx)
{
    const int P = ctx->world;
    PeerRecord mine;
    memset(&mine, 0, sizeof(mine));
    mine.ok = cudaIpcGetMemHandle(&mine.handle, ctx->halo) == cudaSuccess ? 1 : 0;
    if (!mine.ok) cudaGetLastError();
    mine.rows = ctx->halo_rows;
    std::vector<PeerRecord> all;
    TRY(peer_allgather_records(ctx, mine, all));
    for (int q = 0; q < P; ++q) {
        if (!all[q].ok) LGPU_FAIL(ctx, "rank %d could not export its halo buffer (CUDA IPC)", q);
        ctx->peer_halo_rows[q] = all[q].rows;
        if (q == ctx->rank) { ctx->peer_halo[q] = ctx->halo; continue; }
        void *ptr = nullptr;
        CU(ctx, cudaIpcOpenMemHandle(&ptr, all[q].handle, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_halo[q] = (double *)ptr;
    }
    return 0;
}

extern "C" int lgpu_partition_rows(int64_t n, int world, int rank, int64_t *lo, int64_t *hi, int64_t *rows_per_rank)
{
    if (n <= 0 || world <= 0 || rank < 0 || rank >= world) return 1;
    const int64_t rpr = (n + world - 1) / world;
    *rows_per_rank = rpr;
    *lo = std::min<int64_t>((int64_t)rank * rpr, n);
    *hi = std::min<int64_t>(*lo + rpr, n);
    return 0;
}
extern "C" int lgpu_nccl_unique_id(unsigned char id[128])
{
    std::string err;
    if (!nccl_load(err)) return 1;
    lg_ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != 0) return 1;
    memcpy(id, u.internal, 128);
    return 0;
}
extern "C" int lgpu_comm_init(lgpu_ctx *ctx, const unsigned char id[128], int rank, int world)
{
    if (!ctx || world < 1 || rank < 0 || rank >= world) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (world == 1) { ctx->rank = 0; ctx->world = 1; return 0; }
    std::string err;
    if (!nccl_load(err)) LGPU_FAIL(ctx, "%s", err.c_str());
    lg_ncclUniqueId u;
    memcpy(u.internal, id, 128);
    lg_ncclComm_t comm = nullptr;
    NC(ctx, g_nccl.CommInitRank(&comm, world, u, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return peer_setup(ctx);
}

extern "C" int lgpu_uses_peer_exchange(const lgpu_ctx *ctx) { return (ctx && ctx->peer) ? 1 : 0; }
/* by-cone partition: largest cone first onto the least loaded rank (ties: lower cone index first, lower rank first) */
extern "C" int lgpu_cone_owner_map(int ncones, const double *cost, int world, int *owner)
{
    if (ncones < 0 || world < 1 || (ncones > 0 && (!cost || !owner))) return 1;
    std::vector<int> order(ncones);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(world, 0.0);
    for (int k : order) {
        const int q = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        owner[k] = q;
        load[q] += cost[k];
    }
    return 0;
}
extern "C" int lgpu_agree_flag(lgpu_ctx *ctx, int *flag)
{
    if (!ctx || !flag) return 1;
    if (ctx->world <= 1) return 0;
    CU(ctx, cudaSetDevice(ctx->device));
    ctx->hsc[SC_TMP] = (ctx->rank == 0 && *flag) ? 1.0 : 0.0;
    CU(ctx, cudaMemcpyAsync(ctx->dsc + SC_TMP, ctx->hsc + SC_TMP, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(allreduce_scalars(ctx, SC_TMP, 1));
    TRY(fetch_scalars(ctx, SC_TMP, 1));
    *flag = ctx->hsc[SC_TMP] > 0.5 ? 1 : 0;
    return 0;
}

/* constraints this rank owns: all of them on one GPU; in a partitioned run those attached to its rows.  m-vector
 * kernels iterate t in [0, count) and touch entry k = gid ? gid[t] : t, so non-owned entries are never written. */
struct Owned {
    int64_t count;
    const int32_t *gid;
};
static Owned owned(const lgpu_ctx *ctx)
{
    if (ctx->world > 1 && !ctx->cone_par) return {ctx->cones[0].m_loc, ctx->cones[0].rc_gid};
    return {ctx->m, nullptr};
}
static inline int pick_group(int64_t ld)
{
    const int64_t words = ld / 2;
    if (words <= 4) return 4;
    if (words <= 8) return 8;
    if (words <= 16) return 16;
    return 32;
}
static inline int pick_list_group(double avg_len)
{
    if (avg_len <= 2.0) return 1;
    if (avg_len <= 12.0) return 4;
    return 32;
}

#define DISPATCH_G(G, ...)              \
    switch (G) {                        \
    case 1: { constexpr int GG = 1; __VA_ARGS__; } break;   \
    case 4: { constexpr int GG = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int GG = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int GG = 16; __VA_ARGS__; } break; \
    default: { constexpr int GG = 32; __VA_ARGS__; } break; \
    }

static int run_long_rows(lgpu_ctx *ctx, DevCone &c, int64_t ld, const int32_t *slots, const double *vals, const double *Xin,
                         const double *Xhalo, int nsplit, double alpha, double beta, const double *Z, double *T);

/* UVt on the pattern of cone c from two row-major factors */
static void run_uvt(lgpu_ctx *ctx, DevCone &c, int64_t ld, const double *U, const double *V, double *out)
{
    const int same = (U == V) ? 1 : 0;
    if (c.dense_aggregate && ctx->dense_dmma) {
        /* the pattern is the whole triangle: (U V^T + V U^T)/2 as a rank-2k update on the FP64 tensor pipe */
        const int64_t nt = (c.n + 15) / 16, ntiles = nt * (nt + 1) / 2;
        int64_t blocks = (ntiles + 3) / 4;
        const int64_t cap = (int64_t)ctx->num_sms * 8;
        if (blocks > cap) blocks = cap;
        Prof pr(ctx, KC_DENSE);
        k_dense_uvt<<<(unsigned)blocks, 128, 0, ctx->stream>>>(c.n, (int)ld, U, V, same, out);
        return;
    }
    const int G = pick_group(ld);
    Prof pr(ctx, KC_UVT);
    DISPATCH_G(G, k_uvt<GG><<<grid_for(ctx, c.nnzP * GG, (const void *)k_uvt<GG>), LGPU_TPB, 0, ctx->stream>>>(
                      c.nnzP, c.pat_row, c.pat_col, U, V, (int)ld, same, out));
}
/* cv = A_c(uvt) (compact, per non-zero constraint) */
static void run_con_gather(lgpu_ctx *ctx, DevCone &c, const double *uvt, double *cv)
{
    if (c.mA == 0) return;
    const int G = c.con_group;
    {
        Prof pr(ctx, KC_GATHER);
        DISPATCH_G(G, k_con_gather<GG><<<grid_for(ctx, c.mA * GG), LGPU_TPB, 0, ctx->stream>>>(
                          c.mA, c.a_ptr, c.a_slot, c.a_coef, uvt, cv));
    }
    if (c.n_long_con > 0) {
        Prof pr(ctx, KC_GATHER);
        k_con_gather_long<<<(int)std::min<int64_t>(c.n_long_con, (int64_t)ctx->num_sms * 8), LGPU_TPB, 0, ctx->stream>>>(
            c.n_long_con, c.long_con, c.a_ptr, c.a_slot, c.a_coef, uvt, cv);
    }
}
/* <C, uvt> accumulated into dsc[slot] */
static void run_obj_gather(lgpu_ctx *ctx, DevCone &c, const double *uvt, int slot, int accumulate)
{
    const int32_t *cs = c.c_slot;
    const double *cc = c.c_coef;
    launch_reduce<1>(ctx, c.nnzC, [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = fma(cc[i], uvt[cs[i]], acc[0]); },
                     slot1(slot, accumulate));
}
/* S = [C] + sum w A ; w indexed by global id (w_global) or compact id */
static void run_wsum(lgpu_ctx *ctx, DevCone &c, const double *w, bool w_global, bool add_obj, double wscale, double *S)
{
    const int G = pick_list_group(c.nnzP > 0 ? (double)c.nnzA / (double)c.nnzP : 0.0);
    const int32_t *tidx = w_global ? c.t_gid : c.t_loc;
    Prof pr(ctx, KC_WSUM);
    DISPATCH_G(G, k_wsum<GG><<<grid_for(ctx, c.nnzP * GG), LGPU_TPB, 0, ctx->stream>>>(
                      c.nnzP, c.t_ptr, tidx, c.t_val, w, c.cval, add_obj ? 1 : 0, wscale, S));
}
/* Y = alpha S X + beta Z */
static void run_spmm(lgpu_ctx *ctx, DevCone &c, int64_t ld, const double *S, const double *X, double alpha, double beta,
                     const double *Z, double *Y)
{
    if (c.dense_aggregate && ctx->dense_dmma) {
        /* enough warps to fill the GPU: (row tiles) x (column tiles) x (k ranges), k split in multiples of 4 */
        const int64_t base = ((c.n + 7) / 8) * ((ld + 7) / 8);
        int ksplit = 1;
        while (ksplit < 16 && base * ksplit < (int64_t)ctx->num_sms * 32 && c.n / (ksplit * 2) >= 64) ksplit *= 2;
        const int64_t klen = (((c.n + ksplit - 1) / ksplit) + 3) / 4 * 4;
        const int64_t ntiles = base * ksplit;
        int64_t blocks = (ntiles + 3) / 4;
        const int64_t cap = (int64_t)ctx->num_sms * 16;
        if (blocks > cap) blocks = cap;
        double *scratch = nullptr;
        if (ksplit > 1) {
            const int64_t need = (int64_t)ksplit * c.n * ld;
            if (c.symm_scratch_len < need) {
                dev_free(c.symm_scratch);
                if (dev_alloc(ctx, &c.symm_scratch, (size_t)need) != 0) return;
                c.symm_scratch_len = need;
            }
            scratch = c.symm_scratch;
        }
        {
            Prof pr(ctx, KC_DENSE);
            k_dense_symm<<<(unsigned)blocks, 128, 0, ctx->stream>>>(c.n, (int)ld, ksplit, klen, S, X, alpha, beta, Z, Y, scratch);
        }
        if (ksplit > 1) {
            Prof pr(ctx, KC_DENSE);
            k_dense_symm_finish<<<grid_for(ctx, c.n * ld, (const void *)k_dense_symm_finish), LGPU_TPB, 0, ctx->stream>>>(
                c.n * ld, c.n * ld, ksplit, scratch, alpha, beta, Z, Y);
        }
        return;
    }
    const int G = pick_group(ld);
    Prof pr(ctx, KC_SPMM);
    DISPATCH_G(G, k_spmm<GG><<<grid_for(ctx, c.n * GG, (const void *)k_spmm<GG>), LGPU_TPB, 0, ctx->stream>>>(
                      c.n, c.f_ptr, c.f_col, c.f_slot, S, X, (int)ld, alpha, beta, Z, Y));
    run_long_rows(ctx, c, ld, c.f_slot, S, X, nullptr, 0, alpha, beta, Z, Y);
}

/* ------------------------------------------------------------------------------------------------
 * lifecycle
 * ------------------------------------------------------------------------------------------------*/
extern "C" const char *lgpu_version(void) { return "lorads_b200 0.1 (sm_100a)"; }

extern "C" const char *lgpu_last_error(const lgpu_ctx *ctx)
{
    if (!ctx) return g_create_err.c_str();
    return ctx->err.c_str();
}
extern "C" int64_t lgpu_launch_count(const lgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lgpu_create(lgpu_ctx **out, int device)
{
    /* load every kernel of the library when the context is created instead of at each kernel's first launch: lazy loading
     * put 10-20 ms of one-off work inside the timed solve of the small instances (only effective if CUDA is not yet
     * initialised in this process, i.e. in the drop-in binary; harmless otherwise) */
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                       " (liblorads_b200 has no CPU fallback)";
        return 1;
    }
    if (device < 0 || device >= ndev) {
        g_create_err = "device index out of range";
        return 1;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return 1;
    }
    lgpu_ctx *ctx = new lgpu_ctx();
    ctx->device = device;
    if (const char *v = getenv("LORADS_STEP_VARIANT")) ctx->step_variant = atoi(v);
    if (const char *v = getenv("LORADS_FAST_FETCH")) ctx->fast_fetch = atoi(v) != 0;
    if (const char *v = getenv("LORADS_ROWDOTS")) { ctx->rowdots_enabled = atoi(v) != 0; ctx->rowdots_force = atoi(v) >= 2; }
    if (const char *v = getenv("LORADS_STEP_BULK")) ctx->step_bulk = atoi(v);
    if (const char *v = getenv("LORADS_FUSE_PUT")) ctx->fuse_put = atoi(v) != 0;
    if (const char *v = getenv("LORADS_STEP_TILE")) ctx->step_tile_rows = atoi(v);
    if (const char *v = getenv("LORADS_STEP_STAGES")) ctx->step_stages = atoi(v);
    if (const char *v = getenv("LORADS_SPMM_DOT")) ctx->spmm_dot = atoi(v);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&ctx->dsc, sizeof(double) * LGPU_NSCALAR) != cudaSuccess ||
        cudaHostAlloc((void **)&ctx->hsc, sizeof(double) * (LGPU_NSCALAR + 8), cudaHostAllocMapped) != cudaSuccess ||
        cudaMalloc((void **)&ctx->partials, sizeof(double) * LGPU_MAX_REDUCE * LGPU_MAX_PARTIAL_BLOCKS) != cudaSuccess ||
        cudaMalloc((void **)&ctx->counter, sizeof(unsigned int)) != cudaSuccess) {
        g_create_err = std::string("context allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return 1;
    }
    cudaMemsetAsync(ctx->dsc, 0, sizeof(double) * LGPU_NSCALAR, ctx->stream);
    cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream);
    memset(ctx->hsc, 0, sizeof(double) * (LGPU_NSCALAR + 8));
    /* the mirror as the device sees it, and the read-back sequence number behind it (fetch_scalars fast path) */
    ctx->hflag = reinterpret_cast<unsigned long long *>(ctx->hsc + LGPU_NSCALAR);
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, ctx->hsc, 0) == cudaSuccess && dp != nullptr) {
        ctx->hsc_dev = (double *)dp;
        ctx->hflag_dev = reinterpret_cast<unsigned long long *>(ctx->hsc_dev + LGPU_NSCALAR);
    } else {
        cudaGetLastError();
        ctx->hsc_dev = nullptr;
    }
    cudaStreamSynchronize(ctx->stream);
    *out = ctx;
    return 0;
}

/* ---- timing / profiling hooks (bench.py) ----------------------------------------------------------*/
extern "C" int lgpu_sync(lgpu_ctx *ctx)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int lgpu_timer_record(lgpu_ctx *ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= 8) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!ctx->timers[slot]) CU(ctx, cudaEventCreate(&ctx->timers[slot]));
    CU(ctx, cudaEventRecord(ctx->timers[slot], ctx->stream));
    return 0;
}
extern "C" int lgpu_timer_elapsed_ms(lgpu_ctx *ctx, int a, int b, double *ms)
{
    if (!ctx || a < 0 || a >= 8 || b < 0 || b >= 8 || !ctx->timers[a] || !ctx->timers[b]) return 1;
    CU(ctx, cudaEventSynchronize(ctx->timers[b]));
    float f = 0.f;
    CU(ctx, cudaEventElapsedTime(&f, ctx->timers[a], ctx->timers[b]));
    *ms = (double)f;
    return 0;
}
extern "C" int lgpu_profile_enable(lgpu_ctx *ctx, int on)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    prof_flush(ctx);
    for (int k = 0; k < KC_COUNT; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_cnt[k] = 0; }
    ctx->prof = on != 0;
    return 0;
}
extern "C" int lgpu_profile_read(lgpu_ctx *ctx, int ncls, double *ms, int64_t *count)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    prof_flush(ctx);
    for (int k = 0; k < ncls; ++k) {
        ms[k] = k < KC_COUNT ? ctx->prof_ms[k] : 0.0;
        count[k] = k < KC_COUNT ? ctx->prof_cnt[k] : 0;
    }
    return 0;
}
extern "C" int lgpu_profile_num_classes(void) { return KC_COUNT; }
extern "C" const char *lgpu_profile_class_name(int cls)
{
    static const char *names[KC_COUNT] = {"k_uvt", "k_gather", "k_wsum", "k_spmm", "k_vec", "k_reduce", "k_scalar",
                                          "k_layout", "k_mc_spmm", "k_mc_step", "k_mc_dir", "k_dense_dmma", "k_exchange"};
    return (cls >= 0 && cls < KC_COUNT) ? names[cls] : "?";
}

static void free_cone(DevCone &c)
{
    dev_free(c.perm); dev_free(c.iperm); c.reordered = false; c.h_iperm.clear();
    dev_free(c.pat_row); dev_free(c.pat_col); dev_free(c.cval); dev_free(c.c_slot); dev_free(c.c_coef);
    dev_free(c.a_ptr); dev_free(c.a_slot); dev_free(c.a_coef); dev_free(c.con_gid); dev_free(c.long_con); c.n_long_con = 0;
    dev_free(c.t_ptr); dev_free(c.t_loc); dev_free(c.t_gid); dev_free(c.t_val);
    dev_free(c.f_ptr); dev_free(c.f_col); dev_free(c.f_slot); dev_free(c.d_row); dev_free(c.d_val);
    dev_free(c.long_rows); dev_free(c.long_first); dev_free(c.lw_row); dev_free(c.lw_beg); dev_free(c.lw_end);
    dev_free(c.long_scratch); c.n_long = c.n_lwork = c.long_scratch_ld = 0;
    dev_free(c.symm_scratch); c.symm_scratch_len = 0;
    dev_free(c.mc_val); dev_free(c.rc_ptr); dev_free(c.rc_gid); dev_free(c.rc_a);
    dev_free(c.uvt); dev_free(c.S); dev_free(c.cv); dev_free(c.wtmp);
}
static int peer_unmap_halo(lgpu_ctx *ctx, bool barrier);
/* collective = every rank is in the same call (alloc / aug_rank / set_problem): the peers' mappings of this rank's halo
 * are closed, with a barrier, before the buffer is freed */
static void free_vars(lgpu_ctx *ctx, bool collective = true)
{
    if (ctx->peer && ctx->halo != nullptr) peer_unmap_halo(ctx, collective);
    dev_free(ctx->R); dev_free(ctx->U); dev_free(ctx->V); dev_free(ctx->G); dev_free(ctx->M2); dev_free(ctx->bLin);
    dev_free(ctx->cg_r); dev_free(ctx->cg_p); dev_free(ctx->cg_Q); dev_free(ctx->stage);
    dev_free(ctx->CR); dev_free(ctx->CD); dev_free(ctx->rowdots); ctx->rowdots_valid = false; dev_free(ctx->gfull); dev_free(ctx->halo); dev_free(ctx->sendbuf);
    ctx->cr_valid = ctx->cd_valid = false;
    ctx->mc = false;
    for (auto &p : ctx->s) dev_free(p);
    for (auto &p : ctx->y) dev_free(p);
    ctx->s.clear();
    ctx->y.clear();
    ctx->vars_ready = false;
}

extern "C" void lgpu_destroy(lgpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &c : ctx->cones) free_cone(c);
    free_vars(ctx, false);
    for (int q = 0; q < ctx->world; ++q)
        if (ctx->peer && q != ctx->rank && ctx->peer_blk[q]) cudaIpcCloseMemHandle(ctx->peer_blk[q]);
    dev_free(ctx->blk); dev_free(ctx->put_counter); dev_free(ctx->put_dest);
    dev_free(ctx->b); dev_free(ctx->lam); dev_free(ctx->cvs); dev_free(ctx->q1); dev_free(ctx->q2); dev_free(ctx->M1);
    dev_free(ctx->mtmp);
    dev_free(ctx->lp.obj); dev_free(ctx->lp.r_ptr); dev_free(ctx->lp.r_col); dev_free(ctx->lp.r_val);
    dev_free(ctx->lp.c_ptr); dev_free(ctx->lp.c_row); dev_free(ctx->lp.c_val); dev_free(ctx->lp.nrm2sq);
    dev_free(ctx->send_idx); dev_free(ctx->halo_gid);
    dev_free(ctx->dsc); dev_free(ctx->partials); dev_free(ctx->counter);
    prof_flush(ctx);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((lg_ncclComm_t)ctx->comm);
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    for (auto &e : ctx->timers) if (e) cudaEventDestroy(e);
    if (ctx->hsc) cudaFreeHost(ctx->hsc);
    if (ctx->dstage) cudaFree(ctx->dstage);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static int ensure_dstage(lgpu_ctx *ctx, size_t bytes)
{
    if (bytes > ctx->dstage_bytes) {
        if (ctx->dstage) cudaFree(ctx->dstage);
        ctx->dstage = nullptr;
        ctx->dstage_bytes = 0;
        CU(ctx, cudaMalloc(&ctx->dstage, bytes));
        ctx->dstage_bytes = bytes;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * problem upload
 * ------------------------------------------------------------------------------------------------*/
extern "C" int lgpu_set_problem(lgpu_ctx *ctx, int64_t m, const double *b, int n_cones, const int64_t *blk_dims,
                                int64_t n_lp_cols)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (m <= 0 || n_cones < 0 || m >= (int64_t)1 << 31) LGPU_FAIL(ctx, "bad problem sizes");
    for (auto &c : ctx->cones) free_cone(c);
    ctx->cones.clear();
    ctx->cones.resize(n_cones);
    ctx->m = m;
    ctx->ncones = n_cones;
    for (int c = 0; c < n_cones; ++c) {
        if (blk_dims[c] <= 0 || blk_dims[c] >= (int64_t)1 << 31) LGPU_FAIL(ctx, "bad block dimension");
        ctx->cones[c].n = blk_dims[c];
    }
    ctx->lp = DevLp();
    ctx->lp.n = n_lp_cols;
    ctx->h_b.assign(b, b + m);
    dev_free(ctx->b); dev_free(ctx->lam); dev_free(ctx->cvs); dev_free(ctx->q1); dev_free(ctx->q2); dev_free(ctx->M1);
    dev_free(ctx->mtmp);
    TRY(dev_upload(ctx, &ctx->b, ctx->h_b));
    double **mv[] = {&ctx->lam, &ctx->cvs, &ctx->q1, &ctx->q2, &ctx->M1, &ctx->mtmp};
    for (auto p : mv) {
        TRY(dev_alloc(ctx, p, (size_t)m));
        CU(ctx, cudaMemsetAsync(*p, 0, sizeof(double) * m, ctx->stream));
    }
    /* |b| norms.  Quirk Q2 (lorads_solver.c:1468-1472): the reference indexes b with the 1-based result of
     * idamax_, i.e. it reads the element AFTER the first maximum; one past the end reads a heap word that is
     * never a meaningful double (treated as 0 here). */
    double n1 = 0, n2 = 0, mx = -1;
    int64_t imax = 0;
    for (int64_t i = 0; i < m; ++i) {
        n1 += fabs(b[i]);
        n2 += b[i] * b[i];
        if (fabs(b[i]) > mx) { mx = fabs(b[i]); imax = i; }
    }
    ctx->b_nrm1 = n1;
    ctx->b_nrm2 = sqrt(n2);
    ctx->b_nrminf_q = (imax + 1 < m) ? fabs(b[imax + 1]) : 0.0;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* work list of the CSR rows that k_spmm_long_chunks takes over */
static int upload_long_rows(lgpu_ctx *ctx, DevCone &c, const std::vector<int32_t> &ptr)
{
    std::vector<int32_t> lr, first(1, 0), wrow, wbeg, wend;
    for (size_t i = 0; i + 1 < ptr.size(); ++i) {
        if (ptr[i + 1] - ptr[i] <= LGPU_LONG_ROW) continue;
        lr.push_back((int32_t)i);
        for (int32_t e = ptr[i]; e < ptr[i + 1]; e += LGPU_LONG_CHUNK) {
            wrow.push_back((int32_t)i);
            wbeg.push_back(e);
            wend.push_back(std::min<int32_t>(e + LGPU_LONG_CHUNK, ptr[i + 1]));
        }
        first.push_back((int32_t)wrow.size());
    }
    c.n_long = (int64_t)lr.size();
    c.n_lwork = (int64_t)wrow.size();
    dev_free(c.long_rows); dev_free(c.long_first); dev_free(c.lw_row); dev_free(c.lw_beg); dev_free(c.lw_end);
    dev_free(c.long_scratch);
    c.long_scratch_ld = 0;
    if (!lr.empty()) {
        TRY(dev_upload(ctx, &c.long_rows, lr));
        TRY(dev_upload(ctx, &c.long_first, first));
        TRY(dev_upload(ctx, &c.lw_row, wrow));
        TRY(dev_upload(ctx, &c.lw_beg, wbeg));
        TRY(dev_upload(ctx, &c.lw_end, wend));
    }
    return 0;
}
/* the long rows' share of  T = alpha S X + beta Z  (entry values vals[e], or vals[slots[e]]) */
static int run_long_rows(lgpu_ctx *ctx, DevCone &c, int64_t ld, const int32_t *slots, const double *vals, const double *Xin,
                         const double *Xhalo, int nsplit, double alpha, double beta, const double *Z, double *T)
{
    if (c.n_long == 0) return 0;
    if (c.long_scratch_ld < ld) {
        dev_free(c.long_scratch);
        TRY(dev_alloc(ctx, &c.long_scratch, (size_t)c.n_lwork * (size_t)ld));
        c.long_scratch_ld = ld;
    }
    const int G = pick_group(ld);
    const int blocks = (int)std::min<int64_t>(c.n_lwork, (int64_t)ctx->num_sms * 8);
    {
        Prof pr(ctx, KC_SPMM);
        if (Xhalo != nullptr) {
            DISPATCH_G(G, k_spmm_long_chunks<GG, true><<<blocks, LGPU_TPB, 0, ctx->stream>>>(
                              c.n_lwork, c.lw_row, c.lw_beg, c.lw_end, c.f_col, slots, vals, Xin, Xhalo, nsplit, (int)ld, c.long_scratch));
        } else {
            DISPATCH_G(G, k_spmm_long_chunks<GG, false><<<blocks, LGPU_TPB, 0, ctx->stream>>>(
                              c.n_lwork, c.lw_row, c.lw_beg, c.lw_end, c.f_col, slots, vals, Xin, nullptr, 0, (int)ld, c.long_scratch));
        }
    }
    {
        Prof pr(ctx, KC_SPMM);
        k_spmm_long_finish<<<grid_for(ctx, c.n_long * ld, (const void *)k_spmm_long_finish), LGPU_TPB, 0, ctx->stream>>>(
            c.n_long, c.long_rows, c.long_first, (int)ld, c.long_scratch, alpha, beta, Z, T);
    }
    return 0;
}

/* the same facts lgpu_cone_info reports after an upload, computed without a context or a GPU */
extern "C" int lgpu_cone_classify(int64_t n, int64_t m, const int64_t *mat_beg, const int64_t *mat_idx, int64_t out[6])
{
    if (n <= 0 || m <= 0 || !mat_beg || !out) return 1;
    ConeRules r;
    cone_rules(n, m, mat_beg, mat_idx, r);
    out[0] = r.mA;
    out[1] = r.dense ? 1 : 0;
    out[2] = r.sparse_container ? 1 : 0;
    out[3] = r.nnzP;
    out[4] = r.diag_only ? 1 : 0;
    out[5] = r.nnzA;
    return 0;
}

struct lgpu_layout {
    ConeLayout L;
};
extern "C" int lgpu_cone_layout_build(lgpu_layout **out, int64_t n, int64_t m, const int64_t *mat_beg, const int64_t *mat_idx,
                                      const double *mat_elem, int world, int rank)
{
    if (!out || n <= 0 || m <= 0 || !mat_beg || world < 1 || rank < 0 || rank >= world) return 1;
    lgpu_layout *h = new lgpu_layout();
    std::string err;
    if (build_cone_layout(n, m, mat_beg, mat_idx, mat_elem, world, rank, true, h->L, err)) { delete h; return 2; }
    *out = h;
    return 0;
}
extern "C" void lgpu_cone_layout_free(lgpu_layout *layout) { delete layout; }
extern "C" int lgpu_cone_layout_get(const lgpu_layout *layout, const char *name, int64_t cap_bytes, void *out, int64_t *count,
                                    int *elem_bytes)
{
    if (!layout || !name || !count || !elem_bytes) return 1;
    const ConeLayout &L = layout->L;
    *count = 0;
    *elem_bytes = 0;
    const std::string nm(name);
    const void *src = nullptr;
    bool known = false;
    std::vector<double> sc;
#define LQ_ARRAY(field, bytes)                                   \
    if (nm == #field) {                                          \
        src = L.field.data();                                    \
        *count = (int64_t)L.field.size();                        \
        *elem_bytes = bytes;                                     \
        known = true;                                            \
    }
    LQ_ARRAY(pat_row, 4) LQ_ARRAY(pat_col, 4) LQ_ARRAY(cval, 8) LQ_ARRAY(c_slot, 4) LQ_ARRAY(c_coef, 8) LQ_ARRAY(a_ptr, 4)
    LQ_ARRAY(a_slot, 4) LQ_ARRAY(a_coef, 8) LQ_ARRAY(con_gid, 4) LQ_ARRAY(t_ptr, 4) LQ_ARRAY(t_loc, 4) LQ_ARRAY(t_gid, 4)
    LQ_ARRAY(t_val, 8) LQ_ARRAY(f_ptr, 4) LQ_ARRAY(f_col, 4) LQ_ARRAY(f_slot, 4) LQ_ARRAY(d_row, 4) LQ_ARRAY(d_val, 8)
    LQ_ARRAY(mc_val, 8) LQ_ARRAY(rc_ptr, 4) LQ_ARRAY(rc_gid, 4) LQ_ARRAY(rc_a, 8) LQ_ARRAY(lf_ptr, 4) LQ_ARRAY(lf_col, 4)
    LQ_ARRAY(lmc_val, 8) LQ_ARRAY(lrc_ptr, 4) LQ_ARRAY(lrc_gid, 4) LQ_ARRAY(lrc_a, 8) LQ_ARRAY(send_idx, 4) LQ_ARRAY(halo_gid, 4)
    LQ_ARRAY(send_off, -8) LQ_ARRAY(send_cnt, -8) LQ_ARRAY(recv_off, -8) LQ_ARRAY(recv_cnt, -8) LQ_ARRAY(dst_off, -8)
    LQ_ARRAY(perm, 4) LQ_ARRAY(iperm, 4) LQ_ARRAY(cperm, 4) LQ_ARRAY(dev_pat_row, 4) LQ_ARRAY(dev_pat_col, 4)
#undef LQ_ARRAY
    if (nm == "scalars") {
        sc = {(double)L.mA, (double)L.nnzP, (double)L.nnzA, (double)L.nnzC, (double)L.nnzF, (double)L.max_con_len,
              (double)L.max_slot_len, L.c_nrm1, L.c_nrm2sq, L.c_nrminf, L.dense ? 1.0 : 0.0, L.sparse_container ? 1.0 : 0.0,
              L.diag_only ? 1.0 : 0.0, L.use_halo ? 1.0 : 0.0, (double)L.halo_rows, (double)L.send_rows};
        src = sc.data();
        *count = (int64_t)sc.size();
        *elem_bytes = 8;
        known = true;
    }
    if (!known) return 3;
    if (out && src && *count > 0) {
        const int64_t bytes = *count * (int64_t)std::abs(*elem_bytes);
        memcpy(out, src, (size_t)std::min<int64_t>(bytes, cap_bytes));
    }
    return 0;
}

extern "C" int lgpu_cone_upload(lgpu_ctx *ctx, int cone, const int64_t *beg, const int64_t *idx_in, const double *val_in)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    const int64_t n = c.n_glob > 0 ? c.n_glob : c.n, m = ctx->m;
    /* all host-side preprocessing (storage rules, pattern, CSR forms, partition slices, exchange plan): lgpu_layout.h */
    ConeLayout L;
    {
        std::string err;
        if (build_cone_layout(n, m, beg, idx_in, val_in, ctx->world, ctx->rank, ctx->ncones == 1 && ctx->lp.n == 0, L, err, ctx->peer))
            LGPU_FAIL(ctx, "%s", err.c_str());
    }
    /* more than one GPU and not the single MaxCut-type cone the row-block partition handles: by-cone partition (every
     * rank keeps every cone and every vector; the operator work of a cone is done by its owner, see cone_* helpers) */
    if (ctx->world > 1 && !L.partitioned) ctx->cone_par = true;
    free_cone(c);
    c.obj_type = L.obj_type;
    c.mA = L.mA;
    c.sparse_container = L.sparse_container;
    c.dense_aggregate = L.dense;
    c.diag_only = L.diag_only;
    c.nnzP = L.nnzP;
    c.nnzA = L.nnzA;
    c.nnzC = L.nnzC;
    c.nnzF = L.nnzF;
    c.max_con_len = L.max_con_len;
    c.max_slot_len = L.max_slot_len;
    c.c_nrm1 = L.c_nrm1;
    c.c_nrm2sq = L.c_nrm2sq;
    c.c_nrminf = L.c_nrminf;
    c.h_pat_row.swap(L.pat_row);
    c.h_pat_col.swap(L.pat_col);
    c.reordered = L.reordered;
    c.window_before = L.window_hits_before;
    c.window_after = L.window_hits_after;
    ctx->h_cperm.clear();
    if (L.reordered) {
        TRY(dev_upload(ctx, &c.perm, L.perm));
        TRY(dev_upload(ctx, &c.iperm, L.iperm));
        c.h_iperm = L.iperm;
        /* the m-vectors live in the renumbered constraint order from here on (b was uploaded in the caller's order) */
        ctx->h_cperm = L.cperm;
        std::vector<double> bp((size_t)m);
        for (int64_t k = 0; k < m; ++k) bp[L.cperm[k]] = ctx->h_b[k];
        CU(ctx, cudaMemcpyAsync(ctx->b, bp.data(), sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    c.n = n;
    c.n_glob = n;
    c.row_lo = 0;
    c.n_alloc = n;
    c.m_loc = L.mA;
    if (L.partitioned) {
        ctx->send_off = L.send_off; ctx->send_cnt = L.send_cnt; ctx->recv_off = L.recv_off; ctx->recv_cnt = L.recv_cnt;
        ctx->halo_rows = L.halo_rows;
        ctx->send_rows = L.send_rows;
        ctx->use_halo = L.use_halo;
        ctx->dst_off = L.dst_off;
        dev_free(ctx->put_dest);
        if (ctx->peer && L.use_halo && ctx->world > 1) {
            /* per own row and peer: the row's place in that peer's halo, or -1 (k_mc_combine<PUT>) */
            const int np = ctx->world - 1;
            const int64_t nl = L.hi - L.lo;
            std::vector<int32_t> dest((size_t)nl * np, -1);
            for (int q = 0, j = 0; q < ctx->world; ++q) {
                if (q == ctx->rank) continue;
                for (int64_t k = 0; k < L.send_cnt[q]; ++k)
                    dest[(size_t)L.send_idx[L.send_off[q] + k] * np + j] = (int32_t)(L.dst_off[q] + k);
                ++j;
            }
            TRY(dev_upload(ctx, &ctx->put_dest, dest));
        }
        if (L.use_halo) {
            dev_free(ctx->send_idx);
            TRY(dev_upload(ctx, &ctx->send_idx, L.send_idx));
            dev_free(ctx->halo_gid);
            TRY(dev_upload(ctx, &ctx->halo_gid, L.halo_gid));
        }
        TRY(upload_long_rows(ctx, c, L.lf_ptr));
        TRY(dev_upload(ctx, &c.f_ptr, L.lf_ptr));
        TRY(dev_upload(ctx, &c.f_col, L.lf_col));
        TRY(dev_upload(ctx, &c.mc_val, L.lmc_val));
        TRY(dev_upload(ctx, &c.rc_ptr, L.lrc_ptr));
        TRY(dev_upload(ctx, &c.rc_gid, L.lrc_gid));
        TRY(dev_upload(ctx, &c.rc_a, L.lrc_a));
        c.n = L.hi - L.lo;
        c.row_lo = L.lo;
        c.n_alloc = L.rows_per_rank;
        c.m_loc = (int64_t)L.lrc_gid.size();
        c.nnzF = (int64_t)L.lf_col.size();
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    const int64_t mA = L.mA, nnzP = L.nnzP;
    TRY(dev_upload(ctx, &c.pat_row, L.reordered ? L.dev_pat_row : c.h_pat_row));
    TRY(dev_upload(ctx, &c.pat_col, L.reordered ? L.dev_pat_col : c.h_pat_col));
    TRY(dev_upload(ctx, &c.cval, L.cval));
    TRY(dev_upload(ctx, &c.c_slot, L.c_slot));
    TRY(dev_upload(ctx, &c.c_coef, L.c_coef));
    TRY(dev_upload(ctx, &c.a_ptr, L.a_ptr));
    TRY(dev_upload(ctx, &c.a_slot, L.a_slot));
    TRY(dev_upload(ctx, &c.a_coef, L.a_coef));
    TRY(dev_upload(ctx, &c.con_gid, L.con_gid));
    {
        /* lane group of the per-constraint gather from the average list length; lists far above it get their own CTA */
        c.con_group = pick_list_group(mA > 0 ? (double)c.nnzA / (double)mA : 0.0);
        std::vector<int32_t> lc;
        for (int64_t t = 0; t < mA; ++t)
            if (L.a_ptr[t + 1] - L.a_ptr[t] > LGPU_LONG_ROW * c.con_group) lc.push_back((int32_t)t);
        c.n_long_con = (int64_t)lc.size();
        if (!lc.empty()) TRY(dev_upload(ctx, &c.long_con, lc));
    }
    TRY(dev_upload(ctx, &c.t_ptr, L.t_ptr));
    TRY(dev_upload(ctx, &c.t_loc, L.t_loc));
    TRY(dev_upload(ctx, &c.t_gid, L.t_gid));
    TRY(dev_upload(ctx, &c.t_val, L.t_val));
    TRY(upload_long_rows(ctx, c, L.f_ptr));
    TRY(dev_upload(ctx, &c.f_ptr, L.f_ptr));
    TRY(dev_upload(ctx, &c.f_col, L.f_col));
    TRY(dev_upload(ctx, &c.f_slot, L.f_slot));
    if (L.diag_only) {
        TRY(dev_upload(ctx, &c.d_row, L.d_row));
        TRY(dev_upload(ctx, &c.d_val, L.d_val));
        TRY(dev_upload(ctx, &c.mc_val, L.mc_val));
        TRY(dev_upload(ctx, &c.rc_ptr, L.rc_ptr));
        TRY(dev_upload(ctx, &c.rc_gid, L.rc_gid));
        TRY(dev_upload(ctx, &c.rc_a, L.rc_a));
    }
    TRY(dev_alloc(ctx, &c.uvt, (size_t)nnzP));
    TRY(dev_alloc(ctx, &c.S, (size_t)nnzP));
    TRY(dev_alloc(ctx, &c.cv, (size_t)std::max<int64_t>(mA, 1)));
    TRY(dev_alloc(ctx, &c.wtmp, (size_t)std::max<int64_t>(mA, 1)));
    CU(ctx, cudaMemsetAsync(c.cv, 0, sizeof(double) * std::max<int64_t>(mA, 1), ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lgpu_lp_upload(lgpu_ctx *ctx, const int64_t *beg, const int64_t *idx, const double *val)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevLp &lp = ctx->lp;
    const int64_t m = ctx->m, n = lp.n;
    if (n <= 0) LGPU_FAIL(ctx, "no LP block declared in lgpu_set_problem");
    lp.h_obj.assign(n, 0.0);
    for (int64_t e = beg[0]; e < beg[1]; ++e) lp.h_obj[idx[e]] = val[e]; /* lorads_lp_conic.c:38-40 */
    std::vector<int32_t> r_ptr(m + 1, 0), r_col;
    std::vector<double> r_val;
    std::vector<int64_t> cnt(n, 0);
    for (int64_t i = 0; i < m; ++i) {
        for (int64_t e = beg[i + 1]; e < beg[i + 2]; ++e) {
            r_col.push_back((int32_t)idx[e]);
            r_val.push_back(val[e]);
            cnt[idx[e]]++;
        }
        r_ptr[i + 1] = (int32_t)r_col.size();
    }
    lp.nnz = (int64_t)r_col.size();
    lp.h_c_ptr.assign(n + 1, 0);
    for (int64_t j = 0; j < n; ++j) lp.h_c_ptr[j + 1] = lp.h_c_ptr[j] + (int32_t)cnt[j];
    lp.h_c_row.resize(lp.nnz);
    lp.h_c_val.resize(lp.nnz);
    {
        std::vector<int32_t> fill(lp.h_c_ptr.begin(), lp.h_c_ptr.end() - 1);
        for (int64_t i = 0; i < m; ++i)
            for (int32_t e = r_ptr[i]; e < r_ptr[i + 1]; ++e) {
                const int32_t p = fill[r_col[e]]++;
                lp.h_c_row[p] = (int32_t)i;
                lp.h_c_val[p] = r_val[e];
            }
    }
    lp.h_nrm2sq.assign(n, 0.0);
    for (int64_t j = 0; j < n; ++j) {
        double s = 0;
        for (int32_t e = lp.h_c_ptr[j]; e < lp.h_c_ptr[j + 1]; ++e) s += lp.h_c_val[e] * lp.h_c_val[e];
        const double t = sqrt(s); /* nrm2 then squared, lorads_lp_conic.c:112-113 */
        lp.h_nrm2sq[j] = t * t;
        if (lp.h_c_ptr[j + 1] == lp.h_c_ptr[j]) LGPU_FAIL(ctx, "LP column %lld has no constraint entry (the reference rejects it too)", (long long)j);
    }
    /* norms, with the reference's quirks (lorads_lp_conic.c:170-196): |c|_2^2 := |c|_1^2, |c|_inf off by one */
    double n1 = 0, mx = -1;
    int64_t imax = 0;
    for (int64_t j = 0; j < n; ++j) {
        n1 += fabs(lp.h_obj[j]);
        if (fabs(lp.h_obj[j]) > mx) { mx = fabs(lp.h_obj[j]); imax = j; }
    }
    lp.nrm1 = n1;
    lp.nrminf_q = (imax + 1 < n) ? fabs(lp.h_obj[imax + 1]) : 0.0;
    dev_free(lp.obj); dev_free(lp.r_ptr); dev_free(lp.r_col); dev_free(lp.r_val);
    dev_free(lp.c_ptr); dev_free(lp.c_row); dev_free(lp.c_val); dev_free(lp.nrm2sq);
    TRY(dev_upload(ctx, &lp.obj, lp.h_obj));
    TRY(dev_upload(ctx, &lp.r_ptr, r_ptr));
    TRY(dev_upload(ctx, &lp.r_col, r_col));
    TRY(dev_upload(ctx, &lp.r_val, r_val));
    TRY(dev_upload(ctx, &lp.c_ptr, lp.h_c_ptr));
    TRY(dev_upload(ctx, &lp.c_row, lp.h_c_row));
    TRY(dev_upload(ctx, &lp.c_val, lp.h_c_val));
    TRY(dev_upload(ctx, &lp.nrm2sq, lp.h_nrm2sq));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lgpu_cone_info(const lgpu_ctx *ctx, int cone, int64_t out[6])
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    const DevCone &c = ctx->cones[cone];
    out[0] = c.mA;
    out[1] = c.dense_aggregate ? 1 : 0;
    out[2] = c.sparse_container ? 1 : 0;
    out[3] = c.nnzP;
    out[4] = c.diag_only ? 1 : 0;
    out[5] = c.nnzA;
    return 0;
}

/* row relabelling for gather locality (lgpu_layout.h bfs_relabel): out = {applied (0/1), share of the CSR entries within
 * +-65536 rows of the diagonal before, after}; invisible at the ABI (factors go in and out in the caller's row order) */
extern "C" int lgpu_cone_reorder_info(const lgpu_ctx *ctx, int cone, double out[3])
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    const DevCone &c = ctx->cones[cone];
    out[0] = c.reordered ? 1.0 : 0.0;
    out[1] = c.window_before;
    out[2] = c.window_after;
    return 0;
}
extern "C" int lgpu_cone_pattern(const lgpu_ctx *ctx, int cone, int64_t cap, int32_t *row, int32_t *col)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    const DevCone &c = ctx->cones[cone];
    const int64_t k = std::min<int64_t>(cap, c.nnzP);
    memcpy(row, c.h_pat_row.data(), sizeof(int32_t) * k);
    memcpy(col, c.h_pat_col.data(), sizeof(int32_t) * k);
    return 0;
}

extern "C" int lgpu_constants(const lgpu_ctx *ctx, double out[6])
{
    if (!ctx) return 1;
    /* LORADSNrm1Obj / Nrm2Obj / NrmInfObj (lorads_solver.c:222-262, lorads_alg_common.c:481-497) */
    double n1 = 0, n2 = 0, ninf = 0;
    if (ctx->lp.n > 0) {
        n1 += ctx->lp.nrm1;
        n2 += ctx->lp.nrm1 * ctx->lp.nrm1;
        ninf = std::max(ninf, ctx->lp.nrminf_q);
    }
    for (const auto &c : ctx->cones) {
        n1 += c.c_nrm1;
        n2 += c.c_nrm2sq;
        ninf = std::max(ninf, c.c_nrminf);
    }
    out[0] = n1;
    out[1] = sqrt(n2);
    out[2] = ninf;
    out[3] = ctx->b_nrm1;
    out[4] = ctx->b_nrm2;
    out[5] = ctx->b_nrminf_q;
    return 0;
}

extern "C" int lgpu_obj_scale(lgpu_ctx *ctx, double s)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    for (auto &c : ctx->cones) {
        double *cv = c.cval, *cc = c.c_coef;
        if (cv) launch_map(ctx, c.nnzP, [=] __device__(int64_t i) { cv[i] *= s; }); /* (a partitioned rank keeps only mc_val) */
        if (cc) launch_map(ctx, c.nnzC, [=] __device__(int64_t i) { cc[i] *= s; });
        if (c.mc_val) {
            double *mv = c.mc_val;
            launch_map(ctx, c.nnzF, [=] __device__(int64_t i) { mv[i] *= s; });
        }
        c.c_nrm1 *= fabs(s);
        c.c_nrm2sq *= s * s;
        c.c_nrminf *= fabs(s);
    }
    if (ctx->lp.n > 0) {
        double *o = ctx->lp.obj;
        launch_map(ctx, ctx->lp.n, [=] __device__(int64_t i) { o[i] *= s; });
        for (auto &v : ctx->lp.h_obj) v *= s;
    }
    double *lam = ctx->lam;
    {
        const Owned ow = owned(ctx);
        const int32_t *gid = ow.gid;
        launch_map(ctx, ow.count, [=] __device__(int64_t t) { lam[gid ? gid[t] : t] *= s; });
    }
    ctx->cr_valid = ctx->cd_valid = false;
    CHECK_LAUNCH(ctx);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * variables
 * ------------------------------------------------------------------------------------------------*/
static int alloc_flat(lgpu_ctx *ctx, double **p)
{
    TRY(dev_alloc(ctx, p, (size_t)ctx->N));
    CU(ctx, cudaMemsetAsync(*p, 0, sizeof(double) * ctx->N, ctx->stream));
    return 0;
}

static int layout_and_alloc(lgpu_ctx *ctx, const int64_t *rank, int lbfgs_len)
{
    int64_t off = 0;
    for (int c = 0; c < ctx->ncones; ++c) {
        DevCone &cn = ctx->cones[c];
        if (cn.n_glob == 0) { cn.n_glob = cn.n; cn.n_alloc = cn.n; }
        if (rank[c] <= 0 || rank[c] > cn.n_glob) LGPU_FAIL(ctx, "bad rank %lld for cone %d", (long long)rank[c], c);
        cn.r = rank[c];
        cn.ld = (rank[c] + 3) & ~(int64_t)3;
        cn.off = off;
        off += cn.n_alloc * cn.ld;
    }
    ctx->lp.off = off;
    off += ctx->lp.n;
    ctx->N = off;
    double **flat[] = {&ctx->R, &ctx->U, &ctx->V, &ctx->G, &ctx->M2, &ctx->bLin, &ctx->cg_r, &ctx->cg_p, &ctx->cg_Q, &ctx->stage};
    for (auto p : flat) TRY(alloc_flat(ctx, p));
    ctx->mc = ctx->fast_enabled && ctx->ncones == 1 && ctx->lp.n == 0 && ctx->cones[0].diag_only &&
              ctx->cones[0].mA == ctx->m;
    ctx->cr_valid = ctx->cd_valid = false;
    ctx->cr_updates = 0;
    ctx->gram_valid = false;
    ctx->gram_pair_ok[0] = ctx->gram_pair_ok[1] = false;
    ctx->epi_done = false;
    if (ctx->world > 1 && !ctx->mc && !ctx->cone_par) LGPU_FAIL(ctx, "partitioned runs use the fused MaxCut-type path only");
    if (!ctx->mc && ctx->ncones > 0 && ctx->cones[0].reordered)
        LGPU_FAIL(ctx, "this cone was uploaded with the row relabelling of the fused path: keep the fused path on (or set LORADS_REORDER=0)");
    if (ctx->cone_par) {
        if (ctx->mc) LGPU_FAIL(ctx, "internal: by-cone partition with the fused single-cone path");
        /* owner of every cone: greedy balance of the per-iteration operator work ~ (nnz of the pattern + of the constraints
         * + rows) x padded rank; the same map on every rank (lgpu_cone_owner_map) */
        std::vector<double> cost(ctx->ncones);
        for (int k = 0; k < ctx->ncones; ++k) {
            const DevCone &d = ctx->cones[k];
            cost[k] = (double)(d.nnzP + d.nnzA + d.n) * (double)std::max<int64_t>(d.ld, 1);
        }
        ctx->cone_owner.assign(ctx->ncones, 0);
        lgpu_cone_owner_map(ctx->ncones, cost.data(), ctx->world, ctx->cone_owner.data());
    }
    ctx->rowdots_valid = false;
    if (ctx->mc) {
        TRY(alloc_flat(ctx, &ctx->CR));
        TRY(alloc_flat(ctx, &ctx->CD));
        /* carried row products pay where the factor streams from DRAM; on L2-resident problems they only add five
         * reductions and 40 bytes per row to a latency-bound pass (torus n = 2e4: step pass 23 -> 28 us), so they are kept for
         * factors of 64 MB and more (LORADS_ROWDOTS=2 forces them on for tests) */
        const bool big = (double)ctx->cones[0].n_alloc * (double)ctx->cones[0].ld * 8.0 >= 64.0e6;
        if (lbfgs_len == 2 && ctx->rowdots_enabled && (big || ctx->rowdots_force))
            TRY(dev_alloc(ctx, &ctx->rowdots, (size_t)ctx->cones[0].n_alloc * 5));
    }
    if (ctx->world > 1 && !ctx->cone_par) {
        dev_free(ctx->gfull); dev_free(ctx->halo); dev_free(ctx->sendbuf);
        if (ctx->use_halo && ctx->peer) {
            /* two halo buffers, used alternately: a fast peer may already PUT the next exchange's rows while this rank's
             * product still reads the current ones (DESIGN.md "Multi-GPU") */
            const size_t ld = (size_t)ctx->cones[0].ld;
            TRY(dev_alloc(ctx, &ctx->halo, 2 * (size_t)std::max<int64_t>(ctx->halo_rows, 1) * ld));
            TRY(peer_map_halo(ctx));
        } else if (ctx->use_halo) {
            const size_t ld = (size_t)ctx->cones[0].ld;
            TRY(dev_alloc(ctx, &ctx->halo, (size_t)ctx->halo_rows * ld));
            TRY(dev_alloc(ctx, &ctx->sendbuf, (size_t)ctx->send_rows * ld));
        } else {
            TRY(dev_alloc(ctx, &ctx->gfull, (size_t)ctx->world * (size_t)ctx->N));
        }
    }
    ctx->h = lbfgs_len;
    ctx->head = 0;
    ctx->s.assign(lbfgs_len, nullptr);
    ctx->y.assign(lbfgs_len, nullptr);
    for (int k = 0; k < lbfgs_len; ++k) {
        TRY(alloc_flat(ctx, &ctx->s[k]));
        TRY(alloc_flat(ctx, &ctx->y[k]));
    }
    ctx->vars_ready = true;
    return 0;
}

/* A/B switch for the fused MaxCut-type path (default on); takes effect at the next lgpu_alloc_vars / lgpu_aug_rank */
extern "C" int lgpu_set_fused_path(lgpu_ctx *ctx, int on)
{
    if (!ctx) return 1;
    ctx->fast_enabled = on != 0;
    return 0;
}
extern "C" int lgpu_uses_fused_path(const lgpu_ctx *ctx) { return (ctx && ctx->mc) ? 1 : 0; }
/* A/B switch: dense-aggregate cones on the FP64 tensor pipe (default on) vs the pattern-gather kernels */
extern "C" int lgpu_set_dense_tensor_path(lgpu_ctx *ctx, int on)
{
    if (!ctx) return 1;
    ctx->dense_dmma = on != 0;
    return 0;
}
/* A/B switch: carried inner products for the L-BFGS scalars (default on) vs the two-loop recursion's own passes */
extern "C" int lgpu_set_carried_dots(lgpu_ctx *ctx, int on)
{
    if (!ctx) return 1;
    ctx->gram_enabled = on != 0;
    ctx->gram_valid = false;
    ctx->gram_pair_ok[0] = ctx->gram_pair_ok[1] = false;
    return 0;
}

extern "C" int lgpu_alloc_vars(lgpu_ctx *ctx, const int64_t *rank, int lbfgs_len)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (lbfgs_len < 1 || lbfgs_len > 16) LGPU_FAIL(ctx, "lbfgs_len must be in [1,16]");
    free_vars(ctx);
    TRY(layout_and_alloc(ctx, rank, lbfgs_len));
    ctx->cg_last_iter.assign(ctx->ncones, 0);
    CU(ctx, cudaMemsetAsync(ctx->dsc, 0, sizeof(double) * LGPU_NSCALAR, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

static double *flat_of(lgpu_ctx *ctx, int which)
{
    switch (which) {
    case LGPU_R: return ctx->R;
    case LGPU_U: return ctx->U;
    case LGPU_V: return ctx->V;
    case LGPU_GRAD: return ctx->G;
    default: return nullptr;
    }
}
static double *mvec_of(lgpu_ctx *ctx, int which)
{
    switch (which) {
    case LGPU_VEC_DUAL: return ctx->lam;
    case LGPU_VEC_CONSTR_SUM: return ctx->cvs;
    case LGPU_VEC_ARD: return ctx->q1;
    case LGPU_VEC_ADD: return ctx->q2;
    case LGPU_VEC_M1: return ctx->M1;
    case LGPU_VEC_B: return ctx->b;
    default: return nullptr;
    }
}

/* host column-major n x r -> device row-major n x ld at dst */
/* host column-major n_glob x r -> device row-major rows [row_lo, row_lo + n) x ld at dst.  Only the owned rows cross
 * the bus (a strided 2-D copy: r column pieces of n doubles). */
/* pc: the cone whose row relabelling (if any) applies; nullptr = none */
static int upload_factor(lgpu_ctx *ctx, int64_t n, int64_t r, int64_t ld, const double *cm, double *dst, int64_t n_glob = -1,
                         int64_t row_lo = 0, const DevCone *pc = nullptr)
{
    if (n_glob < 0) n_glob = n;
    if (n <= 0) return 0;
    const bool relabel = pc != nullptr && pc->reordered;
    dim3 blk(32, 8);
    if (relabel && n_glob != n) {
        /* partitioned + relabelled: this rank's device rows are scattered rows of the caller's factor: stage it whole */
        const size_t bytes = sizeof(double) * (size_t)n_glob * (size_t)r;
        TRY(ensure_dstage(ctx, bytes));
        CU(ctx, cudaMemcpyAsync(ctx->dstage, cm, bytes, cudaMemcpyHostToDevice, ctx->stream));
        Prof pr(ctx, KC_LAYOUT);
        k_gather_rows_cm<<<grid_for(ctx, n * ld), LGPU_TPB, 0, ctx->stream>>>(n, (int)r, (int)ld, (const double *)ctx->dstage, n_glob,
                                                                              pc->iperm + row_lo, dst);
    } else {
        const size_t bytes = sizeof(double) * (size_t)n * (size_t)r;
        /* straight from the caller's buffer: a pinned buffer goes by DMA, a pageable one is staged by the runtime */
        TRY(ensure_dstage(ctx, bytes));
        if (n_glob == n)
            CU(ctx, cudaMemcpyAsync(ctx->dstage, cm, bytes, cudaMemcpyHostToDevice, ctx->stream));
        else
            CU(ctx, cudaMemcpy2DAsync(ctx->dstage, sizeof(double) * (size_t)n, cm + row_lo, sizeof(double) * (size_t)n_glob,
                                      sizeof(double) * (size_t)n, (size_t)r, cudaMemcpyHostToDevice, ctx->stream));
        Prof pr(ctx, KC_LAYOUT);
        k_col2row<<<(unsigned)((n + 31) / 32), blk, 0, ctx->stream>>>(n, (int)r, (int)ld, (const double *)ctx->dstage, n, 0, dst,
                                                                      relabel ? pc->perm : nullptr);
    }
    CHECK_LAUNCH(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream)); /* the caller may reuse its buffer as soon as this returns */
    return 0;
}
/* the reverse.  In a partitioned run each rank writes ONLY the rows it owns into the caller's n_glob x r array (the
 * other rows are left untouched); the caller assembles the ranks' pieces if it needs the whole factor. */
static int download_factor(lgpu_ctx *ctx, int64_t n, int64_t r, int64_t ld, const double *src, double *cm, int64_t n_glob = -1,
                           int64_t row_lo = 0, const DevCone *pc = nullptr)
{
    if (n_glob < 0) n_glob = n;
    if (n <= 0) return 0;
    const bool relabel = pc != nullptr && pc->reordered;
    const size_t bytes = sizeof(double) * (size_t)n * (size_t)r;
    TRY(ensure_dstage(ctx, bytes));
    dim3 blk(32, 8);
    {
        Prof pr(ctx, KC_LAYOUT);
        k_row2col<<<(unsigned)((n + 31) / 32), blk, 0, ctx->stream>>>(n, (int)r, (int)ld, src, (double *)ctx->dstage, n, 0,
                                                                      (relabel && n_glob == n) ? pc->perm : nullptr);
    }
    CHECK_LAUNCH(ctx);
    if (n_glob == n) {
        CU(ctx, cudaMemcpyAsync(cm, ctx->dstage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (!relabel) {
        CU(ctx, cudaMemcpy2DAsync(cm + row_lo, sizeof(double) * (size_t)n_glob, ctx->dstage, sizeof(double) * (size_t)n,
                                  sizeof(double) * (size_t)n, (size_t)r, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        /* partitioned + relabelled: the owned device rows are scattered rows of the caller's array */
        std::vector<double> tmp((size_t)n * (size_t)r);
        CU(ctx, cudaMemcpyAsync(tmp.data(), ctx->dstage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t col = 0; col < r; ++col)
            for (int64_t t = 0; t < n; ++t) cm[(size_t)col * n_glob + pc->h_iperm[row_lo + t]] = tmp[(size_t)col * n + t];
        return 0;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lgpu_set_factor(lgpu_ctx *ctx, int which, int cone, const double *cm)
{
    if (!ctx || !ctx->vars_ready || cone < 0 || cone >= ctx->ncones || !flat_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    if (which == LGPU_R) ctx->cr_valid = false;
    ctx->rowdots_valid = false; /* R, the gradient or a pair may have been replaced */
    if (which == LGPU_U) { ctx->cd_valid = false; ctx->epi_done = false; }
    return upload_factor(ctx, c.n, c.r, c.ld, cm, flat_of(ctx, which) + c.off, c.n_glob, c.row_lo, &c);
}
extern "C" int lgpu_get_factor(lgpu_ctx *ctx, int which, int cone, double *cm)
{
    if (!ctx || !ctx->vars_ready || cone < 0 || cone >= ctx->ncones || !flat_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    return download_factor(ctx, c.n, c.r, c.ld, flat_of(ctx, which) + c.off, cm, c.n_glob, c.row_lo, &c);
}
extern "C" int lgpu_set_lp(lgpu_ctx *ctx, int which, const double *v)
{
    if (!ctx || !ctx->vars_ready || ctx->lp.n <= 0 || !flat_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(flat_of(ctx, which) + ctx->lp.off, v, sizeof(double) * ctx->lp.n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int lgpu_get_lp(lgpu_ctx *ctx, int which, double *v)
{
    if (!ctx || !ctx->vars_ready || ctx->lp.n <= 0 || !flat_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(v, flat_of(ctx, which) + ctx->lp.off, sizeof(double) * ctx->lp.n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int lgpu_set_vec(lgpu_ctx *ctx, int which, const double *v)
{
    if (!ctx || !mvec_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<double> tmp;
    if (!ctx->h_cperm.empty()) { /* renumbered constraints (row relabelling): caller's order -> device order */
        tmp.resize((size_t)ctx->m);
        for (int64_t k = 0; k < ctx->m; ++k) tmp[ctx->h_cperm[k]] = v[k];
        v = tmp.data();
    }
    CU(ctx, cudaMemcpyAsync(mvec_of(ctx, which), v, sizeof(double) * ctx->m, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
static int get_vec_device_order(lgpu_ctx *ctx, int which, double *v)
{
    if (!ctx || !mvec_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (ctx->world > 1 && !ctx->cone_par && which != LGPU_VEC_B) {
        /* every constraint is owned by exactly one rank: owned entries into a zeroed vector, summed over the ranks */
        double *tmp = ctx->mtmp;
        const double *src = mvec_of(ctx, which);
        CU(ctx, cudaMemsetAsync(tmp, 0, sizeof(double) * ctx->m, ctx->stream));
        const Owned ow = owned(ctx);
        const int32_t *gid = ow.gid;
        launch_map(ctx, ow.count, [=] __device__(int64_t t) { tmp[gid[t]] = src[gid[t]]; });
        NC(ctx, g_nccl.AllReduce(tmp, tmp, (size_t)ctx->m, LG_NCCL_FLOAT64, LG_NCCL_SUM, (lg_ncclComm_t)ctx->comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(v, tmp, sizeof(double) * ctx->m, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    CU(ctx, cudaMemcpyAsync(v, mvec_of(ctx, which), sizeof(double) * ctx->m, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lgpu_get_vec(lgpu_ctx *ctx, int which, double *v)
{
    if (!ctx || !mvec_of(ctx, which)) return 1;
    if (ctx->h_cperm.empty()) return get_vec_device_order(ctx, which, v);
    std::vector<double> tmp((size_t)ctx->m); /* renumbered constraints (row relabelling): device order -> caller's order */
    TRY(get_vec_device_order(ctx, which, tmp.data()));
    for (int64_t k = 0; k < ctx->m; ++k) v[k] = tmp[ctx->h_cperm[k]];
    return 0;
}
extern "C" int lgpu_get_rank(const lgpu_ctx *ctx, int cone, int64_t *rank)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    *rank = ctx->cones[cone].r;
    return 0;
}

extern "C" int lgpu_fill_factor_random(lgpu_ctx *ctx, int which, uint64_t seed)
{
    if (!ctx || !ctx->vars_ready || !flat_of(ctx, which)) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (which == LGPU_R) ctx->cr_valid = false;
    ctx->rowdots_valid = false; /* R, the gradient or a pair may have been replaced */
    if (which == LGPU_U) { ctx->cd_valid = false; ctx->epi_done = false; }
    for (auto &c : ctx->cones) {
        double *p = flat_of(ctx, which) + c.off;
        const int64_t ld = c.ld, r = c.r;
        const uint64_t sd = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(c.off + 1);
        const int64_t goff = c.row_lo * c.ld; /* element index in the whole factor: the same point for any partition */
        launch_map(ctx, c.n * c.ld, [=] __device__(int64_t i) {
            const int64_t col = i % ld;
            uint64_t z = sd + (uint64_t)(i + goff) * 0xBF58476D1CE4E5B9ull;
            z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
            z ^= z >> 27; z *= 0x94D049BB133111EBull;
            z ^= z >> 31;
            const double u1 = (double)(z >> 11) * (1.0 / 9007199254740992.0);
            z = z * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
            z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
            const double u2 = (double)(z >> 11) * (1.0 / 9007199254740992.0);
            p[i] = col < r ? (u1 - u2) : 0.0; /* same distribution as rand()/RAND_MAX - rand()/RAND_MAX */
        });
    }
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_aug_rank(lgpu_ctx *ctx, const int64_t *new_rank)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    /* keep the old R, U, V, Grad alive while the new storage is laid out */
    double *oR = ctx->R, *oU = ctx->U, *oV = ctx->V, *oG = ctx->G;
    ctx->R = ctx->U = ctx->V = ctx->G = nullptr;
    std::vector<int64_t> o_r(ctx->ncones), o_ld(ctx->ncones), o_off(ctx->ncones);
    for (int c = 0; c < ctx->ncones; ++c) {
        o_r[c] = ctx->cones[c].r;
        o_ld[c] = ctx->cones[c].ld;
        o_off[c] = ctx->cones[c].off;
        if (new_rank[c] < o_r[c]) LGPU_FAIL(ctx, "aug_rank cannot shrink");
    }
    const int64_t o_lp_off = ctx->lp.off;
    const int h = ctx->h;
    free_vars(ctx);
    int rc = layout_and_alloc(ctx, new_rank, h);
    if (rc) return rc;
    double *olds[4] = {oR, oU, oV, oG};
    double *news[4] = {ctx->R, ctx->U, ctx->V, ctx->G};
    for (int w = 0; w < 4; ++w) {
        for (int c = 0; c < ctx->ncones; ++c) {
            DevCone &cn = ctx->cones[c];
            Prof pr(ctx, KC_LAYOUT);
            k_restride_aug<<<grid_for(ctx, cn.n * cn.ld), LGPU_TPB, 0, ctx->stream>>>(
                cn.n, (int)o_r[c], (int)o_ld[c], (int)cn.r, (int)cn.ld, olds[w] + o_off[c], news[w] + cn.off, 1, cn.n_glob, cn.row_lo,
                cn.reordered ? cn.iperm : nullptr);
        }
        if (ctx->lp.n > 0)
            CU(ctx, cudaMemcpyAsync(news[w] + ctx->lp.off, olds[w] + o_lp_off, sizeof(double) * ctx->lp.n,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CHECK_LAUNCH(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(oR); cudaFree(oU); cudaFree(oV); cudaFree(oG);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * A(UV^T) and friends
 * ------------------------------------------------------------------------------------------------*/
/* ------------------------------------------------------------------------------------------------
 * fused MaxCut-type path helpers (ctx->mc: one diag_only cone, every constraint non-zero, no LP)
 * ------------------------------------------------------------------------------------------------*/
/* out_k = scale a_k <A_i, B_i> ; with b != nullptr also dsc[SC_PINF] = sum (b - out)^2 */
static void mc_rowdot(lgpu_ctx *ctx, const double *A, const double *B, double scale, double *out, bool with_pinf)
{
    DevCone &c = ctx->cones[0];
    const int G = pick_group(c.ld);
    Prof pr(ctx, KC_GATHER);
    DISPATCH_G(G, k_mc_rowdot<GG><<<grid_for(ctx, c.n * GG, (const void *)k_mc_rowdot<GG>), LGPU_TPB, 0, ctx->stream>>>(
                      c.n, (int)c.ld, A, B, c.rc_ptr, c.rc_gid, c.rc_a, scale, out, with_pinf ? ctx->b : nullptr,
                      ctx->partials, ctx->counter, ctx->dsc, slot1(SC_PINF, 0)));
    if (with_pinf) allreduce_scalars(ctx, SC_PINF, 1);
}

/* make the remote factor rows the sparse product references available: either all rows of every rank
 * (ncclAllGather into gfull, global row order) or only the referenced ones (pack + grouped ncclSend/ncclRecv into
 * `halo`).  One GPU: nothing to do. */
static int mc_exchange(lgpu_ctx *ctx, const double *X)
{
    if (ctx->world <= 1) return 0;
    if (!ctx->use_halo) {
        NC(ctx, g_nccl.AllGather(X, ctx->gfull, (size_t)ctx->N, LG_NCCL_FLOAT64, (lg_ncclComm_t)ctx->comm, ctx->stream));
        return 0;
    }
    DevCone &c = ctx->cones[0];
    const int G = pick_group(c.ld);
    if (ctx->peer && ctx->pending_put_x == X && ctx->pending_put_seq == ctx->xseq) {
        /* the direction pass already PUT these rows (k_mc_combine<PUT>): only wait for the rows of the sources */
        const unsigned long long seq = ctx->pending_put_seq;
        ctx->pending_put_x = nullptr;
        {
            Prof pr(ctx, KC_EXCH);
            k_wait_sources<<<1, 32, 0, ctx->stream>>>(ctx->blk->xflag, ctx->world, ctx->rank, seq);
        }
        ctx->halo_cur = ctx->halo + (size_t)(seq & 1ull) * (size_t)std::max<int64_t>(ctx->halo_rows, 1) * (size_t)c.ld;
        return 0;
    }
    ctx->pending_put_x = nullptr;
    if (ctx->peer) {
        /* PUT the rows every peer's CSR references straight into its halo buffer, then wait for the sources of mine */
        const unsigned long long seq = ++ctx->xseq;
        const size_t buf = (size_t)(seq & 1ull) * (size_t)std::max<int64_t>(ctx->halo_rows, 1) * (size_t)c.ld;
        PeerPut pp;
        pp.world = ctx->world;
        pp.rank = ctx->rank;
        for (int q = 0; q < ctx->world; ++q) {
            pp.send_off[q] = ctx->send_off[q];
            /* rank q's two buffers have q's halo size: its second buffer starts at peer_halo_rows[q] * ld */
            pp.dst[q] = q == ctx->rank ? nullptr
                                       : ctx->peer_halo[q] + (size_t)(seq & 1ull) * (size_t)std::max<int64_t>(ctx->peer_halo_rows[q], 1) * (size_t)c.ld +
                                             (size_t)ctx->dst_off[q] * (size_t)c.ld;
            pp.flag[q] = &ctx->peer_blk[q]->xflag[ctx->rank];
        }
        pp.send_off[ctx->world] = ctx->send_rows;
        {
            Prof pr(ctx, KC_EXCH);
            DISPATCH_G(G, k_put_rows<GG><<<grid_for(ctx, std::max<int64_t>(ctx->send_rows, 1) * GG, (const void *)k_put_rows<GG>), LGPU_TPB, 0,
                                          ctx->stream>>>(ctx->send_rows, (int)c.ld, ctx->send_idx, X, pp, seq, ctx->put_counter));
        }
        {
            Prof pr(ctx, KC_EXCH);
            k_wait_sources<<<1, 32, 0, ctx->stream>>>(ctx->blk->xflag, ctx->world, ctx->rank, seq);
        }
        ctx->halo_cur = ctx->halo + buf;
        return 0;
    }
    ctx->halo_cur = ctx->halo;
    if (ctx->send_rows > 0) {
        Prof pr(ctx, KC_LAYOUT);
        DISPATCH_G(G, k_pack_rows<GG><<<grid_for(ctx, ctx->send_rows * GG, (const void *)k_pack_rows<GG>), LGPU_TPB, 0, ctx->stream>>>(
                          ctx->send_rows, (int)c.ld, ctx->send_idx, X, ctx->sendbuf));
    }
    NC(ctx, g_nccl.GroupStart());
    int bad = 0; /* an open group is always closed, also on the error path */
    for (int q = 0; q < ctx->world && !bad; ++q) {
        if (q == ctx->rank) continue;
        if (ctx->send_cnt[q] > 0)
            bad = g_nccl.Send(ctx->sendbuf + (size_t)ctx->send_off[q] * c.ld, (size_t)ctx->send_cnt[q] * c.ld, LG_NCCL_FLOAT64, q,
                              (lg_ncclComm_t)ctx->comm, ctx->stream);
        if (!bad && ctx->recv_cnt[q] > 0)
            bad = g_nccl.Recv(ctx->halo + (size_t)ctx->recv_off[q] * c.ld, (size_t)ctx->recv_cnt[q] * c.ld, LG_NCCL_FLOAT64, q,
                              (lg_ncclComm_t)ctx->comm, ctx->stream);
    }
    const int end_rc = g_nccl.GroupEnd();
    NC(ctx, bad);
    NC(ctx, end_rc);
    return 0;
}

/* T = C X ; X = this rank's rows (remote rows exchanged first in a partitioned run).  dot_slot >= 0: the product's
 * epilogue also leaves this rank's part of sum_i <X_i, T_i> in dsc[dot_slot] (not all-reduced: the caller's pack is) */
static void mc_spmm_plain(lgpu_ctx *ctx, const double *X, double *T, int dot_slot = -1)
{
    DevCone &c = ctx->cones[0];
    const int G = pick_group(c.ld);
    if (mc_exchange(ctx, X) != 0) return;
    const SlotSpec<1> sp = slot1(dot_slot >= 0 ? dot_slot : SC_TMP, 0);
    const bool halo = ctx->world > 1 && ctx->use_halo;
    const double *Xg = (ctx->world > 1 && !halo) ? ctx->gfull : X;
    const double *Xh = halo ? ctx->halo_cur - (size_t)c.n_alloc * c.ld : nullptr; /* halo row k is addressed as column n_alloc + k */
    const int nsplit = halo ? (int)c.n_alloc : 0;
    const int64_t self_off = (ctx->world > 1 && !halo) ? c.row_lo : 0;
    {
        Prof pr(ctx, KC_MC_SPMM);
#define MC_SPMM(HALO, DOT, MINB)                                                                                                  \
    DISPATCH_G(G, k_mc_spmm<GG, 2, HALO, DOT, MINB><<<grid_for(ctx, c.n * GG, (const void *)k_mc_spmm<GG, 2, HALO, DOT, MINB>),    \
                                                      LGPU_TPB, 0, ctx->stream>>>(c.n, c.f_ptr, c.f_col, c.mc_val, Xg, Xh, nsplit, \
                                                                                  (int)c.ld, T, self_off, ctx->partials,         \
                                                                                  ctx->counter, ctx->dsc, sp))
        const int mode = dot_slot >= 0 ? ctx->spmm_dot : 0;
        if (halo) {
            switch (mode) {
            case 0: MC_SPMM(true, 0, 8); break;
            case 1: MC_SPMM(true, 1, 8); break;
            case 2: MC_SPMM(true, 2, 7); break;
            case 4: MC_SPMM(true, 1, 7); break;
            default: MC_SPMM(true, 2, 6); break; /* measured best at C5: 13.67 ms / iteration vs 13.95 with the separate pass */
            }
        } else {
            switch (mode) {
            case 0: MC_SPMM(false, 0, 8); break;
            case 1: MC_SPMM(false, 1, 8); break;
            case 2: MC_SPMM(false, 2, 7); break;
            case 4: MC_SPMM(false, 1, 7); break;
            default: MC_SPMM(false, 2, 6); break;
            }
        }
#undef MC_SPMM
        run_long_rows(ctx, c, c.ld, nullptr, c.mc_val, Xg, Xh, nsplit, 1.0, 0.0, nullptr, T);
    }
    if (dot_slot >= 0 && c.n_long > 0) {
        /* hub rows were produced by the chunked kernels: add their <X_i, T_i> */
        const int32_t *rows = c.long_rows;
        const int64_t ld = c.ld;
        const bool keep = ctx->defer_allreduce;
        ctx->defer_allreduce = true;
        launch_reduce<1>(ctx, c.n_long * ld, [=] __device__(int64_t q, double(&acc)[1]) {
            const size_t o = (size_t)rows[q / ld] * ld + (size_t)(q % ld);
            acc[0] = fma(X[o], T[o], acc[0]);
        }, slot1(dot_slot, 1));
        ctx->defer_allreduce = keep;
    }
}
static void mc_refresh_cr(lgpu_ctx *ctx)
{
    mc_spmm_plain(ctx, ctx->R, ctx->CR);
    ctx->cr_valid = true;
    ctx->cr_updates = 0;
}
/* Grad = 2 (CR + Diag(A^*(M1)) R), dsc[SC_LAG] = sum Grad^2 ; M1 already formed */
static void mc_grad(lgpu_ctx *ctx)
{
    DevCone &c = ctx->cones[0];
    if (!ctx->cr_valid) mc_refresh_cr(ctx);
    const int G = pick_group(c.ld);
    Prof pr(ctx, KC_MC_STEP);
    DISPATCH_G(G, k_mc_grad<GG><<<grid_for(ctx, c.n * GG, (const void *)k_mc_grad<GG>), LGPU_TPB, 0, ctx->stream>>>(
                      c.n, (int)c.ld, ctx->R, ctx->CR, ctx->G, c.rc_ptr, c.rc_gid, c.rc_a, ctx->M1, ctx->partials, ctx->counter,
                      ctx->dsc, slot1(SC_LAG, 0)));
    allreduce_scalars(ctx, SC_LAG, 1);
}

static void pair_ptrs(lgpu_ctx *ctx, int pair, double **A, double **B)
{
    switch (pair) {
    case LGPU_PAIR_RR: *A = ctx->R; *B = ctx->R; break;
    case LGPU_PAIR_RU: *A = ctx->R; *B = ctx->U; break;
    case LGPU_PAIR_UU: *A = ctx->U; *B = ctx->U; break;
    default: *A = ctx->U; *B = ctx->V; break;
    }
}

/* ---- by-cone partition (several GPUs, problem not of the single MaxCut-type form) ---------------------------------
 * Cones couple only through the m-vector constrValSum and through scalars (lorads_alg_common.c:221-229), so: every rank
 * keeps all cones and all flat vectors (replicated, and updated by the same deterministic kernels on every rank); the
 * OPERATOR work of a cone -- UVt, the constraint gathers, A^*(w), S X, the CG solves -- is done by the cone's owner
 * only; what the others need is summed over the ranks: the length-m constraint values (ncclAllReduce), the objective
 * scalars, the gradient segments (non-owners contribute zeros, so the sum is the owner's value bit for bit). */
static inline bool cone_mine(const lgpu_ctx *ctx, const DevCone &c)
{
    return !ctx->cone_par || ctx->cone_owner[(size_t)(&c - ctx->cones.data())] == ctx->rank;
}
/* terms every rank would add (LP block, replicated m-vector parts) are added by rank 0 only when the result is summed */
static inline bool cone_lead(const lgpu_ctx *ctx) { return !ctx->cone_par || ctx->rank == 0; }
static int cone_allreduce(lgpu_ctx *ctx, double *p, size_t count)
{
    if (!ctx->cone_par || count == 0) return 0;
    NC(ctx, g_nccl.AllReduce(p, p, count, LG_NCCL_FLOAT64, LG_NCCL_SUM, (lg_ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}
/* make `count` doubles at p, valid on `owner`, valid everywhere: zeros elsewhere + sum */
static int cone_bcast(lgpu_ctx *ctx, double *p, size_t count, int owner)
{
    if (!ctx->cone_par || count == 0) return 0;
    if (ctx->rank != owner) CU(ctx, cudaMemsetAsync(p, 0, sizeof(double) * count, ctx->stream));
    return cone_allreduce(ctx, p, count);
}

/* per cone: uvt on the pattern, cv = A_c(.), optional objective accumulate (LORADSInitConstrValObjVal) */
static void cones_auv(lgpu_ctx *ctx, const double *A, const double *B, int obj_slot /* -1: none */)
{
    for (auto &c : ctx->cones) {
        if (!cone_mine(ctx, c)) continue;
        run_uvt(ctx, c, c.ld, A + c.off, B + c.off, c.uvt);
        if (obj_slot >= 0) run_obj_gather(ctx, c, c.uvt, obj_slot, 1);
        run_con_gather(ctx, c, c.uvt, c.cv);
    }
}

/* out (m-vector) = scale * ( LP part + sum over cones of expanded cv )   [InitConstrValSum / ConstrValSumALMtemp] */
static void sum_constr_vals(lgpu_ctx *ctx, const double *A, const double *B, double scale, double *out)
{
    const int64_t m = ctx->m;
    if (!ctx->cone_par && ctx->lp.n == 0 && ctx->ncones == 1 && ctx->cones[0].mA == m) {
        /* one cone in which every constraint is non-zero: its compact constrVal IS the m-vector (con_gid = identity) */
        const double *cv = ctx->cones[0].cv;
        launch_map(ctx, m, [=] __device__(int64_t i) { out[i] = 0.0 + scale * cv[i]; });
        return;
    }
    if (ctx->lp.n > 0 && cone_lead(ctx)) {
        const int32_t *rp = ctx->lp.r_ptr, *rc = ctx->lp.r_col;
        const double *rv = ctx->lp.r_val;
        const double *u = A + ctx->lp.off, *v = B + ctx->lp.off;
        launch_map(ctx, m, [=] __device__(int64_t i) {
            double a = 0.0;
            for (int e = rp[i]; e < rp[i + 1]; ++e) a = fma(rv[e], u[rc[e]] * v[rc[e]], a);
            out[i] = scale * a;
        });
    } else {
        cudaMemsetAsync(out, 0, sizeof(double) * m, ctx->stream);
    }
    for (auto &c : ctx->cones) {
        if (!cone_mine(ctx, c)) continue;
        const int32_t *gid = c.con_gid;
        const double *cv = c.cv;
        launch_map(ctx, c.mA, [=] __device__(int64_t t) { out[gid[t]] += scale * cv[t]; });
    }
    cone_allreduce(ctx, out, (size_t)m); /* by-cone partition: the length-m constraint values are summed over the owners */
}

extern "C" int lgpu_init_constr_val(lgpu_ctx *ctx, int pair)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    double *A, *B;
    pair_ptrs(ctx, pair, &A, &B);
    if (ctx->mc) {
        mc_rowdot(ctx, A, B, 1.0, ctx->cvs, false);
        CHECK_LAUNCH(ctx);
        return 0;
    }
    cones_auv(ctx, A, B, -1);
    sum_constr_vals(ctx, A, B, 1.0, ctx->cvs);
    CHECK_LAUNCH(ctx);
    return 0;
}

/* Grad = 2 (C + A*(M1)) R for all cones (+ LP), with M1 already in ctx->M1; lag = sum |Grad|^2 */
static void grad_from_m1(lgpu_ctx *ctx)
{
    for (auto &c : ctx->cones) {
        if (!cone_mine(ctx, c)) {
            cudaMemsetAsync(ctx->G + c.off, 0, sizeof(double) * (size_t)(c.n_alloc * c.ld), ctx->stream);
            continue;
        }
        run_wsum(ctx, c, ctx->M1, true, true, 1.0, c.S);
        run_spmm(ctx, c, c.ld, c.S, ctx->R + c.off, 2.0, 0.0, nullptr, ctx->G + c.off);
    }
    cone_allreduce(ctx, ctx->G, (size_t)ctx->lp.off); /* the owners' gradient segments, exact (the others hold zeros) */
    if (ctx->lp.n > 0) {
        /* ALMSetGradLP (lorads_alm.c:89-113): grad_j = 2 (c_j + a_j^T M1) r_j */
        const int32_t *cp = ctx->lp.c_ptr, *cr = ctx->lp.c_row;
        const double *cv = ctx->lp.c_val, *obj = ctx->lp.obj, *M1 = ctx->M1;
        const double *r = ctx->R + ctx->lp.off;
        double *g = ctx->G + ctx->lp.off;
        launch_map(ctx, ctx->lp.n, [=] __device__(int64_t j) {
            double w = obj[j];
            for (int e = cp[j]; e < cp[j + 1]; ++e) w = fma(M1[cr[e]], cv[e], w);
            g[j] = 2.0 * w * r[j];
        });
    }
    const double *G = ctx->G;
    launch_reduce<1>(ctx, ctx->N, [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = fma(G[i], G[i], acc[0]); }, slot1(SC_LAG));
}

extern "C" int lgpu_alm_cal_grad(lgpu_ctx *ctx, double rho, double *lag_norm_square)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    {
        /* M1 = -lambda - rho b + rho constrValSum (lorads_alm.c:38-50) */
        double *M1 = ctx->M1;
        const double *lam = ctx->lam, *b = ctx->b, *cvs = ctx->cvs;
        const Owned ow = owned(ctx);
        const int32_t *gid = ow.gid;
        launch_map(ctx, ow.count, [=] __device__(int64_t t) {
            const int64_t i = gid ? gid[t] : t;
            M1[i] = -lam[i] - rho * b[i] + rho * cvs[i];
        });
    }
    if (ctx->mc) mc_grad(ctx);
    else grad_from_m1(ctx);
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_LAG, 1));
    *lag_norm_square = ctx->hsc[SC_LAG];
    ctx->gram.gg = ctx->hsc[SC_LAG];
    ctx->gram_valid = false; /* the gradient changed outside a step: <g, s_j>, <g, y_j> are stale */
    ctx->rowdots_valid = false;
    return 0;
}

/* L-BFGS direction of the fused path with history length 2: the two-loop recursion's scalars come from carried
 * inner products (formed by the step pass, or by one refresh pass when the gradient was recomputed outside a step),
 * so the direction costs ONE pass that also produces q1, q2 and <C, R D^T>.
 * reference: LBFGSDirection + LBFGSDirectionUseGrad, lorads_alm.c:347-508,607-627 */
static int mc_direction_gram(lgpu_ctx *ctx, int nn)
{
    DevCone &c = ctx->cones[0];
    const int h = 2;
    const int j1 = (ctx->head - 1 + h) % h, j0 = j1 ^ 1;
    auto &gm = ctx->gram;
    const bool have = ctx->gram_valid && ctx->gram_pair_ok[j1] && (nn < 2 || ctx->gram_pair_ok[j0]);
    /* the per-row and <C R, .> products of the last bulk step pass describe the current R, g and pairs only if nothing
     * touched them since (same condition as the carried L-BFGS products, which a refresh pass below cannot restore) */
    const bool rowdots = ctx->rowdots_enabled && ctx->rowdots_valid && ctx->rowdots != nullptr && ctx->gram_valid &&
                         (nn == 0 || have) && ctx->cr_valid;
    if (nn > 0 && !have) {
        /* refresh: ten inner products in one pass over (G, s0, y0, s1, y1) */
        const double *G = ctx->G, *s1 = ctx->s[j1], *y1 = ctx->y[j1], *s0 = ctx->s[j0], *y0 = ctx->y[j0];
        SlotSpec<10> sp;
        for (int k = 0; k < 10; ++k) sp.slot[k] = SC_GSN + k;
        sp.accumulate = 0;
        launch_reduce_post<10>(ctx, ctx->N, [=] __device__(int64_t i, double(&acc)[10]) {
            const double g = G[i], a = s1[i], b = y1[i], p = s0[i], q = y0[i];
            acc[0] = fma(g, a, acc[0]); acc[1] = fma(g, b, acc[1]); acc[2] = fma(g, p, acc[2]); acc[3] = fma(g, q, acc[3]);
            acc[4] = fma(p, b, acc[4]); acc[5] = fma(q, b, acc[5]); acc[6] = fma(b, b, acc[6]); acc[7] = fma(q, q, acc[7]);
            acc[8] = fma(b, a, acc[8]); acc[9] = fma(q, p, acc[9]);
        }, sp, NoPost(), KC_MC_DIR);
        CHECK_LAUNCH(ctx);
        TRY(fetch_scalars(ctx, SC_GSN, 10));
        const double *v = ctx->hsc + SC_GSN;
        gm.sg[j1] = v[0]; gm.yg[j1] = v[1]; gm.sg[j0] = v[2]; gm.yg[j0] = v[3];
        gm.so_yn = v[4]; gm.yo_yn = v[5]; gm.yy[j1] = v[6]; gm.yy[j0] = v[7];
        gm.beta[j1] = 1.0 / v[8]; gm.beta[j0] = 1.0 / v[9];
        ctx->gram_valid = true;
        ctx->gram_pair_ok[0] = ctx->gram_pair_ok[1] = true;
    }
    DirCoef cf;
    cf.a1 = cf.a0 = cf.w0 = cf.w1 = 0.0;
    cf.nn = nn;
    if (nn >= 1) {
        /* the recursion's scalars, in the reference's order of operations */
        double dg;
        cf.a1 = gm.beta[j1] * gm.sg[j1];
        if (nn == 1) {
            const double t1 = gm.yg[j1] - cf.a1 * gm.yy[j1];
            cf.w1 = cf.a1 - gm.beta[j1] * t1;
            dg = -(gm.gg - cf.a1 * gm.yg[j1] + cf.w1 * gm.sg[j1]);
        } else {
            cf.a0 = gm.beta[j0] * (gm.sg[j0] - cf.a1 * gm.so_yn);
            const double t0 = gm.yg[j0] - cf.a1 * gm.yo_yn - cf.a0 * gm.yy[j0];
            cf.w0 = cf.a0 - gm.beta[j0] * t0;
            const double t1 = gm.yg[j1] - cf.a1 * gm.yy[j1] - cf.a0 * gm.yo_yn + cf.w0 * gm.so_yn;
            cf.w1 = cf.a1 - gm.beta[j1] * t1;
            dg = -(gm.gg - cf.a1 * gm.yg[j1] - cf.a0 * gm.yg[j0] + cf.w0 * gm.sg[j0] + cf.w1 * gm.sg[j1]);
        }
        if (!(dg < 0.0)) cf.nn = 0; /* <D, Grad> >= 0 (or NaN): fall back to D = -Grad */
    }
    if (!ctx->cr_valid) mc_refresh_cr(ctx);
    const int G = pick_group(c.ld);
    ctx->defer_allreduce = true; /* <C R, D> joins the seven line-search terms' all-reduce */
    PeerDirect pd;
    memset(&pd, 0, sizeof(pd));
    const bool put = ctx->peer && ctx->use_halo && ctx->fuse_put && ctx->put_dest != nullptr;
    if (put) {
        /* the halo exchange of D rides inside this pass; the sparse product that follows only waits for its sources */
        const unsigned long long seq = ++ctx->xseq;
        pd.npeer = ctx->world - 1;
        pd.seq = seq;
        for (int q = 0, j = 0; q < ctx->world; ++q) {
            if (q == ctx->rank) continue;
            pd.dst[j] = ctx->peer_halo[q] + (size_t)(seq & 1ull) * (size_t)std::max<int64_t>(ctx->peer_halo_rows[q], 1) * (size_t)c.ld;
            pd.flag[j] = &ctx->peer_blk[q]->xflag[ctx->rank];
            ++j;
        }
        ctx->pending_put_x = ctx->U;
        ctx->pending_put_seq = seq;
    }
    ctx->p1_host = 0.0;
    if (rowdots) {
        /* <C R, D> for D = -g + a1 y1 + a0 y0 - w0 s0 - w1 s1 from the carried (already rank-reduced) products */
        double p1 = -gm.cr_g;
        if (cf.nn >= 1) {
            p1 += cf.a1 * gm.cr_y[j1];
            if (cf.nn >= 2) { p1 += cf.a0 * gm.cr_y[j0]; p1 -= cf.w0 * gm.cr_s[j0]; }
            p1 -= cf.w1 * gm.cr_s[j1];
        }
        ctx->p1_host = p1;
    }
    {
        Prof pr(ctx, KC_MC_DIR);
#define MC_COMBINE(PUT, RD)                                                                                                          \
    DISPATCH_G(G, k_mc_combine<GG, PUT, RD><<<grid_for(ctx, c.n * GG, (const void *)k_mc_combine<GG, PUT, RD>), LGPU_TPB, 0,           \
                                               ctx->stream>>>(c.n, (int)c.ld, cf, ctx->G, ctx->s[j1], ctx->y[j1], ctx->s[j0], ctx->y[j0], \
                                                              ctx->R, ctx->CR, ctx->U, c.rc_ptr, c.rc_gid, c.rc_a, ctx->q1, ctx->q2,    \
                                                              ctx->partials, ctx->counter, ctx->dsc, slot1(SC_P1, 0),                  \
                                                              (PUT) ? ctx->put_dest : nullptr, pd, ctx->rowdots))
        if (put) {
            if (rowdots) { MC_COMBINE(true, true); } else { MC_COMBINE(true, false); }
        } else {
            if (rowdots) { MC_COMBINE(false, true); } else { MC_COMBINE(false, false); }
        }
#undef MC_COMBINE
    }
    ctx->defer_allreduce = false;
    ctx->epi_done = true;
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_lbfgs_direction(lgpu_ctx *ctx, int64_t inner_iter)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    const int64_t N = ctx->N;
    double *D = ctx->U;
    const double *G = ctx->G;
    double *dsc = ctx->dsc;
    ctx->cd_valid = false;
    ctx->epi_done = false;
    const int h = ctx->h;
    const int nn = (int)((inner_iter <= (ctx->lp.n > 0 ? h : h - 1)) ? inner_iter : h); /* Q4: LP variant uses <= h */
    if (ctx->mc && ctx->gram_enabled && h == 2) return mc_direction_gram(ctx, nn);
    if (inner_iter == 0) {
        launch_map(ctx, N, [=] __device__(int64_t i) { D[i] = -G[i]; });
        CHECK_LAUNCH(ctx);
        return 0; /* the reference returns before the <D,Grad> test changes anything: D = -Grad already */
    }
    /* two-loop recursion (lorads_alm.c:468-505) as 2 nn + 1 fused passes: every pass applies the previous axpy and
     * accumulates the next dot product; the derived scalar (alpha or alpha - beta <y,q>) is formed by the pass's
     * finishing thread.  Same operations in the same order as the plain recursion. */
    int j[16];
    j[0] = (ctx->head - 1 + h) % h;
    for (int k = 1; k < nn; ++k) j[k] = (j[k - 1] - 1 + h) % h;
    {
        const double *s0 = ctx->s[j[0]];
        const int ia = SC_ALPHA0 + j[0], ib = SC_BETA0 + j[0];
        launch_reduce_post<1>(ctx, N, [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = fma(s0[i], G[i], acc[0]); },
                              slot1(SC_TMP), [=] __device__(double *sc) { sc[ia] = sc[ib] * sc[SC_TMP]; }, KC_MC_DIR);
    }
    for (int k = 0; k < nn; ++k) {
        const double *src = (k == 0) ? G : D;
        const double *yk = ctx->y[j[k]];
        const int ia = SC_ALPHA0 + j[k];
        if (k < nn - 1) {
            const double *dv = ctx->s[j[k + 1]];
            const int ia2 = SC_ALPHA0 + j[k + 1], ib2 = SC_BETA0 + j[k + 1];
            launch_reduce_post<1>(ctx, N, [=] __device__(int64_t i, double(&acc)[1]) {
                const double v = fma(-dsc[ia], yk[i], src[i]);
                D[i] = v;
                acc[0] = fma(dv[i], v, acc[0]);
            }, slot1(SC_TMP), [=] __device__(double *sc) { sc[ia2] = sc[ib2] * sc[SC_TMP]; }, KC_MC_DIR);
        } else {
            const int ib = SC_BETA0 + j[k];
            launch_reduce_post<1>(ctx, N, [=] __device__(int64_t i, double(&acc)[1]) {
                const double v = fma(-dsc[ia], yk[i], src[i]);
                D[i] = v;
                acc[0] = fma(yk[i], v, acc[0]);
            }, slot1(SC_TMP), [=] __device__(double *sc) { sc[SC_TMP2] = sc[ia] - sc[ib] * sc[SC_TMP]; }, KC_MC_DIR);
        }
    }
    for (int k = nn - 1; k >= 0; --k) {
        const double *sk = ctx->s[j[k]];
        if (k > 0) {
            const double *dv = ctx->y[j[k - 1]];
            const int ia2 = SC_ALPHA0 + j[k - 1], ib2 = SC_BETA0 + j[k - 1];
            launch_reduce_post<1>(ctx, N, [=] __device__(int64_t i, double(&acc)[1]) {
                const double v = fma(dsc[SC_TMP2], sk[i], D[i]);
                D[i] = v;
                acc[0] = fma(dv[i], v, acc[0]);
            }, slot1(SC_TMP), [=] __device__(double *sc) { sc[SC_TMP2] = sc[ia2] - sc[ib2] * sc[SC_TMP]; }, KC_MC_DIR);
        } else {
            /* last axpy, D = -q and <D, Grad> in one pass */
            launch_reduce_post<1>(ctx, N, [=] __device__(int64_t i, double(&acc)[1]) {
                const double v = -fma(dsc[SC_TMP2], sk[i], D[i]);
                D[i] = v;
                acc[0] = fma(v, G[i], acc[0]);
            }, slot1(SC_DG), NoPost(), KC_MC_DIR);
        }
    }
    /* LBFGSDirectionUseGrad (lorads_alm.c:607-627) */
    {
        Prof pr(ctx, KC_VEC);
        k_neg_if_nonneg<<<grid_for(ctx, N, (const void *)k_neg_if_nonneg), LGPU_TPB, 0, ctx->stream>>>(N, dsc, SC_DG, G, D);
    }
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_alm_linesearch_terms(lgpu_ctx *ctx, double rho, double out[7])
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    double *dsc = ctx->dsc;
    if (ctx->mc) {
        /* one sparse product T = C D; then q1, q2, p1, p2.  If the direction pass already produced q1, q2 and
         * <C, R D^T> = <C R, D> (carried inner products), only p2 = <D, T> is left. */
        DevCone &c = ctx->cones[0];
        const int G = pick_group(c.ld);
        /* p2 = <D, C D> rides in the product's epilogue when the direction pass already produced q1, q2, p1 */
        mc_spmm_plain(ctx, ctx->U, ctx->CD, (ctx->epi_done && ctx->spmm_dot > 0) ? SC_P2 : -1);
        ctx->defer_allreduce = true; /* local sums: p1, p2 and the five terms are all-reduced together below */
        if (ctx->epi_done && ctx->spmm_dot > 0) {
        } else if (ctx->epi_done) {
            const double *D = ctx->U, *T = ctx->CD;
            launch_reduce<1>(ctx, ctx->N, [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = fma(D[i], T[i], acc[0]); },
                             slot1(SC_P2));
        } else {
            SlotSpec<2> sp;
            sp.slot[0] = SC_P1; sp.slot[1] = SC_P2; sp.accumulate = 0;
            Prof pr(ctx, KC_GATHER);
            DISPATCH_G(G, k_mc_epi<GG><<<grid_for(ctx, c.n * GG, (const void *)k_mc_epi<GG>), LGPU_TPB, 0, ctx->stream>>>(
                              c.n, (int)c.ld, ctx->R, ctx->U, ctx->CD, c.rc_ptr, c.rc_gid, c.rc_a, ctx->q1, ctx->q2, ctx->partials,
                              ctx->counter, ctx->dsc, sp));
        }
        ctx->defer_allreduce = false;
        ctx->cd_valid = true;
    } else {
    launch_scalar(ctx, [=] __device__() { dsc[SC_P1] = 0.0; dsc[SC_P2] = 0.0; });
    /* ALMCalq12p12: q1 = 2 A(sym(R D^T)), p1 = 2 <C, R D^T>; q2 = A(D D^T), p2 = <C, D D^T> */
    cones_auv(ctx, ctx->R, ctx->U, SC_P1);
    if (ctx->lp.n > 0 && cone_lead(ctx)) {
        const double *obj = ctx->lp.obj, *u = ctx->R + ctx->lp.off, *v = ctx->U + ctx->lp.off;
        launch_reduce<1>(ctx, ctx->lp.n, [=] __device__(int64_t j, double(&acc)[1]) { acc[0] = fma(obj[j], u[j] * v[j], acc[0]); }, slot1(SC_P1, 1));
    }
    sum_constr_vals(ctx, ctx->R, ctx->U, 2.0, ctx->q1);
    cones_auv(ctx, ctx->U, ctx->U, SC_P2);
    if (ctx->lp.n > 0 && cone_lead(ctx)) {
        const double *obj = ctx->lp.obj, *u = ctx->U + ctx->lp.off;
        launch_reduce<1>(ctx, ctx->lp.n, [=] __device__(int64_t j, double(&acc)[1]) { acc[0] = fma(obj[j], u[j] * u[j], acc[0]); }, slot1(SC_P2, 1));
    }
    sum_constr_vals(ctx, ctx->U, ctx->U, 1.0, ctx->q2);
    }
    {
        /* the five reductions of ALMLineSearch (lorads_alm.c:266-279); q0' = b - constrValSum + lambda / rho */
        const double *b = ctx->b, *cvs = ctx->cvs, *lam = ctx->lam, *q1 = ctx->q1, *q2 = ctx->q2;
        const double rinv = 1.0 / rho;
        SlotSpec<5> sp;
        sp.slot[0] = SC_LS0; sp.slot[1] = SC_LS1; sp.slot[2] = SC_LS2; sp.slot[3] = SC_LS3; sp.slot[4] = SC_LS4;
        sp.accumulate = 0;
        const Owned ow = owned(ctx);
        const int32_t *gid = ow.gid;
        ctx->defer_allreduce = true;
        launch_reduce<5>(ctx, ow.count, [=] __device__(int64_t t, double(&acc)[5]) {
            const int64_t i = gid ? gid[t] : t;
            const double q0 = (b[i] - cvs[i]) + rinv * lam[i];
            const double a = q1[i], c2 = q2[i];
            acc[0] = fma(c2, c2, acc[0]);
            acc[1] = fma(a, c2, acc[1]);
            acc[2] = fma(q0, c2, acc[2]);
            acc[3] = fma(a, a, acc[3]);
            acc[4] = fma(q0, a, acc[4]);
        }, sp);
        ctx->defer_allreduce = false;
        /* row blocks: all seven are partial sums; by cone: only the two objective terms are (the five m-vector terms were
         * formed from replicated vectors) */
        TRY(allreduce_scalars(ctx, SC_P1, ctx->cone_par ? 2 : 7));
    }
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_P1, 7));
    out[0] = 2.0 * (ctx->hsc[SC_P1] + ((ctx->mc && ctx->epi_done) ? ctx->p1_host : 0.0));
    out[1] = ctx->hsc[SC_P2];
    for (int k = 0; k < 5; ++k) out[2 + k] = ctx->hsc[SC_LS0 + k];
    return 0;
}

extern "C" int lgpu_alm_step(lgpu_ctx *ctx, double tau)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (ctx->mc && ctx->cr_valid && ctx->cd_valid) {
        double *yh = ctx->y[ctx->head], *R = ctx->R, *CR = ctx->CR;
        const double *G = ctx->G, *D = ctx->U, *CD = ctx->CD;
        launch_map(ctx, ctx->N, [=] __device__(int64_t i) {
            yh[i] = -G[i];
            R[i] = fma(tau, D[i], R[i]);
            CR[i] = fma(tau, CD[i], CR[i]);
        });
        ctx->cr_updates++;
    } else {
        double *yh = ctx->y[ctx->head], *R = ctx->R;
        const double *G = ctx->G, *D = ctx->U;
        launch_map(ctx, ctx->N, [=] __device__(int64_t i) {
            yh[i] = -G[i];
            R[i] = fma(tau, D[i], R[i]);
        });
        ctx->cr_valid = false;
    ctx->rowdots_valid = false;
        ctx->rowdots_valid = false;
    }
    {
        double *cvs = ctx->cvs;
        const double *q1 = ctx->q1, *q2 = ctx->q2;
        const double t2 = tau * tau;
        const Owned ow = owned(ctx);
        const int32_t *gid = ow.gid;
        launch_map(ctx, ow.count, [=] __device__(int64_t t) {
            const int64_t i = gid ? gid[t] : t;
            cvs[i] = fma(t2, q2[i], fma(tau, q1[i], cvs[i]));
        });
    }
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_lbfgs_push(lgpu_ctx *ctx, double tau)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    double *sh = ctx->s[ctx->head], *yh = ctx->y[ctx->head];
    const double *G = ctx->G, *D = ctx->U;
    const int ib = SC_BETA0 + ctx->head;
    launch_reduce_post<1>(ctx, ctx->N, [=] __device__(int64_t i, double(&acc)[1]) {
        const double sv = tau * D[i];
        const double yv = yh[i] + G[i];
        sh[i] = sv;
        yh[i] = yv;
        acc[0] = fma(yv, sv, acc[0]);
    }, slot1(SC_YS), [=] __device__(double *sc) { sc[ib] = 1.0 / sc[SC_YS]; });
    ctx->gram_pair_ok[ctx->head] = false;
    ctx->head = (ctx->head + 1) % ctx->h;
    ctx->gram_valid = false;
    ctx->rowdots_valid = false;
    CHECK_LAUNCH(ctx);
    return 0;
}

/* Everything of one ALM inner iteration that follows the line search (lorads_alm.c:1342-1357):
 * setAsNegGrad, ALMupdateVar, constrValSum update, ALMCalGrad, setlbfgsHisTwo, updateDimacsALM. */
extern "C" int lgpu_alm_inner_update(lgpu_ctx *ctx, double rho, double tau, double *lag_norm_square, double *pinf_l1)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!(ctx->mc && ctx->cr_valid && ctx->cd_valid)) {
        TRY(lgpu_alm_step(ctx, tau));
        TRY(lgpu_alm_cal_grad(ctx, rho, lag_norm_square));
        TRY(lgpu_lbfgs_push(ctx, tau));
        TRY(lgpu_primal_infeasibility(ctx, LGPU_PAIR_RR, pinf_l1));
        return 0;
    }
    DevCone &c = ctx->cones[0];
    const int G = pick_group(c.ld);
    const bool gram = ctx->gram_enabled && ctx->h == 2;
    const int jn = ctx->head, jo = (ctx->head + 1) % ctx->h;
    /* bulk-copy pipeline (k_mc_step_bulk): tile rows / stages from ld so that a stage holds ~8 KB per stream and the ring
     * fits the 227 KB of shared memory; very wide factors (small problems, L2-resident anyway) keep the register kernel */
    int tr = 0, nstage = 0;
    size_t bulk_smem = 0;
    bool wrote_rowdots = false;
    if (ctx->step_bulk) {
        const int NG = LGPU_TPB / G;
        tr = ctx->step_tile_rows > 0 ? ctx->step_tile_rows : (int)std::max<int64_t>(NG, std::min<int64_t>(64, 8192 / (c.ld * 8)));
        tr = (tr + NG - 1) / NG * NG;
        const size_t sb = step_bulk_stage_bytes(gram ? 7 : 5, tr, (int)c.ld);
        /* 227 KB per CTA in all: the ring + 64 bytes of barriers + the reduction tail's static shared memory (< 1 KB) */
        nstage = (int)std::min<size_t>(LGPU_STEP_MAX_STAGES, (size_t)(225 * 1024) / sb);
        if (ctx->step_stages > 0) nstage = std::min(nstage, ctx->step_stages);
        bulk_smem = (size_t)nstage * sb + 2 * LGPU_STEP_MAX_STAGES * sizeof(uint64_t);
        if (nstage < 2) tr = 0;
    }
    if (tr > 0) {
        const int64_t ntiles = (c.n + tr - 1) / tr;
        const int grid = (int)std::min<int64_t>(ntiles, ctx->num_sms);
        const int threads = LGPU_TPB + 32 * nstage;
        Prof pr(ctx, KC_MC_STEP);
#define MC_STEP_BULK(KERN, SP, ROWPTR)                                                                                              \
    DISPATCH_G(G, {                                                                                                                   \
        auto kern = KERN;                                                                                                             \
        static size_t attr_bytes = 0;                                                                                                 \
        if (bulk_smem > attr_bytes) {                                                                                                 \
            CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));                         \
            attr_bytes = bulk_smem;                                                                                                   \
        }                                                                                                                             \
        kern<<<grid, threads, bulk_smem, ctx->stream>>>(c.n, (int)c.ld, tr, nstage, tau, rho, ctx->R, ctx->U, ctx->CR, ctx->CD, ctx->G, \
                                                       ctx->s[jn], ctx->y[jn], c.rc_ptr, c.rc_gid, c.rc_a, ctx->lam, ctx->b, ctx->cvs, \
                                                       ctx->q1, ctx->q2, ctx->M1, gram ? ctx->s[jo] : nullptr,                        \
                                                       gram ? ctx->y[jo] : nullptr, ctx->partials, ctx->counter, ctx->dsc, SP,         \
                                                       SC_BETA0 + jn, ROWPTR);                                                        \
    })
        if (gram && ctx->rowdots != nullptr) {
            SlotSpec<15> sp;
            for (int k = 0; k < 10; ++k) sp.slot[k] = SC_LAG + k;
            for (int k = 0; k < 5; ++k) sp.slot[10 + k] = SC_CRG + k;
            sp.accumulate = 0;
            wrote_rowdots = true;
            MC_STEP_BULK((k_mc_step_bulk<GG, true, true>), sp, ctx->rowdots);
        } else if (gram) {
            SlotSpec<10> sp;
            for (int k = 0; k < 10; ++k) sp.slot[k] = SC_LAG + k;
            sp.accumulate = 0;
            MC_STEP_BULK((k_mc_step_bulk<GG, true, false>), sp, nullptr);
        } else {
            SlotSpec<3> sp;
            sp.slot[0] = SC_LAG; sp.slot[1] = SC_YS; sp.slot[2] = SC_PINF; sp.accumulate = 0;
            MC_STEP_BULK((k_mc_step_bulk<GG, false, false>), sp, nullptr);
        }
#undef MC_STEP_BULK
    } else if (gram) {
        SlotSpec<10> sp;
        for (int k = 0; k < 10; ++k) sp.slot[k] = SC_LAG + k;
        sp.accumulate = 0;
        Prof pr(ctx, KC_MC_STEP);
#define MC_STEP(MB, CS)                                                                                                         \
    DISPATCH_G(G, k_mc_step<GG, true, MB, CS><<<grid_for(ctx, c.n * GG, (const void *)k_mc_step<GG, true, MB, CS>), LGPU_TPB, 0, \
                                            ctx->stream>>>(c.n, (int)c.ld, tau, rho, ctx->R, ctx->U, ctx->CR, ctx->CD, ctx->G,    \
                                                           ctx->s[jn], ctx->y[jn], c.rc_ptr, c.rc_gid, c.rc_a, ctx->lam, ctx->b, \
                                                           ctx->cvs, ctx->q1, ctx->q2, ctx->M1, ctx->s[jo], ctx->y[jo],          \
                                                           ctx->partials, ctx->counter, ctx->dsc, sp, SC_BETA0 + jn))
        switch (ctx->step_variant) {
        case 1: MC_STEP(3, false); break;
        case 2: MC_STEP(4, false); break;
        case 3: MC_STEP(5, false); break;
        case 4: MC_STEP(3, true); break;
        case 5: MC_STEP(4, true); break;
        default: MC_STEP(2, false); break;
        }
#undef MC_STEP
    } else {
        SlotSpec<3> sp;
        sp.slot[0] = SC_LAG; sp.slot[1] = SC_YS; sp.slot[2] = SC_PINF; sp.accumulate = 0;
        Prof pr(ctx, KC_MC_STEP);
        DISPATCH_G(G, k_mc_step<GG, false, 4, false><<<grid_for(ctx, c.n * GG, (const void *)k_mc_step<GG, false, 4, false>), LGPU_TPB, 0, ctx->stream>>>(
                          c.n, (int)c.ld, tau, rho, ctx->R, ctx->U, ctx->CR, ctx->CD, ctx->G, ctx->s[jn], ctx->y[jn], c.rc_ptr, c.rc_gid,
                          c.rc_a, ctx->lam, ctx->b, ctx->cvs, ctx->q1, ctx->q2, ctx->M1, nullptr, nullptr, ctx->partials,
                          ctx->counter, ctx->dsc, sp, SC_BETA0 + jn));
    }
    /* SC_LAG .. SC_YNYN (10), or with the carried <C R, .> products up to SC_CRYO (18, three refresh-only slots in between) */
    const int nsc = gram ? (wrote_rowdots ? SC_CRYO - SC_LAG + 1 : 10) : 3;
    if (ctx->world > 1) {
        TRY(allreduce_scalars(ctx, SC_LAG, nsc));
        double *dsc = ctx->dsc;
        const int ib = SC_BETA0 + jn;
        launch_scalar(ctx, [=] __device__() { dsc[ib] = 1.0 / dsc[SC_YS]; });
    }
    ctx->head = (ctx->head + 1) % ctx->h;
    ctx->cr_updates++;
    ctx->cd_valid = false;
    ctx->epi_done = false;
    if (ctx->cr_updates >= 64) mc_refresh_cr(ctx); /* bound the rounding drift of the carried C R */
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_LAG, nsc));
    *lag_norm_square = ctx->hsc[SC_LAG];
    *pinf_l1 = sqrt(ctx->hsc[SC_PINF]) / (1.0 + ctx->b_nrm1);
    if (gram) {
        auto &gm = ctx->gram;
        const double *v = ctx->hsc;
        gm.gg = v[SC_LAG];
        gm.beta[jn] = 1.0 / v[SC_YS];
        gm.sg[jn] = v[SC_GSN]; gm.yg[jn] = v[SC_GYN]; gm.sg[jo] = v[SC_GSO]; gm.yg[jo] = v[SC_GYO];
        gm.so_yn = v[SC_SOYN]; gm.yo_yn = v[SC_YOYN]; gm.yy[jn] = v[SC_YNYN];
        if (wrote_rowdots) {
            gm.cr_g = v[SC_CRG];
            gm.cr_s[jn] = v[SC_CRSN]; gm.cr_y[jn] = v[SC_CRYN]; gm.cr_s[jo] = v[SC_CRSO]; gm.cr_y[jo] = v[SC_CRYO];
        }
        ctx->rowdots_valid = wrote_rowdots && ctx->cr_updates > 0; /* a refresh of C R just above makes <C R, .> stale */
        ctx->gram_pair_ok[jn] = true;
        /* <g, .> is fresh for both pairs; the older pair's <y,y> and beta are carried from the step that formed it
         * (gram_pair_ok); otherwise the next two-pair direction refreshes everything in one pass */
        ctx->gram_valid = true;
    }
    return 0;
}

extern "C" int lgpu_primal_infeasibility(lgpu_ctx *ctx, int pair, double *pinf_l1)
{
    if (!ctx || !ctx->vars_ready) return 1;
    if (ctx->mc) {
        CU(ctx, cudaSetDevice(ctx->device));
        double *A, *B;
        pair_ptrs(ctx, pair, &A, &B);
        mc_rowdot(ctx, A, B, 1.0, ctx->cvs, true);
        CHECK_LAUNCH(ctx);
        TRY(fetch_scalars(ctx, SC_PINF, 1));
        *pinf_l1 = sqrt(ctx->hsc[SC_PINF]) / (1.0 + ctx->b_nrm1);
        return 0;
    }
    TRY(lgpu_init_constr_val(ctx, pair));
    const double *b = ctx->b, *cvs = ctx->cvs;
    launch_reduce<1>(ctx, ctx->m, [=] __device__(int64_t i, double(&acc)[1]) {
        const double d = b[i] - cvs[i];
        acc[0] = fma(d, d, acc[0]);
    }, slot1(SC_PINF));
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_PINF, 1));
    *pinf_l1 = sqrt(ctx->hsc[SC_PINF]) / (1.0 + ctx->b_nrm1);
    return 0;
}

extern "C" int lgpu_update_dual_var(lgpu_ctx *ctx, double rho)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    double *lam = ctx->lam;
    const double *b = ctx->b, *cvs = ctx->cvs;
    /* lambda += rho b ; lambda -= rho constrValSum (two axpys, lorads_alg_common.c:511-524) */
    const Owned ow = owned(ctx);
    const int32_t *gid = ow.gid;
    launch_map(ctx, ow.count, [=] __device__(int64_t t) {
        const int64_t i = gid ? gid[t] : t;
        lam[i] = fma(-rho, cvs[i], fma(rho, b[i], lam[i]));
    });
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_average_uv(lgpu_ctx *ctx)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    double *R = ctx->R;
    const double *U = ctx->U, *V = ctx->V;
    launch_map(ctx, ctx->N, [=] __device__(int64_t i) { R[i] = (U[i] + V[i]) / 2.0; });
    ctx->cr_valid = false;
    ctx->rowdots_valid = false;
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_cal_obj(lgpu_ctx *ctx, int admm, double *pobj)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    if (admm) TRY(lgpu_average_uv(ctx));
    double *dsc = ctx->dsc;
    if (ctx->mc) {
        /* <C, R R^T> = <R, C R> */
        if (!ctx->cr_valid) mc_refresh_cr(ctx);
        const double *R = ctx->R, *CR = ctx->CR;
        launch_reduce<1>(ctx, ctx->N, [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = fma(R[i], CR[i], acc[0]); }, slot1(SC_OBJ));
        CHECK_LAUNCH(ctx);
        TRY(fetch_scalars(ctx, SC_OBJ, 1));
        *pobj = ctx->hsc[SC_OBJ];
        return 0;
    }
    launch_scalar(ctx, [=] __device__() { dsc[SC_OBJ] = 0.0; });
    if (ctx->lp.n > 0 && cone_lead(ctx)) {
        const double *obj = ctx->lp.obj, *u = ctx->R + ctx->lp.off;
        launch_reduce<1>(ctx, ctx->lp.n, [=] __device__(int64_t j, double(&acc)[1]) { acc[0] = fma(obj[j], u[j] * u[j], acc[0]); }, slot1(SC_OBJ, 1));
    }
    for (auto &c : ctx->cones) {
        if (!cone_mine(ctx, c)) continue;
        run_uvt(ctx, c, c.ld, ctx->R + c.off, ctx->R + c.off, c.uvt);
        run_obj_gather(ctx, c, c.uvt, SC_OBJ, 1);
    }
    if (ctx->cone_par) TRY(allreduce_scalars(ctx, SC_OBJ, 1));
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_OBJ, 1));
    *pobj = ctx->hsc[SC_OBJ];
    return 0;
}

extern "C" int lgpu_cal_dual_obj(lgpu_ctx *ctx, double *dobj)
{
    if (!ctx) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    const double *b = ctx->b, *lam = ctx->lam;
    const Owned ow = owned(ctx);
    const int32_t *gid = ow.gid;
    launch_reduce<1>(ctx, ow.count, [=] __device__(int64_t t, double(&acc)[1]) {
        const int64_t i = gid ? gid[t] : t;
        acc[0] = fma(b[i], lam[i], acc[0]);
    }, slot1(SC_DOBJ));
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_DOBJ, 1));
    *dobj = ctx->hsc[SC_DOBJ];
    return 0;
}

extern "C" int lgpu_alm_to_admm(lgpu_ctx *ctx)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(ctx->V, ctx->R, sizeof(double) * ctx->N, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->U, ctx->R, sizeof(double) * ctx->N, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->cd_valid = false;
    ctx->epi_done = false;
    return 0;
}
extern "C" int lgpu_copy_r_to_v(lgpu_ctx *ctx)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(ctx->V, ctx->R, sizeof(double) * ctx->N, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * ADMM: CG on x -> x + (sum_i <A_i, sym(x V^T)> A_i) V          lorads_admm.c:442-616, lorads_cgs.c:128-287
 * ------------------------------------------------------------------------------------------------*/
/* res = x + A_V^*(A_V(x)) V for cone c; x, fixed, res are row-major blocks of that cone */
static void cg_mvec(lgpu_ctx *ctx, DevCone &c, const double *x, const double *fixed, double *res)
{
    if (ctx->mc) {
        /* diagonal constraints: the whole operator is row-local, x_i + (sum_k a_k^2) <x_i, V_i> V_i */
        const int G = pick_group(c.ld);
        Prof pr(ctx, KC_MC_STEP);
        DISPATCH_G(G, k_mc_cg_mvec<GG><<<grid_for(ctx, c.n * GG, (const void *)k_mc_cg_mvec<GG>), LGPU_TPB, 0, ctx->stream>>>(c.n, (int)c.ld, x, fixed, c.rc_ptr,
                                                                                             c.rc_a, res));
        return;
    }
    /* LORADSUpdateConstrValCG: weight = A(sym(x V^T)) ; then w_sum = sum weight_i A_i ; res = w_sum V + x */
    run_uvt(ctx, c, c.ld, x, fixed, c.uvt);
    run_con_gather(ctx, c, c.uvt, c.wtmp);
    run_wsum(ctx, c, c.wtmp, false, false, 1.0, c.S);
    run_spmm(ctx, c, c.ld, c.S, fixed, 1.0, 1.0, x, res);
}

static int cg_solve(lgpu_ctx *ctx, int ci, double *x, const double *fixed, const double *bvec, double tol, int64_t maxit,
                    int64_t *iters)
{
    DevCone &c = ctx->cones[ci];
    const int64_t nr = c.n * c.ld;
    double *r = ctx->cg_r + c.off, *p = ctx->cg_p + c.off, *Q = ctx->cg_Q + c.off;
    double *dsc = ctx->dsc;
    /* bNorm = |b|_1 ; r = b - M x ; resiNorm */
    launch_reduce<1>(ctx, nr, [=] __device__(int64_t i, double(&acc)[1]) { acc[0] += fabs(bvec[i]); }, slot1(SC_CG_B1));
    cg_mvec(ctx, c, x, fixed, r);
    launch_reduce<1>(ctx, nr, [=] __device__(int64_t i, double(&acc)[1]) {
        const double v = bvec[i] - r[i];
        r[i] = v;
        acc[0] = fma(v, v, acc[0]);
    }, slot1(SC_CG_RES));
    CHECK_LAUNCH(ctx);
    TRY(fetch_scalars(ctx, SC_CG_RR, SC_CG_BETA - SC_CG_RR + 1));
    const double bnorm = ctx->hsc[SC_CG_B1];
    double res = sqrt(ctx->hsc[SC_CG_RES]);
    if (res / bnorm < tol) {
        *iters = ctx->cg_last_iter[ci]; /* quirk: cg->iter is not reset on this early exit (lorads_cgs.c:204-217) */
        return 0;
    }
    CU(ctx, cudaMemcpyAsync(p, r, sizeof(double) * nr, cudaMemcpyDeviceToDevice, ctx->stream));
    int64_t it = 0;
    bool have_beta = false; /* p <- beta p + r is pending (applied at the start of the next iteration) */
    for (int64_t k = 0; k < maxit; ++k) {
        it += 1;
        SlotSpec<2> sp;
        sp.slot[0] = SC_CG_RR; sp.slot[1] = SC_CG_PQ; sp.accumulate = 0;
        if (ctx->mc) {
            /* fused: [p = beta p + r,] Q = M p, <r,r>, <p,Q>, alpha -- one launch */
            const int G = pick_group(c.ld);
            {
                Prof pr(ctx, KC_MC_STEP);
                DISPATCH_G(G, k_mc_cg_first<GG><<<grid_for(ctx, c.n * GG, (const void *)k_mc_cg_first<GG>), LGPU_TPB, 0, ctx->stream>>>(
                                  c.n, (int)c.ld, have_beta ? SC_CG_BETA : -1, p, r, fixed, c.rc_ptr, c.rc_a, Q, ctx->partials, ctx->counter,
                                  ctx->dsc, sp, SC_CG_ALPHA));
            }
            if (ctx->world > 1) {
                TRY(allreduce_scalars(ctx, SC_CG_RR, 2));
                launch_scalar(ctx, [=] __device__() { dsc[SC_CG_ALPHA] = dsc[SC_CG_RR] / dsc[SC_CG_PQ]; });
            }
        } else {
            if (have_beta) launch_map(ctx, nr, [=] __device__(int64_t i) { p[i] = fma(dsc[SC_CG_BETA], p[i], r[i]); });
            cg_mvec(ctx, c, p, fixed, Q);
            launch_reduce_post<2>(ctx, nr, [=] __device__(int64_t i, double(&acc)[2]) {
                acc[0] = fma(r[i], r[i], acc[0]);
                acc[1] = fma(p[i], Q[i], acc[1]);
            }, sp, [=] __device__(double *sc) { sc[SC_CG_ALPHA] = sc[SC_CG_RR] / sc[SC_CG_PQ]; });
        }
        /* x += alpha p ; r -= alpha Q ; <r,r> ; and, unless the residual is about to be recomputed, beta = <r,r>_new / <r,r>_old */
        const bool restart = (k % 20 == 0);
        auto upd = [=] __device__(int64_t i, double(&acc)[1]) {
            const double a = dsc[SC_CG_ALPHA];
            x[i] = fma(a, p[i], x[i]);
            const double v = fma(-a, Q[i], r[i]);
            r[i] = v;
            acc[0] = fma(v, v, acc[0]);
        };
        if (restart) launch_reduce_post<1>(ctx, nr, upd, slot1(SC_CG_RES), NoPost());
        else launch_reduce_post<1>(ctx, nr, upd, slot1(SC_CG_RES), [=] __device__(double *sc) { sc[SC_CG_BETA] = sc[SC_CG_RES] / sc[SC_CG_RR]; });
        CHECK_LAUNCH(ctx);
        TRY(fetch_scalars(ctx, SC_CG_RES, 1));
        res = sqrt(ctx->hsc[SC_CG_RES]);
        if (res / bnorm < tol) break;
        if (restart) {
            /* residual recomputed from scratch, p = q = r (lorads_cgs.c:242-258); the update that follows runs with
             * beta = <r,r>/<r,r> = 1, i.e. p = 2 r -- reproduced as is */
            cg_mvec(ctx, c, x, fixed, r);
            launch_reduce_post<1>(ctx, nr, [=] __device__(int64_t i, double(&acc)[1]) {
                const double v = bvec[i] - r[i];
                r[i] = v;
                p[i] = v;
                acc[0] = fma(v, v, acc[0]);
            }, slot1(SC_CG_RR), [=] __device__(double *sc) { sc[SC_CG_BETA] = sc[SC_CG_RR] / sc[SC_CG_RR]; });
        }
        have_beta = true;
    }
    /* (the reference also forms p = beta p + r after its last iteration; p is scratch there, nothing reads it again) */
    ctx->cg_last_iter[ci] = it;
    *iters = it;
    return 0;
}

/* LORADSUpdateSDPVarOne for cone ci: update `upd` with `fixed` held (lorads_admm.c:564-616) */
static int admm_update_one(lgpu_ctx *ctx, int ci, double *upd_flat, const double *fixed_flat, double rho, double tol,
                           int64_t maxit, int64_t *iters)
{
    DevCone &c = ctx->cones[ci];
    const int64_t m = ctx->m;
    if (ctx->mc) {
        /* single cone: constrVal_c == constrValSum ; S V = C V + Diag(A^*(M1)) V */
        double *M1 = ctx->M1;
        const double *b = ctx->b, *cvs = ctx->cvs, *lam = ctx->lam;
        const Owned ow = owned(ctx);
        const int32_t *gid = ow.gid;
        launch_map(ctx, ow.count, [=] __device__(int64_t t) {
            const int64_t i = gid ? gid[t] : t;
            M1[i] = rho * ((-b[i] + cvs[i]) - cvs[i]) - lam[i];
        });
        double *upd = upd_flat, *M2 = ctx->M2, *bl = ctx->bLin;
        const double *fixed = fixed_flat;
        mc_spmm_plain(ctx, fixed, M2);
        const int32_t *rp = c.rc_ptr, *rg = c.rc_gid;
        const double *ra = c.rc_a;
        const int64_t ld = c.ld;
        launch_map(ctx, c.n * c.ld, [=] __device__(int64_t i) {
            const int64_t row = i / ld;
            double coef = 0.0;
            for (int t = rp[row]; t < rp[row + 1]; ++t) coef = fma(M1[rg[t]], ra[t], coef);
            const double m2 = fma(coef, fixed[i], M2[i]) - rho * fixed[i];
            M2[i] = m2;
            bl[i] = (-1.0 / rho) * m2;
        });
        TRY(cg_solve(ctx, ci, upd, fixed, bl, tol, maxit, iters));
        return 0;
    }
    {
        /* M1 = rho (-b + constrValSum - constrVal_c) - lambda */
        double *M1 = ctx->M1;
        const double *b = ctx->b, *cvs = ctx->cvs, *lam = ctx->lam;
        launch_map(ctx, m, [=] __device__(int64_t i) { M1[i] = -b[i] + cvs[i]; });
        const int32_t *gid = c.con_gid;
        const double *cv = c.cv;
        launch_map(ctx, c.mA, [=] __device__(int64_t t) { M1[gid[t]] -= cv[t]; });
        launch_map(ctx, m, [=] __device__(int64_t i) { M1[i] = rho * M1[i] - lam[i]; });
    }
    double *upd = upd_flat + c.off;
    const double *fixed = fixed_flat + c.off;
    double *M2 = ctx->M2 + c.off, *bl = ctx->bLin + c.off;
    run_wsum(ctx, c, ctx->M1, true, true, 1.0, c.S);
    /* M2 = S V - rho V ; bLinSys = -M2 / rho */
    run_spmm(ctx, c, c.ld, c.S, fixed, 1.0, -rho, fixed, M2);
    {
        const double sc = -1.0 / rho;
        launch_map(ctx, c.n * c.ld, [=] __device__(int64_t i) { bl[i] = sc * M2[i]; });
    }
    TRY(cg_solve(ctx, ci, upd, fixed, bl, tol, maxit, iters));
    return 0;
}

/* after a U- or V-update of cone c: constrValSum -= old cv ; cv = A(UV^T) ; constrValSum += cv */
static void refresh_cone_cv(lgpu_ctx *ctx, DevCone &c)
{
    if (ctx->mc) {
        mc_rowdot(ctx, ctx->U, ctx->V, 1.0, ctx->cvs, false);
        return;
    }
    const int32_t *gid = c.con_gid;
    double *cv = c.cv, *cvs = ctx->cvs;
    launch_map(ctx, c.mA, [=] __device__(int64_t t) { cvs[gid[t]] -= cv[t]; });
    run_uvt(ctx, c, c.ld, ctx->U + c.off, ctx->V + c.off, c.uvt);
    run_con_gather(ctx, c, c.uvt, c.cv);
    launch_map(ctx, c.mA, [=] __device__(int64_t t) { cvs[gid[t]] += cv[t]; });
}

/* LORADSUpdateLPVarOne sweep (lorads_admm.c:759-792, lorads_alg_common.c:356-374): strictly sequential over the LP
 * columns; one warp on the device, state stays in HBM */
static int admm_lp_sweep(lgpu_ctx *ctx, double rho)
{
    DevLp &lp = ctx->lp;
    Prof pr(ctx, KC_VEC);
    k_lp_admm_sweep<<<1, 32, 0, ctx->stream>>>(lp.n, rho, lp.c_ptr, lp.c_row, lp.c_val, lp.obj, lp.nrm2sq, ctx->b, ctx->lam, ctx->cvs,
                                               ctx->U + lp.off, ctx->V + lp.off);
    CHECK_LAUNCH(ctx);
    return 0;
}

extern "C" int lgpu_admm_update_var(lgpu_ctx *ctx, double rho, double cg_tol, int64_t cg_max_iter, int64_t *cg_iter_total)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    ctx->cd_valid = false;
    ctx->epi_done = false;
    for (int ci = 0; ci < ctx->ncones; ++ci) {
        DevCone &c = ctx->cones[ci];
        /* Gauss-Seidel over the cones, U then V (lorads_alg_common.c:298-326).  By-cone partition: the owner solves and
         * refreshes the cone's constraint values (only it holds the cone's compact constrVal); the updated segment, the
         * length-m constrValSum and the CG count reach the others through sums with zeros (exact). */
        for (int pass = 0; pass < 2; ++pass) {
            double *upd = pass == 0 ? ctx->U : ctx->V;
            const double *fixed = pass == 0 ? ctx->V : ctx->U;
            int64_t it = 0;
            if (cone_mine(ctx, c)) {
                TRY(admm_update_one(ctx, ci, upd, fixed, rho, cg_tol, cg_max_iter, &it));
                refresh_cone_cv(ctx, c);
            }
            if (ctx->cone_par) {
                const int owner = ctx->cone_owner[ci];
                TRY(cone_bcast(ctx, upd + c.off, (size_t)(c.n_alloc * c.ld), owner));
                TRY(cone_bcast(ctx, ctx->cvs, (size_t)ctx->m, owner));
                ctx->hsc[SC_TMP] = ctx->rank == owner ? (double)it : 0.0;
                CU(ctx, cudaMemcpyAsync(ctx->dsc + SC_TMP, ctx->hsc + SC_TMP, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
                TRY(allreduce_scalars(ctx, SC_TMP, 1));
                TRY(fetch_scalars(ctx, SC_TMP, 1));
                it = (int64_t)(ctx->hsc[SC_TMP] + 0.5);
            }
            *cg_iter_total += it;
        }
    }
    CHECK_LAUNCH(ctx);
    if (ctx->lp.n > 0) TRY(admm_lp_sweep(ctx, rho));
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * oracle-rank Gram
 * ------------------------------------------------------------------------------------------------*/
extern "C" int lgpu_gram(lgpu_ctx *ctx, int phase, int cone, double *gram)
{
    if (!ctx || !ctx->vars_ready || cone < 0 || cone >= ctx->ncones) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    const int r = (int)c.r;
    const int nt = (r + 15) / 16;
    /* enough (chunk, tile) units to fill the machine on small cones, 4096-row chunks on large ones */
    int64_t rows_per_chunk = 4096;
    if (ctx->dense_dmma) {
        const int64_t want = (c.n * nt * nt + (int64_t)ctx->num_sms * 8 - 1) / ((int64_t)ctx->num_sms * 8);
        rows_per_chunk = std::min<int64_t>(4096, std::max<int64_t>(64, (want + 3) / 4 * 4));
    }
    int nchunks = (int)((c.n + rows_per_chunk - 1) / rows_per_chunk);
    const size_t part_bytes = sizeof(double) * (size_t)nchunks * nt * nt * 256;
    const size_t gram_bytes = sizeof(double) * (size_t)r * r;
    TRY(ensure_dstage(ctx, part_bytes + gram_bytes));
    double *part = (double *)ctx->dstage;
    double *dg = part + (size_t)nchunks * nt * nt * 256;
    const double *A = (phase == 1 ? ctx->R : ctx->U) + c.off;
    const double *B = ctx->V + c.off;
    dim3 grid(nchunks, nt * nt);
    if (nchunks == 0) {
        /* a partition rank that owns no row of this cone: nothing to launch, k_gram_finish below writes zeros */
    } else if (ctx->dense_dmma) {
        /* FP64 tensor pipe (DMMA), the default; lgpu_set_dense_tensor_path(0) selects the FMA kernel for A/B tests */
        Prof pr(ctx, KC_DENSE);
        const int64_t units = (int64_t)nchunks * nt * nt;
        const int64_t blocks = std::min<int64_t>((units + 3) / 4, (int64_t)ctx->num_sms * 8); /* 4 warps per CTA */
        k_gram_dmma<<<(unsigned)blocks, 128, 0, ctx->stream>>>(c.n, r, (int)c.ld, A, B, phase == 1 ? 0 : 1, rows_per_chunk, nchunks,
                                                               part);
    } else {
        Prof pr(ctx, KC_LAYOUT);
        k_gram_partial<<<grid, 256, 0, ctx->stream>>>(c.n, r, (int)c.ld, A, B, phase == 1 ? 0 : 1, rows_per_chunk, part);
    }
    {
        Prof pr(ctx, KC_LAYOUT);
        k_gram_finish<<<nt * nt, 256, 0, ctx->stream>>>(nchunks, nt * nt, r, part, dg);
    }
    CHECK_LAUNCH(ctx);
    if (ctx->world > 1 && !ctx->cone_par) /* by cone: the factors are replicated, the Gram is already whole */
        NC(ctx, g_nccl.AllReduce(dg, dg, (size_t)r * r, LG_NCCL_FLOAT64, LG_NCCL_SUM, (lg_ncclComm_t)ctx->comm, ctx->stream));
    CU(ctx, cudaMemcpyAsync(gram, dg, gram_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * dual infeasibility: lambda_min(C - A*(lambda)) per cone by Lanczos on the device
 * ------------------------------------------------------------------------------------------------*/
/* extreme eigenvalue of the k x k Lanczos tridiagonal (diagonal a, off-diagonal b) by Sturm bisection;
 * which = -1: smallest, +1: largest */
static double tridiag_extreme_eig(const std::vector<double> &a, const std::vector<double> &b, int k, int which)
{
    double lo = a[0], hi = a[0];
    for (int i = 0; i < k; ++i) {
        double rr = 0.0;
        if (i > 0) rr += fabs(b[i - 1]);
        if (i < k - 1) rr += fabs(b[i]);
        lo = std::min(lo, a[i] - rr);
        hi = std::max(hi, a[i] + rr);
    }
    const int target = which < 0 ? 1 : k; /* number of eigenvalues < x we bracket */
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        int cnt = 0;
        double d = a[0] - mid;
        if (d < 0) cnt++;
        for (int i = 1; i < k; ++i) {
            const double dd = (d == 0.0) ? 1e-300 : d;
            d = a[i] - mid - b[i - 1] * b[i - 1] / dd;
            if (d < 0) cnt++;
        }
        if (cnt >= target) hi = mid; else lo = mid;
        if (hi - lo <= 1e-15 * (fabs(lo) + fabs(hi)) + 1e-300) break;
    }
    return 0.5 * (lo + hi);
}
/* |last component| of the unit eigenvector of T_k for its eigenvalue theta; times beta_k it is the residual norm of the
 * Ritz pair.  Two steps of inverse iteration on T_k - (theta - delta) I with a pivoted tridiagonal elimination (the
 * dgtsv scheme).  The obvious three-term recurrence from the first component is unstable exactly where it matters: once
 * the extreme Ritz pair has converged its last component is ~1e-9, the recurrence returns ~1e-3, and the Lanczos ran to
 * its step limit on every instance (250-300 steps where ~50 suffice). */
static double tridiag_last_component(const std::vector<double> &a, const std::vector<double> &b, int k, double theta)
{
    if (k == 1) return 1.0;
    double scale = 0.0;
    for (int i = 0; i < k; ++i) scale = std::max(scale, fabs(a[i]) + (i < k - 1 ? fabs(b[i]) : 0.0));
    scale = std::max(scale, 1e-300);
    const double shift = theta - 1e-13 * scale;
    std::vector<double> x((size_t)k, 1.0 / sqrt((double)k)), dl((size_t)k), d((size_t)k), du((size_t)k);
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < k; ++i) { d[i] = a[i] - shift; dl[i] = du[i] = (i < k - 1) ? b[i] : 0.0; }
        for (int i = 0; i < k - 1; ++i) {
            if (fabs(d[i]) >= fabs(dl[i])) {
                if (d[i] == 0.0) d[i] = 1e-300;
                const double f = dl[i] / d[i];
                d[i + 1] -= f * du[i];
                x[i + 1] -= f * x[i];
                dl[i] = 0.0; /* from here on dl[i] is the second super-diagonal of the eliminated matrix */
            } else {
                const double f = d[i] / dl[i];
                d[i] = dl[i];
                const double t = d[i + 1];
                d[i + 1] = du[i] - f * t;
                if (i < k - 2) { dl[i] = du[i + 1]; du[i + 1] = -f * dl[i]; }
                else dl[i] = 0.0;
                du[i] = t;
                const double tx = x[i];
                x[i] = x[i + 1];
                x[i + 1] = tx - f * x[i + 1];
            }
        }
        for (int i = 0; i < k; ++i)
            if (fabs(d[i]) < 1e-300 * 1e16) d[i] = (d[i] < 0 ? -1.0 : 1.0) * 1e-16 * scale; /* singular to rounding: fine for inverse iteration */
        x[k - 1] /= d[k - 1];
        x[k - 2] = (x[k - 2] - du[k - 2] * x[k - 1]) / d[k - 2];
        for (int i = k - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - dl[i] * x[i + 2]) / d[i];
        double nrm = 0.0, big = 0.0;
        for (int i = 0; i < k; ++i) big = std::max(big, fabs(x[i]));
        if (!(big > 0.0) || !std::isfinite(big)) return 1.0; /* give up: report "not converged" */
        for (int i = 0; i < k; ++i) { x[i] /= big; nrm += x[i] * x[i]; }
        nrm = sqrt(nrm);
        for (int i = 0; i < k; ++i) x[i] /= nrm;
    }
    return fabs(x[k - 1]);
}

/* Lanczos step: w = S q - bprev qm and alpha = <q, S q>, one warp per row of the symmetric CSR (lanes stride over the
 * row's entries, so a hub row of thousands of entries costs ~n/32 steps instead of n)
 *     reference: sdp_coeff.mv = dataMatSparseMV / dataMatDenseMV, lorads_sdp_data.c:772-787,983-1006 */
__global__ void __launch_bounds__(LGPU_TPB) k_lanczos_symv(int64_t n, const int32_t *__restrict__ fp, const int32_t *__restrict__ fc,
                                                           const int32_t *__restrict__ fs, const double *__restrict__ Sv,
                                                           const double *__restrict__ q, const double *__restrict__ qm,
                                                           const double *__restrict__ bprev_p, double *__restrict__ w,
                                                           double *partials, unsigned int *counter, double *dsc, SlotSpec<1> spec)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double bprev = qm ? *bprev_p : 0.0; /* beta_{k-1} stays on the device between read-backs */
    double red[1] = {0.0};
    for (int64_t i = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < n; i += warps) {
        double a = 0.0;
        for (int e = fp[i] + lane; e < fp[i + 1]; e += 32) a = fma(Sv[fs[e]], q[fc[e]], a);
        a = warp_sum(a);
        if (lane == 0) {
            red[0] = fma(q[i], a, red[0]);
            w[i] = qm ? fma(-bprev, qm[i], a) : a;
        }
    }
    grid_reduce_finish<1>(red, partials, counter, dsc, spec);
}

/* re-orthogonalisation of w against the first k basis vectors in two launches: h = Q^T w (one block per basis
 * vector, fixed summation order), then w -= Q h */
__global__ void __launch_bounds__(LGPU_TPB) k_basis_dots(int64_t n, const double *__restrict__ Q, const double *__restrict__ w,
                                                         double *__restrict__ h)
{
    __shared__ double sh[LGPU_TPB / 32];
    const double *q = Q + (size_t)blockIdx.x * n;
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += LGPU_TPB) a = fma(q[i], w[i], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < LGPU_TPB / 32; ++k) t += sh[k];
        h[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(LGPU_TPB) k_basis_update(int64_t n, int k, const double *__restrict__ Q,
                                                           const double *__restrict__ h, double *__restrict__ w)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = w[i];
        for (int j = 0; j < k; ++j) a = fma(-h[j], Q[(size_t)j * n + i], a);
        w[i] = a;
    }
}

/* The same two kernels with the Lanczos step's own update fused in (w1 = w - alpha q_k is formed on the fly, with the same
 * fma, instead of by a separate launch) and |w|^2 -- hence beta_k and 1 / beta_k, by `post` -- reduced by the update kernel
 * itself: 4 launches per step instead of 6 on the latency-bound cones. */
__global__ void __launch_bounds__(LGPU_TPB) k_basis_dots_fused(int64_t n, const double *__restrict__ Q, const double *__restrict__ w,
                                                               const double *__restrict__ qk, const double *__restrict__ alpha_p,
                                                               double *__restrict__ h)
{
    __shared__ double sh[LGPU_TPB / 32];
    const double *q = Q + (size_t)blockIdx.x * n;
    const double alpha = *alpha_p;
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += LGPU_TPB) a = fma(q[i], fma(-alpha, qk[i], w[i]), a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < LGPU_TPB / 32; ++k) t += sh[k];
        h[blockIdx.x] = t;
    }
}
template <class P>
__global__ void __launch_bounds__(LGPU_TPB) k_basis_update_fused(int64_t n, int k, const double *__restrict__ Q,
                                                                 const double *__restrict__ h, double *__restrict__ w,
                                                                 const double *__restrict__ qk, const double *__restrict__ alpha_p,
                                                                 double *partials, unsigned int *counter, double *dsc, SlotSpec<1> spec,
                                                                 P post)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double alpha = *alpha_p;
    double red[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = fma(-alpha, qk[i], w[i]);
        for (int j = 0; j < k; ++j) a = fma(-h[j], Q[(size_t)j * n + i], a);
        w[i] = a;
        red[0] = fma(a, a, red[0]);
    }
    grid_reduce_finish<1, P>(red, partials, counter, dsc, spec, post);
}

/* S q for the partitioned fused layout: this rank's rows of (C - Diag(sum_k lambda_k a_k)) q into w at their GLOBAL
 * positions; q is a replicated full-length vector.  Column ids are global (all-gather mode) or local/halo ids that
 * halo_gid maps back to global rows. */
__global__ void __launch_bounds__(LGPU_TPB) k_lanczos_symv_part(int64_t nloc, int64_t row_lo, int64_t nsplit,
                                                                const int32_t *__restrict__ fp, const int32_t *__restrict__ fc,
                                                                const double *__restrict__ fv, const int32_t *__restrict__ halo_gid,
                                                                const int32_t *__restrict__ rcptr, const int32_t *__restrict__ rcgid,
                                                                const double *__restrict__ rca, const double *__restrict__ lam,
                                                                const double *__restrict__ q, double *__restrict__ w)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < nloc; i += warps) {
        double a = 0.0;
        for (int e = fp[i] + lane; e < fp[i + 1]; e += 32) {
            const int col = fc[e];
            const int64_t g = halo_gid ? (col < nsplit ? row_lo + col : (int64_t)halo_gid[col - nsplit]) : (int64_t)col;
            a = fma(fv[e], q[g], a);
        }
        a = warp_sum(a);
        if (lane == 0) {
            double d = 0.0;
            for (int t = rcptr[i]; t < rcptr[i + 1]; ++t) d = fma(lam[rcgid[t]], rca[t], d);
            w[row_lo + i] = fma(-d, q[row_lo + i], a);
        }
    }
}

/* Lanczos for the smallest eigenvalue of a symmetric operator of dimension n.  `apply(qk, qm, bprev, w)` must leave
 * w = S qk - bprev qm (qm may be null) and dsc[SC_LANCZOS] = <qk, S qk>.  Stop: Ritz residual |beta_k s_k| <= 1e-6 x
 * (spectral scale of T_k).  Small problems keep the whole Krylov basis and re-orthogonalise against it (two launches per
 * step); large ones run the plain three-term recurrence, whose extreme Ritz value stays accurate without it. */
typedef std::function<int(const double *, const double *, const double *, double *)> LanczosApply;
static int lanczos_min_eig(lgpu_ctx *ctx, int64_t n, int64_t vec_len, const LanczosApply &apply, double *theta_out)
{
    double *dsc = ctx->dsc;
    const int kmax = (int)std::min<int64_t>(n, 300);
    const int batch = 16; /* steps between two read-backs of (alpha, beta) */
    const bool full = (double)vec_len * (double)(kmax + 1) * 8.0 <= 256.0e6;
    const size_t vec = (size_t)vec_len;
    const size_t nvec = full ? (size_t)kmax + 2 : 4;
    TRY(ensure_dstage(ctx, sizeof(double) * (vec * nvec + 3 * (size_t)kmax + 8)));
    double *Q = (double *)ctx->dstage; /* full: q_0 .. q_kmax ; else ring of 3 */
    double *w = Q + vec * (nvec - 1);
    double *hbuf = w + vec;            /* [kmax] basis dot products */
    double *dal = hbuf + kmax + 2;     /* [kmax] alpha_k, on the device until the next read-back */
    double *dbe = dal + kmax;          /* [kmax] beta_k */
    CU(ctx, cudaMemsetAsync(Q, 0, sizeof(double) * vec * nvec, ctx->stream)); /* padding entries stay zero */
    auto qptr = [&](int k) { return Q + vec * (size_t)(full ? k : (k % 3)); };
    {
        double *q0 = qptr(0);
        launch_reduce<1>(ctx, n, [=] __device__(int64_t i, double(&acc)[1]) {
            uint64_t z = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
            z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
            const double v = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
            q0[i] = v;
            acc[0] = fma(v, v, acc[0]);
        }, slot1(SC_LANCZOS));
        launch_map(ctx, n, [=] __device__(int64_t i) { q0[i] /= sqrt(dsc[SC_LANCZOS]); });
    }
    /* The recurrence runs `batch` steps at a time without a host round trip: alpha_k, beta_k and 1 / beta_k are formed
     * by the finishing thread of the |w|^2 reduction and stay on the device.  After a batch the host reads the new
     * (alpha, beta) and applies the stopping test to every prefix in order, so the result is the one the step-by-step
     * test would have returned; the (at most batch - 1) surplus steps cost a few launches. */
    std::vector<double> al((size_t)kmax), be((size_t)kmax);
    double theta = 0.0, theta_prev = 0.0;
    int checked = 0;
    bool done = false, have_prev = false;
    for (int k = 0; k < kmax && !done; ++k) {
        const double *qk = qptr(k);
        const double *qm = k > 0 ? qptr(k - 1) : nullptr;
        TRY(apply(qk, qm, k > 0 ? dbe + (k - 1) : nullptr, w));
        /* w -= alpha_k q_k ; then (small problems) against the whole basis ; beta_k = |w| */
        double *ak = dal + k, *bk = dbe + k;
        auto finish = [=] __device__(double *sc) {
            const double b = sqrt(sc[SC_LANCZOS + 1]);
            *ak = sc[SC_LANCZOS];
            *bk = b;
            sc[SC_LANCZOS + 3] = b > 0.0 ? 1.0 / b : 0.0; /* an exhausted Krylov space yields zeros, not NaNs */
        };
        if (full) {
            {
                Prof pr(ctx, KC_REDUCE);
                k_basis_dots_fused<<<k + 1, LGPU_TPB, 0, ctx->stream>>>(vec_len, Q, w, qk, dsc + SC_LANCZOS, hbuf);
            }
            {
                Prof pr(ctx, KC_VEC);
                k_basis_update_fused<<<grid_for(ctx, n), LGPU_TPB, 0, ctx->stream>>>(vec_len, k + 1, Q, hbuf, w, qk, dsc + SC_LANCZOS,
                                                                                    ctx->partials, ctx->counter, ctx->dsc,
                                                                                    slot1(SC_LANCZOS + 1), finish);
            }
        } else {
            launch_reduce_post<1>(ctx, n, [=] __device__(int64_t i, double(&acc)[1]) {
                const double v = fma(-dsc[SC_LANCZOS], qk[i], w[i]);
                w[i] = v;
                acc[0] = fma(v, v, acc[0]);
            }, slot1(SC_LANCZOS + 1), finish);
        }
        CHECK_LAUNCH(ctx);
        if (k + 1 < kmax) {
            double *qn = qptr(k + 1);
            launch_map(ctx, n, [=] __device__(int64_t i) { qn[i] = w[i] * dsc[SC_LANCZOS + 3]; });
        }
        if ((k + 1) % batch != 0 && k + 1 < kmax) continue;
        /* read the new coefficients back and test the prefixes checked .. k */
        CU(ctx, cudaMemcpyAsync(al.data() + checked, dal + checked, sizeof(double) * (size_t)(k + 1 - checked), cudaMemcpyDeviceToHost,
                                ctx->stream));
        CU(ctx, cudaMemcpyAsync(be.data() + checked, dbe + checked, sizeof(double) * (size_t)(k + 1 - checked), cudaMemcpyDeviceToHost,
                                ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        /* The stopping test is applied to the newest prefix only: the host-side eigenvalue work per test is O(k) per
         * bisection step, and testing all 300 prefixes cost 30-40 ms on the small instances -- several times the device time
         * of the whole recurrence.  Running to the end of a batch costs at most 15 cheap steps and a later Ritz value is only
         * closer to lambda_min (it converges monotonically from above).  An exhausted Krylov space (beta ~ 0) inside the batch
         * is honoured at its own prefix. */
        int jtest = k;
        for (int j = checked; j <= k; ++j)
            if (be[(size_t)j] <= 1e-14 * std::max(fabs(al[(size_t)j]), 1e-300)) { jtest = j; break; }
        {
            const int j = jtest;
            const double bnorm = be[(size_t)j];
            theta = tridiag_extreme_eig(al, be, j + 1, -1);
            const double top = tridiag_extreme_eig(al, be, j + 1, +1);
            const double scale = std::max(std::max(fabs(theta), fabs(top)), 1e-300);
            const double resid = bnorm * tridiag_last_component(al, be, j + 1, theta);
            /* converged: the Ritz residual is small -- or the Ritz VALUE has stopped moving over a whole batch while the
             * residual is inside the reference's own ARPACK tolerance (1e-2, lorads_sdp_conic.c:1668): at a solution the
             * slack matrix has a cluster of ~rank eigenvalues at zero, the vector of one of them never settles (residual
             * stuck near 1e-4) although its value did long ago, and the recurrence ran to its 300-step limit.  1e-7 of the
             * spectral scale per 16 steps is ~1e-9 of the reported figure (it is divided by 1 + |C|_1), four orders inside
             * the solver's tolerances. */
            const bool stalled = have_prev && fabs(theta - theta_prev) <= 1e-7 * scale && resid <= 1e-2 * scale;
            theta_prev = theta;
            have_prev = true;
            if (resid <= 1e-6 * scale || bnorm <= 1e-14 * scale || jtest < k || stalled) done = true;
            else if (k + 1 >= kmax) {
                /* not converged: a Ritz value is only an UPPER bound of lambda_min, which would under-report the dual
                 * infeasibility; report the residual-corrected value (an eigenvalue lies within `resid` of theta) and say so */
                if (kmax < n) {
                    fprintf(stderr, "lorads_b200: warning: Lanczos stopped after %d steps with Ritz residual %.3e (scale %.3e); "
                                    "dual infeasibility uses theta - residual\n", kmax, resid, scale);
                    theta -= resid;
                }
                done = true;
            }
        }
        checked = k + 1;
    }
    *theta_out = theta;
    return 0;
}

/* calculate_dual_infeasibility_solver (lorads_solver.c:1396-1426, lorads_sdp_conic.c:1636-1699): the reference asks
 * ARPACK (dsaupd/dseupd, "SA", nev = 1, tol = 1e-2) for lambda_min(C - A^*(lambda)) of every cone.  Here: Lanczos on the
 * device.  One GPU: S on the pattern, applied through the full symmetric CSR (sdp_coeff.mv, lorads_sdp_data.c:772-787,
 * 983-1006).  Partitioned (fused layout): the Lanczos vectors are replicated full-length vectors, every rank applies its
 * rows of S and the pieces are all-gathered; all other vector work is done redundantly on every rank, so every rank
 * obtains the same scalars without further reductions. */
extern "C" int lgpu_dual_infeasibility(lgpu_ctx *ctx, double *sum_neg_eig)
{
    if (!ctx || !ctx->vars_ready) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    double total = 0.0;
    if (ctx->world > 1 && !ctx->cone_par) {
        DevCone &c = ctx->cones[0];
        const int64_t n = c.n_glob, vec_len = (int64_t)ctx->world * c.n_alloc;
        const int32_t *hg = ctx->use_halo ? ctx->halo_gid : nullptr;
        double theta = 0.0;
        ctx->defer_allreduce = true; /* full-length replicated vectors: the reductions below are already global */
        auto apply = [&](const double *qk, const double *qm, const double *bprev_p, double *w) -> int {
            {
                Prof pr(ctx, KC_SPMM);
                k_lanczos_symv_part<<<grid_for(ctx, c.n * 32, (const void *)k_lanczos_symv_part), LGPU_TPB, 0, ctx->stream>>>(
                    c.n, c.row_lo, c.n_alloc, c.f_ptr, c.f_col, c.mc_val, hg, c.rc_ptr, c.rc_gid, c.rc_a, ctx->lam, qk, w);
            }
            NC(ctx, g_nccl.AllGather(w + (size_t)ctx->rank * c.n_alloc, w, (size_t)c.n_alloc, LG_NCCL_FLOAT64,
                                     (lg_ncclComm_t)ctx->comm, ctx->stream));
            launch_reduce<1>(ctx, n, [=] __device__(int64_t i, double(&acc)[1]) {
                const double a = w[i];
                acc[0] = fma(qk[i], a, acc[0]);
                if (qm) w[i] = fma(-*bprev_p, qm[i], a);
            }, slot1(SC_LANCZOS));
            return 0;
        };
        const int rc = lanczos_min_eig(ctx, n, vec_len, apply, &theta);
        ctx->defer_allreduce = false;
        if (rc) return rc;
        *sum_neg_eig = fabs(std::min(theta, 0.0));
        return 0;
    }
    /* LP part (lorads_solver.c:1404-1412): |min(c_j - a_j^T lambda, 0)| */
    if (ctx->lp.n > 0) {
        const int32_t *cp = ctx->lp.c_ptr, *cr = ctx->lp.c_row;
        const double *cv = ctx->lp.c_val, *obj = ctx->lp.obj, *lam = ctx->lam;
        launch_reduce<1>(ctx, ctx->lp.n, [=] __device__(int64_t j, double(&acc)[1]) {
            double w = obj[j];
            for (int e = cp[j]; e < cp[j + 1]; ++e) w = fma(-lam[cr[e]], cv[e], w);
            acc[0] += fabs(fmin(w, 0.0));
        }, slot1(SC_LANCZOS + 2));
        CHECK_LAUNCH(ctx);
        TRY(fetch_scalars(ctx, SC_LANCZOS + 2, 1));
        if (cone_lead(ctx)) total += ctx->hsc[SC_LANCZOS + 2];
    }
    for (auto &c : ctx->cones) {
        if (!cone_mine(ctx, c)) continue; /* by cone: the owner runs the cone's Lanczos, the sum is all-reduced below */
        /* slack S = C - sum lambda_i A_i on the pattern */
        run_wsum(ctx, c, ctx->lam, true, true, -1.0, c.S);
        const int32_t *fp = c.f_ptr, *fc = c.f_col, *fs = c.f_slot;
        const double *Sv = c.S;
        const int64_t n = c.n;
        double theta = 0.0;
        auto apply = [&](const double *qk, const double *qm, const double *bprev_p, double *w) -> int {
            Prof pr(ctx, KC_SPMM);
            k_lanczos_symv<<<grid_for(ctx, n * 32, (const void *)k_lanczos_symv), LGPU_TPB, 0, ctx->stream>>>(
                n, fp, fc, fs, Sv, qk, qm, bprev_p, w, ctx->partials, ctx->counter, ctx->dsc, slot1(SC_LANCZOS));
            return 0;
        };
        TRY(lanczos_min_eig(ctx, n, n, apply, &theta));
        total += fabs(std::min(theta, 0.0));
    }
    if (ctx->cone_par) {
        ctx->hsc[SC_TMP] = total;
        CU(ctx, cudaMemcpyAsync(ctx->dsc + SC_TMP, ctx->hsc + SC_TMP, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        TRY(allreduce_scalars(ctx, SC_TMP, 1));
        TRY(fetch_scalars(ctx, SC_TMP, 1));
        total = ctx->hsc[SC_TMP];
    }
    *sum_neg_eig = total;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * operator-level entry points on host buffers
 * ------------------------------------------------------------------------------------------------*/
struct TmpLayout {
    int64_t ld = 0;
    double *U = nullptr, *V = nullptr, *Y = nullptr;
};

/* temporary row-major copies of host factors with rank r for cone c (independent of the solver variables) */
static int op_stage_factors(lgpu_ctx *ctx, DevCone &c, int64_t r, const double *U, const double *V, bool needY, TmpLayout *t)
{
    if (r <= 0) LGPU_FAIL(ctx, "bad rank");
    t->ld = (r + 3) & ~(int64_t)3;
    const size_t sz = (size_t)c.n * t->ld;
    CU(ctx, cudaMalloc((void **)&t->U, sizeof(double) * sz));
    TRY(upload_factor(ctx, c.n, r, t->ld, U, t->U, -1, 0, &c));
    if (V != nullptr && V != U) {
        CU(ctx, cudaMalloc((void **)&t->V, sizeof(double) * sz));
        TRY(upload_factor(ctx, c.n, r, t->ld, V, t->V, -1, 0, &c));
    } else {
        t->V = t->U;
    }
    if (needY) CU(ctx, cudaMalloc((void **)&t->Y, sizeof(double) * sz));
    return 0;
}
static void op_free(TmpLayout *t)
{
    if (t->V && t->V != t->U) cudaFree(t->V);
    if (t->U) cudaFree(t->U);
    if (t->Y) cudaFree(t->Y);
}

extern "C" int lgpu_op_uvt(lgpu_ctx *ctx, int cone, int64_t r, const double *U, const double *V, double *uvt)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    TmpLayout t;
    int rc = op_stage_factors(ctx, c, r, U, V, false, &t);
    if (!rc) {
        run_uvt(ctx, c, t.ld, t.U, t.V, c.uvt);
        cudaError_t e = cudaMemcpyAsync(uvt, c.uvt, sizeof(double) * c.nnzP, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = 1; }
    }
    op_free(&t);
    return rc;
}

extern "C" int lgpu_op_auv(lgpu_ctx *ctx, int cone, int64_t r, const double *U, const double *V, double *constr_val, double *obj)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    TmpLayout t;
    int rc = op_stage_factors(ctx, c, r, U, V, false, &t);
    if (!rc) {
        double *dsc = ctx->dsc;
        launch_scalar(ctx, [=] __device__() { dsc[SC_OBJ] = 0.0; });
        run_uvt(ctx, c, t.ld, t.U, t.V, c.uvt);
        run_obj_gather(ctx, c, c.uvt, SC_OBJ, 1);
        run_con_gather(ctx, c, c.uvt, c.cv);
        double *out = ctx->mtmp;
        cudaMemsetAsync(out, 0, sizeof(double) * ctx->m, ctx->stream);
        const int32_t *gid = c.con_gid;
        const double *cv = c.cv;
        launch_map(ctx, c.mA, [=] __device__(int64_t q) { out[gid[q]] = cv[q]; });
        std::vector<double> tmp; /* renumbered constraints (row relabelling): device order -> caller's order below */
        if (!ctx->h_cperm.empty()) tmp.resize((size_t)ctx->m);
        cudaError_t e = cudaMemcpyAsync(tmp.empty() ? constr_val : tmp.data(), out, sizeof(double) * ctx->m, cudaMemcpyDeviceToHost,
                                        ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = 1; }
        if (!rc) rc = fetch_scalars(ctx, SC_OBJ, 1); /* stream-ordered after the copy */
        if (!rc && obj) *obj = ctx->hsc[SC_OBJ];
        if (!rc && !tmp.empty())
            for (int64_t k = 0; k < ctx->m; ++k) constr_val[k] = tmp[ctx->h_cperm[k]];
    }
    op_free(&t);
    return rc;
}

static int op_upload_w(lgpu_ctx *ctx, const double *w)
{
    if (!ctx->h_cperm.empty()) { /* renumbered constraints (row relabelling): caller's order -> device order */
        std::vector<double> tmp((size_t)ctx->m);
        for (int64_t k = 0; k < ctx->m; ++k) tmp[ctx->h_cperm[k]] = w[k];
        CU(ctx, cudaMemcpyAsync(ctx->mtmp, tmp.data(), sizeof(double) * ctx->m, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    CU(ctx, cudaMemcpyAsync(ctx->mtmp, w, sizeof(double) * ctx->m, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

extern "C" int lgpu_op_wsum(lgpu_ctx *ctx, int cone, const double *w, int add_obj, double *S)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    TRY(op_upload_w(ctx, w));
    run_wsum(ctx, c, ctx->mtmp, true, add_obj != 0, 1.0, c.S);
    CHECK_LAUNCH(ctx);
    CU(ctx, cudaMemcpyAsync(S, c.S, sizeof(double) * c.nnzP, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lgpu_op_wsum_mulrk(lgpu_ctx *ctx, int cone, int64_t r, const double *w, int add_obj, const double *X, double *Y)
{
    if (!ctx || cone < 0 || cone >= ctx->ncones) return 1;
    CU(ctx, cudaSetDevice(ctx->device));
    DevCone &c = ctx->cones[cone];
    TmpLayout t;
    int rc = op_stage_factors(ctx, c, r, X, nullptr, true, &t);
    if (!rc) rc = op_upload_w(ctx, w);
    if (!rc) {
        run_wsum(ctx, c, ctx->mtmp, true, add_obj != 0, 1.0, c.S);
        run_spmm(ctx, c, t.ld, c.S, t.U, 1.0, 0.0, nullptr, t.Y);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = 1; }
        if (!rc) rc = download_factor(ctx, c.n, r, t.ld, t.Y, Y, -1, 0, &c);
    }
    op_free(&t);
    return rc;
}
