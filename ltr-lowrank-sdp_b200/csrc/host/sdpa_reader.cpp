/*
 * sdpa_reader.cpp -- SDPA sparse (.dat-s) reader, 64-bit clean, multi-threaded.
 *
 * Produces the same logical data the reference reader hands to the solver (LReadSDPA,
 * lorads/src/src_semi/io/lorads_file_io.c:59-455): per SDP block a CSC over PACKED lower-triangular
 * indices whose column 0 is the (negated) objective and columns 1..m the constraints, plus an LP CSC for
 * a trailing negative-dimension block.
 *
 * The reference reads one line at a time with sscanf into a growing triplet store and compresses it
 * afterwards (:260-346): at n = 1e7 (a 1.8 GB file, 6e7 entries) that is minutes.  Here:
 *   - the file is mapped, the entry section is cut at line boundaries into one piece per thread;
 *   - numbers are converted by an exact fast path (<= 19 significant digits whose value fits 2^53, decimal
 *     exponent within +-22: one correctly rounded multiply or divide, Clinger 1990), anything else goes to
 *     strtod, so every value is bit-identical to the sscanf("%lg") the reference uses;
 *   - the CSC is built by a counting sort over the constraint column (stable in file order, which is the
 *     order the reference's dcs_compress keeps), long columns are placed piece by piece in parallel, short
 *     ones by column range, and columns whose packed indices are not already ascending are merge-sorted
 *     (stable), the long ones by all threads together.
 *
 * File rules mirrored from the reference: comment lines start with '*' or '"' (:104-108); the block
 * dimension line may carry { } ( ) ' , (:143-175); only the LAST block may be LP (:159-190); b is free
 * form with commas (:203-220); entries are `con blk i j val`, 1-based, either triangle (:260-331); the
 * entry section ends at EOF or at the first line that is not an entry (BEGIN.COMMENT); |val| < 1e-12 is
 * dropped with one warning (:288-294); objective entries are negated (:317-319).
 * LORADS_READ_THREADS overrides the thread count (default: online cores, at most 32; 1 for small files).
 */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <thread>
#include <vector>

#include "lorads_host.h"

namespace {

struct Ent {
    int64_t col; /* constraint column (0 = objective) */
    int64_t idx; /* packed lower index (SDP) or LP column id */
    double val;
};
struct IdxVal {
    int64_t idx;
    double val;
};

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

/* "%ld": blanks, optional sign, digits.  Never reads at or past e. */
inline bool parse_i64(const char *&p, const char *e, int64_t *out)
{
    const char *q = p;
    while (q < e && is_blank(*q)) ++q;
    bool neg = false;
    if (q < e && (*q == '-' || *q == '+')) { neg = *q == '-'; ++q; }
    if (q >= e || *q < '0' || *q > '9') return false;
    uint64_t v = 0;
    while (q < e && *q >= '0' && *q <= '9') { v = v * 10 + (uint64_t)(*q - '0'); ++q; }
    *out = neg ? -(int64_t)v : (int64_t)v;
    p = q;
    return true;
}

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

/* the library conversion on a bounded copy of the token (the mapping is not NUL-terminated) */
bool parse_f64_slow(const char *&p, const char *e, double *out)
{
    char tmp[128];
    size_t k = 0;
    while (p + k < e && k + 1 < sizeof(tmp) && p[k] != '\n') { tmp[k] = p[k]; ++k; }
    tmp[k] = '\0';
    char *endp;
    const double v = strtod(tmp, &endp);
    if (endp == tmp) return false;
    *out = v;
    p += endp - tmp;
    return true;
}

/* "%lg": bit-identical to strtod */
inline bool parse_f64(const char *&p, const char *e, double *out)
{
    const char *q = p;
    while (q < e && is_blank(*q)) ++q;
    const char *tok = q;
    bool neg = false;
    if (q < e && (*q == '-' || *q == '+')) { neg = *q == '-'; ++q; }
    uint64_t mant = 0;
    int sig = 0, frac = 0, ndig = 0;
    while (q < e && *q >= '0' && *q <= '9') {
        if (sig > 0 || *q != '0') { if (sig < 19) mant = mant * 10 + (uint64_t)(*q - '0'); ++sig; }
        ++ndig; ++q;
    }
    if (q < e && *q == '.') {
        ++q;
        while (q < e && *q >= '0' && *q <= '9') {
            if (sig > 0 || *q != '0') { if (sig < 19) mant = mant * 10 + (uint64_t)(*q - '0'); ++sig; }
            ++frac; ++ndig; ++q;
        }
    }
    if (ndig == 0) { const char *s = tok; if (!parse_f64_slow(s, e, out)) return false; p = s; return true; } /* inf, nan, junk */
    int ex = 0;
    if (q < e && (*q == 'e' || *q == 'E')) {
        const char *r = q + 1;
        bool eneg = false;
        if (r < e && (*r == '-' || *r == '+')) { eneg = *r == '-'; ++r; }
        if (r < e && *r >= '0' && *r <= '9') {
            int v = 0;
            while (r < e && *r >= '0' && *r <= '9') { if (v < 100000) v = v * 10 + (*r - '0'); ++r; }
            ex = eneg ? -v : v;
            q = r;
        }
    }
    /* hexadecimal floats ("0x..") and over-long mantissas take the library path */
    const bool hexish = q < e && (*q == 'x' || *q == 'X');
    const int e10 = ex - frac;
    if (!hexish && sig <= 19 && mant <= ((uint64_t)1 << 53) && e10 >= -22 && e10 <= 22) {
        double v = (double)mant;
        if (e10 < 0) v /= kPow10[-e10];
        else v *= kPow10[e10];
        *out = neg ? -v : v;
        p = q;
        return true;
    }
    const char *s = tok;
    if (!parse_f64_slow(s, e, out)) return false;
    p = s;
    return true;
}

inline const char *next_line(const char *p, const char *e)
{
    const void *nl = p < e ? memchr(p, '\n', (size_t)(e - p)) : nullptr;
    return nl ? (const char *)nl + 1 : e;
}

struct Piece {
    const char *beg = nullptr, *end = nullptr;
    std::vector<std::vector<Ent>> blk; /* nsdp SDP blocks, then the LP block */
    int stop = 0;                      /* 0 ran to the end, 1 met a non-entry line, 2 index out of range */
    int tiny = 0;
    int64_t kept = 0;
};

struct Shape {
    int64_t m = 0, nblk = 0, nsdp = 0, nlp = 0;
    const int64_t *dims = nullptr;
};

void parse_piece(Piece &pc, const Shape &sh)
{
    const char *p = pc.beg, *end = pc.end;
    const size_t guess = (size_t)(end - p) / 20 + 16;
    if (sh.nsdp + (sh.nlp > 0) == 1) pc.blk[0].reserve(guess);
    while (p < end) {
        const char *line = p;
        while (line < end && is_blank(*line)) ++line;
        if (line >= end) break;
        if (*line == '\n') { p = line + 1; continue; }
        const char *eol = (const char *)memchr(line, '\n', (size_t)(end - line));
        if (!eol) eol = end;
        int64_t con, blk, i, j;
        double v;
        const char *q = line;
        if (!parse_i64(q, eol, &con) || !parse_i64(q, eol, &blk) || !parse_i64(q, eol, &i) || !parse_i64(q, eol, &j) ||
            !parse_f64(q, eol, &v)) {
            pc.stop = 1;
            return;
        }
        p = eol < end ? eol + 1 : end;
        blk -= 1; i -= 1; j -= 1;
        if (con < 0 || con > sh.m || blk < 0 || blk >= sh.nblk) { pc.stop = 2; return; }
        if (std::fabs(v) < 1e-12) { pc.tiny = 1; continue; }
        if (con == 0) v = -v;
        if (sh.nlp > 0 && blk == sh.nsdp) {
            if (i < 0 || i >= sh.nlp) { pc.stop = 2; return; }
            pc.blk[(size_t)sh.nsdp].push_back({con, i, v});
        } else {
            const int64_t n = sh.dims[blk];
            if (i > j) std::swap(i, j);
            if (i < 0 || j >= n) { pc.stop = 2; return; }
            /* lower-triangular entry (row j, col i), column-major packed */
            pc.blk[(size_t)blk].push_back({con, (2 * n - i - 1) * i / 2 + j, v});
        }
        pc.kept++;
    }
}

template <class F> void run_threads(int T, F f)
{
    if (T <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)T - 1);
    for (int t = 1; t < T; ++t) th.emplace_back(f, t);
    f(0);
    for (auto &x : th) x.join();
}

/* stable sort of one column by packed index with T threads: sorted slices, then rounds of pairwise merges */
void sort_long_column(int64_t *idx, double *val, int64_t len, int T)
{
    std::vector<IdxVal> a((size_t)len);
    run_threads(T, [&](int t) {
        const int64_t lo = len * t / T, hi = len * (t + 1) / T;
        for (int64_t k = lo; k < hi; ++k) a[(size_t)k] = {idx[k], val[k]};
        std::stable_sort(a.begin() + lo, a.begin() + hi, [](const IdxVal &x, const IdxVal &y) { return x.idx < y.idx; });
    });
    for (int width = 1; width < T; width *= 2) {
        const int pairs = (T + 2 * width - 1) / (2 * width);
        run_threads(pairs, [&](int q) {
            const int s0 = q * 2 * width, s1 = std::min(T, s0 + width), s2 = std::min(T, s0 + 2 * width);
            if (s1 >= s2) return;
            std::inplace_merge(a.begin() + len * s0 / T, a.begin() + len * s1 / T, a.begin() + len * s2 / T,
                               [](const IdxVal &x, const IdxVal &y) { return x.idx < y.idx; });
        });
    }
    run_threads(T, [&](int t) {
        const int64_t lo = len * t / T, hi = len * (t + 1) / T;
        for (int64_t k = lo; k < hi; ++k) { idx[k] = a[(size_t)k].idx; val[k] = a[(size_t)k].val; }
    });
}

const int64_t kLongColumn = 1 << 16;

/* pieces[0..np) of block b  ->  CSC with ncols columns.  Within a column: file order, then (sort_idx) ascending
 * packed index with ties in file order. */
int build_csc(std::vector<Piece> &pieces, int np, size_t b, int64_t ncols, bool sort_idx, int T, int64_t **beg_out,
              int64_t **idx_out, double **val_out)
{
    int64_t N = 0;
    for (int c = 0; c < np; ++c) N += (int64_t)pieces[(size_t)c].blk[b].size();
    int64_t *beg = (int64_t *)calloc((size_t)ncols + 1, sizeof(int64_t));
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    double *val = (double *)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    *beg_out = beg; *idx_out = idx; *val_out = val;
    if (!beg || !idx || !val) return 1;
    if (N == 0) return 0;
    /* histogram: every thread counts the pieces it takes, one atomic add per RUN of equal columns */
    std::atomic<int> next(0);
    run_threads(std::min(T, np), [&](int) {
        for (int c = next++; c < np; c = next++) {
            const std::vector<Ent> &v = pieces[(size_t)c].blk[b];
            size_t k = 0;
            while (k < v.size()) {
                size_t r = k + 1;
                while (r < v.size() && v[r].col == v[k].col) ++r;
                __atomic_fetch_add(&beg[v[k].col + 1], (int64_t)(r - k), __ATOMIC_RELAXED);
                k = r;
            }
        }
    });
    /* long columns get a per-piece offset table; the others are placed by column range */
    std::vector<int64_t> longs;
    for (int64_t c = 0; c < ncols; ++c)
        if (beg[c + 1] >= kLongColumn) longs.push_back(c);
    for (int64_t c = 0; c < ncols; ++c) beg[c + 1] += beg[c];
    const size_t nl = longs.size();
    std::vector<int64_t> loff((size_t)np * nl, 0); /* entries of long column h in piece c */
    auto long_id = [&](int64_t col) -> int64_t {
        const auto it = std::lower_bound(longs.begin(), longs.end(), col);
        return (it != longs.end() && *it == col) ? (int64_t)(it - longs.begin()) : -1;
    };
    if (nl > 0) {
        next = 0;
        run_threads(std::min(T, np), [&](int) {
            for (int c = next++; c < np; c = next++) {
                const std::vector<Ent> &v = pieces[(size_t)c].blk[b];
                size_t k = 0;
                while (k < v.size()) {
                    size_t r = k + 1;
                    while (r < v.size() && v[r].col == v[k].col) ++r;
                    const int64_t h = (beg[v[k].col + 1] - beg[v[k].col] >= kLongColumn) ? long_id(v[k].col) : -1;
                    if (h >= 0) loff[(size_t)c * nl + (size_t)h] += (int64_t)(r - k);
                    k = r;
                }
            }
        });
        for (size_t h = 0; h < nl; ++h) { /* counts -> offsets inside the column, pieces in file order */
            int64_t run = 0;
            for (int c = 0; c < np; ++c) {
                const int64_t cnt = loff[(size_t)c * nl + h];
                loff[(size_t)c * nl + h] = run;
                run += cnt;
            }
        }
    }
    /* short columns: ranges of columns with about equal numbers of short entries */
    int64_t nshort = N;
    for (int64_t c : longs) nshort -= beg[c + 1] - beg[c];
    std::vector<int64_t> cut((size_t)T + 1, ncols);
    cut[0] = 0;
    {
        int64_t seen = 0;
        int t = 1;
        for (int64_t c = 0; c < ncols && t < T; ++c) {
            const int64_t cnt = beg[c + 1] - beg[c];
            if (cnt < kLongColumn) seen += cnt;
            while (t < T && seen >= nshort * t / T && seen > 0) cut[(size_t)t++] = c + 1;
        }
    }
    run_threads(T, [&](int t) {
        /* (a) the long-column entries of this thread's pieces */
        if (nl > 0)
            for (int c = t; c < np; c += T) {
                const std::vector<Ent> &v = pieces[(size_t)c].blk[b];
                std::vector<int64_t> fill(nl);
                for (size_t h = 0; h < nl; ++h) fill[h] = beg[longs[h]] + loff[(size_t)c * nl + h];
                size_t k = 0;
                while (k < v.size()) {
                    size_t r = k + 1;
                    while (r < v.size() && v[r].col == v[k].col) ++r;
                    const int64_t col = v[k].col;
                    if (beg[col + 1] - beg[col] >= kLongColumn) {
                        int64_t &f = fill[(size_t)long_id(col)];
                        for (size_t e = k; e < r; ++e, ++f) { idx[f] = v[e].idx; val[f] = v[e].val; }
                    }
                    k = r;
                }
            }
        /* (b) every short entry whose column lies in this thread's range, all pieces in file order */
        const int64_t c0 = cut[(size_t)t], c1 = cut[(size_t)t + 1];
        if (c0 >= c1 || nshort == 0) return;
        std::vector<int64_t> fill(beg + c0, beg + c1);
        for (int c = 0; c < np; ++c)
            for (const Ent &en : pieces[(size_t)c].blk[b]) {
                if (en.col < c0 || en.col >= c1) continue;
                if (beg[en.col + 1] - beg[en.col] >= kLongColumn) continue;
                const int64_t f = fill[(size_t)(en.col - c0)]++;
                idx[f] = en.idx;
                val[f] = en.val;
            }
    });
    for (int c = 0; c < np; ++c) std::vector<Ent>().swap(pieces[(size_t)c].blk[b]);
    if (!sort_idx) return 0;
    /* ascending packed index inside every column (stable) */
    run_threads(T, [&](int t) {
        std::vector<IdxVal> tmp;
        for (int64_t c = cut[(size_t)t]; c < cut[(size_t)t + 1]; ++c) {
            const int64_t e0 = beg[c], e1 = beg[c + 1];
            if (e1 - e0 >= kLongColumn || e1 - e0 < 2) continue;
            bool asc = true;
            for (int64_t e = e0 + 1; e < e1; ++e)
                if (idx[e] < idx[e - 1]) { asc = false; break; }
            if (asc) continue;
            tmp.resize((size_t)(e1 - e0));
            for (int64_t e = e0; e < e1; ++e) tmp[(size_t)(e - e0)] = {idx[e], val[e]};
            std::stable_sort(tmp.begin(), tmp.end(), [](const IdxVal &x, const IdxVal &y) { return x.idx < y.idx; });
            for (int64_t e = e0; e < e1; ++e) { idx[e] = tmp[(size_t)(e - e0)].idx; val[e] = tmp[(size_t)(e - e0)].val; }
        }
    });
    for (int64_t c : longs) {
        const int64_t e0 = beg[c], len = beg[c + 1] - beg[c];
        std::atomic<int> unsorted(0);
        run_threads(T, [&](int t) {
            const int64_t lo = std::max<int64_t>(1, len * t / T), hi = len * (t + 1) / T;
            for (int64_t k = lo; k < hi; ++k)
                if (idx[e0 + k] < idx[e0 + k - 1]) { unsorted = 1; break; }
        });
        if (unsorted) sort_long_column(idx + e0, val + e0, len, T);
    }
    return 0;
}

int thread_count(size_t bytes)
{
    if (const char *s = getenv("LORADS_READ_THREADS")) {
        const int v = atoi(s);
        if (v >= 1) return std::min(v, 256);
    }
    if (bytes < ((size_t)4 << 20)) return 1;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    /* the ranks of a multi-GPU run read at the same time and share the cores */
    for (const char *name : {"LORADS_LOCAL_RANKS", "LOCAL_WORLD_SIZE"})
        if (const char *s = getenv(name)) {
            const int p = atoi(s);
            if (p > 1) { n = std::max<long>(1, n / p); break; }
        }
    return (int)std::min<long>(n, 32);
}

/* ---- binary side channel ---------------------------------------------------------------------------------------
 * The parsed problem, array for array (the text of a C5-size instance is ~1.8 GB; .dat-s stays the canonical
 * format, this is a cache of what reading it produced).  Little-endian, 8-byte fields:
 *   magic "LORADSB1" | m nBlks nLpCols nElems | blkDims[nBlks] | b[m] | per block: beg[m+2] idx[N] val[N] | LP: beg[m+2] idx val */
const char kBinaryMagic[8] = {'L', 'O', 'R', 'A', 'D', 'S', 'B', '1'};

template <class T> bool take(const char *&p, const char *end, T *dst, int64_t count)
{
    if (count < 0 || (uint64_t)(end - p) < (uint64_t)count * sizeof(T)) return false;
    /* big arrays: copied by several threads */
    const int T2 = (count * (int64_t)sizeof(T) > ((int64_t)64 << 20)) ? thread_count((size_t)count * sizeof(T)) : 1;
    const char *src = p;
    run_threads(T2, [&](int t) {
        const int64_t lo = count * t / T2, hi = count * (t + 1) / T2;
        memcpy(dst + lo, src + lo * sizeof(T), (size_t)(hi - lo) * sizeof(T));
    });
    p += count * sizeof(T);
    return true;
}

int read_csc(const char *&p, const char *end, int64_t m, int64_t **beg, int64_t **idx, double **val)
{
    *beg = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m + 2));
    if (!*beg || !take(p, end, *beg, m + 2)) return 1;
    const int64_t N = (*beg)[m + 1];
    if (N < 0 || (*beg)[0] != 0) return 1;
    for (int64_t c = 0; c <= m; ++c)
        if ((*beg)[c + 1] < (*beg)[c]) return 1;
    *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    *val = (double *)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    if (!*idx || !*val || !take(p, end, *idx, N) || !take(p, end, *val, N)) return 1;
    return 0;
}

int read_binary(const char *buf, size_t sz, lh_sdpa *out)
{
    const char *p = buf + sizeof(kBinaryMagic), *end = buf + sz;
    int64_t head[4];
    if (!take(p, end, head, 4)) return 1;
    const int64_t m = head[0], nb = head[1], nlp = head[2];
    if (m <= 0 || nb < 0 || nlp < 0 || nb > (int64_t)1 << 40) return 1;
    out->m = m; out->nBlks = nb; out->nLpCols = nlp; out->nElems = head[3];
    const size_t nbs = (size_t)(nb > 0 ? nb : 1);
    out->blkDims = (int64_t *)malloc(sizeof(int64_t) * nbs);
    out->b = (double *)malloc(sizeof(double) * (size_t)m);
    out->matBeg = (int64_t **)calloc(nbs, sizeof(int64_t *));
    out->matIdx = (int64_t **)calloc(nbs, sizeof(int64_t *));
    out->matElem = (double **)calloc(nbs, sizeof(double *));
    if (!out->blkDims || !out->b || !out->matBeg || !out->matIdx || !out->matElem) return 1;
    if (!take(p, end, out->blkDims, nb) || !take(p, end, out->b, m)) return 1;
    for (int64_t k = 0; k < nb; ++k) {
        if (out->blkDims[k] <= 0) return 1;
        if (read_csc(p, end, m, &out->matBeg[k], &out->matIdx[k], &out->matElem[k])) return 1;
        const int64_t tri = out->blkDims[k] * (out->blkDims[k] + 1) / 2, N = out->matBeg[k][m + 1];
        for (int64_t e = 0; e < N; ++e)
            if (out->matIdx[k][e] < 0 || out->matIdx[k][e] >= tri) return 1;
    }
    if (nlp > 0) {
        if (read_csc(p, end, m, &out->lpBeg, &out->lpIdx, &out->lpElem)) return 1;
        for (int64_t e = 0; e < out->lpBeg[m + 1]; ++e)
            if (out->lpIdx[e] < 0 || out->lpIdx[e] >= nlp) return 1;
    }
    return p == end ? 0 : 1;
}

} // namespace

extern "C" int lh_write_sdpa_binary(const char *fname, const lh_sdpa *d)
{
    FILE *f = fopen(fname, "wb");
    if (!f) return 1;
    bool ok = fwrite(kBinaryMagic, 1, sizeof(kBinaryMagic), f) == sizeof(kBinaryMagic);
    auto put = [&](const void *src, size_t bytes) { if (ok && bytes > 0) ok = fwrite(src, 1, bytes, f) == bytes; };
    const int64_t head[4] = {d->m, d->nBlks, d->nLpCols, d->nElems};
    put(head, sizeof(head));
    put(d->blkDims, sizeof(int64_t) * (size_t)d->nBlks);
    put(d->b, sizeof(double) * (size_t)d->m);
    for (int64_t k = 0; k < d->nBlks; ++k) {
        const int64_t N = d->matBeg[k][d->m + 1];
        put(d->matBeg[k], sizeof(int64_t) * (size_t)(d->m + 2));
        put(d->matIdx[k], sizeof(int64_t) * (size_t)N);
        put(d->matElem[k], sizeof(double) * (size_t)N);
    }
    if (d->nLpCols > 0) {
        const int64_t N = d->lpBeg[d->m + 1];
        put(d->lpBeg, sizeof(int64_t) * (size_t)(d->m + 2));
        put(d->lpIdx, sizeof(int64_t) * (size_t)N);
        put(d->lpElem, sizeof(double) * (size_t)N);
    }
    if (fclose(f) != 0) ok = false;
    return ok ? 0 : 1;
}

extern "C" int lh_read_sdpa(const char *fname, lh_sdpa *out, int quiet)
{
    memset(out, 0, sizeof(*out));
    const double t_open = lh_time();
    const int fd = open(fname, O_RDONLY);
    if (fd < 0) return 1;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return 1; }
    const size_t sz = (size_t)st.st_size;
    const char *buf = (const char *)mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (buf == MAP_FAILED) return 1;
    madvise((void *)buf, sz, MADV_WILLNEED);
    if (sz >= sizeof(kBinaryMagic) && memcmp(buf, kBinaryMagic, sizeof(kBinaryMagic)) == 0) {
        const int brc = read_binary(buf, sz, out);
        munmap((void *)buf, sz);
        if (brc) lh_free_sdpa(out);
        return brc;
    }
    const char *p = buf, *end = buf + sz;
    int rc = 1;
    Shape sh;
    std::vector<int64_t> dims;
    std::vector<Piece> pieces;
    do {
        /* comments */
        while (p < end && (*p == '*' || *p == '"')) p = next_line(p, end);
        int64_t m = 0, nblk = 0;
        { const char *q = p; while (q < end && (is_blank(*q) || *q == '\n')) ++q; if (!parse_i64(q, end, &m) || m <= 0) break; p = next_line(q, end); }
        { const char *q = p; while (q < end && (is_blank(*q) || *q == '\n')) ++q; if (!parse_i64(q, end, &nblk) || nblk <= 0) break; p = next_line(q, end); }
        /* block dimensions: numbers separated by anything that is not part of a number */
        dims.assign((size_t)nblk, 0);
        {
            int64_t got = 0;
            while (got < nblk && p < end) {
                while (p < end && !((*p >= '0' && *p <= '9') || *p == '-' || *p == '+')) ++p;
                if (p >= end) break;
                int64_t v;
                const char *q = p;
                if (!parse_i64(q, end, &v)) { ++p; continue; }
                dims[(size_t)got++] = v;
                p = q;
            }
            if (got != nblk) break;
            p = next_line(p, end);
        }
        bool bad = false;
        for (int64_t k = 0; k < nblk; ++k)
            if (dims[(size_t)k] <= 0 && k != nblk - 1) bad = true; /* only the last block may be diagonal */
        if (bad) break;
        int64_t nlp = 0, nsdp = nblk;
        if (dims[(size_t)nblk - 1] < 0) { nlp = -dims[(size_t)nblk - 1]; nsdp = nblk - 1; }
        out->m = m;
        out->nBlks = nsdp;
        out->nLpCols = nlp;
        out->blkDims = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nsdp > 0 ? nsdp : 1));
        if (!out->blkDims) break;
        for (int64_t k = 0; k < nsdp; ++k) out->blkDims[k] = dims[(size_t)k];
        /* right-hand side */
        out->b = (double *)calloc((size_t)m, sizeof(double));
        if (!out->b) break;
        {
            int64_t got = 0;
            while (got < m && p < end) {
                while (p < end && !((*p >= '0' && *p <= '9') || *p == '-' || *p == '+' || *p == '.')) ++p;
                if (p >= end) break;
                double v;
                const char *q = p;
                if (!parse_f64(q, end, &v) || q == p) { ++p; continue; }
                out->b[got++] = v;
                p = q;
            }
            if (got != m) break;
            p = next_line(p, end);
        }
        sh.m = m; sh.nblk = nblk; sh.nsdp = nsdp; sh.nlp = nlp; sh.dims = dims.data();
        /* entries: one piece per thread, cut at line starts */
        const int T = thread_count((size_t)(end - p));
        pieces.resize((size_t)T);
        {
            const char *cur = p;
            for (int t = 0; t < T; ++t) {
                Piece &pc = pieces[(size_t)t];
                pc.beg = cur;
                const char *want = p + (size_t)(end - p) / (size_t)T * (size_t)(t + 1);
                pc.end = (t == T - 1 || want >= end) ? end : next_line(std::max(want, cur), end);
                cur = pc.end;
                pc.blk.resize((size_t)nsdp + 1);
            }
        }
        const bool timing = getenv("LORADS_READ_TIMING") != nullptr;
        const double t_hdr = lh_time();
        run_threads(T, [&](int t) { parse_piece(pieces[(size_t)t], sh); });
        const double t_parse = lh_time();
        /* the section ends at the first piece that met a non-entry line; an out-of-range index before that fails */
        int np = T;
        for (int t = 0; t < T; ++t)
            if (pieces[(size_t)t].stop) { np = t + 1; break; }
        if (pieces[(size_t)np - 1].stop == 2) break;
        int tiny = 0;
        for (int t = 0; t < np; ++t) { tiny |= pieces[(size_t)t].tiny; out->nElems += pieces[(size_t)t].kept; }
        if (tiny && !quiet) printf("[Warning] Entry smaller than 1e-12 is ignored. \n");
        const size_t nb = (size_t)(nsdp > 0 ? nsdp : 1);
        out->matBeg = (int64_t **)calloc(nb, sizeof(int64_t *));
        out->matIdx = (int64_t **)calloc(nb, sizeof(int64_t *));
        out->matElem = (double **)calloc(nb, sizeof(double *));
        if (!out->matBeg || !out->matIdx || !out->matElem) break;
        bool fail = false;
        for (int64_t k = 0; k < nsdp && !fail; ++k)
            fail = build_csc(pieces, np, (size_t)k, m + 1, true, T, &out->matBeg[k], &out->matIdx[k], &out->matElem[k]) != 0;
        if (!fail && nlp > 0) fail = build_csc(pieces, np, (size_t)nsdp, m + 1, false, T, &out->lpBeg, &out->lpIdx, &out->lpElem) != 0;
        if (fail) break;
        if (timing)
            fprintf(stderr, "lorads_b200: reader threads %d  header+b %.3f s  entries %.3f s  csc %.3f s\n", T, t_hdr - t_open,
                    t_parse - t_hdr, lh_time() - t_parse);
        rc = 0;
    } while (0);
    munmap((void *)buf, sz);
    if (rc) lh_free_sdpa(out);
    return rc;
}

extern "C" void lh_free_sdpa(lh_sdpa *d)
{
    if (d->matBeg)
        for (int64_t k = 0; k < d->nBlks; ++k) {
            free(d->matBeg[k]);
            if (d->matIdx) free(d->matIdx[k]);
            if (d->matElem) free(d->matElem[k]);
        }
    free(d->matBeg);
    free(d->matIdx);
    free(d->matElem);
    free(d->lpBeg);
    free(d->lpIdx);
    free(d->lpElem);
    free(d->blkDims);
    free(d->b);
    memset(d, 0, sizeof(*d));
}
