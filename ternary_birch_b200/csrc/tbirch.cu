// tbirch.cu -- libtbirch.so: the C ABI of include/tbirch.h over the sm_100a kernels.
//
// Host responsibilities (everything else runs on the device):
//   * genus table import + device hash table build            Genus.h:248-303, HashMap.h:107-147
//   * per-representative conic base points, which consume the reference's
//     sequential std::mt19937 / std::uniform_int_distribution stream and so
//     define the order of the p+1 lines of every representative
//                                                             Fp.h:10-25,97-168; QuadForm.h:847-935
//   * error translation to the reference's exception texts    NeighborManager.h:350-352, HashMap.h:77
// There is no CPU fallback: without a CUDA device every compute call fails.
#include "../../include/tbirch.h"
#include "tb_kernels.cuh"
#include "tb_build.cuh"
#include "tb_linalg.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>
#include <condition_variable>
#include <mutex>
#include <thread>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

using namespace tb;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches(0);

static int fail(tb_status st, const std::string &msg)
{
    g_last_error = msg;
    return (int)st;
}

#define TB_CUDA(expr)                                                                            \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(TB_ERR_RUNTIME, std::string("CUDA error: ") + cudaGetErrorString(_e) +   \
                                            " at " #expr);                                       \
    } while (0)

static const char *MSG_OVERFLOW =
    "An overflow has occurred. The p-neighbor's discriminant does not match the original.";
static const char *MSG_NOT_FOUND = "Key not found.";
static const char *MSG_BAD_PRIME = "Prime must not divide the discriminant.";
static const char *MSG_NO_MORE = "No more isometries.";

// ---------------------------------------------------------------------------
// Host Fp with the reference's RNG (Fp.h:10-25).  Canonical residues; the
// reference's lazy [0,kp) representation never changes a decision.
// ---------------------------------------------------------------------------
struct HostFp {
    uint64_t p;
    std::mt19937 rng;
    std::uniform_int_distribution<> distr;

    HostFp(uint64_t prime, uint64_t seed) : p(prime), rng(seed), distr(0, (int)(prime - 1)) {}

    uint64_t draw() { return (uint64_t)distr(rng); }
    uint64_t mul(uint64_t a, uint64_t b) const { return (a * b) % p; }   // a, b < p < 2^31
    uint64_t add(uint64_t a, uint64_t b) const { uint64_t s = a + b; return s >= p ? s - p : s; }
    uint64_t sub(uint64_t a, uint64_t b) const { return a >= b ? a - b : a + p - b; }
    uint64_t residue(int64_t x) const
    {
        int64_t v = x % (int64_t)p;
        return (uint64_t)(v < 0 ? v + (int64_t)p : v);
    }
    uint64_t power(uint64_t a, uint64_t e) const
    {
        uint64_t r = 1 % p;
        for (; e; e >>= 1, a = mul(a, a))
            if (e & 1) r = mul(r, a);
        return r;
    }
    int legendre(uint64_t a) const
    {
        a %= p;
        if (!a) return 0;
        return power(a, (p - 1) / 2) == 1 ? 1 : -1;
    }
    uint64_t inverse(uint64_t a) const { return a % p ? power(a % p, p - 2) : 0; }

    // Tonelli-Shanks as the reference runs it (Fp.h:97-145): the non-residue is
    // drawn from the shared RNG, so the draws are part of the stream.
    uint64_t square_root(uint64_t a)
    {
        a %= p;
        if (a <= 1) return a;
        if (legendre(a) != 1) return 0;
        uint64_t odd = p - 1;
        int twos = 0;
        while (!(odd & 1)) { odd >>= 1; ++twos; }
        if (twos == 1) return power(a, (p + 1) / 4);
        uint64_t z = draw();
        while (z == 0 || legendre(z) == 1) z = draw();
        int m = twos;
        uint64_t c = power(z, odd), r = power(a, (odd + 1) / 2), t = power(a, odd);
        while (t != 1)
        {
            int i = 0;
            for (uint64_t t1 = t; t1 != 1; t1 = mul(t1, t1)) ++i;
            uint64_t b = power(c, 1ull << (m - i - 1));
            r = mul(r, b);
            c = mul(b, b);
            t = mul(t, c);
            m = i;
        }
        return r;
    }
};

// One representative's base point (QuadFormFp::isotropic_vector, QuadForm.h:847-907):
// pick a line x = r z' .. through random (r, s) until the conic meets it in a
// rational point.  p = 2 packs the three isotropic vectors instead (915-935).
static void base_point(HostFp &F, const int64_t *q, uint32_t out[3])
{
    const uint64_t p = F.p;
    const uint64_t a = F.residue(q[0]), b = F.residue(q[1]), c = F.residue(q[2]);
    const uint64_t f = F.residue(q[3]), g = F.residue(q[4]), h = F.residue(q[5]);
    if (p == 2)
    {
        // nonzero vectors 1..7 as bits (x y z); keep the first three isotropic ones
        uint32_t found[3] = {0, 0, 0};
        int cnt = 0;
        const uint64_t val[8] = {1, c, b, b ^ f ^ c, a, a ^ g ^ c, a ^ h ^ b, a ^ b ^ c ^ f ^ g ^ h};
        for (uint32_t v = 1; v < 8; ++v)
            if (val[v] == 0 && cnt < 3) found[cnt++] = v;
        out[0] = found[0]; out[1] = found[1]; out[2] = found[2];
        return;
    }
    for (;;)
    {
        uint64_t r = 0, br = 0, alpha = 0;
        while (alpha == 0)
        {
            r = F.draw();
            br = F.mul(b, r);
            alpha = F.add(F.mul(F.add(br, h), r), a);                       // Q(1, r, 0)
        }
        const uint64_t s = F.draw();
        const uint64_t beta = F.add(F.add(F.mul(F.add(F.add(br, br), h), s), F.mul(f, r)), g);
        const uint64_t gamma = F.add(F.mul(F.add(F.mul(b, s), f), s), c);   // Q(0, s, 1)
        uint64_t four_ag = F.mul(gamma, alpha);
        four_ag = F.add(four_ag, four_ag);
        four_ag = F.add(four_ag, four_ag);
        const uint64_t disc = F.sub(F.mul(beta, beta), four_ag);
        if (F.legendre(disc) < 0) continue;
        const uint64_t root = F.sub(F.square_root(disc), beta);
        const uint64_t x = F.mul(root, F.inverse(F.add(alpha, alpha)));
        out[0] = (uint32_t)x;
        out[1] = (uint32_t)F.add(F.mul(r, x), s);
        out[2] = 1;
        return;
    }
}

// ---------------------------------------------------------------------------
// device buffer helper
// ---------------------------------------------------------------------------
// (both helpers release in their destructor, so every early error return of an entry point
//  gives its device memory back; explicit release() calls remain where the order matters)
template <class T>
struct DevBuf {
    T *ptr = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = std::max(n, (size_t)16);
        cudaError_t e = cudaMalloc((void **)&ptr, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

// Stream-ordered allocation from the device's default memory pool (kept warm:
// the release threshold is raised in tb_genus_create), for per-call results.
template <class T>
struct PoolBuf {
    T *ptr = nullptr;
    size_t cap = 0;
    cudaStream_t stream = 0;
    PoolBuf() = default;
    PoolBuf(const PoolBuf &) = delete;
    PoolBuf &operator=(const PoolBuf &) = delete;
    ~PoolBuf() { release(); }
    cudaError_t reserve(size_t n, cudaStream_t st)
    {
        if (n <= cap) return cudaSuccess;
        release();
        stream = st;
        size_t want = std::max(n, (size_t)16);
        cudaError_t e = cudaMallocAsync((void **)&ptr, want * sizeof(T), st);
        if (e == cudaSuccess) cap = want; else ptr = nullptr;
        return e;
    }
    void release()
    {
        if (ptr) cudaFreeAsync(ptr, stream);
        ptr = nullptr;
        cap = 0;
    }
};

struct PrimeState {
    uint64_t p = 0;
    PrimeCtx pc;
    u32 nonres = 0;               // smallest quadratic non-residue mod p (device Tonelli-Shanks)
    bool have_exact = false;
    std::vector<uint32_t> base;   // [N][3] RNG-exact base points of all representatives (lazy)
};

struct Copier;
static void copier_destroy(Copier *c);

// ---------------------------------------------------------------------------
// Which read-back format delivers a large T_p to host memory faster ON THIS RANK: the plain
// int32 DMA (8 bytes per entry over PCIe, no host work) or the 3-byte device format widened by
// host threads (tb_readback.cuh)?  It depends on the host, on how many ranks share it and on
// which PCIe path this GPU sits behind (profiles/r2a_d2h_probe_n4.md: with four GPUs copying at
// once two get 44 GB/s and two 21 GB/s, and host stores compete with the DMA for one memory
// system), so it is measured, not assumed: every large host read-back reports the int32-CSR
// bytes it delivered and the time it was busy, the policy keeps a running rate per format,
// tries each format first and re-tries the losing one now and then.  One policy per device
// (ranks are processes; a process driving several GPUs keeps them apart).
// ---------------------------------------------------------------------------
struct ReadbackPolicy {
    enum { PLAIN = 0, NARROW = 1 };
    std::mutex mx;
    double rate[2] = {0.0, 0.0};      // running delivered GB/s of int32 CSR per format
    int samples[2] = {0, 0};
    int current = NARROW;
    int since_probe = 0;              // read-backs measured since the other format was last tried
    int pending_probe = 0;            // a probe of the other format has been handed out and not measured yet

    int choose()
    {
        std::lock_guard<std::mutex> lk(mx);
        if (samples[NARROW] == 0) return NARROW;          // one measured read-back of each format first
        if (samples[PLAIN] == 0) return PLAIN;
        if (rate[1 - current] > 1.15 * rate[current]) current = 1 - current;   // hysteresis: ranks adapt concurrently
        if (since_probe >= 32 && !pending_probe)          // the host's load changes: look at the other format again
        {
            pending_probe = 1;
            return 1 - current;
        }
        return current;
    }
    void record(int format, double csr_bytes, double seconds)
    {
        if (csr_bytes < 64e6 || seconds <= 0) return;      // small copies measure latency, not the path
        std::lock_guard<std::mutex> lk(mx);
        const double r = csr_bytes / seconds / 1e9;
        rate[format] = samples[format] ? 0.5 * rate[format] + 0.5 * r : r;
        ++samples[format];
        if (format == current) ++since_probe;
        else { since_probe = 0; pending_probe = 0; }
    }
};

static ReadbackPolicy &readback_policy()
{
    static ReadbackPolicy per_device[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return per_device[dev & 63];
}

struct tb_genus {
    int N = 0, K = 0;
    uint64_t seed = 0;
    int64_t disc = 0;
    std::vector<int64_t> primes, conductors, dims, q;
    std::vector<int32_t> lut;

    double lambda_min = 1.0;    // lower bound of the smallest Gram eigenvalue over all representatives
    double log2_coef = 0, log2_iso = 0, log2_scalar = 0;   // log2 of the largest |q|, |s| / |sinv|, scalar entries
    double log2_mid = 0;          // log2 of the largest middle diagonal coefficient b (a <= b <= c on reduced forms)
    double log2_isounit = 99;     // |entry of a reduced neighbor isometry| <= p * 2^log2_isounit (table loop below)
    double log2_mother = 0;                                // log2 of the largest |coefficient| of representative 0
    double mother_lambda = 0;                              // lower bound of the mother form's smallest Gram eigenvalue
    cudaStream_t stream = 0;
    DevBuf<i64> d_cur, d_rep, d_disc;
    DevBuf<int> d_slots, d_lut, d_lutT;
    DevBuf<unsigned char> d_incl;
    GenusCtx gc;

    // per-prime state (valid for cur_p)
    PrimeState ps;
    DevBuf<u32> d_invlut, d_base, d_W;
    DevBuf<RepPrime> d_rp;
    int rp_begin = -1, rp_rows = 0;
    bool rp_exact = false;
    DevBuf<u64> d_err;
    DevBuf<int> d_tab;          // rowbase | rowfirst | rowcount   (row kernels)
    DevBuf<i64> d_nnz;          // nnz | nnzbase
    DevBuf<u64> d_lookback;     // [2^K][rows] look-back words + ticket (fused row kernel)
    DevBuf<int> d_maxmult;      // [2] largest |entry| bound of the last T_p (narrow read-back); rows_acc_kernel's overflow flag
    DevBuf<u32> d_rowscratch;   // row-kernel state in global memory when it does not fit shared memory
    int *h_maxmult = nullptr;   // pinned [2]
    struct Copier *copier = nullptr;   // background thread of the narrow read-back (started on first use)
    bool allow_narrow = true;     // TBIRCH_NO_NARROW_COPY turns the compressed read-back off
    bool force_narrow = false;    // TBIRCH_FORCE_NARROW_COPY: always the 3-byte format where it applies (no measuring)
    int64_t narrow_min_nnz = 1 << 18;   // smaller matrices are copied as they are (TBIRCH_NARROW_MIN_NNZ)
    int64_t narrow_max_ratio = 32;      // narrow row output is attempted while (p+1)/N stays below this (TBIRCH_NARROW_MAX_RATIO)
    bool allow_fused = true;
    bool allow_acc = true;        // TBIRCH_ROWS_SORT: always the sorting row kernel (rows_kernel), never rows_acc_kernel
    bool force_acc = false;       // TBIRCH_ROWS_ACC: rows_acc_kernel wherever its counters fit, whatever the row density
    bool force_checked = false;   // TBIRCH_FORCE_CHECKED: always run the checked C128 kernel in precise mode
    int last_int_mode = -1;       // tb_last_int_mode
    bool force_wide = false;      // TBIRCH_FORCE_WIDE: never run precise launches in the bounded 64-bit mode
    bool no_fp64_reduce = false;  // TBIRCH_NO_FP64_REDUCE: every reduction through the integer loop (parity tests of the fallback)
    size_t device_bytes = 0;
    // per-shard row bookkeeping of the last Hecke call (reused while the shard is unchanged)
    int shard_begin = -1, shard_end = -1;
    std::vector<int> shard_rows, shard_first;
    char *h_stage[2] = {nullptr, nullptr};   // pinned staging for copies into pageable host memory
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    u64 *h_err = nullptr;       // pinned [2]
    i64 *h_nnz = nullptr;       // pinned [2 * 2^K]
    int *h_tab = nullptr;       // pinned [3 * 2^K]
    char *h_table = nullptr;    // pinned staging of the genus table upload
    size_t h_table_bytes = 0;
    uint32_t *h_base = nullptr; // pinned staging
    size_t h_base_cap = 0;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    float last_ms_nbr = 0.f, last_ms_rows = 0.f;
    int64_t last_neighbors = 0;
    bool timing_valid = false;
};

struct tb_hecke {
    tb_genus *g = nullptr;
    int C = 0;
    int rep_begin = 0, rep_end = 0;
    int maxmult = 0x7fffffff;       // bound on |entry| (largest number of neighbors of a row in one class)
    std::vector<int> rows, rowfirst, rowbase;
    std::vector<int64_t> nnz, nnzbase;
    PoolBuf<int> d_indptr, d_data, d_indices;
    bool wide_valid = true;            // d_data / d_indices hold the result (else only the narrow twins do,
    std::vector<char> wide_k;          //   and conductor k's int32 arrays exist once wide_k[k] is set)
    int64_t extent = 0;                // entries spanned by the per-conductor regions
    PoolBuf<unsigned short> d_idx16;   // narrow twins of d_indices / d_data written by the row kernel:
    PoolBuf<signed char> d_dat8;       //   3 bytes per entry for the read-back over PCIe
    // asynchronous host read-backs still in flight (tb_hecke_copy_wait)
    std::mutex copy_mx;
    std::condition_variable copy_cv;
    int copy_pending = 0;
    cudaEvent_t copy_ev = nullptr;     // last plain cudaMemcpyAsync of an asynchronous copy
    bool copy_ev_armed = false;
    // timing of the plain asynchronous host read-back of this T_p (ReadbackPolicy)
    cudaEvent_t plain_ev0 = nullptr, plain_ev1 = nullptr;
    double plain_bytes = 0;
    std::string copy_error;            // failure of a background read-back of THIS result
    ~tb_hecke()
    {
        if (copy_ev) cudaEventDestroy(copy_ev);
        if (plain_ev0) cudaEventDestroy(plain_ev0);
        if (plain_ev1) cudaEventDestroy(plain_ev1);
    }
};

struct tb_isoseq {
    tb_genus *g = nullptr;
    uint64_t p = 0;
    int precise = 0;
    int64_t next_rep = 0;          // next representative to compute
    std::vector<int64_t> chunk;    // records of the current chunk
    size_t pos = 0;                // next record within chunk
    int64_t delivered = 0, total = 0;
};

// ---------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------
extern "C" int tb_abi_version(void) { return TBIRCH_ABI_VERSION; }
extern "C" const char *tb_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t tb_launch_count(void) { return g_launches.load(); }

static int require_device()
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(TB_ERR_RUNTIME, "libtbirch: no CUDA device available (there is no CPU fallback).");
    }
    return TB_OK;
}

static int upload_table(tb_genus *g, const tb_genus_desc *d)
{
    const int N = g->N, K = g->K;
    int lg = 4;
    while ((1ll << lg) < N) ++lg;
    const u32 slots = 1u << (lg + 1);
    // one pinned staging arena: cur | rep | disc | slots | lut
    const size_t rec_bytes = (size_t)N * TB_REC_WORDS * sizeof(i64);
    const size_t disc_bytes = (size_t)N * sizeof(i64);
    const size_t slot_bytes = (size_t)slots * sizeof(int);
    const size_t lut_bytes = ((size_t)N << K) * sizeof(int);
    const size_t incl_bytes = (size_t)(((1 << K) + 7) / 8) * N;
    const size_t total = 2 * rec_bytes + disc_bytes + slot_bytes + 2 * lut_bytes + incl_bytes;
    if (total > g->h_table_bytes)
    {
        if (g->h_table) cudaFreeHost(g->h_table);
        g->h_table = nullptr;
        g->h_table_bytes = 0;
        TB_CUDA(cudaMallocHost((void **)&g->h_table, total));
        g->h_table_bytes = total;
    }
    i64 *cur = (i64 *)g->h_table;
    i64 *rep = (i64 *)(g->h_table + rec_bytes);
    i64 *disc = (i64 *)(g->h_table + 2 * rec_bytes);
    int *table = (int *)(g->h_table + 2 * rec_bytes + disc_bytes);
    int *lut = (int *)(g->h_table + 2 * rec_bytes + disc_bytes + slot_bytes);
    for (int n = 0; n < N; ++n)
    {
        i64 *c = cur + (size_t)n * TB_REC_WORDS, *r = rep + (size_t)n * TB_REC_WORDS;
        for (int i = 0; i < 6; ++i) c[i] = r[i] = d->q[6 * n + i];
        c[6] = r[6] = d->scalar[n];
        for (int i = 0; i < 9; ++i) { c[7 + i] = d->s[9 * n + i]; r[7 + i] = d->sinv[9 * n + i]; }
        // wrap-around discriminant of the representative (NeighborManager.h:16)
        const u64 a = (u64)c[0], b = (u64)c[1], cc = (u64)c[2], f = (u64)c[3], gg = (u64)c[4], h = (u64)c[5];
        disc[n] = (i64)(a * (4ull * b * cc - f * f) - b * gg * gg + h * (f * gg - cc * h));
    }
    // lambda_min(Gram) >= det(Gram) / trace(Gram)^2 for a positive definite 3x3 Gram matrix
    // (det = lambda1 lambda2 lambda3 <= lambda_min * trace^2); det(Gram) = disc / 4
    double lam = 1e300;
    bool definite = true;
    for (int n = 0; n < N; ++n)
    {
        const i64 *c = cur + (size_t)n * TB_REC_WORDS;
        const double a = (double)c[0], b = (double)c[1], cc = (double)c[2], f = (double)c[3], gg = (double)c[4], h = (double)c[5];
        const double det4 = a * (4.0 * b * cc - f * f) - b * gg * gg + h * (f * gg - cc * h);
        const double tr = a + b + cc;
        if (!(a > 0 && b > 0 && cc > 0 && det4 > 0 && 4.0 * a * b - h * h > 0)) { definite = false; break; }
        lam = std::min(lam, 0.25 * det4 / (tr * tr));
        if (n == 0) g->mother_lambda = 0.5 * 0.25 * det4 / (tr * tr);
    }
    g->lambda_min = definite ? lam * 0.5 : 0.0;   // 0 disables the FP64 reduce loop
    {
        double mq = 1, ms = 1, md = 1, mb = 1;
        for (int n = 0; n < N; ++n)
        {
            const i64 *c = cur + (size_t)n * TB_REC_WORDS, *r = rep + (size_t)n * TB_REC_WORDS;
            for (int i = 0; i < 6; ++i) mq = std::max(mq, fabs((double)c[i]));
            mb = std::max(mb, std::max(fabs((double)c[0]), fabs((double)c[1])));
            md = std::max(md, fabs((double)c[6]));
            for (int i = 7; i < 16; ++i) ms = std::max(ms, std::max(fabs((double)c[i]), fabs((double)r[i])));
        }
        g->log2_coef = log2(mq); g->log2_iso = log2(ms); g->log2_scalar = log2(md); g->log2_mid = log2(mb);
        // Reduced neighbor isometry of representative n: column j is a vector v with Q_n(v) = p^2 d_j, d_j a diagonal
        // coefficient of the representative that is hit.  With G_n the Gram matrix (v^T G_n v = 2 Q_n(v)) the largest
        // v_i^2 on that ellipsoid is 2 p^2 d_j (G_n^-1)_ii, so |entry| <= p sqrt(2 D R), D = largest diagonal
        // coefficient in the table, R = largest diagonal entry of any G_n^-1 (cofactor / determinant).
        double R = 0, D = 1;
        for (int n = 0; n < N && definite; ++n)
        {
            const i64 *c = cur + (size_t)n * TB_REC_WORDS;
            const double a = (double)c[0], b = (double)c[1], cc = (double)c[2], f = (double)c[3], gg = (double)c[4], h = (double)c[5];
            const double det = 2.0 * (4.0 * a * b * cc + f * gg * h - a * f * f - b * gg * gg - cc * h * h);
            R = std::max(R, std::max((4.0 * b * cc - f * f) / det, std::max((4.0 * a * cc - gg * gg) / det, (4.0 * a * b - h * h) / det)));
            D = std::max(D, std::max(a, std::max(b, cc)));
        }
        g->log2_isounit = (definite && R > 0) ? 0.5 * log2(2.0 * D * R) + 0.01 : 99.0;
        double mm = 1;
        for (int i = 0; i < 6; ++i) mm = std::max(mm, fabs((double)cur[i]));
        g->log2_mother = log2(mm);
    }
    // open-addressing table of representatives: 2 * 2^ceil(log2 N) slots (>= 32), linear probing,
    // keyed by the reference's FNV hash of the six coefficients (HashMap.h:10-24, QuadForm.cpp:779-803).
    for (u32 i = 0; i < slots; ++i) table[i] = -1;
    for (int n = 0; n < N; ++n)
    {
        u64 fnv = 0x811c9dc5ull;
        for (int i = 0; i < 6; ++i) fnv = (fnv ^ (u64)d->q[6 * n + i]) * 0x01000193ull;
        u32 s = (u32)fnv & (slots - 1);
        while (table[s] != -1) s = (s + 1) & (slots - 1);
        table[s] = n;
    }
    memcpy(lut, d->lut, lut_bytes);
    int *lutT = lut + ((size_t)N << K);
    for (int k = 0; k < (1 << K); ++k)
        for (int n = 0; n < N; ++n) lutT[((size_t)n << K) + k] = d->lut[(size_t)k * N + n];
    // which conductors see each representative, 8 per byte (rows_acc_kernel's count sweep)
    unsigned char *incl = (unsigned char *)(lutT + ((size_t)N << K));
    memset(incl, 0, incl_bytes);
    for (int k = 0; k < (1 << K); ++k)
        for (int n = 0; n < N; ++n)
            if (d->lut[(size_t)k * N + n] != -1) incl[(size_t)(k >> 3) * N + n] |= (unsigned char)(1u << (k & 7));
    TB_CUDA(g->d_incl.reserve(incl_bytes));
    TB_CUDA(cudaMemcpyAsync(g->d_incl.ptr, incl, incl_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(g->d_cur.reserve((size_t)N * TB_REC_WORDS));
    TB_CUDA(g->d_rep.reserve((size_t)N * TB_REC_WORDS));
    TB_CUDA(g->d_disc.reserve(N));
    TB_CUDA(g->d_slots.reserve(slots));
    TB_CUDA(g->d_lut.reserve((size_t)N << K));
    TB_CUDA(g->d_lutT.reserve((size_t)N << K));
    TB_CUDA(cudaMemcpyAsync(g->d_lutT.ptr, lutT, lut_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemcpyAsync(g->d_cur.ptr, cur, rec_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemcpyAsync(g->d_rep.ptr, rep, rec_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemcpyAsync(g->d_disc.ptr, disc, disc_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemcpyAsync(g->d_slots.ptr, table, slot_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemcpyAsync(g->d_lut.ptr, lut, lut_bytes, cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaStreamSynchronize(g->stream));

    GenusCtx &gc = g->gc;
    memset(&gc, 0, sizeof(gc));
    gc.N = N; gc.K = K; gc.slot_mask = slots - 1;
    gc.cur = g->d_cur.ptr; gc.rep = g->d_rep.ptr; gc.disc = g->d_disc.ptr; gc.slots = g->d_slots.ptr;
    for (int i = 0; i < K; ++i)
    {
        const u64 dd = (u64)d->primes[i];
        gc.spin[i] = make_invdiv(dd);
        gc.spin_qmax[i] = ~0ull / dd;
        u64 inv = 0;
        if (dd & 1ull)
        {
            inv = dd;                                   // Newton: 3 correct bits, doubling each step
            for (int it = 0; it < 6; ++it) inv *= 2ull - dd * inv;
        }
        gc.spin_inv[i] = inv;
    }
    return TB_OK;
}

extern "C" int tb_genus_create(const tb_genus_desc *d, tb_genus **out)
{
    if (!d || !out) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_create: null argument");
    *out = nullptr;
    if (int rc = require_device()) return rc;
    if (d->K > 63) return fail(TB_ERR_DOMAIN, "Must have 63 or fewer prime divisors.");
    if (d->N <= 0 || d->K < 0 || d->K > TB_MAX_PRIMES)
        return fail(TB_ERR_RUNTIME, "tb_genus_create: unsupported genus shape (need 1 <= N, K <= 16)");
    int bitsN = 0;
    while ((1ll << bitsN) < d->N) ++bitsN;
    if (bitsN + d->K > 31)
        return fail(TB_ERR_RUNTIME, "tb_genus_create: N * 2^K does not fit the packed 32-bit neighbor word");
    for (int64_t i = 0; i < d->K; ++i)
        if (d->primes[i] < 2) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_create: prime divisor < 2");

    std::unique_ptr<tb_genus> g(new tb_genus);
    g->N = (int)d->N; g->K = (int)d->K; g->seed = d->seed; g->disc = d->disc;
    const int C = 1 << g->K;
    g->primes.assign(d->primes, d->primes + g->K);
    g->conductors.assign(d->conductors, d->conductors + C);
    g->dims.assign(d->dims, d->dims + C);
    g->q.assign(d->q, d->q + 6 * (size_t)g->N);
    g->lut.assign(d->lut, d->lut + ((size_t)g->N << g->K));
    if (int rc = upload_table(g.get(), d)) return rc;
    g->allow_fused = getenv("TBIRCH_TWO_PASS_ROWS") == nullptr;   // debugging / A-B switches
    g->allow_acc = getenv("TBIRCH_ROWS_SORT") == nullptr;
    g->force_checked = getenv("TBIRCH_FORCE_CHECKED") != nullptr;
    g->force_wide = getenv("TBIRCH_FORCE_WIDE") != nullptr;
    g->no_fp64_reduce = getenv("TBIRCH_NO_FP64_REDUCE") != nullptr;
    g->force_acc = getenv("TBIRCH_ROWS_ACC") != nullptr;
    g->allow_narrow = getenv("TBIRCH_NO_NARROW_COPY") == nullptr;   // otherwise ReadbackPolicy decides per T_p,
    g->force_narrow = getenv("TBIRCH_FORCE_NARROW_COPY") != nullptr; // unless this pins the 3-byte format
    if (const char *e = getenv("TBIRCH_NARROW_MIN_NNZ")) g->narrow_min_nnz = std::max(1ll, atoll(e));
    if (const char *e = getenv("TBIRCH_NARROW_MAX_RATIO")) g->narrow_max_ratio = std::max(1ll, atoll(e));
    TB_CUDA(g->d_err.reserve(2));
    TB_CUDA(g->d_tab.reserve(3 * (size_t)C));
    TB_CUDA(g->d_nnz.reserve(2 * (size_t)C));
    TB_CUDA(cudaMallocHost((void **)&g->h_err, 2 * sizeof(u64)));
    TB_CUDA(cudaMallocHost((void **)&g->h_nnz, 2 * (size_t)C * sizeof(i64)));
    TB_CUDA(g->d_maxmult.reserve(2));
    TB_CUDA(cudaMallocHost((void **)&g->h_maxmult, 2 * sizeof(int)));
    TB_CUDA(cudaMallocHost((void **)&g->h_tab, 3 * (size_t)C * sizeof(int)));
    {
        int dev = 0;
        cudaMemPool_t pool;
        TB_CUDA(cudaGetDevice(&dev));
        TB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        TB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    TB_CUDA(cudaEventCreate(&g->ev0));
    TB_CUDA(cudaEventCreate(&g->ev1));
    TB_CUDA(cudaEventCreate(&g->ev2));
    *out = g.release();
    return TB_OK;
}

extern "C" int tb_genus_upload(tb_genus *g, const tb_genus_desc *d)
{
    if (!g || !d) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_upload: null argument");
    if (d->N != g->N || d->K != g->K) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_upload: shape mismatch");
    g->q.assign(d->q, d->q + 6 * (size_t)g->N);
    g->lut.assign(d->lut, d->lut + ((size_t)g->N << g->K));
    g->seed = d->seed;
    g->ps.p = 0;
    g->rp_begin = -1;
    g->shard_begin = -1;
    return upload_table(g, d);
}

extern "C" void tb_genus_destroy(tb_genus *g)
{
    if (!g) return;
    g->d_cur.release(); g->d_rep.release(); g->d_disc.release(); g->d_slots.release(); g->d_lut.release(); g->d_lutT.release();
    g->d_incl.release();
    g->d_invlut.release(); g->d_base.release(); g->d_W.release(); g->d_rp.release(); g->d_err.release();
    g->d_tab.release(); g->d_nnz.release(); g->d_lookback.release();
    for (int i = 0; i < 2; ++i) { if (g->h_stage[i]) cudaFreeHost(g->h_stage[i]); if (g->ev_stage[i]) cudaEventDestroy(g->ev_stage[i]); }
    if (g->h_err) cudaFreeHost(g->h_err);
    if (g->h_nnz) cudaFreeHost(g->h_nnz);
    if (g->h_maxmult) cudaFreeHost(g->h_maxmult);
    g->d_maxmult.release();
    g->d_rowscratch.release();
    copier_destroy(g->copier);
    g->copier = nullptr;
    if (g->h_tab) cudaFreeHost(g->h_tab);
    if (g->h_table) cudaFreeHost(g->h_table);
    if (g->h_base) cudaFreeHost(g->h_base);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->ev2) cudaEventDestroy(g->ev2);
    delete g;
}

extern "C" int tb_genus_set_stream(tb_genus *g, void *s)
{
    if (!g) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_set_stream: null genus");
    g->stream = (cudaStream_t)s;
    return TB_OK;
}
extern "C" int64_t tb_genus_size(const tb_genus *g) { return g ? g->N : 0; }
extern "C" int64_t tb_genus_num_conductors(const tb_genus *g) { return g ? (1ll << g->K) : 0; }
extern "C" int64_t tb_genus_conductor(const tb_genus *g, int64_t k) { return g->conductors[k]; }
extern "C" int64_t tb_genus_dim(const tb_genus *g, int64_t k) { return g->dims[k]; }

extern "C" int tb_last_kernel_time(tb_genus *g, float *ms_nbr, float *ms_rows, int64_t *neighbors)
{
    if (!g || !g->timing_valid) return fail(TB_ERR_RUNTIME, "tb_last_kernel_time: no timed launch yet");
    if (ms_nbr) *ms_nbr = g->last_ms_nbr;
    if (ms_rows) *ms_rows = g->last_ms_rows;
    if (neighbors) *neighbors = g->last_neighbors;
    return TB_OK;
}

extern "C" int tb_last_int_mode(const tb_genus *g) { return g ? g->last_int_mode : -1; }

// ---------------------------------------------------------------------------
// per-prime setup
// ---------------------------------------------------------------------------
static bool divides_disc(const tb_genus *g, uint64_t p)
{
    // Genus.h:334,343: disc % p == 0
    return p != 0 && (g->disc % (int64_t)p) == 0;
}

static int setup_prime(tb_genus *g, uint64_t p)
{
    if (p < 2 || p >= (1ull << 31))
        return fail(TB_ERR_INVALID_ARGUMENT, "libtbirch: prime out of range (2 <= p < 2^31)");
    if (g->ps.p == p) return TB_OK;
    PrimeState &ps = g->ps;
    ps.p = 0;
    g->rp_begin = -1;
    PrimeCtx &pc = ps.pc;
    memset(&pc, 0, sizeof(pc));
    pc.p = (u32)p;
    pc.pp = p * p;
    pc.dp = make_invdiv(p);
    pc.dpp = make_invdiv(p * p);
    pc.use_lut = (p > 2 && p < (1ull << 22)) ? 1u : 0u;
    pc.m32 = (p > 2 && p < (1ull << 16)) ? (u32)((1ull << 32) / p) : 0u;
    // 64-bit trace in the two-limb modes: |composed entry| <= denom * sqrt(c0 / lambda_min(mother))
    {
        const double lden = log2((double)p) + 2 * g->log2_scalar;
        pc.narrow_trace = (g->mother_lambda > 0 && lden + 0.5 * (g->log2_mother - log2(g->mother_lambda)) + 2 < 61) ? 1u : 0u;
    }
    // FP64 reduce loop: isometry entries are bounded by p * sqrt(M / lambda_min) with M < 2^50 the
    // loop's entry guard on the coefficients (tb_neighbor.cuh); require that bound below 2^52.
    pc.fp64_reduce = ((double)p * sqrt(1125899906842624.0 / g->lambda_min) < 4503599627370496.0) ? 1u : 0u;
    if (g->no_fp64_reduce) pc.fp64_reduce = 0u;       // TBIRCH_NO_FP64_REDUCE: the literal integer loop (eisenstein_reduce)
    {
        // What build_neighbor hands to the reduction, bounded from the table and p (the bounds the bounded
        // 64-bit precise mode already rests on, precise_bounds_ok64): form coefficients below 4 A (p + 4)^2,
        // isometry entries below 2 p^2.  Below 2^50 the FP64 loop needs no per-lane magnitude test.
        const double lp4 = log2((double)p + 4.0);
        pc.proven_small = (pc.fp64_reduce && g->lambda_min > 0 && g->log2_coef + 2 * lp4 + 2 < 50 && 2 * lp4 + 1 < 50) ? 1u : 0u;
        // reduced neighbor isometries: |entry| <= p * 2^log2_isounit (tb_genus_create);
        // with the table's own isometries below 2^31 as well, the trace of cur.s * s * rep.sinv needs 32-bit factors only
        const double iso = log2((double)p) + g->log2_isounit;
        pc.iso32 = (pc.proven_small && iso < 31 && g->log2_iso < 31) ? 1u : 0u;
        // binary32 tail of the reduce loop: a finished lane holds a reduced form (|f| <= b, |g|, |h| <= a <= b <= c)
        // and is carried through the remaining trips with zero quotients, so its coefficients must be
        // binary32 integers (c < 2^24) and its running sum a + b + f + g + h (<= 5 b) must stay exact (b < 2^20)
        pc.f32_tail = (pc.fp64_reduce && g->log2_coef < 24 && g->log2_mid < 20) ? 1u : 0u;
        if (getenv("TBIRCH_NO_F32_TAIL")) pc.f32_tail = 0u;                         // A/B switch: binary64 to the end
        // binary64 neighbor setup (fast_setup, tb_neighbor.cuh): bit 0 = exact (p < 2^16 with the inverse table,
        // proven_small's B p^2 < 2^48); bit 1 = the reference's int64 arithmetic provably does not wrap, so int64
        // mode may use it too: the largest intermediate of NeighborManager.h:207-341 on a z = 1, unrotated line is
        // aa + s2 (gg + cc s2) <= B p^2 + p^2 (B (3p/2 + 1) + Bmid p^2) with |s2| < p^2 and cc = a <= Bmid
        pc.pd = (double)p; pc.ppd = (double)(p * p);
        pc.inv_pd = 1.0 / pc.pd; pc.inv_ppd = 1.0 / pc.ppd;
        // p^-1 mod 2^64 (Newton) for recover_row3; it also needs 2c r3 and G31 r1 + G32 r2 + G33 r3 exact in
        // binary64 resp. int64: |r| < 2^31 (iso32) and table coefficients below 2^20
        if ((p & 1) && g->log2_coef < 20) { u64 x = p; for (int it = 0; it < 6; ++it) x *= 2 - p * x; pc.pinv64 = x; }
        if (pc.m32 && pc.use_lut && pc.proven_small && pc.fp64_reduce)
        {
            const unsigned __int128 B = (unsigned __int128)ceil(exp2(g->log2_coef)) + 1, Bm = (unsigned __int128)ceil(exp2(g->log2_mid)) + 1;
            const unsigned __int128 P = p, PP = P * P;
            const unsigned __int128 worst = B * PP + PP * (B * (3 * P / 2 + 2) + Bm * PP);
            pc.fast_setup = 1u | (worst < ((unsigned __int128)1 << 63) ? 2u : 0u);
        }
        if (getenv("TBIRCH_NO_FAST_SETUP")) pc.fast_setup = 0u;                      // A/B switch: integer setup on every lane
        if (getenv("TBIRCH_NO_ISO32")) pc.iso32 = 0u;                                // A/B switch: 64-bit trace products
        if (getenv("TBIRCH_NO_PROVEN_BOUNDS")) pc.proven_small = pc.iso32 = 0u;   // A/B switch: the per-lane tests and 64-bit products
    }
    if (pc.use_lut)
    {
        TB_CUDA(g->d_invlut.reserve(p));
        pc.inv_lut = g->d_invlut.ptr;
        inverse_lut_kernel<<<(unsigned)((p + 255) / 256), 256, 0, g->stream>>>(pc, g->d_invlut.ptr);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
    }
    ps.have_exact = false;
    ps.nonres = 0;
    if (p > 2)
    {
        HostFp F(p, 1);
        u32 z = 2;
        while (F.legendre(z) != -1) ++z;
        ps.nonres = z;
    }
    ps.p = p;
    return TB_OK;
}

// RNG-exact base points of ALL representatives, in order, from one stream: the reference's
// shared Fp across NeighborManager constructions.  They define the ORDER of the p+1 lines of
// each representative, which only isometry_sequence / neighbor records expose.
static void ensure_exact_base(tb_genus *g)
{
    PrimeState &ps = g->ps;
    if (ps.have_exact) return;
    ps.base.resize(3 * (size_t)g->N);
    HostFp F(ps.p, g->seed);
    for (int n = 0; n < g->N; ++n) base_point(F, &g->q[6 * (size_t)n], &ps.base[3 * (size_t)n]);
    ps.have_exact = true;
}

// base points for an explicit list of representatives visited in ascending
// order by ONE Fp (Genus::_eigenvectors builds NeighborManagers only for the
// set-cover representatives, Genus.h:457-462)
static void base_points_for(tb_genus *g, uint64_t p, const std::vector<int> &reps, std::vector<uint32_t> &out)
{
    out.resize(3 * reps.size());
    HostFp F(p, g->seed);
    for (size_t i = 0; i < reps.size(); ++i) base_point(F, &g->q[6 * (size_t)reps[i]], &out[3 * i]);
}

static int stage_rows(tb_genus *g, const uint32_t *base, int rep_begin, int rows)
{
    const size_t words = 3 * (size_t)rows;
    if (words > g->h_base_cap)
    {
        if (g->h_base) cudaFreeHost(g->h_base);
        g->h_base = nullptr;
        g->h_base_cap = 0;
        TB_CUDA(cudaMallocHost((void **)&g->h_base, words * sizeof(uint32_t)));
        g->h_base_cap = words;
    }
    memcpy(g->h_base, base, words * sizeof(uint32_t));
    TB_CUDA(g->d_base.reserve(words));
    TB_CUDA(g->d_rp.reserve(rows));
    TB_CUDA(cudaMemcpyAsync(g->d_base.ptr, g->h_base, words * sizeof(uint32_t), cudaMemcpyHostToDevice, g->stream));
    prep_reps_kernel<<<(rows + 127) / 128, 128, 0, g->stream>>>(g->gc, g->ps.pc, g->d_base.ptr, rep_begin, rows, g->d_rp.ptr);
    ++g_launches;
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

// exact = the reference's RNG-defined base points (host); otherwise any rational point of the
// conic, found on the device (Hecke matrices and eigenvalues do not depend on the choice).
static int prepare_rows(tb_genus *g, int rep_begin, int rows, bool exact)
{
    if (g->rp_begin == rep_begin && g->rp_rows == rows && g->rp_exact == exact) return TB_OK;
    g->rp_begin = -1;
    if (exact)
    {
        ensure_exact_base(g);
        if (int rc = stage_rows(g, &g->ps.base[3 * (size_t)rep_begin], rep_begin, rows)) return rc;
    }
    else
    {
        TB_CUDA(g->d_base.reserve(3 * (size_t)rows));
        TB_CUDA(g->d_rp.reserve(rows));
        base_points_kernel<<<(rows + 127) / 128, 128, 0, g->stream>>>(g->gc, g->ps.pc, g->ps.nonres, rep_begin, rows, g->d_base.ptr);
        prep_reps_kernel<<<(rows + 127) / 128, 128, 0, g->stream>>>(g->gc, g->ps.pc, g->d_base.ptr, rep_begin, rows, g->d_rp.ptr);
        g_launches += 2;
        TB_CUDA(cudaGetLastError());
    }
    g->rp_begin = rep_begin;
    g->rp_rows = rows;
    g->rp_exact = exact;
    return TB_OK;
}

// May the precise mode drop its overflow checks for this prime?  Conservative magnitude
// bounds (log2) on everything the 128-bit pipeline forms, from the genus table and p:
//   build_neighbor   every intermediate is below 8 A p^4          (A = max |coefficient|)
//   reduce           FP64 loop enabled => all loop values < 2^53; isometry entries < 2^52
//   reduced isometry columns v satisfy Q_rep(v) = p^2 * (coefficient of a genus representative),
//                    so |entry| <= p sqrt(A / lambda_min)
//   compose          |cur.s * s * rep.sinv| entries < 9 S^2 p sqrt(A / lambda_min)   (S = max |s|, |sinv|)
//   Spinor::norm     composed case: mother-form minors (<= 4 A0^2) times composed entries / denominators;
//                    r == n case: minors of a representative (<= 4 A^2) times neighbor-isometry entries
//   denominator      p * D^2                                       (D = max scalar)
// When all stay far below 2^127 the unchecked U128 arithmetic is exact, i.e. equal to GMP's.
static bool precise_bounds_ok(const tb_genus *g)
{
    const PrimeCtx &pc = g->ps.pc;
    if (!pc.fp64_reduce) return false;
    const double lp = log2((double)pc.p);
    const double build = 3 + g->log2_coef + 4 * lp;
    const double iso = std::min(1 + lp + 0.5 * (g->log2_coef - log2(g->lambda_min)), lp + g->log2_isounit);
    const double comp = 4 + 2 * g->log2_iso + iso;
    const double denom = lp + 2 * g->log2_scalar;
    const double spin_comp = 6 + 2 * g->log2_mother + std::max(comp, denom);
    const double spin_self = 6 + 2 * g->log2_coef + std::max(iso, lp);
    return build < 120 && comp < 120 && denom < 120 && spin_comp < 120 && spin_self < 120;
}

// May a precise launch run in 64-bit machine arithmetic (X64)?  On top of the U128 conditions:
//   build_neighbor   ring operations are exact modulo 2^64, so only the values that feed a
//                    division, a comparison or the output must truly fit: all are below
//                    4 A (p + 4)^2 once the two large numerators go through muladd_div and
//                    s2 is taken from aa mod p^2 (needs p^3 < 2^63)
//   reduce           FP64 loop, or the integer loop whose values stay below 8x its input
//   trace / compose  composed isometry and denominator below 2^61 (narrow_trace)
//   Spinor::norm     first case in 64 bits, the others widen to U128 on the device
static bool precise_bounds_ok64(const tb_genus *g)
{
    const PrimeCtx &pc = g->ps.pc;
    if (!precise_bounds_ok(g) || !pc.narrow_trace || g->force_wide) return false;
    const double lp = log2((double)pc.p + 4.0);
    const double iso = std::min(1 + lp + 0.5 * (g->log2_coef - log2(g->lambda_min)), lp + g->log2_isounit);
    return lp < 20 && g->log2_coef + 2 * lp < 56 && iso < 60;
}

template <int MODE>
static int launch_neighbors(tb_genus *g, int precise, int rep_begin, int rows, u32 *W, i64 *records,
                            const int *rep_list = nullptr)
{
    NbrLaunch L;
    memset(&L, 0, sizeof(L));
    L.rep_begin = rep_begin;
    L.rows = rows;
    L.P1 = g->ps.pc.p + 1;
    L.total = (u64)rows * L.P1;
    L.dP1 = make_invdiv(L.P1);
    L.mP1 = L.total < (1ull << 32) ? (u64)(((unsigned __int128)1 << 64) / L.P1) + 1 : 0;
    L.rp = g->d_rp.ptr;
    L.W = W;
    L.records = records;
    L.err = g->d_err.ptr;
    L.rep_list = rep_list;
    const u64 blocks = (L.total + TB_NBR_THREADS - 1) / TB_NBR_THREADS;
    if (blocks > 0x7fffffffull) return fail(TB_ERR_RUNTIME, "libtbirch: launch too large");
    TB_CUDA(cudaMemsetAsync(g->d_err.ptr, 0xff, 2 * sizeof(u64), g->stream));
    g->last_int_mode = !precise ? 0 : g->force_checked ? 3 : precise_bounds_ok64(g) ? 1 : precise_bounds_ok(g) ? 2 : 3;
    if (g->last_int_mode == 1)
        neighbors_kernel<X64, MODE><<<(unsigned)blocks, TB_NBR_THREADS, 0, g->stream>>>(g->gc, g->ps.pc, L);
    else if (g->last_int_mode == 2)
        neighbors_kernel<U128, MODE><<<(unsigned)blocks, TB_NBR_THREADS, 0, g->stream>>>(g->gc, g->ps.pc, L);
    else if (precise)
        neighbors_kernel<C128, MODE><<<(unsigned)blocks, TB_NBR_THREADS, 0, g->stream>>>(g->gc, g->ps.pc, L);
    else
        neighbors_kernel<W64, MODE><<<(unsigned)blocks, TB_NBR_THREADS, 0, g->stream>>>(g->gc, g->ps.pc, L);
    ++g_launches;
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

// The reference walks (n, t) in order and throws at the first failure; report
// whichever failure has the smaller neighbor index.
static int check_errors(tb_genus *g)
{
    const u64 ov = g->h_err[0], nf = g->h_err[1];
    if (ov == ~0ull && nf == ~0ull) return TB_OK;
    if (ov <= nf) return fail(TB_ERR_OVERFLOW, MSG_OVERFLOW);
    return fail(TB_ERR_INVALID_ARGUMENT, MSG_NOT_FOUND);
}

// ---------------------------------------------------------------------------
// Hecke matrices
// ---------------------------------------------------------------------------
static size_t rows_smem_bytes(int N, u32 P1)
{
    const size_t mw = (N + 31) / 32;
    const size_t maxcols = std::min<size_t>(P1, N);
    size_t b = (2 * mw + (maxcols + 1) + maxcols + 256) * 4 + (size_t)P1 * 2;
    return (b + 15) & ~(size_t)15;
}

static size_t rows_acc_smem_bytes(int N, u32 P1)
{
    const size_t mw = (N + 31) / 32;
    const size_t maxcols = std::min<size_t>(P1, N);
    size_t b = (2 * mw + 2 * maxcols + 8 * 32 + 40) * 4 + maxcols;
    return (b + 15) & ~(size_t)15;
}

extern "C" int tb_hecke_compute_rows(tb_genus *g, uint64_t p, int precise, int64_t rep_begin, int64_t rep_end,
                                     tb_hecke **out)
{
    if (!g || !out) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_compute: null argument");
    *out = nullptr;
    if (divides_disc(g, p)) return fail(TB_ERR_INVALID_ARGUMENT, MSG_BAD_PRIME);
    if (rep_begin < 0 || rep_end > g->N || rep_begin > rep_end)
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_compute_rows: bad representative range");
    if (int rc = setup_prime(g, p)) return rc;

    const int N = g->N, K = g->K, C = 1 << K;
    const int rb = (int)rep_begin, rows = (int)(rep_end - rep_begin);
    const u32 P1 = (u32)p + 1;

    std::unique_ptr<tb_hecke> h(new tb_hecke);
    h->g = g; h->C = C; h->rep_begin = rb; h->rep_end = (int)rep_end;
    h->rows.assign(C, 0); h->rowfirst.assign(C, 0); h->rowbase.assign(C, 0);
    h->nnz.assign(C, 0); h->nnzbase.assign(C, 0);
    if (g->shard_begin != rb || g->shard_end != (int)rep_end)
    {
        g->shard_rows.assign(C, 0);
        g->shard_first.assign(C, 0);
        for (int k = 0; k < C; ++k)
        {
            const int32_t *lut = &g->lut[(size_t)k * N];
            int cnt = 0, first = -1;
            for (int n = rb; n < (int)rep_end; ++n)
                if (lut[n] != -1) { if (first < 0) first = lut[n]; ++cnt; }
            g->shard_rows[k] = cnt;
            g->shard_first[k] = first < 0 ? 0 : first;
        }
        g->shard_begin = rb;
        g->shard_end = (int)rep_end;
    }
    int total_ptr = 0;
    for (int k = 0; k < C; ++k)
    {
        h->rows[k] = g->shard_rows[k];
        h->rowfirst[k] = g->shard_first[k];
        h->rowbase[k] = total_ptr;
        total_ptr += h->rows[k] + 1;
    }
    TB_CUDA(h->d_indptr.reserve(total_ptr, g->stream));
    TB_CUDA(cudaMemsetAsync(h->d_indptr.ptr, 0, (size_t)total_ptr * sizeof(int), g->stream));
    if (rows == 0)
    {
        TB_CUDA(cudaStreamSynchronize(g->stream));
        *out = h.release();
        return TB_OK;
    }

    if (int rc = prepare_rows(g, rb, rows, false)) return rc;
    TB_CUDA(g->d_W.reserve((size_t)rows * P1));

    // small device-side tables for the row kernels
    DevBuf<int> &d_tab = g->d_tab;
    DevBuf<i64> &d_nnz = g->d_nnz;
    int *tab = g->h_tab;
    for (int k = 0; k < C; ++k) { tab[k] = h->rowbase[k]; tab[C + k] = h->rowfirst[k]; tab[2 * C + k] = h->rows[k]; }
    TB_CUDA(cudaMemcpyAsync(d_tab.ptr, tab, 3 * (size_t)C * sizeof(int), cudaMemcpyHostToDevice, g->stream));

    // Single-pass row kernel when the per-conductor output regions, sized by the upper bound
    // rows_k * min(p+1, dim_k), fit comfortably in HBM; else count / scan / fill.
    int64_t cap_total = 0;
    std::vector<int64_t> cap(C);
    for (int k = 0; k < C; ++k)
    {
        cap[k] = (int64_t)h->rows[k] * std::min<int64_t>(P1, g->dims[k]);
        cap_total += cap[k];
    }
    if (g->device_bytes == 0)
    {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);                // slow call: once per genus
        g->device_bytes = total_b;
    }
    const size_t total_b = g->device_bytes;
    const bool fused = g->allow_fused && (size_t)cap_total * 8 <= std::min<size_t>((size_t)48 << 30, total_b / 3);
    // The row kernel also writes uint16 / int8 twins of indices / data when a host read-back could use
    // them: columns fit 16 bits, the matrices are large, and ~(p+1)/N neighbors per class keeps entries in 8 bits.
    int64_t max_dim = 0;
    for (int k = 0; k < C; ++k) max_dim = std::max<int64_t>(max_dim, g->dims[k]);
    const bool narrow_possible = g->allow_narrow && max_dim <= 65536 && cap_total >= g->narrow_min_nnz &&
                                 (P1 <= 127 || (int64_t)P1 <= g->narrow_max_ratio * (int64_t)N) &&
                                 (g->force_narrow || readback_policy().choose() == ReadbackPolicy::NARROW);

    RowLaunch R;
    memset(&R, 0, sizeof(R));
    R.N = N; R.K = K; R.rep_begin = rb; R.rows = rows; R.P1 = P1; R.mask_words = (N + 31) / 32;
    R.W = g->d_W.ptr; R.lut = g->d_lut.ptr; R.lutT = g->d_lutT.ptr;
    R.rowbase = d_tab.ptr; R.rowfirst = d_tab.ptr + C;
    R.indptr_all = h->d_indptr.ptr;
    R.nnzbase = d_nnz.ptr + C;
    R.nnz_out = d_nnz.ptr;
    R.maxmult = g->d_maxmult.ptr;
    R.overflow = g->d_maxmult.ptr + 1;
    R.incl = g->d_incl.ptr;
    TB_CUDA(cudaMemsetAsync(g->d_maxmult.ptr, 0, 2 * sizeof(int), g->stream));
    size_t smem = rows_smem_bytes(N, P1);
    unsigned row_grid = (unsigned)rows;
    if (smem > 218 * 1024 || getenv("TBIRCH_ROWS_SCRATCH"))
    {
        // the row's sort state does not fit one SM (a genus of > 20000 classes at p > 30000): keep it in a
        // global-memory slot per CTA and let two CTAs per SM walk the rows
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        row_grid = (unsigned)std::min<int>(rows, 2 * sms);
        R.scratch_words = (smem + 3) / 4;
        TB_CUDA(g->d_rowscratch.reserve(R.scratch_words * row_grid));
        R.scratch = g->d_rowscratch.ptr;
        smem = 0;
    }
    const int threads = (int)std::min<u32>(512, std::max<u32>(64, (P1 + 31) / 32 * 32));
    // rows_acc_kernel (accumulate with shared-memory atomics, no sort) when its state fits one SM and a class is
    // not expected to be hit anywhere near 256 times; it reports the rare overflow and the sorting kernel redoes the pass
    const size_t smem_acc = rows_acc_smem_bytes(N, P1);
    // ... and when the row is dense enough: the byte-counter kernel sweeps all N / 32 mask words twice per chunk of 8
    // conductors, the sorting kernel only touches the p + 1 words, so below about N / 4 neighbors per row the sort wins
    // (C2, N = 2016, 32 conductors, p = 101: 140 us against 315; C4, N = 10316, p = 9973: 2.96 ms against 2.25)
    bool use_acc = g->allow_acc && !R.scratch && smem_acc <= 216 * 1024 && (u64)P1 <= 48ull * (u64)N &&
                   (g->force_acc || 4ull * (u64)P1 >= (u64)N);
    // function attributes are per device (context): set them once on each device this process uses
    static std::atomic<uint64_t> attrs_mask(0);
    int attr_dev = 0;
    TB_CUDA(cudaGetDevice(&attr_dev));
    if (!(attrs_mask.load() & (1ull << (attr_dev & 63))))
    {
        const int optin = 218 * 1024;   // 227 KB per CTA minus the kernels' static shared memory (8 KB parity table)
        TB_CUDA(cudaFuncSetAttribute(rows_acc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        TB_CUDA(cudaFuncSetAttribute(rows_acc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TB_CUDA(cudaFuncSetAttribute(rows_acc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        TB_CUDA(cudaFuncSetAttribute(rows_acc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TB_CUDA(cudaFuncSetAttribute(rows_kernel<ROWS_COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        TB_CUDA(cudaFuncSetAttribute(rows_kernel<ROWS_FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        TB_CUDA(cudaFuncSetAttribute(rows_kernel<ROWS_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        TB_CUDA(cudaFuncSetAttribute(rows_kernel<ROWS_COUNT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TB_CUDA(cudaFuncSetAttribute(rows_kernel<ROWS_FILL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TB_CUDA(cudaFuncSetAttribute(rows_kernel<ROWS_FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attrs_mask.fetch_or(1ull << (attr_dev & 63));
    }

    TB_CUDA(cudaEventRecord(g->ev0, g->stream));
    if (int rc = launch_neighbors<MODE_PACKED>(g, precise, rb, rows, g->d_W.ptr, nullptr)) { return rc; }
    TB_CUDA(cudaEventRecord(g->ev1, g->stream));

    if (fused)
    {
        int64_t off = 0;
        for (int k = 0; k < C; ++k) { h->nnzbase[k] = off; g->h_nnz[C + k] = off; off += (cap[k] + 3) & ~(int64_t)3; }
        h->extent = off;
        // when a narrow result is likely (columns fit 16 bits, ~(p+1)/N neighbors per class), write only
        // that: 3 bytes per entry instead of 8; int32 arrays are made on demand (ensure_wide)
        if (narrow_possible)
        {
            TB_CUDA(h->d_idx16.reserve((size_t)std::max<int64_t>(off, 1), g->stream));
            TB_CUDA(h->d_dat8.reserve((size_t)std::max<int64_t>(off, 1), g->stream));
            R.idx16 = h->d_idx16.ptr; R.dat8 = h->d_dat8.ptr;
            h->wide_valid = false;
        }
        else
        {
            TB_CUDA(h->d_data.reserve((size_t)std::max<int64_t>(off, 1), g->stream));
            TB_CUDA(h->d_indices.reserve((size_t)std::max<int64_t>(off, 1), g->stream));
            R.data = h->d_data.ptr; R.indices = h->d_indices.ptr;
        }
        TB_CUDA(g->d_lookback.reserve((size_t)C * rows + 2));
        TB_CUDA(cudaMemsetAsync(g->d_lookback.ptr, 0, ((size_t)C * rows + 2) * sizeof(u64), g->stream));
        TB_CUDA(cudaMemcpyAsync(d_nnz.ptr + C, g->h_nnz + C, C * sizeof(i64), cudaMemcpyHostToDevice, g->stream));
        R.lookback = g->d_lookback.ptr;
        R.ticket = (int *)(g->d_lookback.ptr + (size_t)C * rows);
        R.extent = std::max<int64_t>(off, 1);
        for (;;)
        {
            if (use_acc && R.idx16) rows_acc_kernel<true><<<(unsigned)rows, threads, smem_acc, g->stream>>>(R);
            else if (use_acc) rows_acc_kernel<false><<<(unsigned)rows, threads, smem_acc, g->stream>>>(R);
            else if (R.scratch) rows_kernel<ROWS_FUSED, true><<<row_grid, threads, 0, g->stream>>>(R);
            else rows_kernel<ROWS_FUSED><<<row_grid, threads, smem, g->stream>>>(R);
            ++g_launches;
            TB_CUDA(cudaGetLastError());
            TB_CUDA(cudaEventRecord(g->ev2, g->stream));
            TB_CUDA(cudaMemcpyAsync(g->h_err, g->d_err.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, g->stream));
            TB_CUDA(cudaMemcpyAsync(g->h_nnz, d_nnz.ptr, C * sizeof(i64), cudaMemcpyDeviceToHost, g->stream));
            TB_CUDA(cudaMemcpyAsync(g->h_maxmult, g->d_maxmult.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
            TB_CUDA(cudaStreamSynchronize(g->stream));
            if (int rc = check_errors(g)) { return rc; }
            const bool redo_sort = use_acc && g->h_maxmult[1] != 0;            // a class was hit 256 times or more
            const bool redo_wide = !h->wide_valid && (redo_sort || g->h_maxmult[0] > 127);
            if (!redo_sort && !redo_wide) break;
            // an entry does not fit 8 bits after all (few classes, large p): redo the row pass in int32,
            // with the sorting kernel when the byte counters themselves overflowed
            if (redo_sort) use_acc = false;
            if (redo_wide)
            {
                h->d_idx16.release(); h->d_dat8.release();
                TB_CUDA(h->d_data.reserve((size_t)std::max<int64_t>(off, 1), g->stream));
                TB_CUDA(h->d_indices.reserve((size_t)std::max<int64_t>(off, 1), g->stream));
                R.idx16 = nullptr; R.dat8 = nullptr;
                R.data = h->d_data.ptr; R.indices = h->d_indices.ptr;
                h->wide_valid = true;
            }
            TB_CUDA(cudaMemsetAsync(g->d_lookback.ptr, 0, ((size_t)C * rows + 2) * sizeof(u64), g->stream));
            TB_CUDA(cudaMemsetAsync(g->d_maxmult.ptr, 0, 2 * sizeof(int), g->stream));
        }
        for (int k = 0; k < C; ++k) h->nnz[k] = g->h_nnz[k];
        h->maxmult = g->h_maxmult[0];
    }
    else
    {
        if (R.scratch) rows_kernel<ROWS_COUNT, true><<<row_grid, threads, 0, g->stream>>>(R);
        else rows_kernel<ROWS_COUNT><<<row_grid, threads, smem, g->stream>>>(R);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
        indptr_scan_kernel<<<C, 256, 0, g->stream>>>(h->d_indptr.ptr, d_tab.ptr, d_tab.ptr + 2 * C, d_nnz.ptr);
        ++g_launches;
        TB_CUDA(cudaGetLastError());

        TB_CUDA(cudaMemcpyAsync(g->h_err, g->d_err.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, g->stream));
        TB_CUDA(cudaMemcpyAsync(g->h_nnz, d_nnz.ptr, C * sizeof(i64), cudaMemcpyDeviceToHost, g->stream));
        TB_CUDA(cudaMemcpyAsync(g->h_maxmult, g->d_maxmult.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
        TB_CUDA(cudaStreamSynchronize(g->stream));
        if (int rc = check_errors(g)) { return rc; }
        for (int k = 0; k < C; ++k) h->nnz[k] = g->h_nnz[k];
        h->maxmult = g->h_maxmult[0];

        int64_t total_nnz = 0;
        for (int k = 0; k < C; ++k) { h->nnzbase[k] = total_nnz; total_nnz += h->nnz[k]; }
        h->extent = total_nnz;
        if (narrow_possible && h->maxmult <= 127)            // the count pass already knows the largest entry
        {
            TB_CUDA(h->d_idx16.reserve((size_t)std::max<int64_t>(total_nnz, 1), g->stream));
            TB_CUDA(h->d_dat8.reserve((size_t)std::max<int64_t>(total_nnz, 1), g->stream));
            R.idx16 = h->d_idx16.ptr; R.dat8 = h->d_dat8.ptr;
            h->wide_valid = false;
        }
        else
        {
            TB_CUDA(h->d_data.reserve((size_t)std::max<int64_t>(total_nnz, 1), g->stream));
            TB_CUDA(h->d_indices.reserve((size_t)std::max<int64_t>(total_nnz, 1), g->stream));
            R.data = h->d_data.ptr; R.indices = h->d_indices.ptr;
        }
        for (int k = 0; k < C; ++k) g->h_nnz[C + k] = h->nnzbase[k];
        R.extent = std::max<int64_t>(total_nnz, 1);
        TB_CUDA(cudaMemcpyAsync(d_nnz.ptr + C, g->h_nnz + C, C * sizeof(i64), cudaMemcpyHostToDevice, g->stream));
        if (R.scratch) rows_kernel<ROWS_FILL, true><<<row_grid, threads, 0, g->stream>>>(R);
        else rows_kernel<ROWS_FILL><<<row_grid, threads, smem, g->stream>>>(R);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
        TB_CUDA(cudaEventRecord(g->ev2, g->stream));
        TB_CUDA(cudaStreamSynchronize(g->stream));
    }
    cudaEventElapsedTime(&g->last_ms_nbr, g->ev0, g->ev1);
    cudaEventElapsedTime(&g->last_ms_rows, g->ev1, g->ev2);
    g->last_neighbors = (int64_t)rows * P1;
    g->timing_valid = true;
    *out = h.release();
    return TB_OK;
}

extern "C" int tb_hecke_compute(tb_genus *g, uint64_t p, int precise, tb_hecke **out)
{
    if (!g) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_compute: null genus");
    return tb_hecke_compute_rows(g, p, precise, 0, g->N, out);
}

extern "C" int64_t tb_hecke_num_conductors(const tb_hecke *h) { return h ? h->C : 0; }
extern "C" int64_t tb_hecke_conductor(const tb_hecke *h, int64_t k) { return h->g->conductors[k]; }
extern "C" int64_t tb_hecke_dim(const tb_hecke *h, int64_t k) { return h->g->dims[k]; }
extern "C" int64_t tb_hecke_rows(const tb_hecke *h, int64_t k) { return h->rows[k]; }
extern "C" int64_t tb_hecke_row_begin(const tb_hecke *h, int64_t k) { return h->rowfirst[k]; }
extern "C" int64_t tb_hecke_nnz(const tb_hecke *h, int64_t k) { return h->nnz[k]; }

#include "tb_readback.cuh"

extern "C" int tb_hecke_copy_dense(tb_hecke *h, int64_t k, int32_t *matrix)
{
    if (!h || k < 0 || k >= h->C) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_copy_dense: bad conductor index");
    cudaStream_t st = h->g->stream;
    const int rows = h->rows[k];
    const int dim = (int)h->g->dims[k];
    const size_t cells = (size_t)rows * dim;
    if (cells == 0) return TB_OK;
    if (int rc = ensure_wide(h, k)) return rc;
    cudaPointerAttributes attr;
    bool on_device = cudaPointerGetAttributes(&attr, matrix) == cudaSuccess &&
                     (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    PoolBuf<int> tmp;
    int *dst = matrix;
    if (!on_device)
    {
        TB_CUDA(tmp.reserve(cells, st));
        dst = tmp.ptr;
    }
    TB_CUDA(cudaMemsetAsync(dst, 0, cells * sizeof(int), st));
    if (h->nnz[k])
    {
        dense_scatter_kernel<<<rows, 128, 0, st>>>(h->d_indptr.ptr + h->rowbase[k], h->d_data.ptr + h->nnzbase[k],
                                                   h->d_indices.ptr + h->nnzbase[k], rows, dim, dst);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
    }
    if (!on_device) if (int rc = copy_out(h->g, matrix, dst, cells * sizeof(int))) { tmp.release(); return rc; }
    TB_CUDA(cudaStreamSynchronize(st));
    tmp.release();
    return TB_OK;
}

extern "C" void tb_hecke_destroy(tb_hecke *h)
{
    if (!h) return;
    tb_hecke_copy_wait(h);                       // an asynchronous read-back may still be reading the device arrays
    h->d_indptr.release(); h->d_data.release(); h->d_indices.release(); h->d_idx16.release(); h->d_dat8.release();
    delete h;
}

// ---------------------------------------------------------------------------
// per-neighbor records (parity tests)
// ---------------------------------------------------------------------------
extern "C" int tb_neighbor_records(tb_genus *g, uint64_t p, int precise, int64_t rep_begin, int64_t rep_end,
                                   int64_t *records)
{
    if (!g || !records) return fail(TB_ERR_INVALID_ARGUMENT, "tb_neighbor_records: null argument");
    if (divides_disc(g, p)) return fail(TB_ERR_INVALID_ARGUMENT, MSG_BAD_PRIME);
    if (rep_begin < 0 || rep_end > g->N || rep_begin >= rep_end)
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_neighbor_records: bad representative range");
    if (int rc = setup_prime(g, p)) return rc;
    const int rows = (int)(rep_end - rep_begin);
    const size_t words = (size_t)rows * (p + 1) * 20;
    if (int rc = prepare_rows(g, (int)rep_begin, rows, true)) return rc;
    DevBuf<i64> d_rec;
    TB_CUDA(d_rec.reserve(words));
    TB_CUDA(cudaMemsetAsync(d_rec.ptr, 0, words * sizeof(i64), g->stream));
    if (int rc = launch_neighbors<MODE_RECORDS>(g, precise, (int)rep_begin, rows, nullptr, d_rec.ptr)) { d_rec.release(); return rc; }
    TB_CUDA(cudaMemcpyAsync(g->h_err, g->d_err.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, g->stream));
    TB_CUDA(cudaMemcpyAsync(records, d_rec.ptr, words * sizeof(i64), cudaMemcpyDefault, g->stream));
    TB_CUDA(cudaStreamSynchronize(g->stream));
    d_rec.release();
    return check_errors(g);
}

// ---------------------------------------------------------------------------
// eigenvalues
// ---------------------------------------------------------------------------
extern "C" int tb_eigenvalues(tb_genus *g, uint64_t p, int precise, int64_t V, const int32_t *vecs,
                              const int64_t *cond_mask, const int64_t *rep_index, int32_t *out)
{
    if (!g || (V > 0 && (!vecs || !cond_mask || !rep_index || !out)))
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_eigenvalues: null argument");
    if (V == 0) return TB_OK;
    if (divides_disc(g, p)) return fail(TB_ERR_INVALID_ARGUMENT, MSG_BAD_PRIME);
    const int N = g->N;
    for (int64_t v = 0; v < V; ++v)
        if (rep_index[v] < 0 || rep_index[v] >= N) return fail(TB_ERR_INVALID_ARGUMENT, "tb_eigenvalues: bad representative index");
    if (int rc = setup_prime(g, p)) return rc;
    const u32 P1 = (u32)p + 1;

    // distinct representatives, ascending (EigenvectorManager::indices is sorted, Eigenvector.h:160-161)
    std::vector<int> reps;
    for (int64_t v = 0; v < V; ++v) reps.push_back((int)rep_index[v]);
    std::sort(reps.begin(), reps.end());
    reps.erase(std::unique(reps.begin(), reps.end()), reps.end());
    if (reps.size() > 65535) return fail(TB_ERR_INVALID_ARGUMENT, "tb_eigenvalues: more than 65535 distinct representatives");
    std::vector<uint32_t> base;
    base_points_for(g, p, reps, base);

    // one launch sequence for all set-cover representatives: base points -> prep -> neighbors -> dot
    const int R = (int)reps.size();
    std::vector<int> rep_of_vec(V);
    for (int64_t v = 0; v < V; ++v) rep_of_vec[v] = (int)rep_index[v];
    PoolBuf<int> d_vecs, d_meta, d_out;     // d_meta = rep_list[R] | rep_of_vec[V]
    PoolBuf<i64> d_cond;
    auto cleanup = [&]() { d_vecs.release(); d_meta.release(); d_out.release(); d_cond.release(); };
    TB_CUDA(d_vecs.reserve((size_t)V * N, g->stream));
    TB_CUDA(d_cond.reserve(V, g->stream));
    TB_CUDA(d_meta.reserve((size_t)R + V, g->stream));
    TB_CUDA(d_out.reserve(V, g->stream));
    TB_CUDA(cudaMemcpyAsync(d_vecs.ptr, vecs, (size_t)V * N * sizeof(int), cudaMemcpyDefault, g->stream));
    TB_CUDA(cudaMemcpyAsync(d_cond.ptr, cond_mask, V * sizeof(i64), cudaMemcpyDefault, g->stream));
    TB_CUDA(cudaMemcpyAsync(d_meta.ptr, reps.data(), R * sizeof(int), cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemcpyAsync(d_meta.ptr + R, rep_of_vec.data(), V * sizeof(int), cudaMemcpyHostToDevice, g->stream));
    TB_CUDA(cudaMemsetAsync(d_out.ptr, 0, V * sizeof(int), g->stream));
    TB_CUDA(g->d_W.reserve((size_t)R * P1));
    TB_CUDA(g->d_base.reserve(3 * (size_t)R));
    TB_CUDA(g->d_rp.reserve(R));
    g->rp_begin = -1;                                         // the staged rows below are not a Hecke shard's
    TB_CUDA(cudaMemcpyAsync(g->d_base.ptr, base.data(), base.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, g->stream));
    prep_reps_kernel<<<(R + 127) / 128, 128, 0, g->stream>>>(g->gc, g->ps.pc, g->d_base.ptr, 0, R, g->d_rp.ptr, d_meta.ptr);
    ++g_launches;
    if (int rc = launch_neighbors<MODE_PACKED>(g, precise, 0, R, g->d_W.ptr, nullptr, d_meta.ptr)) { cleanup(); return rc; }
    eigen_dot_kernel<<<dim3((P1 + 255) / 256, R), 256, 0, g->stream>>>(g->d_W.ptr, P1, g->K, N, d_vecs.ptr, d_cond.ptr,
                                                                       d_meta.ptr + R, (int)V, d_meta.ptr, d_out.ptr);
    ++g_launches;
    TB_CUDA(cudaGetLastError());
    std::vector<int> sums(V);
    TB_CUDA(cudaMemcpyAsync(g->h_err, g->d_err.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, g->stream));
    TB_CUDA(cudaMemcpyAsync(sums.data(), d_out.ptr, V * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    TB_CUDA(cudaStreamSynchronize(g->stream));
    cleanup();
    if (int rc = check_errors(g)) return rc;
    // host copy of the pivot coordinates for the final exact division (Genus.h:504-510)
    std::vector<int> pivot(V);
    cudaPointerAttributes attr;
    bool vecs_on_device = cudaPointerGetAttributes(&attr, vecs) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    for (int64_t v = 0; v < V; ++v)
    {
        if (vecs_on_device)
            TB_CUDA(cudaMemcpy(&pivot[v], vecs + (size_t)v * N + rep_index[v], sizeof(int), cudaMemcpyDeviceToHost));
        else
            pivot[v] = vecs[(size_t)v * N + rep_index[v]];
        if (pivot[v] == 0) return fail(TB_ERR_INVALID_ARGUMENT, "tb_eigenvalues: eigenvector vanishes at its representative");
        out[v] = sums[v] / pivot[v];
    }
    return TB_OK;
}

// ---------------------------------------------------------------------------
// isometry sequence
// ---------------------------------------------------------------------------
extern "C" int tb_isoseq_open(tb_genus *g, uint64_t p, int precise, tb_isoseq **out)
{
    if (!g || !out) return fail(TB_ERR_INVALID_ARGUMENT, "tb_isoseq_open: null argument");
    *out = nullptr;
    if (int rc = setup_prime(g, p)) return rc;
    tb_isoseq *it = new tb_isoseq;
    it->g = g; it->p = p; it->precise = precise;
    it->total = (int64_t)g->N * (int64_t)(p + 1);
    *out = it;
    return TB_OK;
}

extern "C" int tb_isoseq_done(const tb_isoseq *it) { return !it || it->delivered >= it->total; }

static int isoseq_fill(tb_isoseq *it)
{
    tb_genus *g = it->g;
    if (int rc = setup_prime(g, it->p)) return rc;
    const int64_t P1 = (int64_t)it->p + 1;
    const int64_t max_records = 1 << 20;
    int rows = (int)std::max<int64_t>(1, std::min<int64_t>(g->N - it->next_rep, max_records / P1));
    const size_t words = (size_t)rows * P1 * 12;
    if (int rc = prepare_rows(g, (int)it->next_rep, rows, true)) return rc;
    DevBuf<i64> d_rec;
    TB_CUDA(d_rec.reserve(words));
    if (int rc = launch_neighbors<MODE_ISOSEQ>(g, it->precise, (int)it->next_rep, rows, nullptr, d_rec.ptr)) { d_rec.release(); return rc; }
    it->chunk.resize(words);
    TB_CUDA(cudaMemcpyAsync(g->h_err, g->d_err.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, g->stream));
    TB_CUDA(cudaMemcpyAsync(it->chunk.data(), d_rec.ptr, words * sizeof(i64), cudaMemcpyDeviceToHost, g->stream));
    TB_CUDA(cudaStreamSynchronize(g->stream));
    d_rec.release();
    if (int rc = check_errors(g)) return rc;
    it->next_rep += rows;
    it->pos = 0;
    return TB_OK;
}

extern "C" int tb_isoseq_read(tb_isoseq *it, int64_t max_records, int64_t *records, int64_t *got)
{
    if (!it || !records || !got) return fail(TB_ERR_INVALID_ARGUMENT, "tb_isoseq_read: null argument");
    *got = 0;
    while (*got < max_records && it->delivered < it->total)
    {
        if (it->pos * 12 >= it->chunk.size())
            if (int rc = isoseq_fill(it)) return rc;
        const int64_t avail = (int64_t)(it->chunk.size() / 12 - it->pos);
        const int64_t take = std::min<int64_t>(avail, max_records - *got);
        memcpy(records + 12 * *got, &it->chunk[12 * it->pos], (size_t)take * 12 * sizeof(int64_t));
        it->pos += take;
        it->delivered += take;
        *got += take;
    }
    return TB_OK;
}

extern "C" int tb_isoseq_next(tb_isoseq *it, int64_t isometry[9], int64_t *denominator, int64_t *src, int64_t *dst)
{
    if (!it) return fail(TB_ERR_INVALID_ARGUMENT, "tb_isoseq_next: null iterator");
    if (tb_isoseq_done(it)) return fail(TB_ERR_DOMAIN, MSG_NO_MORE);
    int64_t rec[12], got = 0;
    if (int rc = tb_isoseq_read(it, 1, rec, &got)) return rc;
    memcpy(isometry, rec, 9 * sizeof(int64_t));
    if (denominator) *denominator = rec[9];
    if (src) *src = rec[10];
    if (dst) *dst = rec[11];
    return TB_OK;
}

extern "C" void tb_isoseq_close(tb_isoseq *it) { delete it; }

// ---------------------------------------------------------------------------
// genus construction
// ---------------------------------------------------------------------------
struct tb_table {
    BuiltGenus g;
};

static std::vector<PrimeSym> to_symbols(const tb_prime_symbol *symbols, int64_t count)
{
    std::vector<PrimeSym> out;
    for (int64_t i = 0; i < count; ++i) out.push_back(PrimeSym{symbols[i].p, (int)symbols[i].power, symbols[i].ramified != 0});
    return out;
}

extern "C" int tb_quad_form(const tb_prime_symbol *symbols, int64_t count, int64_t form[6])
{
    if (!symbols || !form || count <= 0) return fail(TB_ERR_INVALID_ARGUMENT, "tb_quad_form: bad argument");
    if (int rc = require_device()) return rc;
    try { quad_form_for(to_symbols(symbols, count), 0, form); }
    catch (const std::invalid_argument &e) { return fail(TB_ERR_INVALID_ARGUMENT, e.what()); }
    catch (const std::exception &e) { return fail(TB_ERR_RUNTIME, e.what()); }
    return TB_OK;
}

static void make_prime_ctx(uint64_t p, double lambda_min, const u32 *inv_lut, PrimeCtx &pc)
{
    memset(&pc, 0, sizeof(pc));
    pc.p = (u32)p;
    pc.pp = p * p;
    pc.dp = make_invdiv(p);
    pc.dpp = make_invdiv(p * p);
    pc.use_lut = (p > 2 && p < (1ull << 22)) ? 1u : 0u;
    pc.m32 = (p > 2 && p < (1ull << 16)) ? (u32)((1ull << 32) / p) : 0u;
    pc.fp64_reduce = (lambda_min > 0 && (double)p * sqrt(1125899906842624.0 / lambda_min) < 4503599627370496.0) ? 1u : 0u;
    pc.inv_lut = inv_lut;
}

// Genus<R>::Genus, Genus.h:37-246.
static int enumerate_genus(const int64_t form[6], const std::vector<PrimeSym> &symbols, uint64_t seed, BuiltGenus &G)
{
    if (symbols.size() > 63) return fail(TB_ERR_DOMAIN, "Must have 63 or fewer prime divisors.");
    if (symbols.size() > TB_MAX_PRIMES) return fail(TB_ERR_RUNTIME, "libtbirch: more than 16 prime divisors");
    if (seed == 0) { std::random_device rd; seed = rd(); }   // Genus.h:39-43
    const big a0 = form[0], b0 = form[1], c0 = form[2], f0 = form[3], g0 = form[4], h0 = form[5];
    const big disc = a0 * (4 * b0 * c0 - f0 * f0) - b0 * g0 * g0 + h0 * (f0 * g0 - c0 * h0);
    if (disc <= 0 || disc != (big)(int64_t)disc) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_enumerate: discriminant out of range");
    G.K = (int64_t)symbols.size();
    G.seed = seed;
    G.disc = (int64_t)disc;
    for (const PrimeSym &s : symbols) G.primes.push_back(s.p);
    const int64_t C = 1ll << G.K;
    G.conductors.assign(1, 1);                               // Genus.h:62-78
    {
        size_t bits = 0, mask = 1;
        for (int64_t n = 1; n < C; ++n)
        {
            if ((size_t)n == 2 * mask) { ++bits; mask = (size_t)1 << bits; }
            G.conductors.push_back(G.primes[bits] * G.conductors[n ^ mask]);
        }
    }
    const big mass_x24 = genus_mass_x24(form, disc, symbols);

    std::vector<std::array<int64_t, 6>> q;                   // representatives in discovery order
    std::vector<std::array<big, 9>> s;                       // parent -> child isometries
    std::vector<int64_t> parent, rp;
    std::unordered_map<FormKey, int, FormKeyHash> index;
    auto add_rep = [&](const i64 *f, const i64 *iso, int64_t par, int64_t prime) {
        FormKey key;
        memcpy(key.v, f, sizeof(key.v));
        if (!index.emplace(key, (int)q.size()).second) return false;
        q.push_back({(int64_t)f[0], (int64_t)f[1], (int64_t)f[2], (int64_t)f[3], (int64_t)f[4], (int64_t)f[5]});
        std::array<big, 9> m;
        for (int i = 0; i < 9; ++i) m[i] = iso ? (big)iso[i] : (big)((i % 4) == 0);
        s.push_back(m);
        parent.push_back(par);
        rp.push_back(prime);
        return true;
    };
    auto auts_of = [&](const i64 *f) { return (int64_t)(2 * (proper_automorphisms((const int64_t *)f).size() + 1)); };
    add_rep((const i64 *)form, nullptr, -1, 1);
    big sum_x24 = 48 / auts_of((const i64 *)form);
    bool done = (sum_x24 == mass_x24);

    cudaStream_t st = 0;
    DevBuf<i64> d_forms, d_rec;
    DevBuf<u32> d_base, d_lut;
    DevBuf<RepPrime> d_rp;
    auto release = [&]() { d_forms.release(); d_rec.release(); d_base.release(); d_lut.release(); d_rp.release(); };
    int64_t p = 1;
    while (!done)
    {
        do { p = next_prime64(p); } while (disc % p == 0);   // Genus.h:116-123
        if (p >= 65536) { release(); return fail(TB_ERR_RUNTIME, "tb_genus_enumerate: mass not reached below p = 2^16 (inconsistent input?)"); }
        PrimeCtx pc;
        if (p > 2)
        {
            if (d_lut.reserve((size_t)p) != cudaSuccess) { release(); return fail(TB_ERR_RUNTIME, "CUDA allocation failed"); }
        }
        make_prime_ctx((uint64_t)p, 0.0, d_lut.ptr, pc);
        if (pc.use_lut)
        {
            inverse_lut_kernel<<<(unsigned)((p + 255) / 256), 256, 0, st>>>(pc, d_lut.ptr);
            ++g_launches;
        }
        HostFp F((uint64_t)p, seed);                         // one Fp (one RNG stream) per prime
        size_t current = 0;
        while (!done && current < q.size())
        {
            // one wave: every representative not yet expanded at this prime
            const size_t wave_begin = current, wave_end = q.size();
            const int rows = (int)(wave_end - wave_begin);
            const u32 P1 = (u32)p + 1;
            std::vector<i64> forms((size_t)rows * TB_REC_WORDS, 0);
            std::vector<u32> base(3 * (size_t)rows);
            for (int i = 0; i < rows; ++i)
            {
                memcpy(&forms[(size_t)i * TB_REC_WORDS], q[wave_begin + i].data(), 6 * sizeof(i64));
                base_point(F, (const int64_t *)q[wave_begin + i].data(), &base[3 * (size_t)i]);
            }
            const size_t nrec = (size_t)rows * P1;
            if (d_forms.reserve(forms.size()) != cudaSuccess || d_base.reserve(base.size()) != cudaSuccess ||
                d_rp.reserve(rows) != cudaSuccess || d_rec.reserve(nrec * 16) != cudaSuccess)
            { release(); return fail(TB_ERR_RUNTIME, "CUDA allocation failed"); }
            cudaMemcpyAsync(d_forms.ptr, forms.data(), forms.size() * sizeof(i64), cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_base.ptr, base.data(), base.size() * sizeof(u32), cudaMemcpyHostToDevice, st);
            GenusCtx gc;
            memset(&gc, 0, sizeof(gc));
            gc.N = rows; gc.K = 0; gc.cur = d_forms.ptr; gc.rep = d_forms.ptr;
            prep_reps_kernel<<<(rows + 127) / 128, 128, 0, st>>>(gc, pc, d_base.ptr, 0, rows, d_rp.ptr);
            EnumLaunch L;
            memset(&L, 0, sizeof(L));
            L.rows = rows; L.P1 = P1; L.total = nrec; L.dP1 = make_invdiv(P1); L.rp = d_rp.ptr; L.records = d_rec.ptr;
            enum_kernel<<<(unsigned)((nrec + TB_NBR_THREADS - 1) / TB_NBR_THREADS), TB_NBR_THREADS, 0, st>>>(gc, pc, L);
            g_launches += 2;
            std::vector<i64> rec(nrec * 16);
            cudaMemcpyAsync(rec.data(), d_rec.ptr, rec.size() * sizeof(i64), cudaMemcpyDeviceToHost, st);
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { release(); return fail(TB_ERR_RUNTIME, std::string("CUDA error: ") + cudaGetErrorString(e)); }

            for (int i = 0; i < rows && !done; ++i)          // Genus.h:128-175, in order
            {
                for (u32 t = 0; t < P1 && !done; ++t)
                {
                    const i64 *r = &rec[((size_t)i * P1 + t) * 16];
                    if (r[15] != 0) { release(); return fail(TB_ERR_OVERFLOW, MSG_OVERFLOW); }
                    if (add_rep(r, r + 6, (int64_t)(wave_begin + i), p))
                    {
                        sum_x24 += 48 / auts_of(r);
                        done = (sum_x24 == mass_x24);
                    }
                }
                ++current;
            }
            if (sum_x24 > mass_x24) { release(); return fail(TB_ERR_RUNTIME, "tb_genus_enumerate: mass exceeded (inconsistent form / symbols)"); }
        }
    }
    release();

    // isometries to / from the mother form, spinor-prime exponents     Genus.h:189-217
    const size_t N = q.size();
    std::vector<std::array<big, 9>> S(N), Sinv(N);
    std::vector<big> scalar(N, 1);
    G.N = (int64_t)N;
    G.dims.assign(C, 0);
    G.lut.assign((size_t)C * N, -1);
    for (size_t n = 0; n < N; ++n)
    {
        if (n == 0)
        {
            for (int i = 0; i < 9; ++i) S[0][i] = Sinv[0][i] = (i % 4) == 0;
        }
        else
        {
            const std::array<big, 9> &m = s[n];
            const big pr = rp[n];
            std::array<big, 9> adj;                          // Isometry::inverse(p), Isometry.h:33-46
            adj[0] = (m[4] * m[8] - m[5] * m[7]) / pr; adj[1] = (m[2] * m[7] - m[1] * m[8]) / pr; adj[2] = (m[1] * m[5] - m[2] * m[4]) / pr;
            adj[3] = (m[5] * m[6] - m[3] * m[8]) / pr; adj[4] = (m[0] * m[8] - m[2] * m[6]) / pr; adj[5] = (m[2] * m[3] - m[0] * m[5]) / pr;
            adj[6] = (m[3] * m[7] - m[4] * m[6]) / pr; adj[7] = (m[1] * m[6] - m[0] * m[7]) / pr; adj[8] = (m[0] * m[4] - m[1] * m[3]) / pr;
            const size_t par = (size_t)parent[n];
            mat3_mul(adj.data(), Sinv[par].data(), Sinv[n].data());
            mat3_mul(S[par].data(), m.data(), S[n].data());
            scalar[n] = scalar[par] * pr;
        }
        // which conductors see this representative                     Genus.h:219-244
        const std::vector<SmallIso> auts = proper_automorphisms(q[n].data());
        std::vector<char> ignore(C, 0);
        for (const SmallIso &au : auts)
        {
            const u64 vals = host_spinor_norm(q[n].data(), au, G.primes);
            for (int64_t k = 0; k < C; ++k)
                if (!ignore[k] && (__builtin_popcountll(vals & (u64)k) & 1)) ignore[k] = 1;
        }
        G.num_auts.push_back((int64_t)(2 * (auts.size() + 1)));
        for (int64_t k = 0; k < C; ++k)
            if (!ignore[k]) G.lut[(size_t)k * N + n] = (int32_t)G.dims[k]++;
    }
    for (size_t n = 0; n < N; ++n)
    {
        for (int i = 0; i < 6; ++i) G.q.push_back(q[n][i]);
        for (int i = 0; i < 9; ++i)
        {
            if (S[n][i] != (big)(int64_t)S[n][i] || Sinv[n][i] != (big)(int64_t)Sinv[n][i])
                return fail(TB_ERR_OVERFLOW, "tb_genus_enumerate: composed isometry exceeds 64 bits");
            G.s.push_back((int64_t)S[n][i]);
        }
        for (int i = 0; i < 9; ++i) G.sinv.push_back((int64_t)Sinv[n][i]);
        if (scalar[n] != (big)(int64_t)scalar[n]) return fail(TB_ERR_OVERFLOW, "tb_genus_enumerate: scalar exceeds 64 bits");
        G.scalar.push_back((int64_t)scalar[n]);
        G.parent.push_back(parent[n]);
        G.p.push_back(rp[n]);
    }
    return TB_OK;
}

extern "C" int tb_genus_enumerate(const int64_t form[6], const tb_prime_symbol *symbols, int64_t count,
                                  uint64_t seed, tb_table **out)
{
    if (!form || !symbols || !out || count <= 0) return fail(TB_ERR_INVALID_ARGUMENT, "tb_genus_enumerate: bad argument");
    *out = nullptr;
    if (int rc = require_device()) return rc;
    std::unique_ptr<tb_table> t(new tb_table);
    try
    {
        if (int rc = enumerate_genus(form, to_symbols(symbols, count), seed, t->g)) return rc;
    }
    catch (const std::exception &e) { return fail(TB_ERR_RUNTIME, e.what()); }
    *out = t.release();
    return TB_OK;
}

extern "C" int tb_table_desc(const tb_table *t, tb_genus_desc *d, const int64_t **parent, const int64_t **rep_prime,
                             const int64_t **num_auts)
{
    if (!t || !d) return fail(TB_ERR_INVALID_ARGUMENT, "tb_table_desc: null argument");
    const BuiltGenus &g = t->g;
    d->N = g.N; d->K = g.K; d->seed = g.seed; d->disc = g.disc;
    d->primes = g.primes.data(); d->conductors = g.conductors.data(); d->dims = g.dims.data();
    d->q = g.q.data(); d->s = g.s.data(); d->sinv = g.sinv.data(); d->scalar = g.scalar.data(); d->lut = g.lut.data();
    if (parent) *parent = g.parent.data();
    if (rep_prime) *rep_prime = g.p.data();
    if (num_auts) *num_auts = g.num_auts.data();
    return TB_OK;
}

extern "C" void tb_table_destroy(tb_table *t) { delete t; }

// ---------------------------------------------------------------------------
// integer-pipe microbenchmark
// ---------------------------------------------------------------------------
template <int KIND>
static int peak_one(double ops_per_iter_per_thread, double *out)
{
    int dev = 0, sms = 0;
    TB_CUDA(cudaGetDevice(&dev));
    TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    u32 *sink = nullptr;
    TB_CUDA(cudaMalloc((void **)&sink, 4));
    cudaEvent_t e0, e1;
    TB_CUDA(cudaEventCreate(&e0));
    TB_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    int_peak_kernel<KIND><<<blocks, threads>>>(sink, 64, 12345u);   // warm-up
    TB_CUDA(cudaDeviceSynchronize());
    double best = 0;
    for (int rep = 0; rep < 3; ++rep)
    {
        TB_CUDA(cudaEventRecord(e0));
        int_peak_kernel<KIND><<<blocks, threads>>>(sink, iters, 12345u + rep);
        TB_CUDA(cudaEventRecord(e1));
        TB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * iters * ops_per_iter_per_thread;
        best = std::max(best, ops / (ms * 1e-3));
        g_launches += 1;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *out = best;
    return TB_OK;
}

extern "C" int tb_int_peak_detail(double out[7])
{
    if (!out) return fail(TB_ERR_INVALID_ARGUMENT, "tb_int_peak_detail: null argument");
    if (int rc = require_device()) return rc;
    // operations per inner iteration per thread = SASS instructions of the unrolled body (tb_kernels.cuh)
    const double single = (double)TB_PEAK_UNROLL * TB_PEAK_OPS_SINGLE, mixed = (double)TB_PEAK_UNROLL * TB_PEAK_OPS_MIXED;
    if (int rc = peak_one<0>(single, &out[0])) return rc;
    if (int rc = peak_one<1>(single, &out[1])) return rc;
    if (int rc = peak_one<2>(single, &out[2])) return rc;
    if (int rc = peak_one<3>(mixed, &out[3])) return rc;
    if (int rc = peak_one<4>(mixed, &out[4])) return rc;
    if (int rc = peak_one<5>(single, &out[5])) return rc;
    if (int rc = peak_one<6>((double)TB_PEAK_UNROLL * TB_PEAK_OPS_TRIPLE, &out[6])) return rc;
    return TB_OK;
}

extern "C" int tb_int_peak(double *imad, double *iadd, double *mixed)
{
    double v[7];
    if (int rc = tb_int_peak_detail(v)) return rc;
    if (imad) *imad = v[0];
    if (iadd) *iadd = v[1];
    if (mixed) *mixed = std::max(v[2], std::max(v[3], v[4]));
    return TB_OK;
}

#include "tb_linalg_host.cuh"
