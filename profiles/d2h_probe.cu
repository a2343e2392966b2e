// d2h_probe.cu -- what limits device->host read-back when several GPUs of one box ship results at once?
//
// VERDICT r1 "weak" #1: the end-to-end metric (int32 CSR of every T_p delivered into pinned host
// memory) does not scale past one GPU: 18 GB/s per GPU at >= 4 ranks against 54-56 GB/s alone, with a
// multi-second slow phase after a barrier.  nsys is not in the image, so this probe times the DMA
// itself, one OS process per GPU (as bench.py runs), synchronised through a process-shared barrier:
//
//   solo        each rank alone, 1 GiB copies                     -> per-GPU PCIe rate
//   all         all ranks at once, 256 MiB copies for D seconds   -> rate per copy over time (ramp?)
//   all+spin    the same with S busy threads per rank             -> do idle host cores slow the DMA?
//   all+touch   the same with S threads streaming through memory  -> does host memory load speed it up?
//   all+huge    the same into a cudaHostRegister'ed THP region    -> page size of the pinned mapping
//   pairs       rank 0 with each other rank                       -> which GPUs share a link
//   nt-store    host streaming-store bandwidth, all ranks         -> what the widening path can reach
//
// Output: one JSON object per line on stdout (rank 0 prints the merged view).  No torch, no library.
//   nvcc -O2 -std=c++17 -o profiles/_bin/d2h_probe profiles/d2h_probe.cu -lpthread
//   profiles/_bin/d2h_probe <ranks> [seconds]
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <immintrin.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "rank %d: %s: %s\n", g_rank, #x, cudaGetErrorString(e_)); _exit(3); } } while (0)

static int g_rank = 0, g_world = 1;
static const int MAX_RANKS = 8, MAX_SAMPLES = 4096;

struct Shared {
    pthread_barrier_t bar;
    // per rank, per test: samples (time since test start [s], GB/s)
    int count[MAX_RANKS];
    float t[MAX_RANKS][MAX_SAMPLES], gbps[MAX_RANKS][MAX_SAMPLES];
    double scalar[MAX_RANKS];
};
static Shared *S = nullptr;

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void barrier() { pthread_barrier_wait(&S->bar); }

// copy `chunk` bytes device->host repeatedly for `seconds`; record each copy
static void run_copies(char *host, const char *dev, size_t buf, size_t chunk, double seconds, cudaStream_t st, bool active)
{
    S->count[g_rank] = 0;
    barrier();
    const double t0 = now();
    size_t off = 0;
    while (active && now() - t0 < seconds && S->count[g_rank] < MAX_SAMPLES)
    {
        const double a = now();
        CK(cudaMemcpyAsync(host + off, dev + off, chunk, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const double b = now();
        const int i = S->count[g_rank]++;
        S->t[g_rank][i] = (float)(b - t0);
        S->gbps[g_rank][i] = (float)(chunk / (b - a) / 1e9);
        off += chunk;
        if (off + chunk > buf) off = 0;
    }
    barrier();
}

static void report(const char *name, const char *note)
{
    barrier();
    if (g_rank == 0)
    {
        printf("{\"test\": \"%s\", \"note\": \"%s\", \"ranks\": [", name, note);
        double agg_first = 0, agg_last = 0;
        for (int r = 0; r < g_world; ++r)
        {
            const int n = S->count[r];
            // mean rate of the first and the last quarter of the samples, and min / max
            double first = 0, last = 0, mn = 1e9, mx = 0, mean = 0;
            const int q = n / 4 > 0 ? n / 4 : 1;
            for (int i = 0; i < n; ++i)
            {
                const double v = S->gbps[r][i];
                mean += v; if (v < mn) mn = v; if (v > mx) mx = v;
                if (i < q) first += v;
                if (i >= n - q) last += v;
            }
            if (n) { mean /= n; first /= q; last /= q; } else mn = 0;
            agg_first += first; agg_last += last;
            printf("%s{\"rank\": %d, \"copies\": %d, \"mean\": %.1f, \"first_quarter\": %.1f, \"last_quarter\": %.1f, \"min\": %.1f, \"max\": %.1f, \"series\": [",
                   r ? ", " : "", r, n, mean, first, last, mn, mx);
            // thin the time series to <= 40 points
            const int step = n > 40 ? n / 40 : 1;
            bool firstp = true;
            for (int i = 0; i < n; i += step)
            {
                printf("%s[%.2f, %.1f]", firstp ? "" : ", ", S->t[r][i], S->gbps[r][i]);
                firstp = false;
            }
            printf("]}");
        }
        printf("], \"aggregate_first_quarter\": %.1f, \"aggregate_last_quarter\": %.1f, \"unit\": \"GB/s\"}\n", agg_first, agg_last);
        fflush(stdout);
    }
    barrier();
}

static std::atomic<bool> g_stop(false);
static void spin_thread() { while (!g_stop.load(std::memory_order_relaxed)) { _mm_pause(); } }
static void touch_thread(char *mem, size_t bytes)
{
    volatile long sink = 0;
    while (!g_stop.load(std::memory_order_relaxed))
    {
        long s = 0;
        for (size_t i = 0; i < bytes; i += 64) s += mem[i];
        sink += s;
    }
}

__attribute__((target("avx2"))) static void nt_fill(char *dst, size_t bytes)
{
    const __m256i v = _mm256_set1_epi32(0x01020304);
    for (size_t i = 0; i + 32 <= bytes; i += 32) _mm256_stream_si256((__m256i *)(dst + i), v);
    _mm_sfence();
}

static int child(int rank, int world, double seconds, int cores)
{
    g_rank = rank; g_world = world;
    CK(cudaSetDevice(rank));
    const size_t BUF = (size_t)1 << 30, CHUNK = (size_t)256 << 20;
    char *dev = nullptr, *pinned = nullptr;
    CK(cudaMalloc((void **)&dev, BUF));
    CK(cudaMemset(dev, 1, BUF));
    CK(cudaHostAlloc((void **)&pinned, BUF, cudaHostAllocDefault));
    memset(pinned, 0, BUF);
    // transparent-huge-page region, registered
    char *huge = (char *)mmap(nullptr, BUF + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    char *huge_al = (char *)(((uintptr_t)huge + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
    madvise(huge_al, BUF, MADV_HUGEPAGE);
    memset(huge_al, 0, BUF);
    const bool huge_ok = cudaHostRegister(huge_al, BUF, cudaHostRegisterDefault) == cudaSuccess;
    cudaGetLastError();
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaMemcpyAsync(pinned, dev, BUF, cudaMemcpyDeviceToHost, st));   // warm the path once
    CK(cudaStreamSynchronize(st));
    const int S_THREADS = cores / world > 4 ? 4 : (cores / world > 1 ? cores / world - 1 : 1);
    char note[256];

    // ---- solo: one rank at a time, the others asleep in the barrier ----
    for (int r = 0; r < world; ++r)
    {
        run_copies(pinned, dev, BUF, BUF, 0.6, st, rank == r);
        if (rank == r)
        {
            double m = 0;
            for (int i = 0; i < S->count[rank]; ++i) m += S->gbps[rank][i];
            S->scalar[rank] = S->count[rank] ? m / S->count[rank] : 0;
        }
    }
    barrier();
    if (rank == 0)
    {
        printf("{\"test\": \"solo\", \"note\": \"1 GiB copies, one rank at a time, cudaHostAlloc\", \"gbps\": [");
        for (int r = 0; r < world; ++r) printf("%s%.1f", r ? ", " : "", S->scalar[r]);
        printf("]}\n");
        fflush(stdout);
    }
    barrier();

    // ---- all ranks at once, after 2 s of idleness ----
    sleep(2);
    run_copies(pinned, dev, BUF, CHUNK, seconds, st, true);
    report("all", "all ranks, 256 MiB copies into cudaHostAlloc memory, started after 2 s idle");

    // ---- immediately again (no idle gap) ----
    run_copies(pinned, dev, BUF, CHUNK, seconds / 2, st, true);
    report("all_again", "the same, back to back (no idle gap)");

    // ---- with busy threads ----
    sleep(2);
    {
        g_stop = false;
        std::vector<std::thread> th;
        for (int i = 0; i < S_THREADS; ++i) th.emplace_back(spin_thread);
        run_copies(pinned, dev, BUF, CHUNK, seconds, st, true);
        g_stop = true;
        for (auto &t : th) t.join();
        snprintf(note, sizeof note, "all ranks + %d spinning host threads per rank, after 2 s idle", S_THREADS);
        report("all_spin", note);
    }
    sleep(2);
    {
        g_stop = false;
        const size_t TB = (size_t)256 << 20;
        char *mem = (char *)malloc(TB * S_THREADS);
        memset(mem, 1, TB * S_THREADS);
        std::vector<std::thread> th;
        for (int i = 0; i < S_THREADS; ++i) th.emplace_back(touch_thread, mem + TB * i, TB);
        run_copies(pinned, dev, BUF, CHUNK, seconds, st, true);
        g_stop = true;
        for (auto &t : th) t.join();
        free(mem);
        snprintf(note, sizeof note, "all ranks + %d host threads per rank streaming reads through 256 MiB each, after 2 s idle", S_THREADS);
        report("all_touch", note);
    }
    // ---- huge pages ----
    sleep(2);
    if (huge_ok)
    {
        run_copies(huge_al, dev, BUF, CHUNK, seconds, st, true);
        report("all_huge", "all ranks into a cudaHostRegister'ed MADV_HUGEPAGE region, after 2 s idle");
    }
    // ---- small copies: 8 MiB ----
    sleep(2);
    run_copies(pinned, dev, BUF, (size_t)8 << 20, seconds / 2, st, true);
    report("all_8MiB", "all ranks, 8 MiB copies, after 2 s idle");

    // ---- pairs with rank 0 ----
    for (int r = 1; r < world; ++r)
    {
        run_copies(pinned, dev, BUF, CHUNK, 1.0, st, rank == 0 || rank == r);
        snprintf(note, sizeof note, "ranks 0 and %d only", r);
        report("pair", note);
    }
    // ---- first half / second half of the ranks ----
    if (world >= 4)
    {
        run_copies(pinned, dev, BUF, CHUNK, 1.0, st, rank < world / 2);
        report("half_low", "ranks [0, world/2) only");
        run_copies(pinned, dev, BUF, CHUNK, 1.0, st, rank >= world / 2);
        report("half_high", "ranks [world/2, world) only");
    }

    // ---- host streaming stores: cores/world threads per rank, 1 GiB per thread-set ----
    {
        const int T = cores / world > 0 ? cores / world : 1;
        char *mem = (char *)aligned_alloc(4096, BUF);
        memset(mem, 0, BUF);
        barrier();
        const double a = now();
        for (int rep = 0; rep < 4; ++rep)
        {
            std::vector<std::thread> th;
            const size_t per = BUF / T / 4096 * 4096;
            for (int i = 0; i < T; ++i) th.emplace_back(nt_fill, mem + per * i, per);
            for (auto &t : th) t.join();
        }
        const double b = now();
        S->scalar[rank] = 4.0 * (BUF / T / 4096 * 4096) * T / (b - a) / 1e9;
        barrier();
        if (rank == 0)
        {
            double agg = 0;
            printf("{\"test\": \"nt_store\", \"note\": \"AVX2 streaming stores, %d threads per rank, all ranks at once\", \"gbps\": [", T);
            for (int r = 0; r < world; ++r) { printf("%s%.1f", r ? ", " : "", S->scalar[r]); agg += S->scalar[r]; }
            printf("], \"aggregate\": %.1f}\n", agg);
            fflush(stdout);
        }
        barrier();
        // the same while every rank's DMA runs too (plain DMA + host stores compete for host DRAM)
        g_stop = false;
        std::thread filler([&] {
            const size_t per = BUF / T / 4096 * 4096;
            while (!g_stop.load())
            {
                std::vector<std::thread> th;
                for (int i = 0; i < T; ++i) th.emplace_back(nt_fill, mem + per * i, per);
                for (auto &t : th) t.join();
            }
        });
        run_copies(pinned, dev, BUF, CHUNK, seconds / 2, st, true);
        g_stop = true;
        filler.join();
        report("all_with_nt_store", "all ranks' DMA while every rank also runs its streaming-store threads");
        free(mem);
    }
    return 0;
}

int main(int argc, char **argv)
{
    const int world = argc > 1 ? atoi(argv[1]) : 1;
    const double seconds = argc > 2 ? atof(argv[2]) : 5.0;
    if (world < 1 || world > MAX_RANKS) { fprintf(stderr, "ranks must be 1..8\n"); return 2; }
    const int cores = (int)std::thread::hardware_concurrency();
    S = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset(S, 0, sizeof(Shared));
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&S->bar, &at, world);
    printf("{\"test\": \"setup\", \"ranks\": %d, \"host_cores\": %d, \"seconds_per_test\": %.1f}\n", world, cores, seconds);
    fflush(stdout);
    std::vector<pid_t> kids;
    for (int r = 0; r < world; ++r)
    {
        pid_t pid = fork();                 // before any CUDA call: each child gets its own context
        if (pid == 0) _exit(child(r, world, seconds, cores));
        kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (st) rc = 1; }
    return rc;
}
