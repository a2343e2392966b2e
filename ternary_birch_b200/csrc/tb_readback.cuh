// tb_readback.cuh -- host side of getting a T_p out of the device: part of tbirch.cu (included
// there, after tb_genus / tb_hecke and the error helpers are defined; not a standalone header).
//
//   HostPool, parallel_memcpy    a small pool of niced host threads
//   copy_out                     staged copies into pageable memory
//   widen_*, Copier, narrow_copy the narrow read-back
//   tb_hecke_copy_sparse[_async|_all], tb_hecke_copy_wait, tb_hecke_transfer_bytes,
//   tb_host_widen_probe          the C ABI entry points
#pragma once

// A small process-wide pool: parallel_for over [0, n) in `parts` slices.
class HostPool {
public:
    static HostPool &get() { static HostPool p; return p; }
    int size() const { return (int)workers.size() + 1; }
    template <class F>
    void run(int parts, F fn)
    {
        std::function<void(int)> f = fn;
        std::unique_lock<std::mutex> user(user_mx);          // one parallel region at a time
        {
            std::lock_guard<std::mutex> lk(mx);
            job = &f; next = 0; total = parts; pending = parts; ++epoch;
        }
        cv.notify_all();
        work();                                              // the calling thread takes slices too
        std::unique_lock<std::mutex> lk(mx);
        done_cv.wait(lk, [&] { return pending == 0; });
        job = nullptr;
    }
private:
    HostPool()
    {
        int n = 0;
        if (const char *e = getenv("TBIRCH_HOST_THREADS")) n = atoi(e);
        if (n <= 0)
        {
            int hw = (int)std::thread::hardware_concurrency();
            int local = 1;
            if (const char *e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));
            n = std::max(2, std::min(32, hw / local));     // workers are niced: the launching thread preempts them
        }
        for (int i = 1; i < n; ++i) workers.emplace_back([this] { loop(); });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> lk(mx); stop = true; }
        cv.notify_all();
        for (auto &t : workers) t.join();
    }
    void work()
    {
        for (;;)
        {
            int i;
            {
                std::lock_guard<std::mutex> lk(mx);
                if (!job || next >= total) return;
                i = next++;
            }
            (*job)(i);
            std::lock_guard<std::mutex> lk(mx);
            if (--pending == 0) done_cv.notify_all();
        }
    }
    void loop()
    {
        // widening is bulk work: let the thread that launches kernels (and CUDA's own threads) preempt it
        setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10);
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(mx);
                cv.wait(lk, [&] { return stop || epoch != seen; });
                if (stop) return;
                seen = epoch;
            }
            work();
        }
    }
    std::vector<std::thread> workers;
    std::mutex mx, user_mx;
    std::condition_variable cv, done_cv;
    std::function<void(int)> *job = nullptr;
    int next = 0, total = 0, pending = 0;
    uint64_t epoch = 0;
    bool stop = false;
};

// memcpy into pageable memory with the pool: fresh numpy arrays / std::vectors take a page fault per
// 4 KB on first touch, which parallelises across threads
static void parallel_memcpy(void *dst, const void *src, size_t bytes)
{
    const size_t slice = (size_t)1 << 20;
    if (bytes < 4 * slice) { memcpy(dst, src, bytes); return; }
    const int parts = (int)((bytes + slice - 1) / slice);
    HostPool::get().run(parts, [=](int i) {
        const size_t lo = (size_t)i * slice;
        memcpy((char *)dst + lo, (const char *)src + lo, std::min(slice, bytes - lo));
    });
}

// Device -> caller copy.  Device or pinned destinations take one cudaMemcpyAsync; pageable
// host memory (a std::vector, a numpy array) would make the driver stage through its own small
// buffer at a few GB/s, so it is pipelined through two pinned 32 MB buffers instead: the DMA
// of chunk i overlaps the host memcpy of chunk i-1.
static const size_t TB_STAGE_BYTES = (size_t)32 << 20;

static int copy_out(tb_genus *g, void *dst, const void *src_dev, size_t bytes)
{
    if (bytes == 0) return TB_OK;
    cudaStream_t st = g->stream;
    cudaPointerAttributes attr;
    const bool known = cudaPointerGetAttributes(&attr, dst) == cudaSuccess;
    cudaGetLastError();
    if (known && attr.type != cudaMemoryTypeUnregistered)
    {
        TB_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDefault, st));
        return TB_OK;                                  // caller synchronises
    }
    for (int i = 0; i < 2; ++i)
        if (!g->h_stage[i])
        {
            TB_CUDA(cudaMallocHost((void **)&g->h_stage[i], TB_STAGE_BYTES));
            TB_CUDA(cudaEventCreateWithFlags(&g->ev_stage[i], cudaEventDisableTiming));
        }
    const size_t chunks = (bytes + TB_STAGE_BYTES - 1) / TB_STAGE_BYTES;
    for (size_t c = 0; c <= chunks; ++c)
    {
        if (c < chunks)
        {
            const size_t off = c * TB_STAGE_BYTES, len = std::min(TB_STAGE_BYTES, bytes - off);
            TB_CUDA(cudaMemcpyAsync(g->h_stage[c & 1], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, st));
            TB_CUDA(cudaEventRecord(g->ev_stage[c & 1], st));
        }
        if (c > 0)
        {
            const size_t off = (c - 1) * TB_STAGE_BYTES, len = std::min(TB_STAGE_BYTES, bytes - off);
            TB_CUDA(cudaEventSynchronize(g->ev_stage[(c - 1) & 1]));
            parallel_memcpy((char *)dst + off, g->h_stage[(c - 1) & 1], len);
        }
    }
    return TB_OK;
}

// ---------------------------------------------------------------------------
// Narrow read-back.  A T_p at the bench size is 3.6 GB of int32 CSR; PCIe moves it at
// ~54 GB/s, three times longer than the GPU needs to compute it.  When every column index
// fits 16 bits and every |entry| fits 8 (known from the row kernel's max-multiplicity word),
// the row kernel writes indices/entries as 3 bytes per entry, the DMA moves those into pinned
// staging, and host threads widen them into the caller's int32 arrays with streaming stores.
// A background thread per genus (Copier, below) keeps two chunk DMAs in flight while the
// previous chunk is widened; tb_hecke_copy_wait observes completion.
// ---------------------------------------------------------------------------
static size_t narrow_chunk_entries()
{
    static const size_t n = getenv("TBIRCH_NARROW_CHUNK") ? std::max<size_t>(1 << 16, (size_t)atoll(getenv("TBIRCH_NARROW_CHUNK"))) : ((size_t)16 << 20);
    return n;   // entries per chunk (48 MB of staging by default)
}
#define TB_NARROW_CHUNK (narrow_chunk_entries())

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void widen_u16_avx2(const unsigned short *src, int32_t *dst, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8)
        _mm256_stream_si256((__m256i *)(dst + i), _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx2"))) static void widen_i8_avx2(const signed char *src, int32_t *dst, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8)
        _mm256_stream_si256((__m256i *)(dst + i), _mm256_cvtepi8_epi32(_mm_loadl_epi64((const __m128i *)(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
// AVX-512: one streaming store per 64-byte line
__attribute__((target("avx512f"))) static void widen_u16_avx512(const unsigned short *src, int32_t *dst, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16)
        _mm512_stream_si512((__m512i *)(dst + i), _mm512_cvtepu16_epi32(_mm256_loadu_si256((const __m256i *)(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx512f"))) static void widen_i8_avx512(const signed char *src, int32_t *dst, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16)
        _mm512_stream_si512((__m512i *)(dst + i), _mm512_cvtepi8_epi32(_mm_loadu_si128((const __m128i *)(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
static int widen_isa()
{
    static const int isa = getenv("TBIRCH_NO_AVX512") == nullptr && __builtin_cpu_supports("avx512f") ? 2
                           : __builtin_cpu_supports("avx2") ? 1 : 0;
    return isa;
}
#else
static int widen_isa() { return 0; }      // other hosts (aarch64 Grace): the scalar loops below, auto-vectorised
#endif
static void widen_u16(const unsigned short *src, int32_t *dst, size_t n)
{
#if defined(__x86_64__)
    if (widen_isa() == 2) { widen_u16_avx512(src, dst, n); return; }
    if (widen_isa() == 1) { widen_u16_avx2(src, dst, n); return; }
#endif
    for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}
static void widen_i8(const signed char *src, int32_t *dst, size_t n)
{
#if defined(__x86_64__)
    if (widen_isa() == 2) { widen_i8_avx512(src, dst, n); return; }
    if (widen_isa() == 1) { widen_i8_avx2(src, dst, n); return; }
#endif
    for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}

struct WidenJob {
    const unsigned short *idx16;
    const signed char *dat8;
    int32_t *indices, *data;
    size_t n;
};

static void widen_run(const WidenJob &job)
{
    const WidenJob *j = &job;
    const size_t slice = ((size_t)1 << 18);
    const int parts = (int)((j->n + slice - 1) / slice);
    HostPool::get().run(parts, [j, slice](int i) {
        const size_t lo = (size_t)i * slice, len = std::min(slice, j->n - lo);
        if (j->indices) widen_u16(j->idx16 + lo, j->indices + lo, len);
        if (j->data) widen_i8(j->dat8 + lo, j->data + lo, len);
    });
}

static bool host_pointer(const void *p)
{
    if (!p) return true;
    cudaPointerAttributes attr;
    const bool known = cudaPointerGetAttributes(&attr, p) == cudaSuccess;
    cudaGetLastError();
    return !known || attr.type == cudaMemoryTypeUnregistered || attr.type == cudaMemoryTypeHost;
}

// One background thread per genus drives the narrow read-back: it keeps two chunk DMAs in
// flight on its own stream (three pinned staging slots) and widens each chunk as its DMA
// lands, so the copy engine and the host threads work concurrently and the caller's thread
// is free to launch the next T_p.  (Host functions enqueued in CUDA streams serialised the
// DMA with the widening; a plain thread that waits on events does not.)
struct CopyRequest {
    tb_hecke *h;
    const unsigned short *d_idx;
    const signed char *d_dat;
    int32_t *indices, *data;
    size_t nnz;
};

struct Copier {
    static const int SLOTS = 3;
    int device = 0;
    std::thread th;
    std::mutex mx;
    std::condition_variable cv;
    std::deque<CopyRequest> q;
    bool stop = false;
    cudaStream_t stream = nullptr;
    char *stage[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[SLOTS] = {nullptr, nullptr, nullptr};

    struct Chunk { CopyRequest req; size_t off, len; int slot; bool last; };

    void body()
    {
        cudaSetDevice(device);
        ReadbackPolicy &policy = readback_policy();
        std::deque<Chunk> flight;
        CopyRequest cur;
        size_t cur_off = 0;
        bool have = false;
        uint64_t issued = 0;
        // one busy period = from the first DMA issued on an idle copier to the last chunk widened
        std::chrono::steady_clock::time_point busy_t0;
        double busy_bytes = 0;
        bool busy = false;
        for (;;)
        {
            if (!busy && flight.empty()) { busy_t0 = std::chrono::steady_clock::now(); busy_bytes = 0; }
            while (flight.size() < 2 && fill_one(flight, cur, cur_off, have, issued)) {}
            if (flight.empty())
            {
                if (busy)
                {
                    policy.record(ReadbackPolicy::NARROW, busy_bytes,
                                  std::chrono::duration<double>(std::chrono::steady_clock::now() - busy_t0).count());
                    busy = false;
                }
                std::unique_lock<std::mutex> lk(mx);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (stop && q.empty()) return;
                continue;
            }
            busy = true;
            const Chunk c = flight.front();
            flight.pop_front();
            const cudaError_t e = cudaEventSynchronize(ev[c.slot]);
            if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(c.req.h->copy_mx); c.req.h->copy_error = cudaGetErrorString(e); }
            busy_bytes += (double)c.len * 4.0 * ((c.req.indices ? 1 : 0) + (c.req.data ? 1 : 0));
            // issue the next DMA before widening this chunk: its slot (issue order modulo 3) is neither
            // this chunk's nor the one still in flight, and its previous user was widened an iteration ago
            fill_one(flight, cur, cur_off, have, issued);
            const WidenJob job{(const unsigned short *)stage[c.slot], (const signed char *)(stage[c.slot] + TB_NARROW_CHUNK * 2),
                               c.req.indices ? c.req.indices + c.off : nullptr, c.req.data ? c.req.data + c.off : nullptr, c.len};
            widen_run(job);
            if (c.last)
            {
                std::lock_guard<std::mutex> lk(c.req.h->copy_mx);
                --c.req.h->copy_pending;
                c.req.h->copy_cv.notify_all();
            }
        }
    }
    bool fill_one(std::deque<Chunk> &flight, CopyRequest &cur, size_t &cur_off, bool &have, uint64_t &issued)
    {
        if (!have)
        {
            std::lock_guard<std::mutex> lk(mx);
            if (q.empty()) return false;
            cur = q.front();
            q.pop_front();
            cur_off = 0;
            have = true;
        }
        Chunk c;
        c.req = cur;
        c.off = cur_off;
        c.len = std::min(TB_NARROW_CHUNK, cur.nnz - cur_off);
        c.slot = (int)(issued++ % SLOTS);
        cur_off += c.len;
        c.last = cur_off >= cur.nnz;
        if (c.last) have = false;
        unsigned short *s_idx = (unsigned short *)stage[c.slot];
        signed char *s_dat = (signed char *)(stage[c.slot] + TB_NARROW_CHUNK * 2);
        cudaError_t e = cudaSuccess;
        if (cur.indices) e = cudaMemcpyAsync(s_idx, cur.d_idx + c.off, c.len * 2, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess && cur.data) e = cudaMemcpyAsync(s_dat, cur.d_dat + c.off, c.len, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[c.slot], stream);
        if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(cur.h->copy_mx); cur.h->copy_error = cudaGetErrorString(e); }
        flight.push_back(c);
        return true;
    }
};

static void copier_destroy(Copier *c)
{
    if (!c) return;
    {
        std::lock_guard<std::mutex> lk(c->mx);
        c->stop = true;
    }
    c->cv.notify_all();
    if (c->th.joinable()) c->th.join();
    for (int i = 0; i < Copier::SLOTS; ++i)
    {
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// int32 arrays for device consumers (device-pointer copies, dense rows, small matrices) of a result
// that exists in narrow form only.
static int ensure_wide(tb_hecke *h, int64_t k)
{
    if (h->wide_valid) return TB_OK;
    cudaStream_t st = h->g->stream;
    if (h->wide_k.empty())
    {
        h->wide_k.assign(h->C, 0);
        TB_CUDA(h->d_data.reserve((size_t)std::max<int64_t>(h->extent, 1), st));
        TB_CUDA(h->d_indices.reserve((size_t)std::max<int64_t>(h->extent, 1), st));
    }
    if (h->wide_k[k]) return TB_OK;
    const size_t n = (size_t)h->nnz[k];
    if (n)
    {
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
        widen_device_kernel<<<blocks, 256, 0, st>>>(h->d_idx16.ptr + h->nnzbase[k], h->d_dat8.ptr + h->nnzbase[k], n,
                                                    h->d_indices.ptr + h->nnzbase[k], h->d_data.ptr + h->nnzbase[k]);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
        TB_CUDA(cudaStreamSynchronize(st));
    }
    h->wide_k[k] = 1;
    return TB_OK;
}

static bool narrow_applies(const tb_hecke *h, int64_t k, const int32_t *data, const int32_t *indices)
{
    // (whether the row kernel wrote the narrow form at all was ReadbackPolicy's call at compute time)
    return h->d_idx16.ptr && h->maxmult <= 127 && h->nnz[k] >= h->g->narrow_min_nnz &&
           (data || indices) && host_pointer(data) && host_pointer(indices);
}

// data / indices of conductor k -> host arrays through the 3-byte format.  Returns at once;
// tb_hecke_copy_wait (or the synchronous tb_hecke_copy_sparse) observes completion.
static int narrow_copy(tb_hecke *h, int64_t k, int32_t *data, int32_t *indices)
{
    tb_genus *g = h->g;
    if (!g->copier)
    {
        std::unique_ptr<Copier> c(new Copier);
        TB_CUDA(cudaGetDevice(&c->device));
        TB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (int i = 0; i < Copier::SLOTS; ++i)
        {
            TB_CUDA(cudaMallocHost((void **)&c->stage[i], TB_NARROW_CHUNK * 3));
            TB_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
        }
        Copier *raw = c.release();
        raw->th = std::thread([raw] { raw->body(); });
        g->copier = raw;
    }
    CopyRequest r{h, h->d_idx16.ptr + h->nnzbase[k], h->d_dat8.ptr + h->nnzbase[k], indices, data, (size_t)h->nnz[k]};
    {
        std::lock_guard<std::mutex> lk(h->copy_mx);
        ++h->copy_pending;
    }
    {
        std::lock_guard<std::mutex> lk(g->copier->mx);
        g->copier->q.push_back(r);
    }
    g->copier->cv.notify_all();
    return TB_OK;
}

extern "C" int tb_hecke_copy_wait(tb_hecke *h)
{
    if (!h) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_copy_wait: null argument");
    {
        std::unique_lock<std::mutex> lk(h->copy_mx);
        h->copy_cv.wait(lk, [&] { return h->copy_pending == 0; });
    }
    if (h->copy_ev_armed)
    {
        TB_CUDA(cudaEventSynchronize(h->copy_ev));
        h->copy_ev_armed = false;
        if (h->plain_bytes > 0)
        {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->plain_ev0, h->plain_ev1) == cudaSuccess)
                readback_policy().record(ReadbackPolicy::PLAIN, h->plain_bytes, ms * 1e-3);
            cudaGetLastError();
            h->plain_bytes = 0;
        }
    }
    std::lock_guard<std::mutex> lk(h->copy_mx);
    if (!h->copy_error.empty())
    {
        const std::string msg = "libtbirch: read-back failed: " + h->copy_error;
        h->copy_error.clear();
        return fail(TB_ERR_RUNTIME, msg);
    }
    return TB_OK;
}

extern "C" int tb_hecke_copy_sparse_async(tb_hecke *h, int64_t k, int32_t *data, int32_t *indices, int32_t *indptr,
                                          void *cuda_stream)
{
    if (!h || k < 0 || k >= h->C) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_copy_sparse_async: bad conductor index");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t nnz = (size_t)h->nnz[k];
    if (nnz && narrow_applies(h, k, data, indices))
    {
        if (int rc = narrow_copy(h, k, data, indices)) return rc;
    }
    else
    {
        if (int rc = ensure_wide(h, k)) return rc;
        // large copies into host memory are timed for ReadbackPolicy: first copy of this T_p .. last one
        const bool timed = (int64_t)nnz >= h->g->narrow_min_nnz && (data || indices) && host_pointer(data) && host_pointer(indices);
        if (timed && !h->plain_ev0)
        {
            TB_CUDA(cudaEventCreate(&h->plain_ev0));
            TB_CUDA(cudaEventCreate(&h->plain_ev1));
        }
        if (timed && h->plain_bytes == 0) TB_CUDA(cudaEventRecord(h->plain_ev0, st));
        if (nnz && data) TB_CUDA(cudaMemcpyAsync(data, h->d_data.ptr + h->nnzbase[k], nnz * sizeof(int), cudaMemcpyDefault, st));
        if (nnz && indices) TB_CUDA(cudaMemcpyAsync(indices, h->d_indices.ptr + h->nnzbase[k], nnz * sizeof(int), cudaMemcpyDefault, st));
        if (timed)
        {
            TB_CUDA(cudaEventRecord(h->plain_ev1, st));
            h->plain_bytes += (double)nnz * 4.0 * ((data ? 1 : 0) + (indices ? 1 : 0));
        }
    }
    if (indptr) TB_CUDA(cudaMemcpyAsync(indptr, h->d_indptr.ptr + h->rowbase[k], ((size_t)h->rows[k] + 1) * sizeof(int), cudaMemcpyDefault, st));
    if (!h->copy_ev) TB_CUDA(cudaEventCreateWithFlags(&h->copy_ev, cudaEventDisableTiming));
    TB_CUDA(cudaEventRecord(h->copy_ev, st));
    h->copy_ev_armed = true;
    return TB_OK;
}

extern "C" int tb_hecke_copy_sparse(tb_hecke *h, int64_t k, int32_t *data, int32_t *indices, int32_t *indptr)
{
    if (!h || k < 0 || k >= h->C) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_copy_sparse: bad conductor index");
    cudaStream_t st = h->g->stream;
    const size_t nnz = (size_t)h->nnz[k];
    if (nnz && narrow_applies(h, k, data, indices))
    {
        if (int rc = narrow_copy(h, k, data, indices)) return rc;
    }
    else
    {
        if (int rc = ensure_wide(h, k)) return rc;
        const bool timed = (int64_t)nnz >= h->g->narrow_min_nnz && (data || indices) && host_pointer(data) && host_pointer(indices);
        const auto t0 = std::chrono::steady_clock::now();
        if (nnz && data) if (int rc = copy_out(h->g, data, h->d_data.ptr + h->nnzbase[k], nnz * sizeof(int))) return rc;
        if (nnz && indices) if (int rc = copy_out(h->g, indices, h->d_indices.ptr + h->nnzbase[k], nnz * sizeof(int))) return rc;
        if (timed)
        {
            TB_CUDA(cudaStreamSynchronize(st));
            readback_policy().record(ReadbackPolicy::PLAIN, (double)nnz * 4.0 * ((data ? 1 : 0) + (indices ? 1 : 0)),
                                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
    }
    if (indptr) if (int rc = copy_out(h->g, indptr, h->d_indptr.ptr + h->rowbase[k], ((size_t)h->rows[k] + 1) * sizeof(int))) return rc;
    TB_CUDA(cudaStreamSynchronize(st));
    return tb_hecke_copy_wait(h);
}

// Every conductor at once into three concatenated host (or device) arrays: data_all / indices_all hold the
// nnz[k] entries of conductor 0, 1, ... back to back, indptr_all the (rows[k] + 1)-blocks.  One packing
// kernel and three copies instead of 3 * 2^K small ones; large results use the per-conductor narrow
// read-back into the slices instead.
extern "C" int tb_hecke_copy_sparse_all(tb_hecke *h, int32_t *data_all, int32_t *indices_all, int32_t *indptr_all)
{
    if (!h) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_copy_sparse_all: null argument");
    if ((data_all == nullptr) != (indices_all == nullptr))
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_copy_sparse_all: pass both data_all and indices_all, or neither");
    tb_genus *g = h->g;
    cudaStream_t st = g->stream;
    const int C = h->C;
    std::vector<i64> meta(3 * (size_t)C);
    i64 total = 0;
    int total_ptr = 0;
    bool any_narrow = false;
    for (int k = 0; k < C; ++k)
    {
        meta[k] = h->nnzbase[k]; meta[C + k] = h->nnz[k]; meta[2 * C + k] = total;
        if (h->nnz[k] && data_all && indices_all && narrow_applies(h, k, data_all + total, indices_all + total)) any_narrow = true;
        total += h->nnz[k];
        total_ptr += h->rows[k] + 1;
    }
    if (indptr_all) if (int rc = copy_out(g, indptr_all, h->d_indptr.ptr, (size_t)total_ptr * sizeof(int))) return rc;
    if (total && data_all && indices_all)
    {
        if (any_narrow)
        {
            i64 off = 0;
            for (int k = 0; k < C; ++k)
            {
                if (int rc = tb_hecke_copy_sparse_async(h, k, data_all + off, indices_all + off, nullptr, st)) return rc;
                off += h->nnz[k];
            }
        }
        else
        {
            PoolBuf<i64> d_meta;
            PoolBuf<int> d_pack;
            TB_CUDA(d_meta.reserve(meta.size(), st));
            TB_CUDA(d_pack.reserve(2 * (size_t)total, st));
            TB_CUDA(cudaMemcpyAsync(d_meta.ptr, meta.data(), meta.size() * sizeof(i64), cudaMemcpyHostToDevice, st));
            const bool wide = h->wide_valid;
            const unsigned bx = (unsigned)std::min<i64>((total / C + 255) / 256 + 1, 148 * 8);
            pack_all_kernel<<<dim3(bx, C), 256, 0, st>>>(wide ? h->d_data.ptr : nullptr, wide ? h->d_indices.ptr : nullptr,
                                                         h->d_idx16.ptr, h->d_dat8.ptr, d_meta.ptr, d_meta.ptr + C,
                                                         d_meta.ptr + 2 * C, d_pack.ptr, d_pack.ptr + total);
            ++g_launches;
            TB_CUDA(cudaGetLastError());
            int rc = copy_out(g, data_all, d_pack.ptr, (size_t)total * sizeof(int));
            if (!rc) rc = copy_out(g, indices_all, d_pack.ptr + total, (size_t)total * sizeof(int));
            cudaStreamSynchronize(st);
            d_meta.release();
            d_pack.release();
            if (rc) return rc;
        }
    }
    TB_CUDA(cudaStreamSynchronize(st));
    return tb_hecke_copy_wait(h);
}

/* Host-side widening throughput of this machine (no GPU involved): widens `entries` packed
 * entries with the library's thread pool, returns output GB/s, the pool size and whether the
 * AVX2 streaming-store path is in use. */
extern "C" int tb_host_widen_probe(int64_t entries, double *out_gbps, int *threads, int *avx2)
{
    if (entries <= 0) return fail(TB_ERR_INVALID_ARGUMENT, "tb_host_widen_probe: bad size");
    std::vector<unsigned short> idx((size_t)entries);
    std::vector<signed char> dat((size_t)entries);
    for (int64_t i = 0; i < entries; ++i) { idx[i] = (unsigned short)(i * 7919); dat[i] = (signed char)(i * 37 - 3); }
    // deliberately misaligned destinations: the streaming-store loops must handle any alignment
    int32_t *oi_raw = (int32_t *)aligned_alloc(64, (size_t)entries * 4 + 64), *od_raw = (int32_t *)aligned_alloc(64, (size_t)entries * 4 + 64);
    if (!oi_raw || !od_raw) return fail(TB_ERR_RUNTIME, "tb_host_widen_probe: out of memory");
    int32_t *oi = oi_raw + 1, *od = od_raw + 3;
    memset(oi, 0, (size_t)entries * 4);
    memset(od, 0, (size_t)entries * 4);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep)
    {
        const WidenJob job{idx.data(), dat.data(), oi, od, (size_t)entries};
        const auto t0 = std::chrono::steady_clock::now();
        widen_run(job);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::max(best, (double)entries * 8.0 / sec / 1e9);
    }
    bool ok = true;
    for (int64_t i = 0; i < entries && ok; ++i) ok = oi[i] == (int32_t)idx[i] && od[i] == (int32_t)dat[i];
    free(oi_raw);
    free(od_raw);
    if (!ok) return fail(TB_ERR_RUNTIME, "tb_host_widen_probe: wrong result");
    if (out_gbps) *out_gbps = best;
    if (threads) *threads = HostPool::get().size();
    if (avx2) *avx2 = widen_isa();              // 0 scalar, 1 AVX2, 2 AVX-512
    return TB_OK;
}

/* Bytes that cross PCIe for data + indices of conductor k on a host read-back (diagnostic). */
extern "C" int64_t tb_hecke_transfer_bytes(const tb_hecke *h, int64_t k)
{
    if (!h || k < 0 || k >= h->C) return 0;
    const bool narrow = h->d_idx16.ptr && h->maxmult <= 127 && h->nnz[k] >= h->g->narrow_min_nnz;
    return h->nnz[k] * (narrow ? 3 : 8) + ((int64_t)h->rows[k] + 1) * 4;
}

/* ReadbackPolicy's current view on this device (diagnostic for bench.py): running delivered GB/s of
 * int32 CSR per format, how many large read-backs fed each, and the format in use (0 plain, 1 narrow). */
extern "C" int tb_readback_stats(double *plain_gbps, double *narrow_gbps, int *plain_samples, int *narrow_samples, int *current)
{
    ReadbackPolicy &pol = readback_policy();
    std::lock_guard<std::mutex> lk(pol.mx);
    if (plain_gbps) *plain_gbps = pol.rate[0];
    if (narrow_gbps) *narrow_gbps = pol.rate[1];
    if (plain_samples) *plain_samples = pol.samples[0];
    if (narrow_samples) *narrow_samples = pol.samples[1];
    if (current) *current = pol.current;
    return TB_OK;
}

// ---------------------------------------------------------------------------
// Final integer gather of the Hecke rows: the one exchange of the multi-GPU path.
// Each rank packs ALL conductors of its tb_hecke into one contiguous device buffer in the
// 3-byte form (uint16 column, int8 entry; int32 pairs when a column or an entry does not fit),
// the caller moves the buffers rank -> root with grouped send/recv over NVLink
// (ternary_birch_b200/shard.py: torch.distributed is the transport, nothing is padded), and the
// root widens each buffer into the reference's int32 CSR with tb_unpack_rows.
// Layout per conductor k, each section padded to 16 bytes:
//     columns[nnz_k] | entries[nnz_k] | indptr[rows_k + 1] (int32)
// ---------------------------------------------------------------------------
static size_t pad16(size_t x) { return (x + 15) & ~(size_t)15; }

static bool pack_narrow_ok(const tb_hecke *h)
{
    int64_t max_dim = 0;
    for (int k = 0; k < h->C; ++k) max_dim = std::max<int64_t>(max_dim, h->g->dims[k]);
    return max_dim <= 65536 && h->maxmult <= 127;
}

/* Bytes tb_hecke_pack writes for `h`, and through *entry_bytes the form: 3 (uint16 + int8) or 8 (int32 + int32). */
extern "C" int64_t tb_hecke_packed_bytes(const tb_hecke *h, int *entry_bytes)
{
    if (!h) return 0;
    const bool narrow = pack_narrow_ok(h);
    if (entry_bytes) *entry_bytes = narrow ? 3 : 8;
    size_t total = 0;
    for (int k = 0; k < h->C; ++k)
    {
        const size_t n = (size_t)h->nnz[k];
        total += pad16(n * (narrow ? 2 : 4)) + pad16(n * (narrow ? 1 : 4)) + pad16(((size_t)h->rows[k] + 1) * 4);
    }
    return (int64_t)total;
}

extern "C" int tb_hecke_pack(tb_hecke *h, void *dst_device, void *cuda_stream)
{
    if (!h || !dst_device) return fail(TB_ERR_INVALID_ARGUMENT, "tb_hecke_pack: null argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool narrow = pack_narrow_ok(h);
    const bool have_narrow = h->d_idx16.ptr != nullptr;
    char *out = (char *)dst_device;
    for (int k = 0; k < h->C; ++k)
    {
        const size_t n = (size_t)h->nnz[k];
        char *cols = out, *vals = cols + pad16(n * (narrow ? 2 : 4)), *ptr = vals + pad16(n * (narrow ? 1 : 4));
        if (n && narrow && have_narrow)
        {
            TB_CUDA(cudaMemcpyAsync(cols, h->d_idx16.ptr + h->nnzbase[k], n * 2, cudaMemcpyDeviceToDevice, st));
            TB_CUDA(cudaMemcpyAsync(vals, h->d_dat8.ptr + h->nnzbase[k], n, cudaMemcpyDeviceToDevice, st));
        }
        else if (n && narrow)
        {
            const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
            narrow_device_kernel<<<blocks, 256, 0, st>>>(h->d_indices.ptr + h->nnzbase[k], h->d_data.ptr + h->nnzbase[k], n,
                                                        (unsigned short *)cols, (signed char *)vals);
            ++g_launches;
            TB_CUDA(cudaGetLastError());
        }
        else if (n)
        {
            if (int rc = ensure_wide(h, k)) return rc;
            TB_CUDA(cudaMemcpyAsync(cols, h->d_indices.ptr + h->nnzbase[k], n * 4, cudaMemcpyDeviceToDevice, st));
            TB_CUDA(cudaMemcpyAsync(vals, h->d_data.ptr + h->nnzbase[k], n * 4, cudaMemcpyDeviceToDevice, st));
        }
        TB_CUDA(cudaMemcpyAsync(ptr, h->d_indptr.ptr + h->rowbase[k], ((size_t)h->rows[k] + 1) * 4, cudaMemcpyDeviceToDevice, st));
        out = ptr + pad16(((size_t)h->rows[k] + 1) * 4);
    }
    return TB_OK;
}

/* Root side: one rank's packed buffer (device) -> int32 CSR (device), conductors back to back:
 * data_all / indices_all hold sum nnz[k] entries, indptr_all sum (rows[k] + 1). */
extern "C" int tb_unpack_rows(const void *packed_device, int entry_bytes, int64_t C, const int64_t *nnz, const int64_t *rows,
                              int32_t *data_all, int32_t *indices_all, int32_t *indptr_all, void *cuda_stream)
{
    if (!packed_device || !nnz || !rows || (entry_bytes != 3 && entry_bytes != 8))
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_unpack_rows: bad argument");
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool narrow = entry_bytes == 3;
    const char *in = (const char *)packed_device;
    size_t e = 0, r = 0;
    for (int64_t k = 0; k < C; ++k)
    {
        const size_t n = (size_t)nnz[k], np1 = (size_t)rows[k] + 1;
        const char *cols = in, *vals = cols + pad16(n * (narrow ? 2 : 4)), *ptr = vals + pad16(n * (narrow ? 1 : 4));
        if (n && narrow)
        {
            const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
            widen_device_kernel<<<blocks, 256, 0, st>>>((const unsigned short *)cols, (const signed char *)vals, n,
                                                       indices_all + e, data_all + e);
            ++g_launches;
            TB_CUDA(cudaGetLastError());
        }
        else if (n)
        {
            TB_CUDA(cudaMemcpyAsync(indices_all + e, cols, n * 4, cudaMemcpyDeviceToDevice, st));
            TB_CUDA(cudaMemcpyAsync(data_all + e, vals, n * 4, cudaMemcpyDeviceToDevice, st));
        }
        TB_CUDA(cudaMemcpyAsync(indptr_all + r, ptr, np1 * 4, cudaMemcpyDeviceToDevice, st));
        e += n; r += np1;
        in = ptr + pad16(np1 * 4);
    }
    return TB_OK;
}

