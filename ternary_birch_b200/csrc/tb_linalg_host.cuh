// tb_linalg_host.cuh -- host driver of the modular eliminations (tb_linalg.cuh): part of tbirch.cu
// (included at its end; not a standalone header).  C ABI: tb_modkernel_*.
#pragma once

// ---------------------------------------------------------------------------
// Modular kernels for the rational eigenvector search (tb_linalg.cuh)
// ---------------------------------------------------------------------------
struct tb_modkernel {
    int64_t rows = 0, cols = 0, dim = 0, rank = 0;
    uint32_t P = 0;
    float ms = 0;
    std::vector<int64_t> free_cols;
    std::vector<uint32_t> basis;   // [dim][cols]
};

template <class T>
struct LaBuf {
    // stream-ordered pool allocations (default stream): a search runs hundreds of small eliminations, and
    // cudaMalloc / cudaFree would cost more than they do
    T *ptr = nullptr;
    ~LaBuf() { if (ptr) cudaFreeAsync(ptr, 0); }
    cudaError_t alloc(size_t n, bool zero)
    {
        cudaError_t e = cudaMallocAsync((void **)&ptr, std::max<size_t>(n, 1) * sizeof(T), 0);
        if (e != cudaSuccess) { ptr = nullptr; return e; }
        if (zero) e = cudaMemsetAsync(ptr, 0, std::max<size_t>(n, 1) * sizeof(T), 0);
        return e;
    }
};

static void la_keep_pool()
{
    static std::atomic<uint64_t> done_mask(0);          // per device: a process may drive several GPUs
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    const uint64_t bit = 1ull << (dev & 63);
    if (done_mask.load() & bit) return;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
    {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_mask.fetch_or(bit);
}

static bool la_is_prime(uint32_t n)
{
    if (n < 2) return false;
    for (uint32_t d = 2; (uint64_t)d * d <= n; ++d)
        if (n % d == 0) return false;
    return true;
}

// `fill` writes the residues into c.M on stream 0; then eliminate and read the kernel off.
template <class Fill>
static int la_solve(int64_t rows, int64_t cols, uint32_t P, Fill fill, tb_modkernel **out)
{
    if (!out) return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel: null argument");
    *out = nullptr;
    if (rows < 0 || cols < 0 || rows > 0x3fffffff || cols > 0x3fffffff)
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel: bad shape");
    if (P < 3 || P >= (1u << 29) || !la_is_prime(P))
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel: P must be an odd prime below 2^29");
    if (int rc = require_device()) return rc;
    la_keep_pool();
    std::unique_ptr<tb_modkernel> k(new tb_modkernel);
    k->rows = rows; k->cols = cols; k->P = P;
    if (cols == 0) { *out = k.release(); return TB_OK; }

    LaCtx c;
    memset(&c, 0, sizeof(c));
    c.rows = (int)rows; c.cols = (int)cols;
    c.ld = (int)((cols + TB_LA_TILE_COLS - 1) / TB_LA_TILE_COLS * TB_LA_TILE_COLS);
    c.P = P;
    {
        u32 inv = P;                                   // Newton: P * inv = 1 mod 2^32
        for (int i = 0; i < 5; ++i) inv *= 2u - P * inv;
        c.Pneg_inv = 0u - inv;
        c.one = (u32)(((u64)1 << 32) % P);
        c.R2 = (u32)(((u64)c.one * c.one) % P);
        c.magic = (u32)(((u64)1 << 32) / P);
    }
    const int tiles = c.ld / TB_LA_TILE_COLS;
    const int maxpiv = (int)std::min(rows, cols);
    LaBuf<u32> M, L, R, lp, inv;
    LaBuf<int> row_pivot, piv_row, piv_col, rank, panel_piv, cand, tile_free, free_col;
    TB_CUDA(M.alloc((size_t)std::max<int64_t>(rows, 1) * c.ld, true));
    TB_CUDA(L.alloc((size_t)std::max<int64_t>(rows, 1) * TB_LA_PANEL, true));
    TB_CUDA(R.alloc((size_t)TB_LA_PANEL * c.ld, true));
    TB_CUDA(lp.alloc(TB_LA_PANEL * TB_LA_PANEL, true));
    TB_CUDA(inv.alloc(TB_LA_PANEL, true));
    TB_CUDA(row_pivot.alloc(rows, false));
    TB_CUDA(cudaMemsetAsync(row_pivot.ptr, 0xff, std::max<size_t>(rows, 1) * sizeof(int), 0));
    TB_CUDA(piv_row.alloc(maxpiv, true));
    TB_CUDA(piv_col.alloc(maxpiv, true));
    TB_CUDA(rank.alloc(1, true));
    TB_CUDA(panel_piv.alloc(TB_LA_PANEL, true));
    TB_CUDA(cand.alloc(TB_LA_PANEL + 1, true));
    TB_CUDA(tile_free.alloc(tiles, true));
    c.M = M.ptr; c.L = L.ptr; c.R = R.ptr; c.row_pivot = row_pivot.ptr; c.piv_row = piv_row.ptr;
    c.piv_col = piv_col.ptr; c.rank = rank.ptr; c.panel_piv = panel_piv.ptr; c.panel_lp = lp.ptr; c.panel_inv = inv.ptr;

    if (int rc = fill(c)) return rc;

    cudaEvent_t e0, e1;
    TB_CUDA(cudaEventCreate(&e0));
    TB_CUDA(cudaEventCreate(&e1));
    TB_CUDA(cudaEventRecord(e0, 0));
    if (rows > 0)
    {
        int dev = 0, sms = 0, per_sm = 0;
        TB_CUDA(cudaGetDevice(&dev));
        TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, la_panel_kernel, TB_LA_PANEL_THREADS, 0));
        const int warps_per_block = TB_LA_PANEL_THREADS / 32;
        // co-resident by construction; a few CTAs per SM keep the per-warp row loop short
        int grid = (int)std::min<int64_t>((int64_t)sms * std::min(std::max(per_sm, 1), 1), (rows + warps_per_block - 1) / warps_per_block);
        grid = std::max(1, grid);
        const dim3 ugrid(tiles, (unsigned)((rows + TB_LA_TILE_ROWS - 1) / TB_LA_TILE_ROWS));
        const bool use_mma = getenv("TBIRCH_LA_NO_MMA") == nullptr;     // 8-bit limb tensor-core update (default)
        for (int c0 = 0; c0 < c.cols; c0 += TB_LA_PANEL)
        {
            int *cand_ptr = cand.ptr, *tf_ptr = tile_free.ptr;
            void *args[] = {(void *)&c, (void *)&c0, (void *)&cand_ptr, (void *)&tf_ptr};
            TB_CUDA(cudaLaunchCooperativeKernel((void *)la_panel_kernel, dim3(grid), dim3(TB_LA_PANEL_THREADS), args, 0, 0));
            la_pivot_fix_kernel<<<1, TB_LA_PANEL * 32>>>(c, c0);
            la_pivot_rows_kernel<<<(c.cols + 127) / 128, 128>>>(c, c0);
            if (use_mma) la_update_mma_kernel<<<ugrid, 256>>>(c, c0, tile_free.ptr);
            else la_update_kernel<<<ugrid, 256>>>(c, c0, tile_free.ptr);
            g_launches += 4;
        }
        TB_CUDA(cudaGetLastError());
    }
    TB_CUDA(cudaEventRecord(e1, 0));
    TB_CUDA(cudaEventSynchronize(e1));
    TB_CUDA(cudaEventElapsedTime(&k->ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);

    int r = 0;
    TB_CUDA(cudaMemcpy(&r, rank.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> pcol(std::max(r, 1));
    if (r) TB_CUDA(cudaMemcpy(pcol.data(), piv_col.ptr, (size_t)r * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<char> is_piv(cols, 0);
    for (int i = 0; i < r; ++i) is_piv[pcol[i]] = 1;
    std::vector<int> fcols;
    for (int64_t j = 0; j < cols; ++j)
        if (!is_piv[j]) { fcols.push_back((int)j); k->free_cols.push_back(j); }
    k->rank = r;
    k->dim = (int64_t)fcols.size();
    if ((size_t)k->dim * (size_t)cols > ((size_t)1 << 31))
        return fail(TB_ERR_RUNTIME, "tb_modkernel: kernel too large to return (dimension x columns > 2^31)");
    if (k->dim > 0)
    {
        LaBuf<u32> basis;
        TB_CUDA(basis.alloc((size_t)k->dim * cols, true));
        TB_CUDA(free_col.alloc(fcols.size(), false));
        TB_CUDA(cudaMemcpy(free_col.ptr, fcols.data(), fcols.size() * sizeof(int), cudaMemcpyHostToDevice));
        const dim3 bgrid((unsigned)((std::max(r, 1) + 255) / 256), (unsigned)k->dim);
        la_kernel_basis_kernel<<<bgrid, 256>>>(c, free_col.ptr, (int)k->dim, r, basis.ptr);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
        k->basis.resize((size_t)k->dim * cols);
        TB_CUDA(cudaMemcpy(k->basis.data(), basis.ptr, k->basis.size() * sizeof(u32), cudaMemcpyDeviceToHost));
    }
    *out = k.release();
    return TB_OK;
}

extern "C" int tb_modkernel_csr(const int32_t *data, const int32_t *indices, const int32_t *indptr, int64_t n,
                                int64_t shift, uint32_t P, tb_modkernel **out)
{
    if (n < 0 || !indptr || (n > 0 && indptr[n] > 0 && (!data || !indices)))
        return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel_csr: null argument");
    for (int64_t i = 0; i < n; ++i)
    {
        if (indptr[i] > indptr[i + 1]) return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel_csr: indptr not monotone");
        for (int32_t j = indptr[i]; j < indptr[i + 1]; ++j)
            if (indices[j] < 0 || indices[j] >= n) return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel_csr: column index out of range");
    }
    return la_solve(n, n, P, [&](LaCtx &c) -> int {
        if (n == 0) return TB_OK;
        const size_t nnz = (size_t)indptr[n];
        LaBuf<int> d, ix, ip;
        TB_CUDA(d.alloc(nnz, false));
        TB_CUDA(ix.alloc(nnz, false));
        TB_CUDA(ip.alloc((size_t)n + 1, false));
        if (nnz)
        {
            TB_CUDA(cudaMemcpy(d.ptr, data, nnz * sizeof(int), cudaMemcpyHostToDevice));
            TB_CUDA(cudaMemcpy(ix.ptr, indices, nnz * sizeof(int), cudaMemcpyHostToDevice));
        }
        TB_CUDA(cudaMemcpy(ip.ptr, indptr, ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice));
        la_from_csr_kernel<<<(unsigned)n, 128>>>(c, d.ptr, ix.ptr, ip.ptr, (i64)shift);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
        TB_CUDA(cudaDeviceSynchronize());
        return TB_OK;
    }, out);
}

extern "C" int tb_modkernel_dense(const int64_t *Msrc, int64_t rows, int64_t cols, uint32_t P, tb_modkernel **out)
{
    if (rows > 0 && cols > 0 && !Msrc) return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel_dense: null argument");
    return la_solve(rows, cols, P, [&](LaCtx &c) -> int {
        if (rows == 0 || cols == 0) return TB_OK;
        LaBuf<i64> src;
        TB_CUDA(src.alloc((size_t)rows * cols, false));
        TB_CUDA(cudaMemcpy(src.ptr, Msrc, (size_t)rows * cols * sizeof(i64), cudaMemcpyHostToDevice));
        const size_t cells = (size_t)rows * cols;
        la_from_dense_kernel<<<(unsigned)std::min<size_t>((cells + 255) / 256, 148 * 32), 256>>>(c, src.ptr);
        ++g_launches;
        TB_CUDA(cudaGetLastError());
        TB_CUDA(cudaDeviceSynchronize());
        return TB_OK;
    }, out);
}

extern "C" int64_t tb_modkernel_dim(const tb_modkernel *k) { return k ? k->dim : 0; }
extern "C" int64_t tb_modkernel_cols(const tb_modkernel *k) { return k ? k->cols : 0; }
extern "C" float tb_modkernel_ms(const tb_modkernel *k) { return k ? k->ms : 0.f; }
extern "C" int tb_modkernel_free_cols(const tb_modkernel *k, int64_t *free_cols)
{
    if (!k || (k->dim && !free_cols)) return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel_free_cols: null argument");
    std::copy(k->free_cols.begin(), k->free_cols.end(), free_cols);
    return TB_OK;
}
extern "C" int tb_modkernel_basis(const tb_modkernel *k, uint32_t *basis)
{
    if (!k || (k->dim && !basis)) return fail(TB_ERR_INVALID_ARGUMENT, "tb_modkernel_basis: null argument");
    std::copy(k->basis.begin(), k->basis.end(), basis);
    return TB_OK;
}
extern "C" void tb_modkernel_destroy(tb_modkernel *k) { delete k; }
