// tb_linalg.cuh -- exact linear algebra modulo a word-size prime, for the consumer of T_p:
// the rational eigenvector search of the reference's Python layer
// (ternary_birch.pyx:246-379: eigenvalue candidates, kernels of T_p - e, eigenspace
// splitting).  The reference hands these to Sage (charpoly over GF(q), right_kernel over
// Q); here the right kernel of a matrix modulo a prime P < 2^29 is computed on the GPU by
// panel-blocked Gauss-Jordan elimination and lifted to Q on the host (CRT over several P,
// rational reconstruction, exact verification over Z).
//
// Layout: M is rows x ld uint32 (row-major, ld a multiple of 128, residues in [0, P)).
// One panel = TB_LA_PANEL consecutive columns:
//   panel_kernel     a cooperative grid walks the panel's columns: picks a pivot row (any unused row
//                    with a non-zero entry -- exact arithmetic, so any pivot is fine), scales
//                    it, eliminates the column from every other row inside the panel and
//                    records the multipliers L[row][j]
//   pivot_rows_kernel applies the same row operations to the pivot rows' other columns
//                    (a TB_LA_PANEL-step triangular recurrence per column) -> R[j][col]
//   update_kernel    rank-TB_LA_PANEL update of every other row and column:
//                    M[i][c] -= sum_j L[i][j] * R[j][c]  (64-bit accumulation: 32 products
//                    below 2^58 cannot overflow, one reduction per output)
// After the last panel M is in reduced row echelon form; kernel_basis_kernel reads the kernel
// basis off the free columns.
#pragma once

#include "tb_int.cuh"

#include <cooperative_groups.h>

namespace tb {

namespace cg = cooperative_groups;

#define TB_LA_PANEL 32
#define TB_LA_PANEL_THREADS 256
#define TB_LA_TILE_ROWS 64
#define TB_LA_TILE_COLS 128

struct LaCtx {
    u32 *M;          // rows x ld
    u32 *L;          // rows x TB_LA_PANEL multipliers of the current panel
    u32 *R;          // TB_LA_PANEL x ld transformed pivot rows of the current panel
    int *row_pivot;  // [rows] pivot index owned by the row, -1 if none
    int *piv_row;    // [<= min(rows, cols)]
    int *piv_col;
    int *rank;       // device counter
    int *panel_piv;  // [TB_LA_PANEL] pivot row of each panel step, -1 = free column
    u32 *panel_lp;   // [TB_LA_PANEL][TB_LA_PANEL] multipliers of pivot rows at earlier steps
    u32 *panel_inv;  // [TB_LA_PANEL] inverse of each pivot
    int rows, cols, ld;
    u32 P;
    InvDiv dP;
};

__device__ __forceinline__ u32 la_mod(u64 x, const InvDiv &dP)
{
    u64 r;
    udivmod(x, dP, r);
    return (u32)r;
}
__device__ __forceinline__ u32 la_mul(u32 a, u32 b, const InvDiv &dP) { return la_mod((u64)a * b, dP); }
__device__ __forceinline__ u32 la_inv(u32 a, u32 P, const InvDiv &dP)
{
    u32 r = 1, e = P - 2;
    while (e)
    {
        if (e & 1u) r = la_mul(r, a, dP);
        a = la_mul(a, a, dP);
        e >>= 1;
    }
    return r;
}

// dense M from CSR int32 (A - shift * I), residues mod P
__global__ void la_zero_kernel(u32 *M, size_t words)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < words; i += stride) M[i] = 0;
}
__global__ void la_from_csr_kernel(LaCtx c, const int *data, const int *indices, const int *indptr, i64 shift)
{
    const int row = blockIdx.x;
    u32 *Mr = c.M + (size_t)row * c.ld;
    const i64 P = (i64)c.P;
    for (int k = indptr[row] + threadIdx.x; k < indptr[row + 1]; k += blockDim.x)
    {
        i64 v = (i64)data[k] % P;
        if (v < 0) v += P;
        Mr[indices[k]] = (u32)v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && row < c.cols)
    {
        i64 v = ((i64)Mr[row] - shift % P) % P;
        if (v < 0) v += P;
        Mr[row] = (u32)v;
    }
}
__global__ void la_from_dense_kernel(LaCtx c, const i64 *src)
{
    const int row = blockIdx.y;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= c.cols) return;
    const i64 P = (i64)c.P;
    i64 v = src[(size_t)row * c.cols + col] % P;
    if (v < 0) v += P;
    c.M[(size_t)row * c.ld + col] = (u32)v;
}

// Cooperative grid (one CTA per SM at most); global warp gw handles rows gw, gw + GW, ...;
// lane = column inside the panel.  Two grid barriers per pivot step: after the elimination
// pass (so the scaled pivot row may replace the unscaled one everyone read) and after that
// write (the next step eliminates the old pivot row like any other row).
__device__ __forceinline__ void la_search_column(const LaCtx &c, int col, int gw, int GW, int lane, int *cand)
{
    // smallest unused row with a non-zero entry in `col`: lanes test 32 rows at a time
    for (int base = gw * 32; base < c.rows; base += GW * 32)
    {
        const int row = base + lane;
        const bool nz = row < c.rows && c.row_pivot[row] < 0 && c.M[(size_t)row * c.ld + col] != 0;
        const u32 b = __ballot_sync(0xffffffffu, nz);
        if (b)
        {
            if (lane == 0) atomicMin(cand, base + __ffs(b) - 1);
            break;
        }
    }
}

__global__ void __launch_bounds__(TB_LA_PANEL_THREADS) la_panel_kernel(LaCtx c, int c0, int *cand, int *tile_free)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 s_prow[TB_LA_PANEL];
    __shared__ u32 s_inv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp, GW = gridDim.x * wpb;
    const int width = min(TB_LA_PANEL, c.cols - c0);
    const bool live = lane < width;

    // clear this panel's state
    for (int row = gw; row < c.rows; row += GW) c.L[(size_t)row * TB_LA_PANEL + lane] = 0;
    if (blockIdx.x == 0)
    {
        if (threadIdx.x <= TB_LA_PANEL) cand[threadIdx.x] = 0x7fffffff;
        if (threadIdx.x < TB_LA_PANEL) c.panel_piv[threadIdx.x] = -1;
        for (int i = threadIdx.x; i < TB_LA_PANEL * TB_LA_PANEL; i += blockDim.x) c.panel_lp[i] = 0;
    }
    grid.sync();
    la_search_column(c, c0, gw, GW, lane, &cand[0]);
    grid.sync();

    for (int j = 0; j < width; ++j)
    {
        const int piv = *(volatile int *)&cand[j];
        if (piv == 0x7fffffff)                                   // free column
        {
            if (gw == 0 && lane == 0) atomicAdd(&tile_free[(c0 + j) / TB_LA_TILE_COLS], 1);
            if (j + 1 < width) la_search_column(c, c0 + j + 1, gw, GW, lane, &cand[j + 1]);
            grid.sync();
            continue;
        }

        // every CTA derives the scaled pivot row from the (still unscaled) row in M
        if (warp == 0)
        {
            const u32 pv = c.M[(size_t)piv * c.ld + c0 + j];
            const u32 inv = la_inv(pv, c.P, c.dP);
            const u32 x = live ? c.M[(size_t)piv * c.ld + c0 + lane] : 0u;
            s_prow[lane] = la_mul(x, inv, c.dP);
            if (lane == 0) s_inv = inv;
        }
        __syncthreads();
        const u32 pr = s_prow[lane];

        // eliminate column j from every other row inside the panel; note candidates for column j + 1
        int next = 0x7fffffff;
        for (int row = gw; row < c.rows; row += GW)
        {
            if (row == piv) continue;
            u32 *Mr = c.M + (size_t)row * c.ld + c0;
            u32 x = live ? Mr[lane] : 0u;
            const u32 f = __shfl_sync(0xffffffffu, x, j);
            if (f != 0)
            {
                if (lane == j) c.L[(size_t)row * TB_LA_PANEL + j] = f;
                const u32 sub = la_mul(f, pr, c.dP);
                x = x >= sub ? x - sub : x + c.P - sub;
                if (live) Mr[lane] = x;
            }
            if (j + 1 < width && next == 0x7fffffff)
            {
                const u32 nx = __shfl_sync(0xffffffffu, x, j + 1);
                if (nx != 0 && c.row_pivot[row] < 0) next = row;
            }
        }
        if (lane == 0 && next != 0x7fffffff) atomicMin(&cand[j + 1], next);
        grid.sync();

        if (gw == 0)
        {
            // scaled pivot row replaces the unscaled one; the multipliers it received at earlier
            // steps feed the pivot-row recurrence and it takes no part in the rank update
            if (live) c.M[(size_t)piv * c.ld + c0 + lane] = pr;
            const u32 lp = c.L[(size_t)piv * TB_LA_PANEL + lane];
            c.panel_lp[j * TB_LA_PANEL + lane] = lane < j ? lp : 0u;
            c.L[(size_t)piv * TB_LA_PANEL + lane] = 0;
            if (lane == 0)
            {
                const int k = *c.rank;
                c.piv_row[k] = piv;
                c.piv_col[k] = c0 + j;
                c.row_pivot[piv] = k;
                *c.rank = k + 1;
                c.panel_piv[j] = piv;
                c.panel_inv[j] = s_inv;
            }
        }
        grid.sync();
    }
}

// R[j][col] = the panel's pivot rows after the panel's row operations, for the columns
// outside the panel; the pivot rows of M are overwritten with them.
__global__ void la_pivot_rows_kernel(LaCtx c, int c0)
{
    __shared__ u32 s_lp[TB_LA_PANEL * TB_LA_PANEL];
    __shared__ int s_piv[TB_LA_PANEL];
    __shared__ u32 s_inv[TB_LA_PANEL];
    for (int i = threadIdx.x; i < TB_LA_PANEL * TB_LA_PANEL; i += blockDim.x) s_lp[i] = c.panel_lp[i];
    if (threadIdx.x < TB_LA_PANEL)
    {
        s_piv[threadIdx.x] = c.panel_piv[threadIdx.x];
        s_inv[threadIdx.x] = c.panel_inv[threadIdx.x];
    }
    __syncthreads();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= c.cols) return;
    const bool inside = col >= c0 && col < c0 + TB_LA_PANEL;
    u32 xp[TB_LA_PANEL];
#pragma unroll
    for (int k = 0; k < TB_LA_PANEL; ++k)
    {
        const int piv = s_piv[k];
        u32 v = 0;
        if (piv >= 0 && !inside)
        {
            u64 acc = 0;
#pragma unroll
            for (int j = 0; j < k; ++j) acc += (u64)s_lp[k * TB_LA_PANEL + j] * xp[j];
            const u32 sub = la_mod(acc, c.dP);
            u32 x = c.M[(size_t)piv * c.ld + col];
            x = x >= sub ? x - sub : x + c.P - sub;
            v = la_mul(x, s_inv[k], c.dP);
            c.M[(size_t)piv * c.ld + col] = v;
        }
        xp[k] = v;
        c.R[(size_t)k * c.ld + col] = v;
    }
}

// M[i][c] -= sum_j L[i][j] R[j][c] for all rows and all columns outside the panel.
// 256 threads: a 64 x 128 tile, 4 rows x 8 columns per thread.
// Column tiles wholly left of the panel hold pivot columns (unit vectors the update cannot change)
// unless a free column fell into them: tile_free counts those.
__global__ void __launch_bounds__(256) la_update_kernel(LaCtx c, int c0, const int *tile_free)
{
    if ((blockIdx.x + 1) * TB_LA_TILE_COLS <= c0 && tile_free[blockIdx.x] == 0) return;
    __shared__ u32 s_L[TB_LA_TILE_ROWS][TB_LA_PANEL + 1];
    __shared__ __align__(16) u32 s_R[TB_LA_PANEL][TB_LA_TILE_COLS];
    const int col0 = blockIdx.x * TB_LA_TILE_COLS, row0 = blockIdx.y * TB_LA_TILE_ROWS;
    for (int i = threadIdx.x; i < TB_LA_TILE_ROWS * TB_LA_PANEL; i += 256)
    {
        const int r = i / TB_LA_PANEL, j = i % TB_LA_PANEL;
        s_L[r][j] = row0 + r < c.rows ? c.L[(size_t)(row0 + r) * TB_LA_PANEL + j] : 0u;
    }
    for (int i = threadIdx.x; i < TB_LA_PANEL * TB_LA_TILE_COLS / 4; i += 256)
    {
        const int j = i / (TB_LA_TILE_COLS / 4), q = i % (TB_LA_TILE_COLS / 4);
        *reinterpret_cast<uint4 *>(&s_R[j][q * 4]) = *reinterpret_cast<const uint4 *>(c.R + (size_t)j * c.ld + col0 + q * 4);
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    u64 acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0;
#pragma unroll 8
    for (int j = 0; j < TB_LA_PANEL; ++j)
    {
        const uint4 r0 = *reinterpret_cast<const uint4 *>(&s_R[j][tx * 4]);
        const uint4 r1 = *reinterpret_cast<const uint4 *>(&s_R[j][64 + tx * 4]);
        const u32 rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
        {
            const u32 l = s_L[ty * 4 + a][j];
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] += (u64)l * rv[b];
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
        const int row = row0 + ty * 4 + a;
        if (row >= c.rows) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h)
        {
            const int col = col0 + h * 64 + tx * 4;
            if (col >= c.ld) continue;
            uint4 *ptr = reinterpret_cast<uint4 *>(c.M + (size_t)row * c.ld + col);
            uint4 m = *ptr;
            u32 mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int b = 0; b < 4; ++b)
            {
                const int cc = col + b;
                if (cc >= c0 && cc < c0 + TB_LA_PANEL) continue;      // panel columns are final already
                const u32 sub = la_mod(acc[a][h * 4 + b], c.dP);
                mv[b] = mv[b] >= sub ? mv[b] - sub : mv[b] + c.P - sub;
            }
            *ptr = make_uint4(mv[0], mv[1], mv[2], mv[3]);
        }
    }
}

// basis[m][free_col[m]] = 1, basis[m][piv_col[k]] = -M[piv_row[k]][free_col[m]]
__global__ void la_kernel_basis_kernel(LaCtx c, const int *free_col, int dim, int rank, u32 *basis)
{
    const int m = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= dim) return;
    const int fc = free_col[m];
    if (k == 0) basis[(size_t)m * c.cols + fc] = 1;
    if (k >= rank) return;
    const u32 v = c.M[(size_t)c.piv_row[k] * c.ld + fc];
    basis[(size_t)m * c.cols + c.piv_col[k]] = v ? c.P - v : 0u;
}

}  // namespace tb
