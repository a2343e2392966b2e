// tb_linalg.cuh -- exact linear algebra modulo a word-size prime, for the consumer of T_p:
// the rational eigenvector search of the reference's Python layer
// (ternary_birch.pyx:246-379: eigenvalue candidates, kernels of T_p - e, eigenspace
// splitting).  The reference hands these to Sage (charpoly over GF(q), right_kernel over
// Q); here the right kernel of a matrix modulo a prime P < 2^29 is computed on the GPU by
// panel-blocked Gauss-Jordan elimination and lifted to Q on the host (CRT over several P,
// rational reconstruction, exact verification over Z).
//
// Layout: M is rows x ld uint32 (row-major, ld a multiple of 128), Montgomery residues x 2^32 mod P.
// One panel = TB_LA_PANEL consecutive columns:
//   panel_kernel     a cooperative grid walks the panel's columns: picks a pivot row (any unused row
//                    with a non-zero entry -- exact arithmetic, so any pivot is fine),
//                    eliminates the column from every other row inside the panel and
//                    records the multipliers L[row][j]; one grid barrier per column
//   pivot_fix_kernel scales the panel's pivot rows, sorts out their multipliers
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
#define TB_LA_PANEL_THREADS 1024
#define TB_LA_TILE_ROWS 64
#define TB_LA_TILE_COLS 128

struct LaCtx {
    u32 *M;          // rows x ld, Montgomery residues
    u32 *L;          // rows x TB_LA_PANEL multipliers of the current panel
    u32 *R;          // TB_LA_PANEL x ld transformed pivot rows of the current panel
    int *row_pivot;  // [rows] pivot index owned by the row, -1 if none
    int *piv_row;    // [<= min(rows, cols)]
    int *piv_col;
    int *rank;       // device counter
    int *panel_piv;  // [TB_LA_PANEL] pivot row of each panel step, -1 = free column
    u32 *panel_lp;   // [TB_LA_PANEL][TB_LA_PANEL] multipliers a pivot row received before its own step
    u32 *panel_inv;  // [TB_LA_PANEL] inverse of each pivot
    int rows, cols, ld;
    u32 P;
    u32 Pneg_inv;    // -P^-1 mod 2^32
    u32 R2;          // 2^64 mod P
    u32 one;         // 2^32 mod P (Montgomery 1)
    u32 magic;       // floor(2^32 / P)
};

// Montgomery arithmetic modulo P < 2^29 with R = 2^32: residues are kept as x R mod P, a
// product costs IMAD.WIDE + IMAD.LO + IMAD.WIDE + compare/subtract.
__device__ __forceinline__ u32 la_redc(u64 t, const LaCtx &c)      // t < P 2^32  ->  t / R mod P
{
    const u32 m = (u32)t * c.Pneg_inv;
    const u32 r = (u32)((t + (u64)m * c.P) >> 32);
    return r >= c.P ? r - c.P : r;
}
__device__ __forceinline__ u32 la_mul(u32 a, u32 b, const LaCtx &c) { return la_redc((u64)a * b, c); }
__device__ __forceinline__ u32 la_sub(u32 a, u32 b, const LaCtx &c) { return a >= b ? a - b : a + c.P - b; }
// any 64-bit accumulator of Montgomery products -> its Montgomery residue
__device__ __forceinline__ u32 la_fold(u64 acc, const LaCtx &c)
{
    const u32 r = (u32)(acc >> 32) + la_redc((u64)(u32)acc, c);     // < 2^32 + P fits: acc >> 32 < 2^31 here
    const u32 q = __umulhi(r, c.magic);
    const u32 v = r - q * c.P;                                       // < 2 P
    return v >= c.P ? v - c.P : v;
}
__device__ __forceinline__ u32 la_inv(u32 a, const LaCtx &c)        // Fermat, in the Montgomery domain
{
    u32 r = c.one, e = c.P - 2;
    while (e)
    {
        if (e & 1u) r = la_mul(r, a, c);
        a = la_mul(a, a, c);
        e >>= 1;
    }
    return r;
}
__device__ __forceinline__ u32 la_to_mont(i64 v, const LaCtx &c)
{
    v %= (i64)c.P;
    if (v < 0) v += (i64)c.P;
    return la_mul((u32)v, c.R2, c);
}

// dense M from CSR int32 (A - shift * I)
__global__ void la_from_csr_kernel(LaCtx c, const int *data, const int *indices, const int *indptr, i64 shift)
{
    const int row = blockIdx.x;
    u32 *Mr = c.M + (size_t)row * c.ld;
    __shared__ i64 s_diag;
    if (threadIdx.x == 0) s_diag = 0;
    __syncthreads();
    for (int k = indptr[row] + threadIdx.x; k < indptr[row + 1]; k += blockDim.x)
    {
        if (indices[k] == row) s_diag = (i64)data[k];
        else Mr[indices[k]] = la_to_mont((i64)data[k], c);
    }
    __syncthreads();
    if (threadIdx.x == 0 && row < c.cols) Mr[row] = la_to_mont(s_diag % (i64)c.P - shift % (i64)c.P, c);
}
__global__ void la_from_dense_kernel(LaCtx c, const i64 *src)
{
    // grid-stride over all cells: no limit on the number of rows
    const size_t cells = (size_t)c.rows * c.cols, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride)
    {
        const size_t row = i / c.cols, col = i - row * c.cols;
        c.M[row * c.ld + col] = la_to_mont(src[i], c);
    }
}

// Cooperative grid; global warp gw handles rows gw, gw + GW, ...; lane = column inside the
// panel.  ONE grid barrier per pivot step: pivot rows stay unscaled in M for the whole panel
// (every CTA derives the scaled row it eliminates with from the unscaled one, which nobody
// writes during that step), and rows are eliminated uniformly whether or not they were a
// pivot before: u -= u[j] * scaled_pivot_row is the same update on a scaled or unscaled row.
// The final scaling of the panel's pivot rows happens in la_pivot_fix_kernel.
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

        if (warp == 0)
        {
            const u32 x = live ? c.M[(size_t)piv * c.ld + c0 + lane] : 0u;
            const u32 inv = la_inv(__shfl_sync(0xffffffffu, x, j), c);
            s_prow[lane] = la_mul(x, inv, c);
            if (gw == 0 && lane == 0)
            {
                const int k = *c.rank;
                c.piv_row[k] = piv;
                c.piv_col[k] = c0 + j;
                c.row_pivot[piv] = k;
                *c.rank = k + 1;
                c.panel_piv[j] = piv;
                c.panel_inv[j] = inv;
            }
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
                x = la_sub(x, la_mul(f, pr, c), c);
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
    }
}

// After the panel: scale the panel's pivot rows inside the panel, and split the multipliers a
// pivot row collected into those it received before its own step (they drive the pivot-row
// recurrence) and those after it (scaled, they take part in the rank update like any row's).
__global__ void la_pivot_fix_kernel(LaCtx c, int c0)
{
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;     // one warp per panel step
    const int piv = c.panel_piv[k];
    u32 lp = 0;
    if (piv >= 0)
    {
        const u32 inv = c.panel_inv[k];
        const u32 l = c.L[(size_t)piv * TB_LA_PANEL + lane];
        lp = lane < k ? l : 0u;
        c.L[(size_t)piv * TB_LA_PANEL + lane] = lane > k ? la_mul(l, inv, c) : 0u;
        if (c0 + lane < c.cols)
        {
            u32 *m = c.M + (size_t)piv * c.ld + c0 + lane;
            *m = la_mul(*m, inv, c);
        }
    }
    c.panel_lp[k * TB_LA_PANEL + lane] = lp;
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
            const u32 x = la_sub(c.M[(size_t)piv * c.ld + col], la_fold(acc, c), c);
            v = la_mul(x, s_inv[k], c);
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
                mv[b] = la_sub(mv[b], la_fold(acc[a][h * 4 + b], c), c);
            }
            *ptr = make_uint4(mv[0], mv[1], mv[2], mv[3]);
        }
    }
}

// The same update on the tensor cores.  A 29-bit residue is four 8-bit limbs, so the 32-term
// product sum is sum_k 2^(8k) S_k with S_k = sum over limb pairs a + b = k of an 8-bit GEMM:
// 16 `mma.m16n8k32.u8.u8.s32` per 16 x 8 output tile (K = 32 is exactly the panel width), seven
// int32 accumulators (each at most 4 * 32 * 255^2 < 2^23), recombined in the epilogue into the
// 64-bit sum the scalar kernel forms, then the same Montgomery fold.  64 x 128 tile per CTA,
// eight warps of 16 rows x 64 columns.  Limb planes live in shared memory with a 48-byte row
// stride (conflict-free fragment loads).
#define TB_LA_LIMB_STRIDE 48

__device__ __forceinline__ void la_mma_u8(int (&d)[4], const u32 (&a)[4], const u32 (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// four values -> their limb-th bytes packed into one word
__device__ __forceinline__ u32 la_pack_limb(u32 v0, u32 v1, u32 v2, u32 v3, int limb)
{
    const u32 sel = 0x0040u + 0x0011u * (u32)limb;             // byte `limb` of the first and of the second operand
    const u32 x = __byte_perm(v0, v1, sel), y = __byte_perm(v2, v3, sel);
    return __byte_perm(x, y, 0x5410);
}

__global__ void __launch_bounds__(256) la_update_mma_kernel(LaCtx c, int c0, const int *tile_free)
{
    if ((blockIdx.x + 1) * TB_LA_TILE_COLS <= c0 && tile_free[blockIdx.x] == 0) return;
    __shared__ __align__(16) unsigned char s_A[4][TB_LA_TILE_ROWS][TB_LA_LIMB_STRIDE];
    __shared__ __align__(16) unsigned char s_B[4][TB_LA_TILE_COLS][TB_LA_LIMB_STRIDE];
    const int col0 = blockIdx.x * TB_LA_TILE_COLS, row0 = blockIdx.y * TB_LA_TILE_ROWS;
    // L tile: 64 rows x 32 values, four consecutive k per thread-item
    for (int i = threadIdx.x; i < TB_LA_TILE_ROWS * (TB_LA_PANEL / 4); i += 256)
    {
        const int r = i / (TB_LA_PANEL / 4), q = i % (TB_LA_PANEL / 4);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row0 + r < c.rows) v = *reinterpret_cast<const uint4 *>(c.L + (size_t)(row0 + r) * TB_LA_PANEL + q * 4);
#pragma unroll
        for (int limb = 0; limb < 4; ++limb)
            *reinterpret_cast<u32 *>(&s_A[limb][r][q * 4]) = la_pack_limb(v.x, v.y, v.z, v.w, limb);
    }
    // R tile: 32 k x 128 columns, transposed so that k runs along the bytes of a column's row
    for (int i = threadIdx.x; i < TB_LA_TILE_COLS * (TB_LA_PANEL / 4); i += 256)
    {
        const int col = i % TB_LA_TILE_COLS, q = i / TB_LA_TILE_COLS;
        const u32 *src = c.R + (size_t)(q * 4) * c.ld + col0 + col;
        const u32 v0 = src[0], v1 = src[c.ld], v2 = src[2 * (size_t)c.ld], v3 = src[3 * (size_t)c.ld];
#pragma unroll
        for (int limb = 0; limb < 4; ++limb)
            *reinterpret_cast<u32 *>(&s_B[limb][col][q * 4]) = la_pack_limb(v0, v1, v2, v3, limb);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int wrow = (warp & 3) * 16, wcol = (warp >> 2) * 64;
    u32 A[4][4];
#pragma unroll
    for (int limb = 0; limb < 4; ++limb)
    {
        A[limb][0] = *reinterpret_cast<const u32 *>(&s_A[limb][wrow + g][tq * 4]);
        A[limb][1] = *reinterpret_cast<const u32 *>(&s_A[limb][wrow + g + 8][tq * 4]);
        A[limb][2] = *reinterpret_cast<const u32 *>(&s_A[limb][wrow + g][16 + tq * 4]);
        A[limb][3] = *reinterpret_cast<const u32 *>(&s_A[limb][wrow + g + 8][16 + tq * 4]);
    }
#pragma unroll 1
    for (int nt = 0; nt < 8; ++nt)
    {
        const int ncol = wcol + nt * 8;
        u32 B[4][2];
#pragma unroll
        for (int limb = 0; limb < 4; ++limb)
        {
            B[limb][0] = *reinterpret_cast<const u32 *>(&s_B[limb][ncol + g][tq * 4]);
            B[limb][1] = *reinterpret_cast<const u32 *>(&s_B[limb][ncol + g][16 + tq * 4]);
        }
        int S[7][4];
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) S[k][e] = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) la_mma_u8(S[a + b], A[a], B[b]);
        // fragment element e: row = g + 8 * (e >> 1), column = 2 * tq + (e & 1)
#pragma unroll
        for (int half = 0; half < 2; ++half)
        {
            const int row = row0 + wrow + g + 8 * half;
            const int col = col0 + ncol + 2 * tq;
            if (row >= c.rows || col >= c.ld) continue;
            uint2 *ptr = reinterpret_cast<uint2 *>(c.M + (size_t)row * c.ld + col);
            uint2 m = *ptr;
            u32 mv[2] = {m.x, m.y};
#pragma unroll
            for (int b = 0; b < 2; ++b)
            {
                const int cc = col + b;
                if (cc >= c0 && cc < c0 + TB_LA_PANEL) continue;      // panel columns are final already
                const int e = half * 2 + b;
                const u32 lo = (u32)S[0][e] + ((u32)S[1][e] << 8);                 // weights 2^0, 2^8    (< 2^31)
                const u32 mid = (u32)S[2][e] + ((u32)S[3][e] << 8);                // weights 2^16, 2^24  (< 2^31)
                const u32 hi = (u32)S[4][e] + ((u32)S[5][e] << 8) + ((u32)S[6][e] << 16);   // weights 2^32, 2^40, 2^48
                const u64 acc = (u64)lo + ((u64)mid << 16) + ((u64)hi << 32);
                mv[b] = la_sub(mv[b], la_fold(acc, c), c);
            }
            *ptr = make_uint2(mv[0], mv[1]);
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
    const u32 v = la_redc((u64)c.M[(size_t)c.piv_row[k] * c.ld + fc], c);      // out of the Montgomery domain
    basis[(size_t)m * c.cols + c.piv_col[k]] = v ? c.P - v : 0u;
}

}  // namespace tb
