// tb_kernels.cuh -- the CUDA kernels of the p-neighbor Hecke path (sm_100a).
//
//   inverse_lut_kernel   Fp::make_inverse_lut                      Fp.h:228-269
//   prep_reps_kernel     NeighborManager ctor constants            NeighborManager.h:13-48
//   neighbors_kernel     the (rep, t) loop body                    Genus.h:562-618
//   rows_kernel          per-conductor row accumulation + CSR      Genus.h:620-657
//   indptr_scan_kernel   CSR row pointers                          Genus.h:656
//   dense_scatter_kernel dense rows                                Genus.h:785-803
//   eigen_dot_kernel     eigenvalue accumulation                   Genus.h:490-499
//
// The path is integer-issue bound with an L2-resident working set; no tensor
// cores, no TMA (there is no bulk tile to move: every thread's operands live
// in registers and the genus table is a few MB of random 128-byte lines).
#pragma once

// compute-sanitizer is closed on the GPU pool this was developed on, so the memory-safety evidence is a
// second build with -DTB_DEBUG_BOUNDS: every computed shared-memory index and every global output
// position of the row kernels (and the table index of the lookup) is asserted before use, a violation
// traps the kernel ("device-side assert triggered" -> the call fails), and the GPU test-suite is run
// against that build (profiles/run_bounds_build.sh).
#ifdef TB_DEBUG_BOUNDS
#include <cassert>
#define TB_CHECK(cond) assert(cond)
#else
#define TB_CHECK(cond) ((void)0)
#endif

#include "tb_neighbor.cuh"

namespace tb {

#define TB_NBR_THREADS 128
#ifndef TB_NBR_MINBLOCKS
#define TB_NBR_MINBLOCKS 7
#endif

// ---------------------------------------------------------------------------
// Inverse table mod p (one launch per prime).
// ---------------------------------------------------------------------------
__global__ void inverse_lut_kernel(PrimeCtx pc, u32 *lut)
{
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= pc.p) return;
    PrimeCtx local = pc;
    local.use_lut = 0;
    lut[a] = a ? fp_pow(a, pc.p - 2u, local) : 0u;
}

// ---------------------------------------------------------------------------
// Per-representative constants for this prime.  base[n] = (x0, y0, z0) comes
// from the host (it consumes the reference's sequential RNG stream).
// ---------------------------------------------------------------------------
__global__ void prep_reps_kernel(GenusCtx gc, PrimeCtx pc, const u32 *base, int rep_begin, int rows, RepPrime *out,
                                 const int *rep_list = nullptr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int n = rep_list ? rep_list[i] : rep_begin + i;
    const i64 *cur = gc.cur + (size_t)n * TB_REC_WORDS;
    const u32 p = pc.p;
    RepPrime rp;
    rp.pad = 0;
    if (p == 2u)
    {
        rp.a = rp.b = rp.h = 0;
        rp.x0 = base[3 * i + 0]; rp.y0 = base[3 * i + 1]; rp.a0 = base[3 * i + 2];
        rp.delta = 0;
        out[i] = rp;
        return;
    }
    const u32 a = fp_mod_i64(cur[0], pc), b = fp_mod_i64(cur[1], pc);
    const u32 f = fp_mod_i64(cur[3], pc), g = fp_mod_i64(cur[4], pc), h = fp_mod_i64(cur[5], pc);
    const u32 x0 = base[3 * i + 0], y0 = base[3 * i + 1];
    // a0 = a - 2 a x0 - g - h y0 ; delta = b - 2 b y0 - f + h - h x0      NeighborManager.h:34-47
    u32 t = fp_mul(a, x0, pc);
    u32 a0 = fp_sub(fp_sub(fp_sub(a, fp_add(t, t, p), p), g, p), fp_mul(h, y0, pc), p);
    t = fp_mul(b, y0, pc);
    u32 delta = fp_sub(fp_add(fp_sub(fp_sub(b, fp_add(t, t, p), p), f, p), h, p), fp_mul(h, x0, pc), p);
    rp.a = a; rp.b = b; rp.h = h; rp.x0 = x0; rp.y0 = y0; rp.a0 = a0; rp.delta = delta;
    out[i] = rp;
}

// ---------------------------------------------------------------------------
// Conic base points on the device, one thread per representative, for the calls whose
// result does not depend on WHICH rational point parametrises the p+1 lines (Hecke
// matrices, eigenvalues: the set of lines is the same, only their order changes).  The
// order-defining RNG-exact points (isometry_sequence) stay on the host.  Same construction
// as QuadFormFp::isotropic_vector (QuadForm.h:847-907) with a deterministic scan instead of
// random draws: take r with Q(1,r,0) != 0, then the first s whose line meets the conic in a
// rational point, and solve the quadratic (Tonelli-Shanks with the launch's non-residue).
// ---------------------------------------------------------------------------
__device__ __forceinline__ u32 fp_sqrt_dev(u32 a, const PrimeCtx &pc, u32 nonres)
{
    const u32 p = pc.p;
    if (a <= 1u) return a;
    u32 odd = p - 1u;
    int twos = 0;
    while (!(odd & 1u)) { odd >>= 1; ++twos; }
    if (twos == 1) return fp_pow(a, (p + 1u) / 4u, pc);
    int m = twos;
    u32 c = fp_pow(nonres, odd, pc), r = fp_pow(a, (odd + 1u) / 2u, pc), t = fp_pow(a, odd, pc);
    while (t != 1u)
    {
        int i = 0;
        for (u32 t1 = t; t1 != 1u; t1 = fp_mul(t1, t1, pc)) ++i;
        u32 b = c;
        for (int j = 0; j < m - i - 1; ++j) b = fp_mul(b, b, pc);
        r = fp_mul(r, b, pc);
        c = fp_mul(b, b, pc);
        t = fp_mul(t, c, pc);
        m = i;
    }
    return r;
}

__global__ void base_points_kernel(GenusCtx gc, PrimeCtx pc, u32 nonres, int rep_begin, int rows, u32 *base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const i64 *cur = gc.cur + (size_t)(rep_begin + i) * TB_REC_WORDS;
    const u32 p = pc.p;
    PrimeCtx fc = pc;
    fc.use_lut = 0;                                          // Fermat inverses: independent of the table build
    if (p == 2u)
    {
        const u32 a = (u32)cur[0] & 1u, b = (u32)cur[1] & 1u, c = (u32)cur[2] & 1u;
        const u32 f = (u32)cur[3] & 1u, g = (u32)cur[4] & 1u, h = (u32)cur[5] & 1u;
        const u32 val[8] = {1u, c, b, b ^ f ^ c, a, a ^ g ^ c, a ^ h ^ b, a ^ b ^ c ^ f ^ g ^ h};
        u32 found[3] = {0u, 0u, 0u};
        int cnt = 0;
        for (u32 v = 1; v < 8; ++v) if (val[v] == 0u && cnt < 3) found[cnt++] = v;
        base[3 * i + 0] = found[0]; base[3 * i + 1] = found[1]; base[3 * i + 2] = found[2];
        return;
    }
    const u32 a = fp_mod_i64(cur[0], fc), b = fp_mod_i64(cur[1], fc), c = fp_mod_i64(cur[2], fc);
    const u32 f = fp_mod_i64(cur[3], fc), g = fp_mod_i64(cur[4], fc), h = fp_mod_i64(cur[5], fc);
    const u32 half_exp = (p - 1u) / 2u;
    for (u32 r = 0; r < p; ++r)
    {
        const u32 br = fp_mul(b, r, fc);
        const u32 alpha = fp_add(fp_mul(fp_add(br, h, p), r, fc), a, p);              // Q(1, r, 0)
        if (alpha == 0u) continue;
        const u32 two_alpha_inv = fp_inv(fp_add(alpha, alpha, p), fc);
        const u32 fr_g = fp_add(fp_mul(f, r, fc), g, p);
        const u32 slope = fp_add(fp_add(br, br, p), h, p);
        for (u32 s = 0; s < p; ++s)
        {
            const u32 beta = fp_add(fp_mul(slope, s, fc), fr_g, p);
            const u32 gamma = fp_add(fp_mul(fp_add(fp_mul(b, s, fc), f, p), s, fc), c, p);   // Q(0, s, 1)
            u32 four_ag = fp_mul(gamma, alpha, fc);
            four_ag = fp_add(four_ag, four_ag, p);
            four_ag = fp_add(four_ag, four_ag, p);
            const u32 disc = fp_sub(fp_mul(beta, beta, fc), four_ag, p);
            if (disc != 0u && fp_pow(disc, half_exp, fc) != 1u) continue;             // not a square
            const u32 root = fp_sub(fp_sqrt_dev(disc, fc, nonres), beta, p);
            const u32 x = fp_mul(root, two_alpha_inv, fc);
            base[3 * i + 0] = x;
            base[3 * i + 1] = fp_add(fp_mul(r, x, fc), s, p);
            base[3 * i + 2] = 1u;
            return;
        }
    }
    base[3 * i + 0] = 0u; base[3 * i + 1] = 0u; base[3 * i + 2] = 1u;                 // unreachable for p not dividing disc
}

// ---------------------------------------------------------------------------
// neighbors_kernel: one thread per (representative, t).
//   MODE_PACKED   W[i] = (r << K) | spin                         Genus.h:617
//   MODE_RECORDS  20 int64: line[3], reduced form[6], isometry[9], r, spin
//   MODE_ISOSEQ   12 int64: composed isometry[9], denominator, src, dst
// err[0] / err[1]: smallest neighbor index that overflowed / missed the table.
// ---------------------------------------------------------------------------
enum { MODE_PACKED = 0, MODE_RECORDS = 1, MODE_ISOSEQ = 2 };

struct NbrLaunch {
    int rep_begin;
    int rows;
    u32 P1;           // p + 1
    u32 pad;
    u64 total;        // rows * P1
    InvDiv dP1;
    u64 mP1;          // floor(2^64 / P1) + 1 when total < 2^32 (two 32-bit multiplies give i / P1 exactly), else 0
    const RepPrime *rp;
    u32 *W;
    i64 *records;
    u64 *err;
    const int *rep_list;   // optional: row -> representative (eigenvalue path), else rep_begin + row
};

// register budget per integer mode: 72 registers (7 CTAs/SM) suit the int64 kernel, the
// two-limb kernels need ~3x the live state
template <class Z> struct NbrBlocks { static const int value = 3; };
template <> struct NbrBlocks<W64> { static const int value = TB_NBR_MINBLOCKS; };
template <> struct NbrBlocks<X64> { static const int value = TB_NBR_MINBLOCKS; };

template <class Z, int MODE>
__global__ void __launch_bounds__(TB_NBR_THREADS, NbrBlocks<Z>::value)
neighbors_kernel(const __grid_constant__ GenusCtx gc, const __grid_constant__ PrimeCtx pc,
                 const __grid_constant__ NbrLaunch L)
{
    if (threadIdx.x == 0) *cta_flag_ptr() = 0u;
    __syncthreads();

    // lanes past the end recompute the last neighbor (no early exit: the W64 reduce votes
    // warp-wide) and skip the writes
    const u64 i_raw = (u64)blockIdx.x * TB_NBR_THREADS + threadIdx.x;
    const bool live = i_raw < L.total;
    const u64 i = live ? i_raw : L.total - 1;
    {
        // (row, t) = divmod(i, p + 1).  Below 2^32 neighbors: i * M >> 64 with M = floor(2^64 / P1) + 1 is exact
        // (error i (M P1 - 2^64) / (P1 2^64) < 2^-32 < 1 / P1)
        u64 row;
        u32 t;
        if (L.mP1)
        {
            const u32 i32 = (u32)i;
            const u64 mid = (u64)i32 * (u32)(L.mP1 >> 32) + (((u64)i32 * (u32)L.mP1) >> 32);
            row = mid >> 32;
            t = i32 - (u32)row * L.P1;
        }
        else
        {
            u64 t64;
            row = udivmod(i, L.dP1, t64);
            t = (u32)t64;
        }
        const int n = L.rep_list ? L.rep_list[row] : L.rep_begin + (int)row;
        const RepPrime rp = L.rp[row];

        Neighbor<Z> o;
        const int status = one_neighbor<Z>(n, t, rp, pc, gc, MODE == MODE_ISOSEQ, o);
        if (status == NBR_OVERFLOW && live) atomicMin(L.err + 0, i);
        else if (status == NBR_NOT_FOUND && live) atomicMin(L.err + 1, i);

        if (!live)
        {
        }
        else if (MODE == MODE_PACKED)
        {
            L.W[i] = status == NBR_OK ? (((u32)o.r << gc.K) | o.spin) : 0xffffffffu;
        }
        else if (MODE == MODE_RECORDS)
        {
            i64 *out = L.records + i * 20;
            out[0] = o.vx; out[1] = o.vy; out[2] = o.vz;
            if (status == NBR_OK)
            {
                out[3] = o.q.a.low64(); out[4] = o.q.b.low64(); out[5] = o.q.c.low64();
                out[6] = o.q.f.low64(); out[7] = o.q.g.low64(); out[8] = o.q.h.low64();
                out[9] = o.s.a11.low64(); out[10] = o.s.a12.low64(); out[11] = o.s.a13.low64();
                out[12] = o.s.a21.low64(); out[13] = o.s.a22.low64(); out[14] = o.s.a23.low64();
                out[15] = o.s.a31.low64(); out[16] = o.s.a32.low64(); out[17] = o.s.a33.low64();
                out[18] = o.r; out[19] = o.spin;
            }
        }
        else
        {
            i64 *out = L.records + i * 12;
            if (status == NBR_OK)
            {
                const Iso<Z> &c = o.comp;
                if (!(c.a11.fits_i64() && c.a12.fits_i64() && c.a13.fits_i64() && c.a21.fits_i64() &&
                      c.a22.fits_i64() && c.a23.fits_i64() && c.a31.fits_i64() && c.a32.fits_i64() &&
                      c.a33.fits_i64() && o.denom.fits_i64()))
                    raise_overflow();
                out[0] = c.a11.low64(); out[1] = c.a12.low64(); out[2] = c.a13.low64();
                out[3] = c.a21.low64(); out[4] = c.a22.low64(); out[5] = c.a23.low64();
                out[6] = c.a31.low64(); out[7] = c.a32.low64(); out[8] = c.a33.low64();
                out[9] = o.denom.low64(); out[10] = n; out[11] = o.r;
            }
        }
    }

    // no closing barrier (it held every warp of the CTA until the slowest one finished): a
    // thread that raised the sticky flag sees its own write, so it reports it itself
    if (*cta_flag_ptr()) atomicMin(L.err + 0, (u64)blockIdx.x * TB_NBR_THREADS);
}

// ---------------------------------------------------------------------------
// Block-wide exclusive prefix sum of one int per thread (blockDim <= 1024).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        const int nw = (blockDim.x + 31) >> 5;
        int w = lane < nw ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += y;
        }
        warp_sums[lane] = wi - w;          // exclusive warp offsets
        if (lane == 31) warp_sums[32] = wi;   // grand total
    }
    __syncthreads();
    const int out = warp_sums[warp] + incl - v;
    total = warp_sums[32];
    __syncthreads();
    return out;
}

// ---------------------------------------------------------------------------
// rows_kernel: one CTA per representative row; all 2^K conductors.
//
// The p+1 packed words of the row are counting-sorted by r with a bitmask over
// the genus (rank = popcount prefix), so each distinct r becomes one column j
// holding a run of spins.  For conductor k the row entry of column j is
//     sum_i chi_k(spin_i) = run - 2 * #{i : popc(spin_i & k) odd}        birch_util.h:81-91
// kept if lut[k][r] != -1 and the sum is nonzero.                         Genus.h:628-653
// MODE ROWS_COUNT counts the nonzeros per (k, row); ROWS_FILL writes them at indptr
// offsets (ascending columns, zeros dropped); ROWS_FUSED does both in one pass: rows are
// taken in ticket order and each CTA obtains its per-conductor output offset by a
// decoupled look-back over the rows before it (single-pass chained scan), writing into
// per-conductor regions sized by an upper bound.
// ---------------------------------------------------------------------------
struct RowLaunch {
    int N, K;
    int rep_begin, rows;
    u32 P1;
    int mask_words;        // ceil(N / 32)
    const u32 *W;          // [rows][P1]
    const int *lut;        // [2^K][N]
    const int *lutT;       // [N][2^K] (transposed: one sector serves all conductors of a column)
    const int *rowbase;    // [2^K] offset of conductor k's indptr block in indptr_all
    const int *rowfirst;   // [2^K] lut[k][first included rep >= rep_begin] (row index of the shard's first row)
    int *indptr_all;       // concatenated (rows_k + 1) blocks
    const i64 *nnzbase;    // [2^K] offset of conductor k in data/indices (FILL, FUSED)
    u64 *lookback;         // [2^K][rows] (flag << 62) | count, zeroed before launch (FUSED)
    int *ticket;           // row dispenser, zeroed before launch (FUSED)
    i64 *nnz_out;          // [2^K] written by the last row (FUSED)
    int *maxmult;          // max number of neighbors of one row in one class = max |entry| (narrow read-back)
    u32 *scratch;          // global-memory stand-in for the per-row shared-memory state when it does not fit one SM:
    size_t scratch_words;  //   one slot of scratch_words per CTA, and each CTA then walks several rows
    i64 extent;            // entries allocated for data / indices (bounds checks of the debug build)
    const unsigned char *incl;   // [ceil(2^K / 8)][N] bit kk: lut[8 c + kk][n] != -1   (rows_acc_kernel)
    int *overflow;         // set when a class is hit 256 times or more (rows_acc_kernel; the host falls back)
    unsigned short *idx16; // optional narrow twins of indices / data for the read-back over PCIe
    signed char *dat8;     //   (valid when every dim <= 65536 and maxmult <= 127), same offsets
    int *data;
    int *indices;
};

enum { ROWS_COUNT = 0, ROWS_FILL = 1, ROWS_FUSED = 2 };

template <int MODE, bool SCRATCH = false>
__global__ void __launch_bounds__(512, 2) rows_kernel(const __grid_constant__ RowLaunch L)
{
    extern __shared__ u32 smem[];
    // layout: mask[mw] | wprefix[mw] | start[maxcols+1] | colr[maxcols] | scratch[256] | spins (u16)[P1]
    // (in shared memory; for a very large genus at a very large prime the same layout lives in a
    //  per-CTA slot of global memory and the CTA walks rows persistently)
    const int mw = L.mask_words;
    const int maxcols = min((int)L.P1, L.N);
    u32 *mask = SCRATCH ? L.scratch + (size_t)blockIdx.x * L.scratch_words : smem;
    u32 *wprefix = mask + mw;
    int *start = (int *)(wprefix + mw);
    int *colr = start + maxcols + 1;
    int *warp_sums = colr + maxcols;
    unsigned short *spins = (unsigned short *)(warp_sums + 256);

    __shared__ int s_row;
    __shared__ i64 s_base[8];
    for (int walk = 0;; ++walk)
    {
    if (MODE == ROWS_FUSED)
    {
        // tickets are handed out in launch order, so every row a CTA waits on below belongs
        // to a CTA that is already running or finished: the look-back cannot deadlock
        if (threadIdx.x == 0) s_row = atomicAdd(L.ticket, 1);
        __syncthreads();
    }
    const int lrow = MODE == ROWS_FUSED ? s_row : (int)(blockIdx.x + (size_t)walk * gridDim.x);
    if (lrow >= L.rows) break;
    const int n = L.rep_begin + lrow;
    const u32 *Wrow = L.W + (size_t)lrow * L.P1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const u32 spin_mask = (1u << L.K) - 1u;

    for (int i = tid; i < mw; i += nt) mask[i] = 0u;
    for (int i = tid; i <= maxcols; i += nt) start[i] = 0;
    __syncthreads();

    // touched representatives
    // (a failed neighbor carries 0xffffffff: r >= N, skipped everywhere below)
    for (u32 t = tid; t < L.P1; t += nt)
    {
        u32 r = Wrow[t] >> L.K;
        if (r < (u32)L.N) atomicOr(&mask[r >> 5], 1u << (r & 31));
    }
    __syncthreads();

    // popcount prefix over mask words -> rank of every touched r
    int carry = 0;
    for (int base = 0; base < mw; base += nt)
    {
        int i = base + tid;
        int c = i < mw ? __popc(mask[i]) : 0;
        int total;
        int ex = block_exclusive_scan(c, warp_sums, total);
        if (i < mw) wprefix[i] = carry + ex;
        carry += total;
    }
    const int ncols = carry;
    __syncthreads();

    // run lengths, then run starts
    for (u32 t = tid; t < L.P1; t += nt)
    {
        u32 r = Wrow[t] >> L.K;
        if (r >= (u32)L.N) continue;
        int j = wprefix[r >> 5] + __popc(mask[r >> 5] & ((1u << (r & 31)) - 1u));
        TB_CHECK(j >= 0 && j < maxcols);
        atomicAdd(&start[j + 1], 1);
    }
    // column -> representative
    for (int i = tid; i < mw; i += nt)
    {
        u32 m = mask[i];
        int j = wprefix[i];
        while (m)
        {
            int b = __ffs(m) - 1;
            colr[j++] = i * 32 + b;
            m &= m - 1;
        }
    }
    __syncthreads();
    carry = 0;
    for (int base = 0; base < ncols; base += nt)
    {
        int j = base + tid;
        int c = j < ncols ? start[j + 1] : 0;
        int total;
        int ex = block_exclusive_scan(c, warp_sums, total);
        if (j < ncols) start[j + 1] = carry + ex + c;   // inclusive end
        carry += total;
    }
    const int nvalid = carry;
    __syncthreads();
    // scatter spins into their runs, filling each run from its end backwards
    for (u32 t = tid; t < L.P1; t += nt)
    {
        u32 w = Wrow[t];
        u32 r = w >> L.K;
        if (r >= (u32)L.N) continue;
        int j = wprefix[r >> 5] + __popc(mask[r >> 5] & ((1u << (r & 31)) - 1u));
        int slot = atomicSub(&start[j + 1], 1) - 1;
        TB_CHECK(j >= 0 && j < maxcols && slot >= 0 && slot < (int)L.P1);
        spins[slot] = (unsigned short)(w & spin_mask);
    }
    __syncthreads();
    // now start[j+1] == begin of run j; the end of run j is the begin of run j+1, nvalid for the last
    // (start[0] is unused by the scatter and still 0)

    // ---- per-conductor row entries, ascending columns, zeros dropped ----------
    // Each warp owns a contiguous block of columns, so ordered compaction needs no
    // block-wide scan per tile: warps count their nonzeros with ballots, one sync
    // combines the per-warp totals, and (FILL) a second sweep writes at the offsets.
    // Conductors are handled eight at a time (counters live in registers).
    const int C = 1 << L.K;
    const int lane = tid & 31, w = tid >> 5, nw = nt >> 5;
    const int B = (((ncols + nw - 1) / nw) + 31) & ~31;       // columns per warp, multiple of 32
    const int jbeg = min(ncols, w * B), jend = min(ncols, jbeg + B);
    int *warp_cnt = warp_sums;                                 // [8][32], reuses the scan scratch
    __shared__ int s_npos[8];
    __shared__ u64 s_partab[1024];                             // parity table of the current conductor chunk (K <= 10)
    const bool use_tab = L.K <= 10;

    if (MODE != ROWS_FILL && L.maxmult)
    {
        // the largest run bounds every |entry| of the row in every conductor
        int mx = 0;
        for (int j = tid; j < ncols; j += nt)
            mx = max(mx, ((j + 1 < ncols) ? start[j + 2] : nvalid) - start[j + 1]);
        for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx > *(volatile int *)L.maxmult) atomicMax(L.maxmult, mx);
    }

    for (int kc = 0; kc < C; kc += 8)
    {
        const int nk = min(8, C - kc);
        __syncthreads();
        if (tid < 8) s_npos[tid] = tid < nk ? L.lut[(size_t)(kc + tid) * L.N + n] : -1;
        if (use_tab)
        {
            for (int sp = tid; sp < C; sp += nt)
            {
                u64 v = 0;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) v |= (u64)(__popc((u32)sp & (u32)(kc + kk)) & 1) << (8 * kk);
                s_partab[sp] = v;
            }
        }
        __syncthreads();

        int cnt[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) cnt[kk] = 0;

        for (int pass = 0; pass < (MODE == ROWS_COUNT ? 1 : 2); ++pass)
        {
            i64 wbase[8];
            if (pass == 1)
            {
                // exclusive prefix of the per-warp totals -> this warp's first output slot per conductor
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                {
                    int before = 0;
                    for (int ww = 0; ww < w; ++ww) before += warp_cnt[kk * 32 + ww];
                    wbase[kk] = 0;
                    if (kk < nk && s_npos[kk] != -1)
                    {
                        const int k = kc + kk;
                        if (MODE == ROWS_FUSED) wbase[kk] = s_base[kk] + before;
                        else
                        {
                            const int out_row = s_npos[kk] - L.rowfirst[k];
                            wbase[kk] = L.nnzbase[k] + (L.indptr_all + L.rowbase[k])[out_row] + before;
                        }
                    }
                    cnt[kk] = 0;
                }
            }
            for (int j0 = jbeg; j0 < jend; j0 += 32)
            {
                const int j = j0 + lane;
                const bool valid = j < jend;
                int r = 0, b = 0, e = 0;
                if (valid)
                {
                    r = colr[j];
                    b = start[j + 1];
                    e = (j + 1 < ncols) ? start[j + 2] : nvalid;
                }
                // odd[kk] = #{entries of the run with chi_k = -1}.  Run lengths vary (the tail
                // iterations run with a handful of lanes), so the per-entry work is one table
                // add: s_partab[spin] holds the eight parities spread over the bytes of a u64.
                int odd[8];
                if (use_tab && e - b < 256)
                {
                    u64 acc = 0;
                    for (int i = b; i < e; ++i) acc += s_partab[spins[i]];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) odd[kk] = (int)((acc >> (8 * kk)) & 0xffull);
                }
                else
                {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) odd[kk] = 0;
                    for (int i = b; i < e; ++i)
                    {
                        const u32 sp = spins[i];
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) odd[kk] += __popc(sp & (u32)(kc + kk)) & 1;
                    }
                }
                // row positions of this column for the eight conductors: 32 contiguous bytes
                int rp[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) rp[kk] = -1;
                if (valid)
                {
                    const int *src = L.lutT + ((size_t)r << L.K) + kc;
                    if (nk == 8)
                    {
                        const int4 x = __ldg(reinterpret_cast<const int4 *>(src));
                        const int4 y = __ldg(reinterpret_cast<const int4 *>(src) + 1);
                        rp[0] = x.x; rp[1] = x.y; rp[2] = x.z; rp[3] = x.w;
                        rp[4] = y.x; rp[5] = y.y; rp[6] = y.z; rp[7] = y.w;
                    }
                    else
                    {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) if (kk < nk) rp[kk] = __ldg(src + kk);
                    }
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                {
                    if (kk >= nk || s_npos[kk] == -1) continue;          // Genus.h:624-625 (warp-uniform)
                    int value = 0;
                    const int rpos = rp[kk];
                    if (rpos != -1) value = (e - b) - 2 * odd[kk];
                    const unsigned ballot = __ballot_sync(0xffffffffu, value != 0);
                    if (pass == 1 && value != 0)
                    {
                        const i64 at = wbase[kk] + cnt[kk] + __popc(ballot & ((1u << lane) - 1u));
                        TB_CHECK(at >= 0 && at < L.extent && rpos >= 0);
                        if (L.data)
                        {
                            L.data[at] = value;
                            L.indices[at] = rpos;
                        }
                        if (L.idx16)
                        {
                            L.idx16[at] = (unsigned short)rpos;
                            L.dat8[at] = (signed char)value;
                        }
                    }
                    cnt[kk] += __popc(ballot);
                }
            }
            if (pass == 0)
            {
                if (lane == 0)
                {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) warp_cnt[kk * 32 + w] = cnt[kk];
                }
                __syncthreads();
                if (MODE == ROWS_COUNT && tid < nk && s_npos[tid] != -1)
                {
                    const int k = kc + tid;
                    int total = 0;
                    for (int ww = 0; ww < nw; ++ww) total += warp_cnt[tid * 32 + ww];
                    (L.indptr_all + L.rowbase[k])[s_npos[tid] - L.rowfirst[k] + 1] = total;
                }
                if (MODE == ROWS_FUSED)
                {
                    if (tid < nk)
                    {
                        const int k = kc + tid;
                        i64 total = 0;
                        if (s_npos[tid] != -1)
                            for (int ww = 0; ww < nw; ++ww) total += warp_cnt[tid * 32 + ww];
                        u64 *state = L.lookback + (size_t)k * L.rows;
                        const u64 FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VALUE = (1ull << 62) - 1;
                        if (lrow > 0) atomicExch(state + lrow, FLAG_AGG | (u64)total);
                        i64 excl = 0;
                        for (int j = lrow - 1; j >= 0; )
                        {
                            const u64 v = *(volatile u64 *)(state + j);
                            const u64 flag = v >> 62;
                            if (flag == 0) continue;                 // predecessor still counting
                            excl += (i64)(v & VALUE);
                            if (flag == 2) break;                    // inclusive prefix: done
                            --j;
                        }
                        const i64 incl = excl + total;
                        atomicExch(state + lrow, FLAG_INC | (u64)incl);
                        s_base[tid] = L.nnzbase[k] + excl;
                        if (s_npos[tid] != -1)
                            (L.indptr_all + L.rowbase[k])[s_npos[tid] - L.rowfirst[k] + 1] = (int)incl;
                        if (lrow == L.rows - 1) L.nnz_out[k] = incl;
                    }
                    __syncthreads();
                }
            }
        }
    }
    if (!SCRATCH) break;         // shared-memory form: one row per CTA
    __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// rows_acc_kernel: the row phase for the common case (every class of a row is hit fewer than
// 256 times), single pass like rows_kernel<ROWS_FUSED> but without the counting sort.
//
// A Hecke row entry is  run - 2 * #{neighbors of the class with chi_k = -1}, so all a row needs
// per touched class is a handful of small counters -- and those are accumulated directly with
// shared-memory atomics while the packed words stream by, instead of first sorting the spins
// into runs and then walking the runs (ncu, profiles/r1k: 75 % of rows_kernel's instructions sat
// in that walk and in its two-sweep compaction).  Per row, one CTA:
//   A  mask[r]      bitmask of the classes hit (atomicOr), rank of a class = popcount prefix
//   C  acc[rank]    two 32-bit words of 4 byte-counters each: for the 8 conductors of the
//                   current chunk, how many neighbors of the class have chi = -1; byte 0 of
//                   chunk 0 (conductor index 0: chi = +1 always) counts the run itself
//   D  mask words in order (lane = bit = class): counters -> entries -> ballots; a count sweep
//      gives every warp its offset per conductor, a second sweep stores (uint16 column + int8
//      entry, or int32 pairs).  Rows chain through the same decoupled look-back.
// More than 8 conductors: C and D repeat per chunk of 8 (the row's words are L2-resident).
// A class hit 256 times or more overflows its run counter: the kernel raises *overflow and
// the host redoes the row phase with rows_kernel.
// ---------------------------------------------------------------------------
__device__ __forceinline__ u64 parity_bytes(u32 spin, int kc)
{
    u64 v = 0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) v |= (u64)(__popc(spin & (u32)(kc + kk)) & 1) << (8 * kk);
    return v;
}

template <bool NARROW>
__global__ void __launch_bounds__(512, 2) rows_acc_kernel(const __grid_constant__ RowLaunch L)
{
    extern __shared__ u32 smem[];
    // layout: mask[mw] | wprefix[mw] | acc_lo[maxcols] | acc_hi[maxcols] | warp_cnt[8][32] | runlen (u8)[maxcols]
    const int mw = L.mask_words;
    const int maxcols = min((int)L.P1, L.N);
    u32 *mask = smem;
    u32 *wprefix = mask + mw;
    u32 *acc_lo = wprefix + mw;
    u32 *acc_hi = acc_lo + maxcols;
    int *warp_cnt = (int *)(acc_hi + maxcols);
    unsigned char *runlen = (unsigned char *)(warp_cnt + 8 * 32 + 40);
    int *scan_tmp = warp_cnt;                                  // block_exclusive_scan scratch (40 ints), before phase D

    __shared__ int s_row;
    __shared__ i64 s_base[8];
    __shared__ int s_npos[8];
    __shared__ u64 s_partab[1024];
    __shared__ int s_maxrun;

    // rows are taken in ticket order: every row a CTA waits on in the look-back belongs to a CTA
    // that is already running or finished
    if (threadIdx.x == 0) { s_row = atomicAdd(L.ticket, 1); s_maxrun = 0; }
    __syncthreads();
    const int lrow = s_row;
    if (lrow >= L.rows) return;
    const int n = L.rep_begin + lrow;
    const u32 *Wrow = L.W + (size_t)lrow * L.P1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, w = tid >> 5, nw = nt >> 5;
    const u32 spin_mask = (1u << L.K) - 1u;
    const int C = 1 << L.K;
    const bool use_tab = L.K <= 10;
    const u32 lt = (1u << lane) - 1u;

    // ---- A: classes hit --------------------------------------------------------
    for (int i = tid; i < mw; i += nt) mask[i] = 0u;
    __syncthreads();
    for (u32 t = tid; t < L.P1; t += nt)
    {
        const u32 r = Wrow[t] >> L.K;                          // a failed neighbor carries 0xffffffff: r >= N, skipped
        if (r < (u32)L.N) atomicOr(&mask[r >> 5], 1u << (r & 31));
    }
    __syncthreads();
    int carry = 0;
    for (int base = 0; base < mw; base += nt)
    {
        const int i = base + tid;
        const int c = i < mw ? __popc(mask[i]) : 0;
        int total;
        const int ex = block_exclusive_scan(c, scan_tmp, total);
        if (i < mw) wprefix[i] = carry + ex;
        carry += total;
    }
    const int ncols = carry;

    // this warp's block of mask words (so of ranks, so of output positions): contiguous, ascending
    const int wpw = (mw + nw - 1) / nw;
    const int ibeg = min(mw, w * wpw), iend = min(mw, ibeg + wpw);

    for (int kc = 0; kc < C; kc += 8)
    {
        const int nk = min(8, C - kc);
        const bool both = nk > 4;
        __syncthreads();
        // ---- C: counters ---------------------------------------------------------
        for (int j = tid; j < ncols; j += nt) { acc_lo[j] = 0u; acc_hi[j] = 0u; }
        if (tid < 8) s_npos[tid] = tid < nk ? L.lut[(size_t)(kc + tid) * L.N + n] : -1;
        if (use_tab)
            for (int sp = tid; sp < C; sp += nt) s_partab[sp] = parity_bytes((u32)sp, kc) | (kc == 0 ? 1ull : 0ull);
        __syncthreads();
        for (u32 t = tid; t < L.P1; t += nt)
        {
            const u32 word = Wrow[t];
            const u32 r = word >> L.K;
            if (r >= (u32)L.N) continue;
            const u32 sp = word & spin_mask;
            const int j = wprefix[r >> 5] + __popc(mask[r >> 5] & ((1u << (r & 31)) - 1u));
            TB_CHECK(j >= 0 && j < ncols && ncols <= maxcols);
            const u64 add = use_tab ? s_partab[sp] : (parity_bytes(sp, kc) | (kc == 0 ? 1ull : 0ull));
            const u32 old = atomicAdd(&acc_lo[j], (u32)add);
            if (kc == 0 && (old & 0xffu) == 0xffu) *L.overflow = 1;         // 256th neighbor of one class
            if (both && (u32)(add >> 32)) atomicAdd(&acc_hi[j], (u32)(add >> 32));
        }
        __syncthreads();
        if (kc == 0 && C > 8)
            for (int j = tid; j < ncols; j += nt) runlen[j] = (unsigned char)(acc_lo[j] & 0xffu);

        // ---- D: entries, ascending columns, zeros dropped -----------------------------
        // A class of column r has, per conductor kk of the chunk, the entry run - 2 odd_kk; it is stored iff
        // lut[kc+kk][r] != -1 (inc bit) and 2 odd_kk != run.  All eight tests are done at once on the counter
        // bytes: with hf = run / 2 in every byte, a byte of (counters ^ hf) is zero exactly where the entry
        // is (an odd run has none), and the exact zero-byte mask ~(((x & 0x7f..) + 0x7f..) | x | 0x7f..)
        // marks those.  The count sweep just adds the resulting 0/1 bytes per lane (no ballots); the store
        // sweep takes one ballot per conductor for the positions.
        u32 rowmask = 0u;                                                    // conductors of the chunk that contain row n (uniform)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) if (kk < nk && s_npos[kk] != -1) rowmask |= 1u << kk;
        const u32 rowb_lo = ((rowmask & 15u) * 0x00204081u) & 0x01010101u;   // bit kk -> byte kk
        const u32 rowb_hi = ((rowmask >> 4) * 0x00204081u) & 0x01010101u;
        // -- count
        {
            u32 cl = 0u, ch = 0u;                                            // byte lanes, folded into cnt every 255 words
            int cnt[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) cnt[kk] = 0;
            int mx = 0;
            for (int i0 = ibeg; i0 < iend; i0 += 255)
            {
                const int i1 = min(iend, i0 + 255);
                for (int i = i0; i < i1; ++i)
                {
                    const u32 m = mask[i];
                    if (!((m >> lane) & 1u)) continue;
                    const int j = wprefix[i] + __popc(m & lt);
                    TB_CHECK(j >= 0 && j < ncols && i * 32 + lane < L.N);
                    u32 lo = acc_lo[j];
                    const u32 hi = both ? acc_hi[j] : 0u;
                    const u32 run = kc == 0 ? (lo & 0xffu) : (u32)runlen[j];
                    if (kc == 0) lo &= ~0xffu;
                    const u32 inc = L.incl[(size_t)(kc >> 3) * L.N + i * 32 + lane];     // bit kk: lut[kc + kk][r] != -1
                    mx = max(mx, (int)run);
                    const u32 hf = (run >> 1) * 0x01010101u;
                    const u32 xl = lo ^ hf, xh = hi ^ hf;
                    u32 zl = ~(((xl & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xl | 0x7f7f7f7fu);  // 0x80 where 2 odd == run (even run)
                    u32 zh = ~(((xh & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xh | 0x7f7f7f7fu);
                    if (run & 1u) { zl = 0u; zh = 0u; }
                    cl += ~(zl >> 7) & ((inc & 15u) * 0x00204081u) & rowb_lo;
                    ch += ~(zh >> 7) & ((inc >> 4) * 0x00204081u) & rowb_hi;
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { cnt[kk] += (int)((cl >> (8 * kk)) & 0xffu); cnt[4 + kk] += (int)((ch >> (8 * kk)) & 0xffu); }
                cl = 0u; ch = 0u;
            }
            // per-lane counts -> per-warp totals -> this row's totals -> chained offsets
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                for (int o = 16; o; o >>= 1) cnt[kk] += __shfl_xor_sync(0xffffffffu, cnt[kk], o);
            if (lane == 0)
            {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) warp_cnt[kk * 32 + w] = cnt[kk];
            }
            if (kc == 0)
            {
                for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (lane == 0 && mx > s_maxrun) atomicMax(&s_maxrun, mx);
            }
        }
        __syncthreads();
        if (tid < nk)
        {
            const int k = kc + tid;
            i64 total = 0;
            if (s_npos[tid] != -1)
                for (int ww = 0; ww < nw; ++ww) total += warp_cnt[tid * 32 + ww];
            u64 *state = L.lookback + (size_t)k * L.rows;
            const u64 FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VALUE = (1ull << 62) - 1;
            if (lrow > 0) atomicExch(state + lrow, FLAG_AGG | (u64)total);
            i64 excl = 0;
            for (int j = lrow - 1; j >= 0; )
            {
                const u64 v = *(volatile u64 *)(state + j);
                const u64 flag = v >> 62;
                if (flag == 0) continue;                 // predecessor still counting
                excl += (i64)(v & VALUE);
                if (flag == 2) break;                    // inclusive prefix: done
                --j;
            }
            const i64 incl_total = excl + total;
            atomicExch(state + lrow, FLAG_INC | (u64)incl_total);
            s_base[tid] = L.nnzbase[k] + excl;
            if (s_npos[tid] != -1)
                (L.indptr_all + L.rowbase[k])[s_npos[tid] - L.rowfirst[k] + 1] = (int)incl_total;
            if (lrow == L.rows - 1) L.nnz_out[k] = incl_total;
        }
        if (kc == 0 && tid == 32 && L.maxmult && s_maxrun > *(volatile int *)L.maxmult) atomicMax(L.maxmult, s_maxrun);
        __syncthreads();
        // -- store
        {
            i64 wbase[8];                                        // this warp's next position in conductor kk's output
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
            {
                int before = 0;
                for (int ww = 0; ww < w; ++ww) before += warp_cnt[kk * 32 + ww];
                wbase[kk] = s_base[kk] + before;
            }
            for (int i = ibeg; i < iend; ++i)
            {
                const u32 m = mask[i];
                if (m == 0u) continue;                                   // warp-uniform
                const bool hit = (m >> lane) & 1u;
                const int r = i * 32 + lane;
                u32 lo = 0u, hi = 0u, run = 0u, inc = 0u;
                if (hit)
                {
                    const int j = wprefix[i] + __popc(m & lt);
                    lo = acc_lo[j];
                    hi = both ? acc_hi[j] : 0u;
                    run = kc == 0 ? (lo & 0xffu) : (u32)runlen[j];
                    if (kc == 0) lo &= ~0xffu;
                    inc = L.incl[(size_t)(kc >> 3) * L.N + r];
                }
                const u32 hf = (run >> 1) * 0x01010101u;
                const u32 xl = lo ^ hf, xh = hi ^ hf;
                u32 zl = ~(((xl & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xl | 0x7f7f7f7fu);
                u32 zh = ~(((xh & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xh | 0x7f7f7f7fu);
                if (run & 1u) { zl = 0u; zh = 0u; }
                const u32 nl = ~(zl >> 7) & ((inc & 15u) * 0x00204081u) & rowb_lo;      // inc = 0 on lanes without a class
                const u32 nh = ~(zh >> 7) & ((inc >> 4) * 0x00204081u) & rowb_hi;
                int rp[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) rp[kk] = -1;
                if (nl | nh)
                {
                    const int *src = L.lutT + ((size_t)r << L.K) + kc;
                    if (nk == 8)
                    {
                        const int4 x = __ldg(reinterpret_cast<const int4 *>(src));
                        const int4 y = __ldg(reinterpret_cast<const int4 *>(src) + 1);
                        rp[0] = x.x; rp[1] = x.y; rp[2] = x.z; rp[3] = x.w;
                        rp[4] = y.x; rp[5] = y.y; rp[6] = y.z; rp[7] = y.w;
                    }
                    else
                    {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) if (kk < nk) rp[kk] = __ldg(src + kk);
                    }
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                {
                    if (kk >= nk || !((rowmask >> kk) & 1u)) continue;               // Genus.h:624-625 (warp-uniform)
                    const bool nz = ((kk < 4 ? nl : nh) >> (8 * (kk & 3))) & 1u;
                    const unsigned ballot = __ballot_sync(0xffffffffu, nz);
                    if (nz)
                    {
                        const u32 odd = ((kk < 4 ? lo : hi) >> (8 * (kk & 3))) & 0xffu;
                        const int value = (int)run - 2 * (int)odd;
                        const i64 at = wbase[kk] + (i64)__popc(ballot & lt);
                        TB_CHECK(at >= 0 && at < L.extent && rp[kk] >= 0 && value != 0);
                        if (NARROW) { L.idx16[at] = (unsigned short)rp[kk]; L.dat8[at] = (signed char)value; }
                        else { L.data[at] = value; L.indices[at] = rp[kk]; }
                    }
                    wbase[kk] += (i64)__popc(ballot);
                }
            }
        }
    }
}

// In-place inclusive scan of each conductor's (rows_k + 1) block; one CTA per
// conductor.  Also emits nnz[k].
__global__ void indptr_scan_kernel(int *indptr_all, const int *rowbase, const int *rowcount, i64 *nnz)
{
    __shared__ int warp_sums[36];
    const int k = blockIdx.x;
    int *indptr = indptr_all + rowbase[k];
    const int rows = rowcount[k];
    int carry = 0;
    for (int base = 0; base < rows; base += blockDim.x)
    {
        int i = base + threadIdx.x;
        int c = i < rows ? indptr[i + 1] : 0;
        int total;
        int ex = block_exclusive_scan(c, warp_sums, total);
        if (i < rows) indptr[i + 1] = carry + ex + c;
        carry += total;
    }
    if (threadIdx.x == 0) { indptr[0] = 0; nnz[k] = carry; }
}

// int32 arrays from the narrow ones (device consumers of a result the row kernel wrote narrow only).
__global__ void widen_device_kernel(const unsigned short *idx16, const signed char *dat8, size_t n, int *indices, int *data)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        indices[i] = (int)idx16[i];
        data[i] = (int)dat8[i];
    }
}

// The 3-byte form from int32 arrays (final gather of a result the row kernel wrote wide).
__global__ void narrow_device_kernel(const int *indices, const int *data, size_t n, unsigned short *idx16, signed char *dat8)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        idx16[i] = (unsigned short)indices[i];
        dat8[i] = (signed char)data[i];
    }
}

// All conductors' data / indices, packed back to back (conductor-major) for a single read-back.
// grid.y = conductor; src offsets are the per-conductor regions, dst offsets the running totals.
__global__ void pack_all_kernel(const int *data, const int *indices, const unsigned short *idx16, const signed char *dat8,
                                const i64 *nnzbase, const i64 *nnz, const i64 *dstbase, int *data_out, int *indices_out)
{
    const int k = blockIdx.y;
    const i64 n = nnz[k], src = nnzbase[k], dst = dstbase[k];
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        if (data)
        {
            data_out[dst + i] = data[src + i];
            indices_out[dst + i] = indices[src + i];
        }
        else
        {
            data_out[dst + i] = (int)dat8[src + i];
            indices_out[dst + i] = (int)idx16[src + i];
        }
    }
}

// Dense rows from the CSR rows of one conductor.
__global__ void dense_scatter_kernel(const int *indptr, const int *data, const int *indices, int rows, int dim, int *matrix)
{
    const int row = blockIdx.x;
    if (row >= rows) return;
    for (int i = indptr[row] + threadIdx.x; i < indptr[row + 1]; i += blockDim.x)
        matrix[(size_t)row * dim + indices[i]] = data[i];
}

// ---------------------------------------------------------------------------
// Eigenvalue accumulation, all set-cover representatives in one launch.  Genus.h:490-499.
//   out[v] += chi_{cond[v]}(spin) * vecs[v][r]   over the neighbors of rep_of_vec[v]
// grid.y = row (one representative of rep_list per row), threads over t.
// ---------------------------------------------------------------------------
__global__ void eigen_dot_kernel(const u32 *W, u32 P1, int K, int N, const int *vecs, const i64 *cond,
                                 const int *rep_of_vec, int V, const int *rep_list, int *out)
{
    const int row = blockIdx.y;
    const int n = rep_list[row];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = t < P1;
    const u32 w = live ? W[(size_t)row * P1 + t] : 0u;
    const u32 r = w >> K, spin = w & ((1u << K) - 1u);
    live = live && r < (u32)N;
    for (int v = 0; v < V; ++v)
    {
        if (rep_of_vec[v] != n) continue;                    // block-uniform
        int term = 0;
        if (live)
        {
            const int coord = vecs[(size_t)v * N + r];
            term = (__popc(spin & (u32)cond[v]) & 1) ? -coord : coord;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) term += __shfl_xor_sync(0xffffffffu, term, d);
        if ((threadIdx.x & 31) == 0 && term != 0) atomicAdd(out + v, term);
    }
}

// ---------------------------------------------------------------------------
// Integer-pipe microbenchmark: the roofline denominator.  Every step below is ONE PTX
// instruction in an asm volatile statement, which ptxas turns into exactly one SASS
// instruction (tests/test_cabi.py disassembles the library and checks it), so the operation
// count of a launch is the number of statements executed, not an estimate from the source.
//   KIND 0  IMAD only                      (fma pipe: 16 lanes/clk/SMSP)
//   KIND 1  LOP3 only                      (alu pipe: 16 lanes/clk/SMSP)
//   KIND 2  ADD only: ptxas itself splits these between IADD3 (alu) and IMAD.IADD (fma)
//   KIND 3  IMAD and ADD alternating
//   KIND 4  IMAD and LOP3 alternating
//   KIND 5  DFMA only                      (fp64 pipe)
//   KIND 6  DFMA, IMAD and LOP3 in turn    (three pipes: does a DFMA cost more than one issue slot?)
// KIND 2-4 use both pipes and are bounded by the issue slot (1 warp instruction/clk/SMSP =
// 148 x 4 x 32 x 1.965 GHz = 37.2 T int32-op/s nominal).  16 independent chains per thread.
// The reduce loop of neighbors_kernel is half DFMA/DADD/DSETP, so KIND 5-6 say what the issue
// slot can deliver for THAT mix.
// ---------------------------------------------------------------------------
#define TB_PEAK_DFMA(z) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(z) : "d"(dm), "d"(dc))
#define TB_PEAK_IMAD(y) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y) : "r"(m), "r"(c))
#define TB_PEAK_ADD(x, z) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(z))
#define TB_PEAK_LOP(x, z) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(z), "r"(m))
#define TB_PEAK_UNROLL 16          /* steps per loop iteration */
#define TB_PEAK_OPS_SINGLE 8       /* instructions per step, KIND 0-2 */
#define TB_PEAK_OPS_MIXED 16       /* instructions per step, KIND 3-4 */
#define TB_PEAK_OPS_TRIPLE 24      /* instructions per step, KIND 6 */

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(u32 *sink, int iters, u32 seed)
{
    u32 x0 = seed + threadIdx.x, x1 = x0 * 3u + 1u, x2 = x0 * 5u + 2u, x3 = x0 * 7u + 3u;
    u32 x4 = x0 * 11u + 4u, x5 = x0 * 13u + 5u, x6 = x0 * 17u + 6u, x7 = x0 * 19u + 7u;
    u32 y0 = x0 ^ 1u, y1 = x1 ^ 2u, y2 = x2 ^ 3u, y3 = x3 ^ 4u, y4 = x4 ^ 5u, y5 = x5 ^ 6u, y6 = x6 ^ 7u, y7 = x7 ^ 8u;
    const u32 m = seed | 1u, c = seed ^ 0x9e3779b9u;
    double z0 = (double)x0, z1 = (double)x1, z2 = (double)x2, z3 = (double)x3, z4 = (double)x4, z5 = (double)x5, z6 = (double)x6, z7 = (double)x7;
    const double dm = 1.0 - 1.0 / (double)(seed | 1u), dc = (double)(seed & 7u);
#pragma unroll 1
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int u = 0; u < TB_PEAK_UNROLL; ++u)
        {
            if (KIND == 5)
            {
                TB_PEAK_DFMA(z0); TB_PEAK_DFMA(z1); TB_PEAK_DFMA(z2); TB_PEAK_DFMA(z3);
                TB_PEAK_DFMA(z4); TB_PEAK_DFMA(z5); TB_PEAK_DFMA(z6); TB_PEAK_DFMA(z7);
            }
            if (KIND == 6)
            {
                TB_PEAK_DFMA(z0); TB_PEAK_IMAD(y0); TB_PEAK_LOP(x0, x1); TB_PEAK_DFMA(z1); TB_PEAK_IMAD(y1); TB_PEAK_LOP(x1, x2);
                TB_PEAK_DFMA(z2); TB_PEAK_IMAD(y2); TB_PEAK_LOP(x2, x3); TB_PEAK_DFMA(z3); TB_PEAK_IMAD(y3); TB_PEAK_LOP(x3, x4);
                TB_PEAK_DFMA(z4); TB_PEAK_IMAD(y4); TB_PEAK_LOP(x4, x5); TB_PEAK_DFMA(z5); TB_PEAK_IMAD(y5); TB_PEAK_LOP(x5, x6);
                TB_PEAK_DFMA(z6); TB_PEAK_IMAD(y6); TB_PEAK_LOP(x6, x7); TB_PEAK_DFMA(z7); TB_PEAK_IMAD(y7); TB_PEAK_LOP(x7, x0);
            }
            if (KIND == 0)
            {
                TB_PEAK_IMAD(y0); TB_PEAK_IMAD(y1); TB_PEAK_IMAD(y2); TB_PEAK_IMAD(y3);
                TB_PEAK_IMAD(y4); TB_PEAK_IMAD(y5); TB_PEAK_IMAD(y6); TB_PEAK_IMAD(y7);
            }
            if (KIND == 1)
            {
                TB_PEAK_LOP(x0, x1); TB_PEAK_LOP(x1, x2); TB_PEAK_LOP(x2, x3); TB_PEAK_LOP(x3, x4);
                TB_PEAK_LOP(x4, x5); TB_PEAK_LOP(x5, x6); TB_PEAK_LOP(x6, x7); TB_PEAK_LOP(x7, x0);
            }
            if (KIND == 2)
            {
                TB_PEAK_ADD(x0, x1); TB_PEAK_ADD(x1, x2); TB_PEAK_ADD(x2, x3); TB_PEAK_ADD(x3, x4);
                TB_PEAK_ADD(x4, x5); TB_PEAK_ADD(x5, x6); TB_PEAK_ADD(x6, x7); TB_PEAK_ADD(x7, x0);
            }
            if (KIND == 3)
            {
                TB_PEAK_IMAD(y0); TB_PEAK_ADD(x0, x1); TB_PEAK_IMAD(y1); TB_PEAK_ADD(x1, x2);
                TB_PEAK_IMAD(y2); TB_PEAK_ADD(x2, x3); TB_PEAK_IMAD(y3); TB_PEAK_ADD(x3, x4);
                TB_PEAK_IMAD(y4); TB_PEAK_ADD(x4, x5); TB_PEAK_IMAD(y5); TB_PEAK_ADD(x5, x6);
                TB_PEAK_IMAD(y6); TB_PEAK_ADD(x6, x7); TB_PEAK_IMAD(y7); TB_PEAK_ADD(x7, x0);
            }
            if (KIND == 4)
            {
                TB_PEAK_IMAD(y0); TB_PEAK_LOP(x0, x1); TB_PEAK_IMAD(y1); TB_PEAK_LOP(x1, x2);
                TB_PEAK_IMAD(y2); TB_PEAK_LOP(x2, x3); TB_PEAK_IMAD(y3); TB_PEAK_LOP(x3, x4);
                TB_PEAK_IMAD(y4); TB_PEAK_LOP(x4, x5); TB_PEAK_IMAD(y5); TB_PEAK_LOP(x5, x6);
                TB_PEAK_IMAD(y6); TB_PEAK_LOP(x6, x7); TB_PEAK_IMAD(y7); TB_PEAK_LOP(x7, x0);
            }
        }
    }
    u32 r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
    r ^= (u32)__double2hiint(((z0 + z1) + (z2 + z3)) + ((z4 + z5) + (z6 + z7)));
    if (r == 0x12345678u) sink[0] = r;
}

}  // namespace tb
