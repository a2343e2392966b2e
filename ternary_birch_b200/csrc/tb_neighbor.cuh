// tb_neighbor.cuh -- one p-neighbor, start to finish, in registers (sm_100a).
//
// Device-side building blocks of the hot path:
//   line_for_t        t -> isotropic line mod p      NeighborManager.h:50-144, 370-394
//   build_neighbor    lattice + isometry             NeighborManager.h:207-356
//   fast_setup        both of the above in binary64 on the generic lane, under host-proven bounds
//   eisenstein_reduce unique reduced representative  QuadForm.h:530-767   (the literal integer loop)
//   reduce_loop       the same loop, exact in binary64 then binary32; isometry as DIso or IIso
//   reduce_boundary   post-loop fix-ups              QuadForm.h:693-764
//   genus_lookup      open-addressing probe          HashMap.h:68-89, QuadForm.cpp:779-803
//   spinor_norm       spinor/Kronecker class         Spinor.h:19-66, Genus.h:582-615
//   one_neighbor      the loop body of Genus.h:567-618
// One thread owns one neighbor.  The integer code is templated on the integer mode Z
// (W64 | X64 | C128 | U128, tb_int.cuh).
#pragma once

#include "tb_int.cuh"

#ifndef TB_CHECK            /* defined by tb_kernels.cuh: assert under -DTB_DEBUG_BOUNDS, nothing otherwise */
#define TB_CHECK(cond) ((void)0)
#endif

namespace tb {

// ---------------------------------------------------------------------------
// Launch-invariant context.
// ---------------------------------------------------------------------------
#define TB_MAX_PRIMES 16

// Per-representative record, one 128-byte line each.
//   cur[n]: q[6], scalar, s[9]      (data of the representative being expanded)
//   rep[r]: q[6], scalar, sinv[9]   (data of the representative that is hit)
#define TB_REC_WORDS 16

// Per (representative, prime) record written by prep_reps_kernel.
struct RepPrime {
    u32 a, b, h;      // coefficients mod p
    u32 x0, y0;       // conic base point (x0, y0, 1); p = 2: packed vectors 0 and 1
    u32 a0, delta;    // NeighborManager.h:34-47; p = 2: a0 = packed vector 2
    u32 pad;
};

struct PrimeCtx {
    u32 p;
    u32 use_lut;
    u32 fp64_reduce;  // host-verified: isometry entries stay below 2^52 in the FP64 reduce loop
    u32 m32;          // floor(2^32 / p) when p < 2^16 (32-bit Barrett), else 0
    u32 narrow_trace; // two-limb modes: composed isometry / denominator proven below 2^61 (64-bit trace)
    u32 proven_small; // every value build_neighbor hands to the reduction is proven below 2^50 (no per-lane test)
    u32 iso32;        // table isometries and reduced neighbor isometries proven below 2^31 (32-bit products, int32 reduce isometry)
    u32 f32_tail;     // fp64_reduce and small table forms (c < 2^24, b < 2^20): the reduce loop finishes in binary32
    u32 fast_setup;   // bit 0: the binary64 neighbor setup (fast_setup below) is exact for this launch; bit 1: the
    u32 pad4;         //   reference's int64 arithmetic provably does not wrap, so int64 mode may use it (setup_prime)
    u64 pp;
    double pd, ppd;   // p, p^2
    double inv_pd, inv_ppd;   // RN(1 / p), RN(1 / p^2)
    u64 pinv64;       // p^-1 mod 2^64 (odd p), else 0
    InvDiv dp;        // division by p
    InvDiv dpp;       // division by p^2
    const u32 *inv_lut;   // [p] inverses mod p, or nullptr
};

struct GenusCtx {
    int N;
    int K;
    u32 slot_mask;
    u32 pad;
    const i64 *cur;       // [N][16]
    const i64 *rep;       // [N][16]
    const i64 *disc;      // [N] wrap-around discriminant of each representative
    const int *slots;     // [slot_mask+1] representative index or -1
    InvDiv spin[TB_MAX_PRIMES];   // the level's prime divisors
    u64 spin_inv[TB_MAX_PRIMES];  // d^-1 mod 2^64 (0 for d = 2)
    u64 spin_qmax[TB_MAX_PRIMES]; // floor((2^64 - 1) / d)
};

template <class Z>
struct Form {
    Z a, b, c, f, g, h;
};

// 3x3 integer matrix, row-major (Isometry.h:350-354).  The methods are the
// column operations reduce/build apply; the comment gives the reference's name.
template <class Z>
struct Iso {
    Z a11, a12, a13, a21, a22, a23, a31, a32, a33;

    // col3 += col1 + col2                                    A101011001
    __device__ __forceinline__ void c3_add_c1_c2() { a13 += a11 + a12; a23 += a21 + a22; a33 += a31 + a32; }
    // col2 += t col1                                          A1t0010001
    __device__ __forceinline__ void c2_add_c1(Z t) { a12 += t * a11; a22 += t * a21; a32 += t * a31; }
    // col3 += t col2                                          A10001t001
    __device__ __forceinline__ void c3_add_c2(Z t) { a13 += t * a12; a23 += t * a22; a33 += t * a32; }
    // col3 += t col1                                          A10t010001
    __device__ __forceinline__ void c3_add_c1(Z t) { a13 += t * a11; a23 += t * a21; a33 += t * a31; }
    // (c1, c2, c3) <- (-c2, -c1, -c3)                         A0n0n0000n
    __device__ __forceinline__ void swap12_neg()
    {
        Z t;
        t = -a11; a11 = -a12; a12 = t; a13 = -a13;
        t = -a21; a21 = -a22; a22 = t; a23 = -a23;
        t = -a31; a31 = -a32; a32 = t; a33 = -a33;
    }
    // (c1, c2, c3) <- (-c1, -c3, -c2)                         An0000n0n0
    __device__ __forceinline__ void swap23_neg()
    {
        Z t;
        t = -a12; a12 = -a13; a13 = t; a11 = -a11;
        t = -a22; a22 = -a23; a23 = t; a21 = -a21;
        t = -a32; a32 = -a33; a33 = t; a31 = -a31;
    }
    __device__ __forceinline__ void neg_c1() { a11 = -a11; a21 = -a21; a31 = -a31; }   // An00010001
    __device__ __forceinline__ void neg_c2() { a12 = -a12; a22 = -a22; a32 = -a32; }   // A1000n0001
    __device__ __forceinline__ void neg_c3() { a13 = -a13; a23 = -a23; a33 = -a33; }   // A10001000n
    // c3 += c1 + c2; c1 = -c1; c2 = -c2                       An010n1001
    __device__ __forceinline__ void fix_sum()
    {
        a13 += a11 + a12; a11 = -a11; a12 = -a12;
        a23 += a21 + a22; a21 = -a21; a22 = -a22;
        a33 += a31 + a32; a31 = -a31; a32 = -a32;
    }
    // c2 = -(c2 + c1); c1 = -c1                               Ann00n0001
    __device__ __forceinline__ void fix_ah()
    {
        a12 += a11; a12 = -a12; a11 = -a11;
        a22 += a21; a22 = -a22; a21 = -a21;
        a32 += a31; a32 = -a32; a31 = -a31;
    }
    // c3 = -(c3 + c1); c1 = -c1                               An0n01000n
    __device__ __forceinline__ void fix_ag()
    {
        a13 += a11; a13 = -a13; a11 = -a11;
        a23 += a21; a23 = -a23; a21 = -a21;
        a33 += a31; a33 = -a33; a31 = -a31;
    }
    // c3 = -(c3 + c2); c2 = -c2                               A1000nn00n
    __device__ __forceinline__ void fix_bf()
    {
        a13 += a12; a13 = -a13; a12 = -a12;
        a23 += a22; a23 = -a23; a22 = -a22;
        a33 += a32; a33 = -a33; a32 = -a32;
    }
    // c2 -= c1; c1 = -c1; c3 = -c3                            Ann001000n
    __device__ __forceinline__ void edge_ah()
    {
        a12 -= a11; a11 = -a11; a13 = -a13;
        a22 -= a21; a21 = -a21; a23 = -a23;
        a32 -= a31; a31 = -a31; a33 = -a33;
    }
    // c3 -= c1; c1 = -c1; c2 = -c2                            An0n0n0001
    __device__ __forceinline__ void edge_ag()
    {
        a13 -= a11; a11 = -a11; a12 = -a12;
        a23 -= a21; a21 = -a21; a22 = -a22;
        a33 -= a31; a31 = -a31; a32 = -a32;
    }
    // c3 -= c2; c2 = -c2; c1 = -c1                            An000nn001
    __device__ __forceinline__ void edge_bf()
    {
        a13 -= a12; a12 = -a12; a11 = -a11;
        a23 -= a22; a22 = -a22; a21 = -a21;
        a33 -= a32; a32 = -a32; a31 = -a31;
    }
    // (c2, c3) <- (-c3, c2)                                   A1000010n0
    __device__ __forceinline__ void rot23()
    {
        Z t;
        t = a13; a13 = a12; a12 = -t;
        t = a23; a23 = a22; a22 = -t;
        t = a33; a33 = a32; a32 = -t;
    }
    // col2 += t col3                                          A1000100t1
    __device__ __forceinline__ void c2_add_c3(Z t) { a12 += a13 * t; a22 += a23 * t; a32 += a33 * t; }
    // col1 += t col3                                          A100010t01
    __device__ __forceinline__ void c1_add_c3(Z t) { a11 += a13 * t; a21 += a23 * t; a31 += a33 * t; }
    // col2 *= p, col3 *= p^2                                  A1000p000p2
    __device__ __forceinline__ void scale_p(Z p, Z pp)
    {
        a12 *= p; a22 *= p; a32 *= p;
        a13 *= pp; a23 *= pp; a33 *= pp;
    }
};

// ---------------------------------------------------------------------------
// Fp on canonical residues, p < 2^31.
// ---------------------------------------------------------------------------
// The general paths (p >= 2^16: 64-bit product and 2-by-1 division; p >= 2^22 or no table: Fermat
// inverse) live out of line: inlined they put ~50 resp. ~110 cold instructions after every hot
// multiplication / inversion, and the kernel is bound by instruction fetch.
__device__ __noinline__ u32 fp_mul_wide(u32 a, u32 b, const PrimeCtx &pc)
{
    u64 r;
    udivmod((u64)a * (u64)b, pc.dp, r);
    return (u32)r;
}
__device__ __forceinline__ u32 fp_mul(u32 a, u32 b, const PrimeCtx &pc)
{
    if (pc.m32)
    {
        // p < 2^16: the product fits 32 bits; Barrett quotient is floor(x/p) or one less
        const u32 x = a * b;
        u32 r = x - __umulhi(x, pc.m32) * pc.p;
        return r >= pc.p ? r - pc.p : r;
    }
    return fp_mul_wide(a, b, pc);
}
__device__ __forceinline__ u32 fp_add(u32 a, u32 b, u32 p) { u32 s = a + b; return s >= p ? s - p : s; }
__device__ __forceinline__ u32 fp_sub(u32 a, u32 b, u32 p) { return a >= b ? a - b : a + p - b; }
__device__ __forceinline__ u32 fp_pow(u32 a, u32 e, const PrimeCtx &pc)
{
    u32 r = 1;
    while (e)
    {
        if (e & 1u) r = fp_mul(r, a, pc);
        a = fp_mul(a, a, pc);
        e >>= 1;
    }
    return r;
}
// Fp::inverse (Fp.h:147-150): table below 2^22 (built per prime by
// inverse_lut_kernel), Fermat above.  F2::inverse for p = 2.
__device__ __noinline__ u32 fp_inv_general(u32 a, const PrimeCtx &pc)
{
    if (pc.p == 2u) return a & 1u;
    return a ? fp_pow(a, pc.p - 2u, pc) : 0u;
}
__device__ __forceinline__ u32 fp_inv(u32 a, const PrimeCtx &pc)
{
    if (pc.use_lut) return __ldg(pc.inv_lut + a);      // use_lut implies p > 2
    return fp_inv_general(a, pc);
}
// Fp::mod of a signed 64-bit coefficient (Fp.h:29-35).
__device__ __forceinline__ u32 fp_mod_i64(i64 x, const PrimeCtx &pc)
{
    u64 r;
    udivmod(x < 0 ? 0ull - (u64)x : (u64)x, pc.dp, r);
    if (x < 0 && r != 0) r = pc.p - r;
    return (u32)r;
}

// ---------------------------------------------------------------------------
// t -> isotropic line.  NeighborManager::isotropic_vector, NeighborManager.h:50-144;
// p = 2: isotropic_vector_p2, 370-394.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void line_for_t(const RepPrime &rp, u32 t, const PrimeCtx &pc,
                                           u32 &vx, u32 &vy, u32 &vz)
{
    const u32 p = pc.p;
    if (p == 2u)
    {
        u32 w = t == 0 ? rp.x0 : (t == 1 ? rp.y0 : rp.a0);
        vz = w & 1u; vy = (w >> 1) & 1u; vx = (w >> 2) & 1u;
        return;
    }
    if (t == p)
    {
        u32 at = fp_sub(rp.delta, rp.h, p);
        if (at == 0) { vx = rp.x0; vy = rp.y0 ? rp.y0 - 1 : p - 1; vz = 1; }
        else if (rp.b == 0) { vx = 0; vy = 1; vz = 0; }
        else
        {
            u32 inv = fp_mul(fp_inv(rp.b, pc), at, pc);
            vx = rp.x0; vy = fp_add(rp.y0, fp_sub(inv, 1, p), p); vz = 1;
        }
        return;
    }
    // at = ((t-1) b + delta) t + a0   (equals a0 at t = 0 and a0 + delta at t = 1)
    u32 at = fp_mul(fp_add(fp_mul(fp_sub(t, 1, p), rp.b, pc), rp.delta, p), t, pc);
    at = fp_add(at, rp.a0, p);
    if (at == 0)
    {
        vx = rp.x0 ? rp.x0 - 1 : p - 1; vy = fp_sub(rp.y0, t, p); vz = 1;
        return;
    }
    // ct = (b t + h) t + a = Q(1, t, 0)
    u32 ct = fp_add(fp_mul(fp_add(fp_mul(rp.b, t, pc), rp.h, p), t, pc), rp.a, p);
    if (ct == 0)
    {
        if (t == 0) { vx = 1; vy = 0; vz = 0; }
        else { vx = fp_inv(t, pc); vy = 1; vz = 0; }
        return;
    }
    u32 inv = fp_mul(at, fp_inv(ct, pc), pc);
    vx = fp_add(rp.x0, fp_sub(inv, 1, p), p);
    vy = fp_add(fp_sub(rp.y0, t, p), fp_mul(t, inv, pc), p);
    vz = 1;
}

// ---------------------------------------------------------------------------
// Neighbor lattice.  NeighborManager::build_neighbor, NeighborManager.h:207-356.
// Returns false when the int64 discriminant check (344-354) fails.
// ---------------------------------------------------------------------------
template <class Z>
struct CheckDisc { static const bool value = false; };
template <>
struct CheckDisc<W64> { static const bool value = true; };
// X64: 64-bit arithmetic under host-proven bounds (exact, so no discriminant check)
template <class Z>
struct IsExact64 { static const bool value = false; };
template <>
struct IsExact64<X64> { static const bool value = true; };

template <class Z>
__device__ __forceinline__ Z form_disc(const Form<Z> &q)
{
    // QuadForm.h:27-32
    return q.a * (Z(4) * q.b * q.c - q.f * q.f) - q.b * q.g * q.g + q.h * (q.f * q.g - q.c * q.h);
}

template <class Z>
__device__ __forceinline__ bool build_neighbor(const Form<Z> &q, u32 vx, u32 vy, u32 vz, i64 disc64,
                                               const PrimeCtx &pc, Form<Z> &out, Iso<Z> &s)
{
    const Z p = Z((i64)pc.p);
    const Z pp = Z((i64)pc.pp);
    const i64 half = (i64)(pc.p >> 1);
    Z aa, bb, cc, ff, gg, hh;

    // balanced lifts of the line's coordinates (223-224)
    i64 ux = (i64)vx, uy = (i64)vy;
    if (ux > half) ux -= (i64)pc.p;
    if (uy > half) uy -= (i64)pc.p;
    const Z zero = Z(0), one = Z(1), mone = Z(-1);

    if (vz == 1u)                                            // 226-245
    {
        Z u = Z(ux), v = Z(uy);
        s.a11 = u; s.a12 = zero; s.a13 = mone;
        s.a21 = v; s.a22 = one; s.a23 = zero;
        s.a31 = one; s.a32 = zero; s.a33 = zero;
        Z au = q.a * u, hu = q.h * u, hv = q.h * v, bv = q.b * v;
        Z temp = au + hv + q.g;
        aa = q.c + u * temp + v * (bv + q.f);
        bb = q.b;
        cc = q.a;
        ff = -q.h;
        gg = -(temp + au);
        hh = q.f + hu + bv.twice();
    }
    else if (uy == 1)                                        // 246-261
    {
        Z u = Z(ux);
        s.a11 = u; s.a12 = zero; s.a13 = one;
        s.a21 = one; s.a22 = zero; s.a23 = zero;
        s.a31 = zero; s.a32 = one; s.a33 = zero;
        Z au = q.a * u, gu = q.g * u;
        Z temp = au + q.h;
        aa = q.b + u * temp;
        bb = q.c;
        cc = q.a;
        ff = q.g;
        gg = temp + au;
        hh = q.f + gu;
    }
    else                                                     // 262-272
    {
        s.a11 = one; s.a12 = zero; s.a13 = zero;
        s.a21 = zero; s.a22 = zero; s.a23 = mone;
        s.a31 = zero; s.a32 = one; s.a33 = zero;
        aa = q.a; bb = q.c; cc = q.b;
        ff = -q.f; gg = -q.h; hh = q.g;
    }

    Z gmodp = trunc_rem(gg, pc.dp);                          // 279-295
    if (gmodp.is_zero())
    {
        s.rot23();
        Z t = bb; bb = cc; cc = t;
        t = gg; gg = hh; hh = -t;
        ff = -ff;
        gmodp = trunc_rem(gg, pc.dp);
    }
    if (gmodp.is_neg()) gmodp += p;                          // 302
    const Z ginv = Z((i64)fp_inv((u32)gmodp.low64(), pc));   // 303

    Z s1 = trunc_rem((-hh) * ginv, pc.dp);                   // 309-312
    // X64: reduce aa first so the product stays below p^3 (same truncated remainder)
    Z s2 = IsExact64<Z>::value ? trunc_rem(trunc_rem(-aa, pc.dpp) * ginv, pc.dpp) : trunc_rem((-aa) * ginv, pc.dpp);
    if (s1 < Z(-half)) s1 += p;
    if (s2 < Z(-(i64)(pc.pp >> 1))) s2 += pp;

    s.c2_add_c3(s1);                                         // 314-319
    Z temp1 = cc * s1;
    bb += s1 * (ff + temp1);
    ff += temp1.twice();
    hh += s1 * gg;

    s.c1_add_c3(s2);                                         // 325-329
    Z temp2 = cc * s2;
    aa = muladd_div(aa, s2, gg + temp2, pc.dpp);             // aa += s2 * (gg + temp2); ... aa /= pp  (341)
    gg += temp2.twice();
    hh = muladd_div(hh, s2, ff, pc.dp);                      // hh += s2 * ff; ... hh /= p

    s.scale_p(p, pp);                                        // 337-341
    cc *= pp;
    ff *= p;

    out.a = aa; out.b = bb; out.c = cc; out.f = ff; out.g = gg; out.h = hh;

    if (CheckDisc<Z>::value)                                 // 344-354
        return form_disc(out).low64() == disc64;
    return true;
}

// ---------------------------------------------------------------------------
// Eisenstein reduction with isometry tracking.  QuadForm<R>::reduce,
// QuadForm.h:530-767 (GMP twin: QuadForm.cpp:69-379).  The statement order is
// the reference's: the isometry (not the reduced form) depends on it.
// Returns false when the iteration cap trips (wrapped garbage only).
// ---------------------------------------------------------------------------
#define TB_REDUCE_CAP 256

template <class Z>
__device__ __forceinline__ Z shear_quotient(Z a, Z h)
{
    // t = a >= h ? (a-h) div 2a : -((a+h-1) div 2a)         551-558
    Z den = a.twice();
    if (a >= h) return ref_quotient(a - h, den);
    return -ref_quotient(a + h - Z(1), den);
}

template <class Z>
__device__ __noinline__ bool eisenstein_reduce(Form<Z> &q, Iso<Z> &s)
{
    Z a = q.a, b = q.b, c = q.c, f = q.f, g = q.g, h = q.h;
    Z t;
    const Z zero = Z(0);
    int iters = 0;
    bool again = true;
    while (again)
    {
        if (++iters > TB_REDUCE_CAP) return false;

        t = a + b + f + g + h;                               // 542-549
        if (t < zero)
        {
            s.c3_add_c1_c2();
            c += t;
            f += h + b.twice();
            g += h + a.twice();
        }

        t = shear_quotient(a, h);                            // 551-567
        if (!t.is_zero())
        {
            s.c2_add_c1(t);
            Z temp = a * t;
            h += temp;
            b += h * t;
            f += g * t;
            h += temp;
        }

        t = shear_quotient(b, f);                            // 569-585
        if (!t.is_zero())
        {
            s.c3_add_c2(t);
            Z temp = b * t;
            f += temp;
            c += f * t;
            g += h * t;
            f += temp;
        }

        t = shear_quotient(a, g);                            // 587-603
        if (!t.is_zero())
        {
            s.c3_add_c1(t);
            Z temp = a * t;
            g += temp;
            c += g * t;
            f += h * t;
            g += temp;
        }

        if (a > b || (a == b && f.abs() > g.abs()))          // 605-610
        {
            s.swap12_neg();
            t = a; a = b; b = t;
            t = f; f = g; g = t;
        }
        if (b > c || (b == c && g.abs() > h.abs()))          // 612-617
        {
            s.swap23_neg();
            t = b; b = c; c = t;
            t = g; g = h; h = t;
        }
        if (a > b || (a == b && f.abs() > g.abs()))          // 619-624
        {
            s.swap12_neg();
            t = a; a = b; b = t;
            t = f; f = g; g = t;
        }

        // sign normalisation                                  626-687
        const bool fneg = f.is_neg(), gneg = g.is_neg(), hneg = h.is_neg();
        const bool fz = f.is_zero(), gz = g.is_zero(), hz = h.is_zero();
        bool fgh = !fz && !gz && !hz;
        if (fgh) fgh = ((int)fneg + (int)gneg + (int)hneg) % 2 == 0;
        bool n1, n2, n3;
        if (fgh) { n1 = fneg; n2 = gneg; n3 = hneg; }
        else
        {
            int s1 = !fneg && !fz, s2 = !gneg && !gz, s3 = !hneg && !hz;
            if ((s1 + s2 + s3) % 2 == 1)
            {
                if (fz) s1 = 1;
                else if (gz) s2 = 1;
                else if (hz) s3 = 1;
            }
            n1 = s1 == 1; n2 = s2 == 1; n3 = s3 == 1;
        }
        if (n1) { s.neg_c1(); f = -f; }
        if (n2) { s.neg_c2(); g = -g; }
        if (n3) { s.neg_c3(); h = -h; }

        again = !(f.abs() <= b && g.abs() <= a && h.abs() <= a && a + b + f + g + h >= zero);   // 689-690
    }

    Z sum = a + b + f + g + h;
    if (sum.is_zero() && a.twice() + g.twice() + h > zero)   // 693-700
    {
        s.fix_sum();
        c += sum;
        f += h + b.twice(); f = -f;
        g += h + a.twice(); g = -g;
    }
    if (a == -h && !g.is_zero())                             // 702-708
    {
        s.fix_ah();
        f += g; f = -f;
        g = -g;
        h = -h;
    }
    if (a == -g && !h.is_zero())                             // 710-716
    {
        s.fix_ag();
        f += h; f = -f;
        h = -h;
        g += a.twice();
    }
    if (b == -f && !h.is_zero())                             // 718-724
    {
        s.fix_bf();
        g += h; g = -g;
        h = -h;
        f += b.twice();
    }
    if (a == h && g > f.twice()) { s.edge_ah(); f = g - f; }   // 726-730
    if (a == g && h > f.twice()) { s.edge_ag(); f = h - f; }   // 732-736
    if (b == f && h > g.twice()) { s.edge_bf(); g = h - g; }   // 738-742
    if (a == b && f.abs() > g.abs())                         // 744-750
    {
        s.swap12_neg();
        t = a; a = b; b = t;
        t = g; g = f; f = t;
    }
    if (b == c && g.abs() > h.abs())                         // 752-757
    {
        s.swap23_neg();
        t = g; g = h; h = t;
    }
    if (a == b && f.abs() > g.abs())                         // 759-764
    {
        s.swap12_neg();
        t = g; g = f; f = t;
    }

    q.a = a; q.b = b; q.c = c; q.f = f; q.g = g; q.h = h;
    return true;
}

// ---------------------------------------------------------------------------
// Floating-point form of the same reduction (all integer modes): reduce_trip / reduce_loop below.
//
// B200's FP64 pipe issues DFMA/DADD/DSETP at the same rate as each integer pipe
// and is otherwise idle in this kernel, while 64-bit integer compares, selects
// and negations saturate the ALU pipe.  Every quantity in the reduction loop is
// an integer, so it is computed EXACTLY in binary64 as long as it stays below
// 2^53: a multiply-add is one DFMA (the product is never rounded separately), a
// compare one DSETP (|x| is a free operand modifier), a sign flip one LOP3.
// Once the coefficients are below 2^20 the form part continues in binary32 (see
// "Binary32 tail" further down); the isometry is carried in binary64 with lazy
// signs (DIso) or, under a host-proven bound, in int32 ring arithmetic (IIso).
//
// Why the values stay small (positive definite form, exact arithmetic):
//   * diagonal coefficients are values Q(v) of the current basis vectors; each
//     shear replaces one by a not-larger value, swaps permute them and the
//     "sum < 0" step replaces c by Q(1,1,1) < c: they never grow;
//   * off-diagonals obey f^2 < 4bc, g^2 < 4ac, h^2 < 4ab, so |f|,|g|,|h| < 2 max(a,b,c);
//   * a shear's intermediates are a t (|t| <= (|h|+a)/2a + 1, so |a t| <= 2.5 max), h + a t (<= 4.5 max)
//     and FMA results that are themselves new coefficients; the fix step forms h + 2b (< 4 max) and
//     the running sum a + b + f + g + h stays below 8 max(a,b,c);
//   * isometry columns are lattice vectors v with Q_rep(v) = p^2 * (diagonal
//     coefficient), so |v_i| <= p sqrt(max(a,b,c) / lambda_min(Q_rep)); the host
//     enables this path only when that bound is below 2^52 (PrimeCtx::fp64_reduce).
// The loop is entered only when every lane's 15 inputs are below 2^50 in
// magnitude (warp vote); otherwise the warp takes the integer loop above.
//
// SIMT shape: the body is branch-free (shears with t = 0 are no-ops, swaps and
// sign flips are selects or predicated moves), so a warp diverges only on the
// trip count; finished lanes are frozen by zeroing their reciprocal seeds (all
// three quotients become 0) and masking the swap predicates.  The operation order
// is the reference's (QuadForm.h:540-691), so the isometry is bit-identical.
// ---------------------------------------------------------------------------
// 1 / d to ~2^-36 relative: the hardware's ~20-bit seed and ONE Newton step.  That is all the shear
// quotients need: on a positive definite form |h| < 2 sqrt(ab), |f| < 2 sqrt(bc), |g| < 2 sqrt(ac), so
// with coefficients below 2^50 every quotient (a-h)/2a, (b-f)/2b, (a-g)/2a is below 2^26 in magnitude
// and its estimate is within 2^-10 of the true value -- shear_quot's exact remainder then decides
// between floor and floor + 1.
// x with its sign bit XORed by m (m = 0 or 0x80000000): one LOP3
__device__ __forceinline__ double xflip(double x, u32 m)
{
    return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x));
}
// conditional swap, arithmetic form (FP64 pipe: DADD + 2 DFMA), m = 1.0 to swap, 0.0 to keep;
// exact for integer-valued operands with |y - x| < 2^53
__device__ __forceinline__ void dswap_fma(double m, double &x, double &y)
{
    const double d = y - x;
    x = fma(m, d, x);
    y = fma(-m, d, y);
}
// The reference's shear multiplier (QuadForm.h:551-558)
//     t = a >= h ? (a-h) div 2a : -((a+h-1) div 2a)
// is floor((a-h)/2a) in both branches (for h > a: -ceil((h-a)/2a) = -floor((h-a+2a-1)/2a)).
// q = RN(num * rden) by the 1.5*2^52 trick is floor or floor+1 (the estimate is within 1/2 of
// num/den for |num/den| < 2^49); the exact remainder tells which (shear_quot below).

// 2^50 magnitude test on an int64
__device__ __forceinline__ bool small50(i64 v) { return (u64)(v + (1ll << 50)) < (1ull << 51); }
template <int TAG>
__device__ __forceinline__ bool fits50(I64T<TAG> x) { return small50(x.v); }
template <bool CHK>
__device__ __forceinline__ bool fits50(I128<CHK> x) { return x.fits_i64() && small50((i64)x.lo); }


// Real-type helpers of the loop body: the form part of a trip runs in binary64 or, once every value
// is a small integer, in binary32 (reduce_loop_fp64's tail); the isometry stays binary64 throughout.
__device__ __forceinline__ double rmagic(double) { return 6755399441055744.0; }   // 1.5 * 2^52
__device__ __forceinline__ float rmagic(float) { return 12582912.0f; }            // 1.5 * 2^23
__device__ __forceinline__ double rcp_seeded(double d, bool on)
{
    // hardware seed with a finished lane's zeroed, one Newton step (0 stays 0 through the Newton step)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = __hiloint2double(on ? __double2hiint(r) : 0, 0);
    const double e = fma(-d, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float rcp_seeded(float d, bool on)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return on ? r : 0.0f;
}
__device__ __forceinline__ double radd(double x, double y) { return __dadd_rn(x, y); }
__device__ __forceinline__ float radd(float x, float y) { return __fadd_rn(x, y); }
__device__ __forceinline__ double rfma(double x, double y, double z) { return __fma_rn(x, y, z); }
__device__ __forceinline__ float rfma(float x, float y, float z) { return __fmaf_rn(x, y, z); }
__device__ __forceinline__ int rsign(double r) { return __double2hiint(r) >> 31; }    // -1 if r < 0 (r is never -0)
__device__ __forceinline__ int rsign(float r) { return __float_as_int(r) >> 31; }
__device__ __forceinline__ int rquot(double qm) { return __double2loint(qm); }        // int32(qm - magic)
__device__ __forceinline__ int rquot(float qm) { return __float_as_int(qm) - 0x4B400000; }
__device__ __forceinline__ double rcorr(double, int m) { return __hiloint2double(m & (int)0xBFF00000, 0); }   // -1.0 or 0.0
__device__ __forceinline__ float rcorr(float, int m) { return __int_as_float(m & (int)0xBF800000); }
// -x if p else x: a sign-bit XOR in binary64 (ALU, off the FP64 pipe), a predicated negation in binary32
__device__ __forceinline__ double rnegif(double x, bool p) { return xflip(x, p ? 0x80000000u : 0u); }
__device__ __forceinline__ float rnegif(float x, bool p) { return p ? -x : x; }
__device__ __forceinline__ double rabs(double x) { return fabs(x); }
__device__ __forceinline__ float rabs(float x) { return fabsf(x); }
__device__ __forceinline__ double rwide(double t) { return t; }
__device__ __forceinline__ double rwide(float t) { return (double)t; }

// floor((x - y) / den), see above; a finished lane has rden = 0, hence quotient 0 (its remainder
// x - y is a - h, b - f or a - g of a form that passed the exit test: never negative)
template <class F>
__device__ __forceinline__ F shear_quot(F x, F y, F den, F rden, int &ti)
{
    const F num = x - y;
    const F qm = rfma(num, rden, rmagic(F(0)));
    const F q = radd(qm, -rmagic(F(0)));
    const F r = rfma(-q, den, num);            // exact
    const int m = rsign(r);
    ti = rquot(qm) + m;                        // the same quotient as an int32 (free: the low word of q + 1.5 * 2^52)
    return q + rcorr(F(0), m);                 // q - 1 if r < 0 (sign mask of r ANDed into the bits of -1.0)
}
template <class F>
__device__ __forceinline__ void rswap(bool p, F &x, F &y)
{
    const F t = p ? y : x;
    y = p ? x : y;
    x = t;
}

// Isometry of the loop: binary64 entries with lazy column signs (actual column i = stored column i
// with its sign bit XOR sg_i).  The shears only ever need the pairwise products of the signs, so
// those are what is carried -- x12 = sg1^sg2, x23 = sg2^sg3, x13 = sg1^sg3, sign-bit masks -- next
// to sg1 itself; sg2 = sg1^x12 and sg3 = sg1^x13 are rebuilt once at the end.  A swap-and-negate
// of two columns permutes the pairwise masks (the common negation cancels in them):
//   (c1,c2,c3) <- (-c2,-c1,-c3):  x12 stays, x23 <-> x13, sg1 <- sg2 ^ S = sg1 ^ x12 ^ S
//   (c1,c2,c3) <- (-c1,-c3,-c2):  x23 stays, x12 <-> x13, sg1 <- sg1 ^ S
struct DIso {
    double s11, s12, s13, s21, s22, s23, s31, s32, s33;
    u32 sg1, x12, x23, x13;

    __device__ __forceinline__ void load(const double *v)
    {
        s11 = v[0]; s12 = v[1]; s13 = v[2]; s21 = v[3]; s22 = v[4]; s23 = v[5]; s31 = v[6]; s32 = v[7]; s33 = v[8];
        sg1 = x12 = x23 = x13 = 0u;
    }
    __device__ __forceinline__ void fix(bool on)                         // col3 += col1 + col2
    {
        const double fd = on ? 1.0 : 0.0;
        const double m13 = xflip(fd, x13), m23 = xflip(fd, x23);
        s13 = fma(m23, s12, fma(m13, s11, s13));
        s23 = fma(m23, s22, fma(m13, s21, s23));
        s33 = fma(m23, s32, fma(m13, s31, s33));
    }
    __device__ __forceinline__ void shear12(double t, int) { const double ts = xflip(t, x12); s12 = fma(ts, s11, s12); s22 = fma(ts, s21, s22); s32 = fma(ts, s31, s32); }
    __device__ __forceinline__ void shear23(double t, int) { const double ts = xflip(t, x23); s13 = fma(ts, s12, s13); s23 = fma(ts, s22, s23); s33 = fma(ts, s32, s33); }
    __device__ __forceinline__ void shear13(double t, int) { const double ts = xflip(t, x13); s13 = fma(ts, s11, s13); s23 = fma(ts, s21, s23); s33 = fma(ts, s31, s33); }
    __device__ __forceinline__ void swap12(bool p)                       // (c1,c2,c3) <- (-c2,-c1,-c3)
    {
        const double m = p ? 1.0 : 0.0;
        dswap_fma(m, s11, s12); dswap_fma(m, s21, s22); dswap_fma(m, s31, s32);
        const u32 o23 = x23;
        sg1 ^= p ? (x12 ^ 0x80000000u) : 0u; x23 = p ? x13 : x23; x13 = p ? o23 : x13;
    }
    __device__ __forceinline__ void swap23(bool p)                       // (c1,c2,c3) <- (-c1,-c3,-c2)
    {
        const double m = p ? 1.0 : 0.0;
        dswap_fma(m, s12, s13); dswap_fma(m, s22, s23); dswap_fma(m, s32, s33);
        const u32 o12 = x12;
        sg1 ^= p ? 0x80000000u : 0u; x12 = p ? x13 : x12; x13 = p ? o12 : x13;
    }
    __device__ __forceinline__ void negate(bool f1, bool f2, bool f3)    // which columns to negate
    {
        const u32 n1 = f1 ? 0x80000000u : 0u, n2 = f2 ? 0x80000000u : 0u, n3 = f3 ? 0x80000000u : 0u;
        sg1 ^= n1; x12 ^= n1 ^ n2; x23 ^= n2 ^ n3; x13 ^= n1 ^ n3;
    }
};

// The same isometry in int32 ring arithmetic (launches with PrimeCtx::iso32).  The column operations
// are ring operations (add, multiply by an integer), so they can run modulo 2^32: the host has proven
// that every entry of the isometry the loop ENDS with is below 2^31 in magnitude, hence the wrapped
// result is the result, whatever the intermediates did.  That takes 42 of the 115 FP64-pipe
// instructions of a binary64 trip (the pipe that stage saturates) to IMAD on the otherwise idle FMA
// pipe, halves the isometry's registers and needs no lazy signs: swaps, negations and the fix step are
// predicated integer moves and adds.
// Only rows 1 and 2 are carried: the rows of S are independent under column operations, and row 3 follows
// from the other two at the end (recover_row3), which saves a third of the isometry's work in every trip.
struct IIso {
    int s11, s12, s13, s21, s22, s23;

    __device__ __forceinline__ void load(const double *v)
    {
        s11 = __double2int_rn(v[0]); s12 = __double2int_rn(v[1]); s13 = __double2int_rn(v[2]);
        s21 = __double2int_rn(v[3]); s22 = __double2int_rn(v[4]); s23 = __double2int_rn(v[5]);
    }
    __device__ __forceinline__ void fix(bool on)
    {
        s13 = on ? s13 + s11 + s12 : s13; s23 = on ? s23 + s21 + s22 : s23;
    }
    __device__ __forceinline__ void shear12(double, int t) { s12 += t * s11; s22 += t * s21; }
    __device__ __forceinline__ void shear23(double, int t) { s13 += t * s12; s23 += t * s22; }
    __device__ __forceinline__ void shear13(double, int t) { s13 += t * s11; s23 += t * s21; }
    static __device__ __forceinline__ void swapneg(bool p, int &ci, int &cj, int &ck)
    {
        const int ni = -ci, nj = -cj, nk = -ck;
        ci = p ? nj : ci; cj = p ? ni : cj; ck = p ? nk : ck;
    }
    __device__ __forceinline__ void swap12(bool p) { swapneg(p, s11, s12, s13); swapneg(p, s21, s22, s23); }
    __device__ __forceinline__ void swap23(bool p) { swapneg(p, s12, s13, s11); swapneg(p, s22, s23, s21); }
    __device__ __forceinline__ void negate(bool n1, bool n2, bool n3)
    {
        s11 = n1 ? -s11 : s11; s21 = n1 ? -s21 : s21;
        s12 = n2 ? -s12 : s12; s22 = n2 ? -s22 : s22;
        s13 = n3 ? -s13 : s13; s23 = n3 ? -s23 : s23;
    }
};

// Row 3 of the neighbor isometry S from rows r1, r2.  With G the Gram matrix of the representative
// (2a h g / h 2b f / g f 2c) and G' that of the current form, S^T G S = p^2 G' and det S = +p^3 hold after
// build_neighbor and are kept by every step of the reduction (shears, paired swap-negations, sign flips in
// pairs), so cof(S) = p^3 S^-T and  p G S = cof(S) G'.  The third row of cof(S) is r1 x r2, hence
//     2c r3 = (r1 x r2) G' / p - g r1 - f r2.
// (r1 x r2) G' can exceed 64 bits, but it is p times an integer below 2^63, so it is formed modulo 2^64 and
// divided exactly by multiplying with p^-1 mod 2^64; the rest is small (|2c r3| < 2^51 under iso32 and the
// table bound) and the last division goes through binary64.
__device__ __forceinline__ void recover_row3(const int *r, const i64 *cur, i64 qa, i64 qb, i64 qc, i64 qf, i64 qg, i64 qh,
                                             u64 pinv, int *r3)
{
    const i64 s11 = r[0], s12 = r[1], s13 = r[2], s21 = r[3], s22 = r[4], s23 = r[5];
    const u64 x1 = (u64)(s12 * s23 - s13 * s22), x2 = (u64)(s13 * s21 - s11 * s23), x3 = (u64)(s11 * s22 - s12 * s21);
    const u64 A = (u64)qa, B = (u64)qb, C = (u64)qc, F = (u64)qf, G = (u64)qg, H = (u64)qh;
    const u64 w1 = x1 * (A + A) + x2 * H + x3 * G;
    const u64 w2 = x1 * H + x2 * (B + B) + x3 * F;
    const u64 w3 = x1 * G + x2 * F + x3 * (C + C);
    const i64 g0 = __ldg(cur + 4), f0 = __ldg(cur + 3), c0 = __ldg(cur + 2);
    const i64 t1 = (i64)(w1 * pinv) - (g0 * s11 + f0 * s21);
    const i64 t2 = (i64)(w2 * pinv) - (g0 * s12 + f0 * s22);
    const i64 t3 = (i64)(w3 * pinv) - (g0 * s13 + f0 * s23);
    const double inv = 1.0 / __ll2double_rn(c0 + c0);
    r3[0] = __double2int_rn(__ll2double_rn(t1) * inv);
    r3[1] = __double2int_rn(__ll2double_rn(t2) * inv);
    r3[2] = __double2int_rn(__ll2double_rn(t3) * inv);
}

// One trip on every lane of the warp; finished lanes are frozen: their reciprocals (hence shear
// quotients) and swap predicates are masked, the fix step and the sign normalisation cannot fire on
// them (exit test: sum >= 0; normalisation is idempotent).
template <class F, class ISO>
__device__ __forceinline__ void reduce_trip(F &a, F &b, F &c, F &f, F &g, F &h, F &sum, ISO &S, bool &active)
{
    const F zero = F(0), one = F(1);
    // a + b + f + g + h < 0                                          542-549
    // (some lane of nearly every trip needs it, so it runs unconditionally with a 0/1 multiplier
    //  instead of under a vote and a divergent branch)
    {
        const bool fix = sum < zero;
        const F fm = fix ? one : zero;
        S.fix(fix);
        c = rfma(fm, sum, c);
        f = rfma(fm, h + (b + b), f);
        g = rfma(fm, h + (a + a), g);
    }
    // shear 1: col2 += t col1                                        551-567
    int ti;
    const F den = a + a;
    const F rden = rcp_seeded(den, active);
    {
        const F t = shear_quot(a, h, den, rden, ti);
        S.shear12(rwide(t), ti);
        const F temp = a * t;
        h += temp;
        b = rfma(h, t, b);
        f = rfma(g, t, f);
        h += temp;
    }
    // shear 2: col3 += t col2                                        569-585
    {
        const F den2 = b + b;
        const F rden2 = rcp_seeded(den2, active);
        const F t = shear_quot(b, f, den2, rden2, ti);
        S.shear23(rwide(t), ti);
        const F temp = b * t;
        f += temp;
        c = rfma(f, t, c);
        g = rfma(h, t, g);
        f += temp;
    }
    // shear 3: col3 += t col1 (a, hence 2a, is unchanged)             587-603
    {
        const F t = shear_quot(a, g, den, rden, ti);
        S.shear13(rwide(t), ti);
        const F temp = a * t;
        g += temp;
        c = rfma(g, t, c);
        f = rfma(h, t, f);
        g += temp;
    }
    // sort a <= b <= c with the reference's tie-breaks                605-624
    // (form pairs swap by select on the ALU pipe, isometry columns arithmetically: FP64 pipe for DIso, FMA pipe for IIso)
    {
        const bool p = active & ((a > b) | ((a == b) & (rabs(f) > rabs(g))));
        rswap(p, a, b); rswap(p, f, g);
        S.swap12(p);
    }
    {
        const bool p = active & ((b > c) | ((b == c) & (rabs(g) > rabs(h))));
        rswap(p, b, c); rswap(p, g, h);
        S.swap23(p);
    }
    {
        const bool p = active & ((a > b) | ((a == b) & (rabs(f) > rabs(g))));
        rswap(p, a, b); rswap(p, f, g);
        S.swap12(p);
    }
    // sign normalisation                                              626-687
    // Restated: if f g h > 0 (all nonzero, evenly many negative) flip the negative ones;
    // otherwise flip the positive ones and, when that is an odd number of flips, also the
    // first zero among f, g, h (column flips must come in pairs: det stays +p^2).
    {
        const bool fneg = f < zero, gneg = g < zero, hneg = h < zero;
        const bool fpos = f > zero, gpos = g > zero, hpos = h > zero;
        const bool caseA = (f * g) * h > zero;
        const bool oddB = !caseA & ((fpos != gpos) != hpos);
        const bool fz = !fneg & !fpos, gz = !gneg & !gpos, hz = !hneg & !hpos;
        const bool f1 = (caseA ? fneg : fpos) | (oddB & fz);
        const bool f2 = (caseA ? gneg : gpos) | (oddB & !fz & gz);
        const bool f3 = (caseA ? hneg : hpos) | (oddB & !fz & !gz & hz);
        f = rnegif(f, f1); g = rnegif(g, f2); h = rnegif(h, f3);
        S.negate(f1, f2, f3);
    }
    // exit test                                                       689-690
    sum = (((a + b) + f) + g) + h;
    const bool done = (rabs(f) <= b) & (rabs(g) <= a) & (rabs(h) <= a) & (sum >= zero);
    active = active & !done;
}

// Does any post-loop fix-up of QuadForm.h:693-764 fire on the form as the loop left it?  The ten
// fix-ups run in sequence, each on the state its predecessors left; when none of their conditions
// holds on the initial state, the first changes nothing, so the second sees the same state, and so
// on: nothing fires (759 repeats 744).  Exact in binary64 and, in the tail, in binary32 (every
// operand is an integer below 2^23).
template <class F>
__device__ __forceinline__ bool boundary_fixup_fires(F a, F b, F c, F f, F g, F h)
{
    const F zero = F(0);
    const F sum = (((a + b) + f) + g) + h;
    return ((sum == zero) & ((a + a) + (g + g) + h > zero)) |      // 693
           ((a == -h) & (g != zero)) |                             // 702
           ((a == -g) & (h != zero)) |                             // 710
           ((b == -f) & (h != zero)) |                             // 718
           ((a == h) & (g > f + f)) |                              // 726
           ((a == g) & (h > f + f)) |                              // 732
           ((b == f) & (h > g + g)) |                              // 738
           ((a == b) & (rabs(f) > rabs(g))) |                      // 744, 759
           ((b == c) & (rabs(g) > rabs(h)));                       // 752
}

// Binary32 tail (f32_tail).  Diagonal coefficients never grow and after a trip's sort c is the
// largest; |f|, |g|, |h| < 2 c, the running sum < 8 c, shear intermediates (a t, h + a t) < 4.5 c.
// Once every unfinished lane of the warp has c < 2^20 -- after ~5 of the ~12.6 trips a warp makes
// at p ~ 10^4 -- every value of every later trip is an integer below 2^23 and binary32
// FFMA/FADD/FMUL are exact on it (one rounding per operation, of an integer below 2^24).  The
// quotient estimate num * rcp(den) is within 2^-10 of num / den (|num / den| < 2^12), so RN gives
// floor or floor + 1 and the exact remainder decides, as in binary64.  Finished lanes hold reduced
// forms and are carried along with zero quotients: the host requires c < 2^24 and b < 2^20 on
// every table form (their running sum is at most 5 b), so their values convert exactly as well.
//
// Returns the trip count; on_boundary says whether some fix-up of QuadForm.h:693-764 fires.
struct DReduce {
    double a, b, c, f, g, h;
    double s[9];                  // in: the neighbor isometry, row-major; out (DIso): the reduced one
    int si[6];                    // out (IIso): rows 1 and 2 of the reduced isometry
    bool on_boundary;
};

template <class ISO>
__device__ __forceinline__ int reduce_loop(DReduce &d, bool f32_tail, ISO &S)
{
    const unsigned warp = 0xffffffffu;
    double a = d.a, b = d.b, c = d.c, f = d.f, g = d.g, h = d.h;
    S.load(d.s);
    bool active = true;
    int iters = 0;
    double sum = (((a + b) + f) + g) + h;
    // binary64 while some unfinished lane still has a coefficient of 2^20 or more (the first trip
    // always: the neighbor's coefficients are not sorted yet)
    const double big = f32_tail ? 1048576.0 : 0.0;
    bool more = true;
    while (__any_sync(warp, more))
    {
        if (++iters > TB_REDUCE_CAP) break;
        reduce_trip<double>(a, b, c, f, g, h, sum, S, active);
        more = active & (c >= big);
    }
    if (f32_tail)
    {
        float a4 = (float)a, b4 = (float)b, c4 = (float)c, f4 = (float)f, g4 = (float)g, h4 = (float)h;
        float sum4 = (float)sum;
        while (__any_sync(warp, active))
        {
            if (++iters > TB_REDUCE_CAP) break;
            reduce_trip<float>(a4, b4, c4, f4, g4, h4, sum4, S, active);
        }
        d.on_boundary = boundary_fixup_fires(a4, b4, c4, f4, g4, h4);
        d.a = (double)a4; d.b = (double)b4; d.c = (double)c4; d.f = (double)f4; d.g = (double)g4; d.h = (double)h4;
    }
    else
    {
        d.on_boundary = boundary_fixup_fires(a, b, c, f, g, h);
        d.a = a; d.b = b; d.c = c; d.f = f; d.g = g; d.h = h;
    }
    return iters;
}

// iso_int: the isometry runs in int32 ring arithmetic (PrimeCtx::iso32) and comes back in d.si
__device__ __forceinline__ int reduce_loop_fp64(DReduce &d, bool f32_tail, bool iso_int = false)
{
    if (iso_int)
    {
        IIso S;
        const int iters = reduce_loop(d, f32_tail, S);
        d.si[0] = S.s11; d.si[1] = S.s12; d.si[2] = S.s13; d.si[3] = S.s21; d.si[4] = S.s22; d.si[5] = S.s23;
        return iters;                                    // row 3: reduce_finish
    }
    DIso S;
    const int iters = reduce_loop(d, f32_tail, S);
    const u32 sg2 = S.sg1 ^ S.x12, sg3 = S.sg1 ^ S.x13;
    d.s[0] = xflip(S.s11, S.sg1); d.s[3] = xflip(S.s21, S.sg1); d.s[6] = xflip(S.s31, S.sg1);
    d.s[1] = xflip(S.s12, sg2); d.s[4] = xflip(S.s22, sg2); d.s[7] = xflip(S.s32, sg2);
    d.s[2] = xflip(S.s13, sg3); d.s[5] = xflip(S.s23, sg3); d.s[8] = xflip(S.s33, sg3);
    return iters;
}

// Post-loop boundary fix-ups (QuadForm.h:693-764): rare, so plain branches, and out of line: the
// hot path only asks whether any of them fires (boundary_fixup_fires) and calls this when one does.
template <class Z>
__device__ __noinline__ void reduce_boundary(Form<Z> &q, Iso<Z> &s)
{
    Z a = q.a, b = q.b, c = q.c, f = q.f, g = q.g, h = q.h, t;
    const Z zero = Z(0);
    Z sum = a + b + f + g + h;
    if (sum.is_zero() && a.twice() + g.twice() + h > zero)   // 693-700
    {
        s.fix_sum();
        c += sum;
        f += h + b.twice(); f = -f;
        g += h + a.twice(); g = -g;
    }
    if (a == -h && !g.is_zero()) { s.fix_ah(); f += g; f = -f; g = -g; h = -h; }              // 702-708
    if (a == -g && !h.is_zero()) { s.fix_ag(); f += h; f = -f; h = -h; g += a.twice(); }      // 710-716
    if (b == -f && !h.is_zero()) { s.fix_bf(); g += h; g = -g; h = -h; f += b.twice(); }      // 718-724
    if (a == h && g > f.twice()) { s.edge_ah(); f = g - f; }                                  // 726-730
    if (a == g && h > f.twice()) { s.edge_ag(); f = h - f; }                                  // 732-736
    if (b == f && h > g.twice()) { s.edge_bf(); g = h - g; }                                  // 738-742
    if (a == b && f.abs() > g.abs()) { s.swap12_neg(); t = a; a = b; b = t; t = g; g = f; f = t; }   // 744-750
    if (b == c && g.abs() > h.abs()) { s.swap23_neg(); t = g; g = h; h = t; }                 // 752-757
    if (a == b && f.abs() > g.abs()) { s.swap12_neg(); t = g; g = f; f = t; }                 // 759-764
    q.a = a; q.b = b; q.c = c; q.f = f; q.g = g; q.h = h;
}

// The integer loop lives out of line (it is the fallback); copies keep the caller's state in registers.
template <class Z>
__device__ __forceinline__ bool reduce_out_of_line(Form<Z> &q, Iso<Z> &s)
{
    Form<Z> tq = q;
    Iso<Z> ts = s;
    const bool ok = eisenstein_reduce(tq, ts);
    q = tq;
    s = ts;
    return ok;
}

// The reduction in three steps, so that lanes whose neighbor was set up directly in binary64
// (fast_setup) and lanes that went through the integer setup share one copy of the loop:
//   reduce_prepare   integer form + isometry -> binary64 state, or (launch / warp not eligible) the
//                    whole reduction in the integer code; every lane of the warp calls it (it votes)
//   reduce_loop_fp64 the loop
//   reduce_finish    binary64 state -> integers, boundary fix-ups
template <class Z>
__device__ __forceinline__ bool reduce_prepare(Form<Z> &q, Iso<Z> &s, bool fp64_allowed, bool proven_small,
                                               DReduce &d, bool &ok)
{
    // proven_small: the host bounded everything build_neighbor can hand over below 2^50 for this launch
    // (PrimeCtx::proven_small), so the per-lane magnitude test and its vote are skipped
    ok = true;
    if (!proven_small)
    {
        bool small = fits50(q.a) && fits50(q.b) && fits50(q.c) && fits50(q.f) && fits50(q.g) && fits50(q.h) &&
                     fits50(s.a11) && fits50(s.a12) && fits50(s.a13) && fits50(s.a21) && fits50(s.a22) &&
                     fits50(s.a23) && fits50(s.a31) && fits50(s.a32) && fits50(s.a33) &&
                     q.a > Z(0) && q.b > Z(0);
        __syncwarp(0xffffffffu);
        if (!(fp64_allowed && __all_sync(0xffffffffu, small))) { ok = reduce_out_of_line(q, s); return false; }
    }
    else if (!fp64_allowed) { ok = reduce_out_of_line(q, s); return false; }
    d.a = __ll2double_rn(q.a.low64()); d.b = __ll2double_rn(q.b.low64()); d.c = __ll2double_rn(q.c.low64());
    d.f = __ll2double_rn(q.f.low64()); d.g = __ll2double_rn(q.g.low64()); d.h = __ll2double_rn(q.h.low64());
    d.s[0] = __ll2double_rn(s.a11.low64()); d.s[1] = __ll2double_rn(s.a12.low64()); d.s[2] = __ll2double_rn(s.a13.low64());
    d.s[3] = __ll2double_rn(s.a21.low64()); d.s[4] = __ll2double_rn(s.a22.low64()); d.s[5] = __ll2double_rn(s.a23.low64());
    d.s[6] = __ll2double_rn(s.a31.low64()); d.s[7] = __ll2double_rn(s.a32.low64()); d.s[8] = __ll2double_rn(s.a33.low64());
    return true;
}
template <class Z>
__device__ __forceinline__ void reduce_finish(const DReduce &d, Form<Z> &q, Iso<Z> &s, bool iso_int = false,
                                              const i64 *cur = nullptr, u64 pinv = 0)
{
    const i64 qa = __double2ll_rn(d.a), qb = __double2ll_rn(d.b), qc = __double2ll_rn(d.c);
    const i64 qf = __double2ll_rn(d.f), qg = __double2ll_rn(d.g), qh = __double2ll_rn(d.h);
    q.a = Z(qa); q.b = Z(qb); q.c = Z(qc); q.f = Z(qf); q.g = Z(qg); q.h = Z(qh);
    if (iso_int)
    {
        int r3[3];
        recover_row3(d.si, cur, qa, qb, qc, qf, qg, qh, pinv, r3);
        s.a11 = Z((i64)d.si[0]); s.a12 = Z((i64)d.si[1]); s.a13 = Z((i64)d.si[2]);
        s.a21 = Z((i64)d.si[3]); s.a22 = Z((i64)d.si[4]); s.a23 = Z((i64)d.si[5]);
        s.a31 = Z((i64)r3[0]); s.a32 = Z((i64)r3[1]); s.a33 = Z((i64)r3[2]);
    }
    else
    {
        s.a11 = Z(__double2ll_rn(d.s[0])); s.a12 = Z(__double2ll_rn(d.s[1])); s.a13 = Z(__double2ll_rn(d.s[2]));
        s.a21 = Z(__double2ll_rn(d.s[3])); s.a22 = Z(__double2ll_rn(d.s[4])); s.a23 = Z(__double2ll_rn(d.s[5]));
        s.a31 = Z(__double2ll_rn(d.s[6])); s.a32 = Z(__double2ll_rn(d.s[7])); s.a33 = Z(__double2ll_rn(d.s[8]));
    }
    if (d.on_boundary)
    {
        Form<Z> tq = q;          // copies: only these are address-taken by the out-of-line call
        Iso<Z> ts = s;
        reduce_boundary(tq, ts);
        q = tq;
        s = ts;
    }
}
// Reduction entry point of the genus-enumeration kernels: FP64 loop when the launch allows it and
// the whole warp's operands are small, else the integer code.  Every lane of the warp must call this.
template <class Z>
__device__ __forceinline__ bool reduce_neighbor(Form<Z> &q, Iso<Z> &s, bool fp64_allowed, bool proven_small,
                                                bool f32_tail = false)
{
    DReduce d;
    bool ok;
    if (!reduce_prepare(q, s, fp64_allowed, proven_small, d, ok)) return ok;
    const int iters = reduce_loop_fp64(d, f32_tail);
    reduce_finish(d, q, s);
    return iters <= TB_REDUCE_CAP;
}

// ---------------------------------------------------------------------------
// Genus lookup.  QuadForm::hash_value (QuadForm.cpp:779-803) + HashMap::indexof
// (HashMap.h:68-89).  Returns the representative index or -1 ("Key not found.").
// ---------------------------------------------------------------------------
template <class Z>
__device__ __forceinline__ int genus_lookup(const Form<Z> &q, const GenusCtx &gc)
{
    const bool narrow = q.a.fits_i64() && q.b.fits_i64() && q.c.fits_i64() &&
                        q.f.fits_i64() && q.g.fits_i64() && q.h.fits_i64();
    const i64 k0 = q.a.low64(), k1 = q.b.low64(), k2 = q.c.low64();
    const i64 k3 = q.f.low64(), k4 = q.g.low64(), k5 = q.h.low64();
    u64 fnv = 0x811c9dc5ull;
    fnv = (fnv ^ (u64)k0) * 0x01000193ull;
    fnv = (fnv ^ (u64)k1) * 0x01000193ull;
    fnv = (fnv ^ (u64)k2) * 0x01000193ull;
    fnv = (fnv ^ (u64)k3) * 0x01000193ull;
    fnv = (fnv ^ (u64)k4) * 0x01000193ull;
    fnv = (fnv ^ (u64)k5) * 0x01000193ull;
    u32 slot = (u32)fnv & gc.slot_mask;
    int found = -1;
    // single exit, so the warp reconverges right after the probe loop (avg 1.3 probes);
    // every lane of the warp calls this
    for (u32 probes = 0; narrow && probes <= gc.slot_mask; ++probes)
    {
        const int idx = __ldg(gc.slots + slot);
        if (idx < 0) break;
        TB_CHECK(idx < gc.N);
        const longlong2 *key = reinterpret_cast<const longlong2 *>(gc.rep + (size_t)idx * TB_REC_WORDS);
        const longlong2 x = __ldg(key), y = __ldg(key + 1), z = __ldg(key + 2);
        if (x.x == k0 && x.y == k1 && y.x == k2 && y.y == k3 && z.x == k4 && z.y == k5) { found = idx; break; }
        slot = (slot + 1) & gc.slot_mask;
    }
    __syncwarp(0xffffffffu);
    return found;
}

// ---------------------------------------------------------------------------
// Spinor norm class.  Spinor::compute_vals (Spinor.h:52-66) and Spinor::norm
// (Spinor.h:19-41).
// ---------------------------------------------------------------------------
// Parity of the d-adic valuation of a nonzero 64-bit magnitude, d an odd prime:
// d | y  <=>  y * d^-1 (mod 2^64) <= floor((2^64-1)/d), and then that product is y / d.
__device__ __forceinline__ u32 parities_u64(u64 y, const GenusCtx &gc)
{
    u32 val = 0;
    for (int i = 0; i < gc.K; ++i)
    {
        const u64 dinv = gc.spin_inv[i];
        u32 odd;
        if (dinv == 0)
        {
            odd = (u32)__ffsll((i64)y) - 1u;          // d = 2: trailing zeros
        }
        else
        {
            const u64 qmax = gc.spin_qmax[i];
            odd = 0;
            u64 q = y * dinv;
            while (q <= qmax) { y = q; odd ^= 1u; q = y * dinv; }   // valuation is 0 for most inputs
        }
        val |= (odd & 1u) << i;
    }
    return val;
}
template <int TAG>
__device__ __forceinline__ u32 prime_parities(I64T<TAG> x, const GenusCtx &gc)
{
    if (x.is_zero()) { raise_overflow(); return 0; }   // the reference's loop would not terminate
    return parities_u64(x.umag(), gc);
}

template <class Z>
__device__ __forceinline__ u32 prime_parities(Z x, const GenusCtx &gc)
{
    u32 val = 0;
    if (x.is_zero()) { raise_overflow(); return 0; }   // the reference's loop would not terminate
    {
        u64 mh, ml;
        x.umag(mh, ml);
        if (mh == 0) return parities_u64(ml, gc);
    }
    for (int i = 0; i < gc.K; ++i)
    {
        const InvDiv &iv = gc.spin[i];
        while (strip_factor(x, iv)) val ^= (1u << i);
    }
    return val;
}

template <class Z>
__device__ __forceinline__ u32 spinor_norm(const Form<Z> &q, const Iso<Z> &s, Z scalar, const GenusCtx &gc)
{
    const u32 twist = (1u << gc.K) - 1u;
    Z tr = s.a11 + s.a22 + s.a33;
    if (tr != -scalar) return prime_parities(tr + scalar, gc);

    Z delta = (q.a * (s.a22 + s.a33)).twice() - (q.g * s.a31 + q.h * s.a21);
    if (!delta.is_zero()) return prime_parities(delta, gc) ^ twist;

    Z abh = Z(4) * q.a * q.b - q.h * q.h;
    Z abhm33 = abh * s.a33 + s.a32 * (q.g * q.h - (q.a * q.f).twice()) + s.a31 * (q.f * q.h - (q.b * q.g).twice());
    if (abhm33 != abh * scalar)
        return prime_parities(abhm33 - abh * scalar, gc) ^ prime_parities(q.a.twice(), gc) ^ twist;

    return prime_parities(abh * scalar, gc);
}

// X64: the first case is exact in 64 bits (|tr + scalar| < 2^63 by the host's bounds); the
// later cases multiply form minors into the isometry, so they run in U128.  They are rare.
template <>
__device__ __forceinline__ u32 spinor_norm<X64>(const Form<X64> &q, const Iso<X64> &s, X64 scalar, const GenusCtx &gc)
{
    X64 tr = s.a11 + s.a22 + s.a33;
    if (tr != -scalar) return prime_parities(tr + scalar, gc);
    Form<U128> qw;
    Iso<U128> sw;
    qw.a = U128(q.a.v); qw.b = U128(q.b.v); qw.c = U128(q.c.v); qw.f = U128(q.f.v); qw.g = U128(q.g.v); qw.h = U128(q.h.v);
    sw.a11 = U128(s.a11.v); sw.a12 = U128(s.a12.v); sw.a13 = U128(s.a13.v);
    sw.a21 = U128(s.a21.v); sw.a22 = U128(s.a22.v); sw.a23 = U128(s.a23.v);
    sw.a31 = U128(s.a31.v); sw.a32 = U128(s.a32.v); sw.a33 = U128(s.a33.v);
    return spinor_norm<U128>(qw, sw, U128(scalar.v), gc);
}

template <class Z>
__device__ __forceinline__ Iso<Z> iso_mul(const Iso<Z> &x, const Iso<Z> &y)
{
    // Isometry.h:63-76
    Iso<Z> t;
    t.a11 = x.a11 * y.a11 + x.a12 * y.a21 + x.a13 * y.a31;
    t.a12 = x.a11 * y.a12 + x.a12 * y.a22 + x.a13 * y.a32;
    t.a13 = x.a11 * y.a13 + x.a12 * y.a23 + x.a13 * y.a33;
    t.a21 = x.a21 * y.a11 + x.a22 * y.a21 + x.a23 * y.a31;
    t.a22 = x.a21 * y.a12 + x.a22 * y.a22 + x.a23 * y.a32;
    t.a23 = x.a21 * y.a13 + x.a22 * y.a23 + x.a23 * y.a33;
    t.a31 = x.a31 * y.a11 + x.a32 * y.a21 + x.a33 * y.a31;
    t.a32 = x.a31 * y.a12 + x.a32 * y.a22 + x.a33 * y.a32;
    t.a33 = x.a31 * y.a13 + x.a32 * y.a23 + x.a33 * y.a33;
    return t;
}

template <class Z>
__device__ __forceinline__ Form<Z> load_form(const i64 *rec)
{
    Form<Z> q;
    q.a = Z(__ldg(rec + 0)); q.b = Z(__ldg(rec + 1)); q.c = Z(__ldg(rec + 2));
    q.f = Z(__ldg(rec + 3)); q.g = Z(__ldg(rec + 4)); q.h = Z(__ldg(rec + 5));
    return q;
}
template <class Z>
__device__ __forceinline__ Iso<Z> load_iso(const i64 *m)
{
    Iso<Z> s;
    s.a11 = Z(__ldg(m + 0)); s.a12 = Z(__ldg(m + 1)); s.a13 = Z(__ldg(m + 2));
    s.a21 = Z(__ldg(m + 3)); s.a22 = Z(__ldg(m + 4)); s.a23 = Z(__ldg(m + 5));
    s.a31 = Z(__ldg(m + 6)); s.a32 = Z(__ldg(m + 7)); s.a33 = Z(__ldg(m + 8));
    return s;
}

// ---------------------------------------------------------------------------
// Neighbor setup in binary64 (launches with PrimeCtx::fast_setup): t -> isotropic line -> neighbor
// form and isometry, for the generic lane -- the line has z = 1 (NeighborManager.h:96-143 with
// at != 0, ct != 0, t != p) and g' != 0 mod p (no rotation, NeighborManager.h:279-295) -- which is
// all but about four lanes in p + 1.  The other lanes return false and take the integer code
// (setup_general).  The integer code spends ~800 instructions per warp here, mostly emulating
// 64-bit products and 2-by-1 divisions with IMAD; every quantity is an integer below 2^53 under
// the launch's bounds, so DFMA computes it exactly:
//   p < 2^16, table coefficients |q| <= B with B p^2 < 2^48 (proven_small):
//     u, v <= p/2;  au, hu, hv, bv <= B p / 2;  temp <= B (p + 1);  aa <= B (p^2 + p + 1) 3/4;
//     gg, hh <= B (3p/2 + 1);  (-hh) ginv < 2 B p^2;  (aa mod p^2) ginv < p^3 < 2^48;
//     bb <= B (p^2 + p + 1);  hh + s1 gg < 2 B p^2;  gg + cc s2 < 2 B p^2.
//   The two numerators that exceed 2^53, aa + s2 (gg + cc s2) < B p^4 and hh + s2 ff < 3 B p^3, are
//   carried as an error-free product (hi, lo = fma(x, y, -hi)): the quotient estimate RN((hi + rest) / d)
//   is off by a few units at most, the exact remainder hi - q d + rest + lo is a small multiple of d
//   (both divisions are exact in the reference: NeighborManager.h:338-341) and corrects it.
//   x mod d for |x| < 2^53: q = RN(x * RN(1/d)) by the 1.5 * 2^52 trick, r = fma(-q, d, x) exact in
//   (-d, d), plus d if negative.  The reference's remainders are C-truncated and then shifted by
//   "if (s < -d/2) s += d" (NeighborManager.h:309-312), i.e. with c = x mod d in [0, d):
//   s = c - d if x < 0 and c >= d - floor(d/2), else c.
// In int64 mode the reference's own arithmetic wraps modulo 2^64 (and its discriminant check then
// fails, NeighborManager.h:344-354); fast_setup is only set when the host has bounded the largest
// intermediate, s2 (gg + cc s2) + aa, below 2^63, so the exact value IS the reference's value and
// the check passes by construction.
// ---------------------------------------------------------------------------
__device__ __forceinline__ u32 fp_mul16(u32 a, u32 b, const PrimeCtx &pc)
{
    const u32 x = a * b;
    const u32 r = x - __umulhi(x, pc.m32) * pc.p;
    return r >= pc.p ? r - pc.p : r;
}
__device__ __forceinline__ double rint52(double x)      // nearest integer, |x| < 2^51
{
    return __dadd_rn(__dadd_rn(x, 6755399441055744.0), -6755399441055744.0);
}
// x mod d in [0, d) for an integer |x| < 2^53 (d, inv = RN(1/d) launch constants)
__device__ __forceinline__ double fmod_pos(double x, double d, double inv)
{
    const double q = rint52(x * inv);
    const double r = __fma_rn(-q, d, x);
    return r < 0.0 ? r + d : r;
}
// (hi + lo + rest) / d for an exact division: hi, lo an error-free product, |rest| < 2^52
__device__ __forceinline__ double exact_div(double hi, double lo, double rest, double d, double inv)
{
    const double q = rint52((hi + rest) * inv);
    const double e = (__fma_rn(-q, d, hi) + rest) + lo;      // exact: a small multiple of d
    return q + rint52(e * inv);
}

__device__ __forceinline__ bool fast_setup(const i64 *cur, const RepPrime &rp, u32 t, const PrimeCtx &pc,
                                           DReduce &d, u32 &vx, u32 &vy, u32 &vz)
{
    const u32 p = pc.p;
    // line_for_t's generic branch
    u32 at = fp_mul16(fp_add(fp_mul16(fp_sub(t, 1u, p), rp.b, pc), rp.delta, p), t, pc);
    at = fp_add(at, rp.a0, p);
    const u32 ct = fp_add(fp_mul16(fp_add(fp_mul16(rp.b, t, pc), rp.h, p), t, pc), rp.a, p);
    bool generic = (t != p) & (at != 0u) & (ct != 0u);
    const u32 ctc = generic ? ct : 1u;                      // keep the table index in range on the other lanes
    const u32 tc = generic ? t : 0u;
    const u32 inv = fp_mul16(at, __ldg(pc.inv_lut + ctc), pc);
    vx = fp_add(rp.x0, fp_sub(inv, 1u, p), p);
    vy = fp_add(fp_sub(rp.y0, tc, p), fp_mul16(tc, inv, pc), p);
    vz = 1u;

    // build_neighbor, z = 1 (NeighborManager.h:226-245)
    const int half = (int)(p >> 1);
    int ux = (int)vx, uy = (int)vy;
    if (ux > half) ux -= (int)p;
    if (uy > half) uy -= (int)p;
    const double u = (double)ux, v = (double)uy;
    const double a = __ll2double_rn(__ldg(cur + 0)), b = __ll2double_rn(__ldg(cur + 1)), c = __ll2double_rn(__ldg(cur + 2));
    const double f = __ll2double_rn(__ldg(cur + 3)), g = __ll2double_rn(__ldg(cur + 4)), h = __ll2double_rn(__ldg(cur + 5));
    const double P = pc.pd, PP = pc.ppd, iP = pc.inv_pd, iPP = pc.inv_ppd;
    const double au = a * u, hu = h * u, hv = h * v, bv = b * v;
    const double temp = (au + hv) + g;
    const double aa = fma(v, bv + f, fma(u, temp, c));
    const double gg = -(temp + au);
    const double hh = fma(2.0, bv, f + hu);
    double bb = b;
    const double cc = a;
    double ff = -h;

    // g' mod p and its inverse (279-303)
    const double gm = fmod_pos(gg, P, iP);
    generic &= gm != 0.0;
    const u32 gi = (u32)__double2loint(__dadd_rn(gm, 6755399441055744.0));
    const u32 ginv_u = __ldg(pc.inv_lut + gi);
    const double ginv = __hiloint2double(0x43300000, (int)ginv_u) - 4503599627370496.0;

    // s1 = (-hh ginv) mod p, s2 = (-aa ginv) mod p^2, the reference's truncated / shifted residues (309-312)
    const double c1 = fmod_pos(-hh * ginv, P, iP);
    const double s1 = (hh > 0.0 && c1 >= P - (double)half) ? c1 - P : c1;
    const double c2 = fmod_pos(fmod_pos(-aa, PP, iPP) * ginv, PP, iPP);
    const double hpp = (double)(pc.pp >> 1);
    const double s2 = (c2 >= PP - hpp) ? c2 - PP : c2;        // aa = Q(u, v, 1) > 0, so -aa ginv < 0

    // 314-341
    const double temp1 = cc * s1;
    bb = fma(s1, ff + temp1, bb);
    ff = fma(2.0, temp1, ff);
    const double hh1 = fma(s1, gg, hh);
    const double temp2 = cc * s2;
    const double x2 = gg + temp2;
    const double hi_a = s2 * x2, lo_a = __fma_rn(s2, x2, -hi_a);
    const double hi_h = s2 * ff, lo_h = __fma_rn(s2, ff, -hi_h);
    d.a = exact_div(hi_a, lo_a, aa, PP, iPP);
    d.b = bb;
    d.c = cc * PP;
    d.f = ff * P;
    d.g = fma(2.0, temp2, gg);
    d.h = exact_div(hi_h, lo_h, hh1, P, iP);
    // isometry: [[u, 0, -1], [v, 1, 0], [1, 0, 0]], col2 += s1 col3, col1 += s2 col3, col2 *= p, col3 *= p^2
    d.s[0] = u - s2; d.s[1] = -s1 * P; d.s[2] = -PP;
    d.s[3] = v;      d.s[4] = P;       d.s[5] = 0.0;
    d.s[6] = 1.0;    d.s[7] = 0.0;     d.s[8] = 0.0;
    return generic;
}

// ---------------------------------------------------------------------------
// One neighbor, start to finish: the loop body of Genus.h:567-618.
// ---------------------------------------------------------------------------
enum NeighborStatus { NBR_OK = 0, NBR_OVERFLOW = 1, NBR_NOT_FOUND = 2 };

template <class Z> struct IsWide { static const bool value = true; };
template <> struct IsWide<W64> { static const bool value = false; };
template <> struct IsWide<X64> { static const bool value = false; };

template <class Z>
struct Neighbor {
    u32 vx, vy, vz;   // isotropic line
    Form<Z> q;        // reduced neighbor
    Iso<Z> s;         // representative -> neighbor
    Iso<Z> comp;      // cur.s * s * rep.sinv (valid when composed)
    Z denom;          // p * scalar(cur) * scalar(rep) (valid when composed)
    int r;
    u32 spin;
    bool composed;
};

// Integer setup of one neighbor (any prime, any line): the representative's form, the isotropic line
// for t and the neighbor lattice; a neighbor that overflowed int64 is replaced by the representative's
// own form with the identity.  Out of line: in launches with fast_setup only a few lanes come here.
template <class Z>
__device__ __noinline__ int setup_general_nl(int n, u32 t, const RepPrime &rp, const PrimeCtx &pc,
                                             const GenusCtx &gc, Form<Z> &oq, Iso<Z> &os, u32 *line)
{
    int status = 0;
    const i64 *cur = gc.cur + (size_t)n * TB_REC_WORDS;
    const Form<Z> q = load_form<Z>(cur);
    u32 vx, vy, vz;
    line_for_t(rp, t, pc, vx, vy, vz);
    Form<Z> nq;
    Iso<Z> s;
    if (!build_neighbor(q, vx, vy, vz, __ldg(gc.disc + n), pc, nq, s))
    {
        status = 1;                                          // NBR_OVERFLOW
        const Z zero = Z(0), one = Z(1);
        nq = q;
        s.a11 = one; s.a12 = zero; s.a13 = zero;
        s.a21 = zero; s.a22 = one; s.a23 = zero;
        s.a31 = zero; s.a32 = zero; s.a33 = one;
    }
    oq = nq;
    os = s;
    line[0] = vx; line[1] = vy; line[2] = vz;
    return status;
}
// (the caller's copies: only these are address-taken by the out-of-line call, the Neighbor stays in registers)
template <class Z>
__device__ __forceinline__ int setup_general(int n, u32 t, const RepPrime &rp, const PrimeCtx &pc,
                                             const GenusCtx &gc, Neighbor<Z> &o)
{
    Form<Z> tq;
    Iso<Z> ts;
    u32 line[3];
    const int status = setup_general_nl<Z>(n, t, rp, pc, gc, tq, ts, line);
    o.q = tq;
    o.s = ts;
    o.vx = line[0]; o.vy = line[1]; o.vz = line[2];
    return status;
}

// want_comp: also produce the composed isometry when r == n (IsometrySequence
// has no r == n shortcut, IsometrySequence.h:63-69).
template <class Z>
__device__ __forceinline__ int one_neighbor(int n, u32 t, const RepPrime &rp, const PrimeCtx &pc,
                                            const GenusCtx &gc, bool want_comp, Neighbor<Z> &o)
{
    // Straight-line: every lane of the warp runs every phase (the reduce loop and the lookup
    // vote / resynchronise warp-wide).  A lane whose neighbor overflowed carries the
    // representative's own, already reduced, form through the remaining phases and only its
    // status differs.
    int status = NBR_OK;
    const i64 *cur = gc.cur + (size_t)n * TB_REC_WORDS;
    DReduce d;
    bool in_fp64;
    // int64 mode additionally needs the no-wrap bit (its reference arithmetic wraps, the exact one does not)
    if (CheckDisc<Z>::value ? pc.fast_setup == 3u : (pc.fast_setup & 1u) != 0u)     // launch-uniform
    {
        in_fp64 = true;
        if (!fast_setup(cur, rp, t, pc, d, o.vx, o.vy, o.vz))
        {
            // the few special lanes (t = p, a zero of the line's conic, a rotated neighbor): integer setup
            bool ok;
            status = setup_general<Z>(n, t, rp, pc, gc, o);
            reduce_prepare(o.q, o.s, true, true, d, ok);
        }
    }
    else
    {
        bool ok;
        status = setup_general<Z>(n, t, rp, pc, gc, o);
        in_fp64 = reduce_prepare(o.q, o.s, pc.fp64_reduce != 0, pc.proven_small != 0, d, ok);
        if (!ok) status = NBR_OVERFLOW;
    }
    if (in_fp64)                                             // warp-uniform
    {
        // the isometry in int32 ring arithmetic when its final entries are proven below 2^31 -- but the neighbor's
        // starting entries must convert too (below 2^31: p^2 <= 2^30 covers -p^2, s1 p and u - s2)
        const bool iso_int = pc.iso32 != 0 && pc.p < 32768u && pc.pinv64 != 0;
        if (reduce_loop_fp64(d, pc.f32_tail != 0, iso_int) > TB_REDUCE_CAP) status = NBR_OVERFLOW;
        reduce_finish(d, o.q, o.s, iso_int, cur, pc.pinv64);
    }

    int r = genus_lookup(o.q, gc);                           // Genus.h:575
    if (r < 0)
    {
        if (status == NBR_OK) status = NBR_NOT_FOUND;
        r = n;
    }
    o.r = r;
    o.composed = false;

    const Z pz = Z((i64)pc.p);
    if (r == n && !want_comp)                                // Genus.h:582-585
    {
        o.spin = spinor_norm(o.q, o.s, pz, gc);
        return status;
    }
    const i64 *rep = gc.rep + (size_t)r * TB_REC_WORDS;      // Genus.h:588-615
    if (IsWide<Z>::value && pc.narrow_trace && !want_comp &&
        o.s.a11.fits_i64() && o.s.a12.fits_i64() && o.s.a13.fits_i64() && o.s.a21.fits_i64() && o.s.a22.fits_i64() &&
        o.s.a23.fits_i64() && o.s.a31.fits_i64() && o.s.a32.fits_i64() && o.s.a33.fits_i64())
    {
        // Two-limb modes: the composed isometry cur.s * s * rep.sinv maps the mother form to
        // itself with scalar denom, so its entries are below denom * sqrt(c0 / lambda_min(mother));
        // the host set narrow_trace because that (and denom) stay below 2^61.  Trace and
        // denominator are then exact in wrap-around 64-bit ring arithmetic.
        Iso<W64> s64;
        s64.a11 = W64(o.s.a11.low64()); s64.a12 = W64(o.s.a12.low64()); s64.a13 = W64(o.s.a13.low64());
        s64.a21 = W64(o.s.a21.low64()); s64.a22 = W64(o.s.a22.low64()); s64.a23 = W64(o.s.a23.low64());
        s64.a31 = W64(o.s.a31.low64()); s64.a32 = W64(o.s.a32.low64()); s64.a33 = W64(o.s.a33.low64());
        const Iso<W64> cs64 = load_iso<W64>(cur + 7), ri64 = load_iso<W64>(rep + 7);
        const Iso<W64> m64 = iso_mul(cs64, s64);
        const W64 tr64 = m64.a11 * ri64.a11 + m64.a12 * ri64.a21 + m64.a13 * ri64.a31 +
                         m64.a21 * ri64.a12 + m64.a22 * ri64.a22 + m64.a23 * ri64.a32 +
                         m64.a31 * ri64.a13 + m64.a32 * ri64.a23 + m64.a33 * ri64.a33;
        const W64 den64 = W64((i64)pc.p) * W64(__ldg(cur + 6)) * W64(__ldg(rep + 6));
        if (tr64 != -den64)
        {
            o.spin = prime_parities(tr64 + den64, gc);
            return status;
        }
    }
    if (!IsWide<Z>::value && pc.iso32 && !want_comp)
    {
        // 64-bit modes under PrimeCtx::iso32: every factor of cur.s * s * rep.sinv fits 32 bits, so the trace
        //     tr = sum_{k,j} s[k][j] * X[j][k],   X = rep.sinv * cur.s
        // is 27 32x32->64 multiply-adds for X and nine 64x32 ones for the sum (ring arithmetic modulo 2^64,
        // exactly what the 64-bit products below compute, at half the instructions)
        int cs[9], ri[9], sv[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { cs[i] = (int)__ldg(cur + 7 + i); ri[i] = (int)__ldg(rep + 7 + i); }
        sv[0] = (int)o.s.a11.low64(); sv[1] = (int)o.s.a12.low64(); sv[2] = (int)o.s.a13.low64();
        sv[3] = (int)o.s.a21.low64(); sv[4] = (int)o.s.a22.low64(); sv[5] = (int)o.s.a23.low64();
        sv[6] = (int)o.s.a31.low64(); sv[7] = (int)o.s.a32.low64(); sv[8] = (int)o.s.a33.low64();
        u64 tr = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                const i64 X = (i64)ri[3 * j] * cs[k] + (i64)ri[3 * j + 1] * cs[3 + k] + (i64)ri[3 * j + 2] * cs[6 + k];
                tr += (u64)X * (u64)(i64)sv[3 * k + j];
            }
        const u64 den = (u64)pc.p * (u64)__ldg(cur + 6) * (u64)__ldg(rep + 6);
        if (tr != 0ull - den)                                // Spinor.h:21-25
        {
            o.spin = prime_parities(Z((i64)(tr + den)), gc);
            return status;
        }
    }
    const Iso<Z> cs = load_iso<Z>(cur + 7);
    const Iso<Z> rsinv = load_iso<Z>(rep + 7);
    const Iso<Z> m = iso_mul(cs, o.s);
    o.denom = pz * Z(__ldg(cur + 6)) * Z(__ldg(rep + 6));
    if (!want_comp)
    {
        // Spinor::norm's first case needs only the trace of cur.s * s * rep.sinv
        const Z tr = m.a11 * rsinv.a11 + m.a12 * rsinv.a21 + m.a13 * rsinv.a31 +
                     m.a21 * rsinv.a12 + m.a22 * rsinv.a22 + m.a23 * rsinv.a32 +
                     m.a31 * rsinv.a13 + m.a32 * rsinv.a23 + m.a33 * rsinv.a33;
        if (tr != -o.denom)                                  // Spinor.h:21-25
        {
            o.spin = prime_parities(tr + o.denom, gc);
            return status;
        }
    }
    o.comp = iso_mul(m, rsinv);
    o.composed = true;
    if (r == n)
    {
        o.spin = spinor_norm(o.q, o.s, pz, gc);
    }
    else
    {
        const Form<Z> mother = load_form<Z>(gc.rep);
        o.spin = spinor_norm(mother, o.comp, o.denom, gc);
    }
    return status;
}

}  // namespace tb
