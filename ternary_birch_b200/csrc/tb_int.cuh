// tb_int.cuh -- integer types of the p-neighbor kernels (sm_100a).
//
// Two arithmetic modes, selected by template parameter:
//   W64   int64 with wrap-around ring operations: the reference's Genus<Z64>
//         (pyx precise=False).  Signed overflow is UB in C++, wrap in practice
//         on the reference's x86 build; here every ring op goes through
//         unsigned arithmetic so the wrap is defined and identical.
//   C128  overflow-CHECKED two-limb signed integer replacing GMP (the
//         reference's Genus<Z>, pyx precise=True).  Any operation whose true
//         result leaves [-2^127, 2^127) raises the CTA's sticky overflow flag,
//         which the host turns into the reference's std::overflow_error.
//
// Division by the launch-invariant divisors p and p^2 (NeighborManager.h:279,
// 309-310, 338-341) and by the level's prime divisors (Spinor.h:58-62) uses the
// Moeller-Granlund 2-by-1 reciprocal step, exact for every input.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace tb {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

// ---------------------------------------------------------------------------
// Sticky per-CTA overflow flag of the checked mode.  Zeroed by the kernel
// prologue (flag_reset + __syncthreads), folded into the launch's error words
// by the epilogue.
// ---------------------------------------------------------------------------
__device__ __forceinline__ u32 *cta_flag_ptr()
{
    __shared__ u32 flag;
    return &flag;
}
__device__ __forceinline__ void raise_overflow() { *cta_flag_ptr() = 1u; }

// ---------------------------------------------------------------------------
// Invariant unsigned division  (Moeller & Granlund, "Improved division by
// invariant integers", Alg. 4).  dn = d << s is normalised, v = floor((2^128-1)
// / dn) - 2^64.
// ---------------------------------------------------------------------------
struct InvDiv {
    u64 d;
    u64 dn;
    u64 v;
    int s;
    int pad;
};

inline InvDiv make_invdiv(u64 d)
{
    InvDiv iv;
    iv.d = d;
    iv.s = __builtin_clzll(d);
    iv.dn = d << iv.s;
    unsigned __int128 all = ~(unsigned __int128)0;
    iv.v = (u64)((all / iv.dn) - ((unsigned __int128)1 << 64));
    iv.pad = 0;
    return iv;
}

// (u1:u0) / dn with u1 < dn  ->  q, remainder in r.
__device__ __forceinline__ u64 div2by1(u64 u1, u64 u0, u64 dn, u64 v, u64 &r)
{
    u64 q0 = v * u1;
    u64 q1 = __umul64hi(v, u1);
    q0 += u0;
    q1 += u1 + (q0 < u0 ? 1ull : 0ull);
    q1 += 1;
    u64 rr = u0 - q1 * dn;
    if (rr > q0) { q1 -= 1; rr += dn; }
    if (rr >= dn) { q1 += 1; rr -= dn; }
    r = rr;
    return q1;
}

// x / d and x % d for a 64-bit unsigned x.
__device__ __forceinline__ u64 udivmod(u64 x, const InvDiv &iv, u64 &rem)
{
    u64 u1 = iv.s ? (x >> (64 - iv.s)) : 0ull;
    u64 u0 = x << iv.s;
    u64 r;
    u64 q = div2by1(u1, u0, iv.dn, iv.v, r);
    rem = r >> iv.s;
    return q;
}

// (hi:lo) / d for a 128-bit unsigned value below 2^127; quotient in (qh:ql).
__device__ __forceinline__ void udivmod128(u64 hi, u64 lo, const InvDiv &iv, u64 &qh, u64 &ql, u64 &rem)
{
    u64 u2 = iv.s ? (hi >> (64 - iv.s)) : 0ull;
    u64 u1 = iv.s ? ((hi << iv.s) | (lo >> (64 - iv.s))) : hi;
    u64 u0 = lo << iv.s;
    u64 r;
    if (u2 >= iv.dn) u2 -= iv.dn;   // cannot happen for inputs < 2^127 (u2 < 2^s <= dn); keeps the step well defined
    qh = div2by1(u2, u1, iv.dn, iv.v, r);
    ql = div2by1(r, u0, iv.dn, iv.v, r);
    rem = r >> iv.s;
}

// ---------------------------------------------------------------------------
// I64T<TAG>: int64 with wrap-around ring operations.  W64 = I64T<0> is the reference's
// Genus<Z64>.  X64 = I64T<1> is the same machine arithmetic used for precise = 1 launches
// whose every intermediate the host has PROVEN to stay inside int64 (precise_bounds_ok64),
// except the one term of build_neighbor that is carried in 128 bits (wide_muladd_div) and
// the rare spinor-norm branches, which widen to U128.
// ---------------------------------------------------------------------------
template <int TAG>
struct I64T {
    typedef I64T W64;
    i64 v;

    __device__ __forceinline__ I64T() {}
    __device__ __forceinline__ I64T(i64 x) : v(x) {}
    __device__ __forceinline__ static W64 from_i64(i64 x) { return W64(x); }

    __device__ __forceinline__ i64 low64() const { return v; }
    __device__ __forceinline__ bool fits_i64() const { return true; }
    __device__ __forceinline__ bool is_neg() const { return v < 0; }
    __device__ __forceinline__ bool is_zero() const { return v == 0; }

    friend __device__ __forceinline__ W64 operator+(W64 a, W64 b) { return W64((i64)((u64)a.v + (u64)b.v)); }
    friend __device__ __forceinline__ W64 operator-(W64 a, W64 b) { return W64((i64)((u64)a.v - (u64)b.v)); }
    friend __device__ __forceinline__ W64 operator*(W64 a, W64 b) { return W64((i64)((u64)a.v * (u64)b.v)); }
    friend __device__ __forceinline__ W64 operator-(W64 a) { return W64((i64)(0ull - (u64)a.v)); }
    __device__ __forceinline__ W64 &operator+=(W64 b) { *this = *this + b; return *this; }
    __device__ __forceinline__ W64 &operator-=(W64 b) { *this = *this - b; return *this; }
    __device__ __forceinline__ W64 &operator*=(W64 b) { *this = *this * b; return *this; }

    friend __device__ __forceinline__ bool operator==(W64 a, W64 b) { return a.v == b.v; }
    friend __device__ __forceinline__ bool operator!=(W64 a, W64 b) { return a.v != b.v; }
    friend __device__ __forceinline__ bool operator<(W64 a, W64 b) { return a.v < b.v; }
    friend __device__ __forceinline__ bool operator<=(W64 a, W64 b) { return a.v <= b.v; }
    friend __device__ __forceinline__ bool operator>(W64 a, W64 b) { return a.v > b.v; }
    friend __device__ __forceinline__ bool operator>=(W64 a, W64 b) { return a.v >= b.v; }

    // |x| with the reference's std::abs wrap at INT64_MIN
    __device__ __forceinline__ W64 abs() const { return v < 0 ? -(*this) : *this; }
    // magnitude as unsigned (INT64_MIN -> 2^63)
    __device__ __forceinline__ u64 umag() const { return v < 0 ? 0ull - (u64)v : (u64)v; }

    // times 2 / times 4 (ring ops)
    __device__ __forceinline__ W64 twice() const { return W64((i64)((u64)v << 1)); }
};
typedef I64T<0> W64;
typedef I64T<1> X64;

// C truncating x % d and x / d by an invariant positive divisor.
template <int TAG>
__device__ __forceinline__ I64T<TAG> trunc_rem(I64T<TAG> x, const InvDiv &iv)
{
    u64 r;
    udivmod(x.umag(), iv, r);
    return x.v < 0 ? I64T<TAG>(-(i64)r) : I64T<TAG>((i64)r);
}
template <int TAG>
__device__ __forceinline__ I64T<TAG> trunc_div(I64T<TAG> x, const InvDiv &iv)
{
    u64 r;
    u64 q = udivmod(x.umag(), iv, r);
    return x.v < 0 ? I64T<TAG>((i64)(0ull - q)) : I64T<TAG>((i64)q);
}
// x is a multiple of d?  If so replace x by x / d.
template <int TAG>
__device__ __forceinline__ bool strip_factor(I64T<TAG> &x, const InvDiv &iv)
{
    u64 r;
    u64 q = udivmod(x.umag(), iv, r);
    if (r != 0) return false;
    x = x.v < 0 ? I64T<TAG>((i64)(0ull - q)) : I64T<TAG>((i64)q);
    return true;
}

// birch_util::dumb_div (birch_util.h:102-133) as used by QuadForm::reduce:
// a compare ladder for quotients 0..7, else a truncating division.  For the
// inputs reduce produces on a valid form (0 <= num, 0 < den) this is
// floor(num/den); the general branch keeps the reference's exact behaviour on
// wrapped garbage, except that den == 0 (a SIGFPE in the reference) raises the
// overflow flag.
template <int TAG>
__device__ __forceinline__ I64T<TAG> ref_quotient(I64T<TAG> num, I64T<TAG> den)
{
    const i64 n = num.v, d = den.v;
    if (n >= 0 && d > 0 && n < (1ll << 50))
    {
        // |n/d - fl(n)/fl(d)| < 1/2 here, so the truncated estimate is off by at most one.
        i64 q = __double2ll_rz(__ddiv_rn((double)n, (double)d));
        i64 r = n - q * d;
        if (r < 0) { q -= 1; r += d; }
        if (r >= d) { q += 1; }
        return I64T<TAG>(q);
    }
    if (n < d) return I64T<TAG>(0);
    I64T<TAG> D(d);
    if (num < I64T<TAG>(5) * D)
    {
        if (num < I64T<TAG>(3) * D) return (num < I64T<TAG>(2) * D) ? I64T<TAG>(1) : I64T<TAG>(2);
        return (num < I64T<TAG>(4) * D) ? I64T<TAG>(3) : I64T<TAG>(4);
    }
    if (num < I64T<TAG>(7) * D) return (num < I64T<TAG>(6) * D) ? I64T<TAG>(5) : I64T<TAG>(6);
    if (num < I64T<TAG>(8) * D) return I64T<TAG>(7);
    if (d == 0) { raise_overflow(); return I64T<TAG>(0); }
    if (d == -1) return -num;          // INT64_MIN / -1 traps on x86; wrap instead
    return I64T<TAG>(n / d);
}

// ---------------------------------------------------------------------------
// I128<CHK>: two-limb signed 128-bit.  C128 = I128<true> checks every operation for
// overflow (sticky CTA flag); U128 = I128<false> is the same arithmetic without the
// checks, used when the host has PROVEN from the launch's bounds (max coefficient, p,
// isometry sizes) that no intermediate can leave 127 bits -- see precise_bounds_ok().
// ---------------------------------------------------------------------------
template <bool CHK>
struct I128 {
    typedef I128<CHK> C128;   // local alias: the body below is written in terms of C128
    static __device__ __forceinline__ void ovf() { if (CHK) ovf(); }
    u64 lo;
    i64 hi;

    __device__ __forceinline__ I128() {}
    __device__ __forceinline__ I128(i64 x) : lo((u64)x), hi(x < 0 ? -1ll : 0ll) {}
    __device__ __forceinline__ I128(u64 l, i64 h) : lo(l), hi(h) {}
    __device__ __forceinline__ static C128 from_i64(i64 x) { return C128(x); }

    __device__ __forceinline__ i64 low64() const { return (i64)lo; }
    __device__ __forceinline__ bool fits_i64() const { return hi == ((i64)lo >> 63); }
    __device__ __forceinline__ bool is_neg() const { return hi < 0; }
    __device__ __forceinline__ bool is_zero() const { return (lo | (u64)hi) == 0; }

    friend __device__ __forceinline__ C128 operator+(C128 a, C128 b)
    {
        u64 l = a.lo + b.lo;
        u64 c = l < a.lo ? 1ull : 0ull;
        i64 h = (i64)((u64)a.hi + (u64)b.hi + c);
        if (((a.hi ^ h) & (b.hi ^ h)) < 0) ovf();
        return C128(l, h);
    }
    friend __device__ __forceinline__ C128 operator-(C128 a)
    {
        u64 l = 0ull - a.lo;
        i64 h = (i64)(0ull - (u64)a.hi - (a.lo != 0 ? 1ull : 0ull));
        if (a.hi < 0 && h < 0 && a.lo == 0) ovf();   // -(-2^127)
        return C128(l, h);
    }
    friend __device__ __forceinline__ C128 operator-(C128 a, C128 b)
    {
        u64 l = a.lo - b.lo;
        u64 br = a.lo < b.lo ? 1ull : 0ull;
        i64 h = (i64)((u64)a.hi - (u64)b.hi - br);
        if (((a.hi ^ b.hi) & (a.hi ^ h)) < 0) ovf();
        return C128(l, h);
    }
    friend __device__ __forceinline__ C128 operator*(C128 a, C128 b)
    {
        if (!CHK)
        {
            // low 128 bits of the product (exact: the host proved it fits)
            const u64 l = a.lo * b.lo;
            const u64 h = __umul64hi(a.lo, b.lo) + a.lo * (u64)b.hi + (u64)a.hi * b.lo;
            return C128(l, (i64)h);
        }
        if (a.fits_i64() && b.fits_i64())
        {
            // |a|,|b| <= 2^63  =>  |ab| <= 2^126: always representable
            i64 x = (i64)a.lo, y = (i64)b.lo;
            return C128((u64)x * (u64)y, __mul64hi(x, y));
        }
        return mul_slow(a, b);
    }
    __device__ __forceinline__ C128 &operator+=(C128 b) { *this = *this + b; return *this; }
    __device__ __forceinline__ C128 &operator-=(C128 b) { *this = *this - b; return *this; }
    __device__ __forceinline__ C128 &operator*=(C128 b) { *this = *this * b; return *this; }

    friend __device__ __forceinline__ bool operator==(C128 a, C128 b) { return a.lo == b.lo && a.hi == b.hi; }
    friend __device__ __forceinline__ bool operator!=(C128 a, C128 b) { return !(a == b); }
    friend __device__ __forceinline__ bool operator<(C128 a, C128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
    friend __device__ __forceinline__ bool operator>(C128 a, C128 b) { return b < a; }
    friend __device__ __forceinline__ bool operator<=(C128 a, C128 b) { return !(b < a); }
    friend __device__ __forceinline__ bool operator>=(C128 a, C128 b) { return !(a < b); }

    __device__ __forceinline__ C128 abs() const { return hi < 0 ? -(*this) : *this; }
    __device__ __forceinline__ void umag(u64 &mh, u64 &ml) const
    {
        if (hi < 0) { ml = 0ull - lo; mh = 0ull - (u64)hi - (lo != 0 ? 1ull : 0ull); }
        else { ml = lo; mh = (u64)hi; }
    }
    __device__ __forceinline__ static C128 from_mag(u64 mh, u64 ml, bool neg)
    {
        if ((i64)mh < 0) { ovf(); }
        C128 r(ml, (i64)mh);
        return neg ? -r : r;
    }
    __device__ __forceinline__ C128 twice() const { return *this + *this; }

    static __device__ __noinline__ C128 mul_slow(C128 a, C128 b)
    {
        u64 ah, al, bh, bl;
        a.umag(ah, al);
        b.umag(bh, bl);
        bool neg = (a.hi < 0) != (b.hi < 0);
        if (ah != 0 && bh != 0) { ovf(); return C128(0); }
        u64 cross_hi = ah ? __umul64hi(ah, bl) : __umul64hi(al, bh);
        u64 cross = ah ? ah * bl : al * bh;
        u64 l = al * bl;
        u64 h = __umul64hi(al, bl);
        u64 top = h + cross;
        if (cross_hi != 0 || top < h || (i64)top < 0) { ovf(); return C128(0); }
        C128 r(l, (i64)top);
        return neg ? -r : r;
    }
};

template <bool CHK>
__device__ __forceinline__ I128<CHK> trunc_rem(I128<CHK> x, const InvDiv &iv)
{
    u64 mh, ml, qh, ql, r;
    x.umag(mh, ml);
    if (mh == 0) udivmod(ml, iv, r);
    else udivmod128(mh, ml, iv, qh, ql, r);
    I128<CHK> out((i64)r);                      // r < d < 2^63
    return x.hi < 0 ? -out : out;
}
template <bool CHK>
__device__ __forceinline__ I128<CHK> trunc_div(I128<CHK> x, const InvDiv &iv)
{
    u64 mh, ml, qh, ql, r;
    x.umag(mh, ml);
    if (mh == 0) { qh = 0; ql = udivmod(ml, iv, r); }
    else udivmod128(mh, ml, iv, qh, ql, r);
    return I128<CHK>::from_mag(qh, ql, x.hi < 0);
}
template <bool CHK>
__device__ __forceinline__ bool strip_factor(I128<CHK> &x, const InvDiv &iv)
{
    u64 mh, ml, qh, ql, r;
    x.umag(mh, ml);
    if (mh == 0) { qh = 0; ql = udivmod(ml, iv, r); }
    else udivmod128(mh, ml, iv, qh, ql, r);
    if (r != 0) return false;
    x = I128<CHK>::from_mag(qh, ql, x.hi < 0);
    return true;
}

// floor(num/den) for the non-negative numerator / positive denominator that
// reduce forms on a positive definite ternary form (QuadForm.cpp:104-117 uses
// mpz_fdiv_q).  Wider operands take a shift-subtract division.
template <bool CHK>
__device__ __noinline__ I128<CHK> quotient_slow(I128<CHK> num, I128<CHK> den)
{
    // binary long division on magnitudes (num >= 0, den > 0 established by the caller)
    u64 nh = (u64)num.hi, nl = num.lo, dh = (u64)den.hi, dl = den.lo;
    u64 qh = 0, ql = 0, rh = 0, rl = 0;
    for (int bit = 127; bit >= 0; --bit)
    {
        rh = (rh << 1) | (rl >> 63);
        rl = (rl << 1) | ((bit >= 64 ? (nh >> (bit - 64)) : (nl >> bit)) & 1ull);
        bool ge = rh > dh || (rh == dh && rl >= dl);
        if (ge)
        {
            u64 t = rl - dl;
            rh = rh - dh - (rl < dl ? 1ull : 0ull);
            rl = t;
            if (bit >= 64) qh |= 1ull << (bit - 64); else ql |= 1ull << bit;
        }
    }
    return I128<CHK>(ql, (i64)qh);
}

template <bool CHK>
__device__ __forceinline__ I128<CHK> ref_quotient(I128<CHK> num, I128<CHK> den)
{
    if (num.hi == 0 && den.hi == 0 && (i64)num.lo >= 0 && (i64)den.lo > 0 && num.lo < (1ull << 50))
    {
        i64 n = (i64)num.lo, d = (i64)den.lo;
        i64 q = __double2ll_rz(__ddiv_rn((double)n, (double)d));
        i64 r = n - q * d;
        if (r < 0) { q -= 1; r += d; }
        if (r >= d) { q += 1; }
        return I128<CHK>(q);
    }
    if (num.is_neg() || den.is_neg() || den.is_zero())
    {
        // not reachable with exact arithmetic on a positive definite form
        I128<CHK>::ovf();
        return I128<CHK>(0);
    }
    if (num.hi == 0 && den.hi == 0) return I128<CHK>(num.lo / den.lo, 0);
    return quotient_slow(num, den);
}

typedef I128<true> C128;
typedef I128<false> U128;

// trunc((acc + x * y) / d): the two terms of build_neighbor (NeighborManager.h:325-341) whose
// numerators are the largest values the build forms.  X64 carries the numerator in 128 bits;
// the quotient is proven to fit int64, so the high limb is below d and one 2-by-1 step suffices.
template <class Z>
__device__ __forceinline__ Z muladd_div(Z acc, Z x, Z y, const InvDiv &iv)
{
    return trunc_div(acc + x * y, iv);
}
template <>
__device__ __forceinline__ X64 muladd_div<X64>(X64 acc, X64 x, X64 y, const InvDiv &iv)
{
    const u64 lo = (u64)x.v * (u64)y.v;
    const i64 hi = __mul64hi(x.v, y.v);
    u64 slo = lo + (u64)acc.v;
    u64 shi = (u64)hi + (u64)(acc.v >> 63) + (slo < lo ? 1ull : 0ull);
    const bool neg = (i64)shi < 0;
    if (neg)
    {
        slo = 0ull - slo;
        shi = ~shi + (slo == 0 ? 1ull : 0ull);
    }
    const u64 u1 = iv.s ? ((shi << iv.s) | (slo >> (64 - iv.s))) : shi;
    u64 r;
    const u64 q = div2by1(u1, slo << iv.s, iv.dn, iv.v, r);
    return X64(neg ? (i64)(0ull - q) : (i64)q);
}

}  // namespace tb
