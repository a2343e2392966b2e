// tb_build.cuh -- genus construction (the caller side of the hot path).
//
//   quad_form_for        QuadForm<Z>::get_quad_form             QuadForm.cpp:444-777
//   hilbert_symbol       Math<Z>::hilbert_symbol                Math.cpp:20-57
//   genus_mass_x24       Genus<R>::get_mass                     Genus.h:517-531
//   proper_automorphisms QuadForm<R>::proper_automorphisms      QuadForm.h:289-528 (table: Isometry.h:359-644)
//   enumerate_genus      Genus<R>::Genus                        Genus.h:37-246
//
// The enumeration is the reference's sequential discovery (the order of the
// representatives is part of the contract: it indexes every Hecke matrix), driven
// from the host; the neighbors themselves -- isotropic line, neighbor lattice,
// Eisenstein reduction with isometry -- are computed on the GPU by the same device
// code as the Hecke path, in overflow-checked 128-bit arithmetic, one wave of
// not-yet-expanded representatives per launch.  Host-side big integers are
// __int128 (levels up to ~2^60).
#pragma once

#include "tb_neighbor.cuh"

#include <algorithm>
#include <array>
#include <map>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace tb {

typedef __int128 big;

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// One thread per (wave representative, t): 16 int64 per neighbor = reduced form[6],
// isometry[9], status.  gc.cur points at the wave's forms (record layout of the
// genus table, only q[0..5] is read).
struct EnumLaunch {
    int rows;
    u32 P1;
    u64 total;
    InvDiv dP1;
    const RepPrime *rp;
    i64 *records;
};

__global__ void __launch_bounds__(TB_NBR_THREADS)
enum_kernel(const __grid_constant__ GenusCtx gc, const __grid_constant__ PrimeCtx pc,
            const __grid_constant__ EnumLaunch L)
{
    if (threadIdx.x == 0) *cta_flag_ptr() = 0u;
    __syncthreads();
    const u64 i_raw = (u64)blockIdx.x * TB_NBR_THREADS + threadIdx.x;
    const bool live = i_raw < L.total;
    const u64 i = live ? i_raw : L.total - 1;
    u64 t64;
    const u64 row = udivmod(i, L.dP1, t64);
    const RepPrime rp = L.rp[row];
    const i64 *cur = gc.cur + (size_t)row * TB_REC_WORDS;
    const Form<C128> q = load_form<C128>(cur);
    u32 vx, vy, vz;
    line_for_t(rp, (u32)t64, pc, vx, vy, vz);
    Form<C128> nq;
    Iso<C128> s;
    build_neighbor(q, vx, vy, vz, 0, pc, nq, s);
    const bool ok = reduce_neighbor(nq, s, pc.fp64_reduce != 0, false);
    bool fits = nq.a.fits_i64() && nq.b.fits_i64() && nq.c.fits_i64() && nq.f.fits_i64() && nq.g.fits_i64() &&
                nq.h.fits_i64() && s.a11.fits_i64() && s.a12.fits_i64() && s.a13.fits_i64() && s.a21.fits_i64() &&
                s.a22.fits_i64() && s.a23.fits_i64() && s.a31.fits_i64() && s.a32.fits_i64() && s.a33.fits_i64();
    if (live)
    {
        i64 *out = L.records + i * 16;
        out[0] = nq.a.low64(); out[1] = nq.b.low64(); out[2] = nq.c.low64();
        out[3] = nq.f.low64(); out[4] = nq.g.low64(); out[5] = nq.h.low64();
        out[6] = s.a11.low64(); out[7] = s.a12.low64(); out[8] = s.a13.low64();
        out[9] = s.a21.low64(); out[10] = s.a22.low64(); out[11] = s.a23.low64();
        out[12] = s.a31.low64(); out[13] = s.a32.low64(); out[14] = s.a33.low64();
        out[15] = (ok && fits && !*cta_flag_ptr()) ? 0 : 1;
    }
}

// QuadForm::reduce of a single form (used twice by get_quad_form, QuadForm.cpp:744-755).
__global__ void reduce_form_kernel(i64 *form /* in/out [6] */, i64 *status)
{
    if (threadIdx.x == 0) *cta_flag_ptr() = 0u;
    __syncwarp();
    Form<C128> q = load_form<C128>(form);
    Iso<C128> s;
    s.a11 = C128(1); s.a12 = C128(0); s.a13 = C128(0);
    s.a21 = C128(0); s.a22 = C128(1); s.a23 = C128(0);
    s.a31 = C128(0); s.a32 = C128(0); s.a33 = C128(1);
    const bool ok = reduce_neighbor(q, s, false, false);
    if (threadIdx.x == 0)
    {
        form[0] = q.a.low64(); form[1] = q.b.low64(); form[2] = q.c.low64();
        form[3] = q.f.low64(); form[4] = q.g.low64(); form[5] = q.h.low64();
        const bool fits = q.a.fits_i64() && q.b.fits_i64() && q.c.fits_i64() && q.f.fits_i64() && q.g.fits_i64() && q.h.fits_i64();
        *status = (ok && fits && !*cta_flag_ptr()) ? 0 : 1;
    }
}

// ---------------------------------------------------------------------------
// host number theory
// ---------------------------------------------------------------------------
struct PrimeSym {
    int64_t p;
    int power;
    bool ramified;
};

static inline big big_abs(big x) { return x < 0 ? -x : x; }

static int64_t powmod64(int64_t a, int64_t e, int64_t m)
{
    big r = 1, b = ((big)a % m + m) % m;
    while (e) { if (e & 1) r = r * b % m; b = b * b % m; e >>= 1; }
    return (int64_t)r;
}
static int legendre64(big a, int64_t p)
{
    int64_t r = (int64_t)(((a % p) + p) % p);
    if (r == 0) return 0;
    return powmod64(r, (p - 1) / 2, p) == 1 ? 1 : -1;
}
static bool is_prime64(int64_t n)
{
    if (n < 2) return false;
    for (int64_t d = 2; d * d <= n; ++d) if (n % d == 0) return false;
    return true;
}
static int64_t next_prime64(int64_t n) { do { ++n; } while (!is_prime64(n)); return n; }

// Hilbert symbol (a, b)_p by the standard tables on valuations and unit classes (Math.cpp:20-57).
static int hilbert_symbol(big a, big b, int64_t p)
{
    int av = 0, bv = 0;
    while (a % p == 0) { ++av; a /= p; }
    while (b % p == 0) { ++bv; b /= p; }
    if (p == 2)
    {
        // (a,b)_2 = (-1)^{eps(u)eps(v) + alpha w(v) + beta w(u)}, a = 2^alpha u, b = 2^beta v
        auto eps = [](big u) { int r = (int)(((u % 8) + 8) % 8); return (r == 3 || r == 7) ? 1 : 0; };
        auto omg = [](big u) { int r = (int)(((u % 8) + 8) % 8); return (r == 3 || r == 5) ? 1 : 0; };
        int e = eps(a) * eps(b) + (av & 1) * omg(b) + (bv & 1) * omg(a);
        return (e & 1) ? -1 : 1;
    }
    // odd p: (a,b)_p = (-1)^{alpha beta eps(p)} (u/p)^beta (v/p)^alpha
    int sign = 1;
    if ((av & 1) && (bv & 1) && (p % 4 == 3)) sign = -sign;
    if ((bv & 1) && legendre64(a, p) == -1) sign = -sign;
    if ((av & 1) && legendre64(b, p) == -1) sign = -sign;
    return sign;
}

struct HostForm {
    big a, b, c, f, g, h;
    big disc() const { return a * (4 * b * c - f * f) - b * g * g + h * (f * g - c * h); }
    big eval(big x, big y, big z) const { return x * (a * x + g * z + h * y) + y * (b * y + f * z) + z * z * c; }
};

static int reduce_on_device(HostForm &q, cudaStream_t st)
{
    i64 host[7];
    const big v[6] = {q.a, q.b, q.c, q.f, q.g, q.h};
    for (int i = 0; i < 6; ++i)
    {
        if (v[i] != (big)(i64)v[i]) return 1;
        host[i] = (i64)v[i];
    }
    i64 *dev = nullptr;
    if (cudaMalloc((void **)&dev, 7 * sizeof(i64)) != cudaSuccess) return 2;
    cudaMemcpyAsync(dev, host, 6 * sizeof(i64), cudaMemcpyHostToDevice, st);
    reduce_form_kernel<<<1, 32, 0, st>>>(dev, dev + 6);
    cudaMemcpyAsync(host, dev, 7 * sizeof(i64), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(dev);
    if (e != cudaSuccess || host[6] != 0) return 2;
    q.a = host[0]; q.b = host[1]; q.c = host[2]; q.f = host[3]; q.g = host[4]; q.h = host[5];
    return 0;
}

// QuadForm<Z>::get_quad_form (QuadForm.cpp:444-777): a positive definite ternary form of
// the given discriminant, ramified exactly at the flagged primes.  A quaternion-algebra
// style search: pick b = -prod(subset of a factor base) with the prescribed Hilbert symbols
// (a GF(2) system), start from the diagonal form <1, b, det>, strip the squares of auxiliary
// primes from the discriminant with p^2-isotropic vectors, reduce, scale, reduce.
static void quad_form_for(const std::vector<PrimeSym> &input, cudaStream_t st, int64_t out[6])
{
    big det = 1, disc = 1;
    bool two_specified = false;
    int num_ramified = 0;
    for (const PrimeSym &s : input)
    {
        if (s.power % 2 == 0) throw std::invalid_argument("Prime powers must be odd.");
        if (s.p == 2) two_specified = true;
        if (s.ramified) ++num_ramified;
        det *= s.p;
        for (int k = 0; k < s.power; ++k) disc *= s.p;
    }
    if (num_ramified % 2 == 0) throw std::invalid_argument("Must specify an odd number of ramified primes.");

    std::vector<PrimeSym> primes;
    u64 target = 1;                                          // ramified at the infinite place
    if (!two_specified) { primes.push_back(PrimeSym{2, 0, false}); target <<= 1; }
    for (const PrimeSym &s : input) { primes.push_back(s); target <<= 1; if (s.ramified) target |= 1; }

    auto sign_vector = [&](big x) {                          // QuadForm.cpp:381-390
        u64 vec = (x == -1);
        for (const PrimeSym &s : primes) vec = (vec << 1) | (hilbert_symbol(x, -det, s.p) == -1 ? 1u : 0u);
        return vec;
    };
    auto solve = [](const std::vector<u64> &vecs, u64 start, u64 want) {   // GF2_solve_naive, 394-412
        const u64 upper = 1ull << vecs.size();
        for (u64 i = start + 1; i < upper; ++i)
        {
            u64 x = 0, mask = upper >> 1;
            for (size_t j = 0; j < vecs.size(); ++j) { if (i & mask) x ^= vecs[j]; mask >>= 1; }
            if (x == want) return i;
        }
        return (u64)0;
    };

    std::vector<big> fullbase;
    std::vector<u64> signs;
    fullbase.push_back(-1);
    signs.push_back(sign_vector(-1));
    for (const PrimeSym &s : primes) { signs.push_back(sign_vector(s.p)); fullbase.push_back(s.p); }

    int64_t p = 2;
    u64 solution = 0;
    bool done = false, added_to_end = false;
    while (!done)
    {
        solution = 0;
        do
        {
            solution = solve(signs, solution, target);
            if (solution)
            {
                u64 mask = 1ull << fullbase.size();
                big b = -1;
                for (const big &q : fullbase) { mask >>= 1; if (solution & mask) b *= q; }
                mask = 1ull << primes.size();
                for (const PrimeSym &s : primes)
                {
                    mask >>= 1;
                    const int sign = (target & mask) ? -1 : 1;
                    if (hilbert_symbol(-b, -disc, s.p) != sign)
                        throw std::runtime_error("Hilbert symbols do not check out. How did this happen?");
                }
                bool good = true;
                for (size_t n = primes.size() + 1; n < fullbase.size(); ++n)
                    if (hilbert_symbol(-b, -disc, (int64_t)fullbase[n]) == -1) good = false;
                if (good) { done = true; break; }
            }
        } while (solution != 0);
        if (done) break;
        p = next_prime64(p);
        while (disc % p == 0) p = next_prime64(p);
        if (added_to_end) { signs.pop_back(); fullbase.pop_back(); }
        signs.push_back(sign_vector(p));
        fullbase.push_back(p);
        added_to_end = true;
    }

    big b = -1;
    u64 mask = 1ull << fullbase.size();
    std::vector<big> base;
    base.push_back(2);
    for (const big &q : fullbase)
    {
        mask >>= 1;
        if (solution & mask)
        {
            b *= q;
            if (q > 0)
            {
                if (det % q == 0) det /= q;
                else { det *= q; base.push_back(q); }
            }
        }
    }

    HostForm q{1, b, det, 0, 0, 0};
    big N = q.disc();
    for (const big &bp : base)                               // QuadForm.cpp:663-737
    {
        const big pp = bp * bp;
        while (N % pp == 0)
        {
            // Z_isotropic_mod_pp (QuadForm.cpp:417-442): (0,0,1), then (0,1,z), then (1,y,z)
            big vx = 0, vy = 0, vz = 1;
            bool found = q.eval(0, 0, 1) % pp == 0;
            for (big z = 0; z < pp && !found; ++z)
                if (q.eval(0, 1, z) % pp == 0) { vx = 0; vy = 1; vz = z; found = true; }
            for (big y = 0; y < pp && !found; ++y)
                for (big z = 0; z < pp && !found; ++z)
                    if (q.eval(1, y, z) % pp == 0) { vx = 1; vy = y; vz = z; found = true; }
            if (!found) break;
            if (vx == 0 && vy == 0) { q.c /= pp; q.f /= bp; q.g /= bp; }
            else if (vx == 0)
            {
                q.b += q.c * vz * vz + q.f * vz;
                q.f += 2 * q.c * vz;
                q.h += q.g * vz;
                q.b /= pp; q.f /= bp; q.h /= bp;
            }
            else
            {
                q.a += q.b * vy * vy + q.c * vz * vz + q.f * vy * vz + q.g * vz + q.h * vy;
                q.g += 2 * q.c * vz + q.f * vy;
                q.h += 2 * q.b * vy + q.f * vz;
                q.a /= pp; q.g /= bp; q.h /= bp;
            }
            N = q.disc();
        }
    }

    if (reduce_on_device(q, st)) throw std::runtime_error("get_quad_form: intermediate form out of range");
    for (const PrimeSym &s : primes)
        for (int j = 3; j <= s.power; j += 2) { q.f *= s.p; q.g *= s.p; q.c *= (big)s.p * s.p; }
    if (reduce_on_device(q, st)) throw std::runtime_error("get_quad_form: intermediate form out of range");

    const big x = 4 * q.a * q.b - q.h * q.h;                 // final verification, 757-771
    mask = 1ull << primes.size();
    for (const PrimeSym &s : primes)
    {
        mask >>= 1;
        const int sign = (target & mask) ? -1 : 1;
        if (hilbert_symbol(-x, -disc, s.p) != sign)
            throw std::runtime_error("Hilbert symbols do not check out. How did this happen?");
    }
    out[0] = (int64_t)q.a; out[1] = (int64_t)q.b; out[2] = (int64_t)q.c;
    out[3] = (int64_t)q.f; out[4] = (int64_t)q.g; out[5] = (int64_t)q.h;
}

// Genus<R>::get_mass (Genus.h:517-531): 24 * mass of the genus.
static big genus_mass_x24(const int64_t q[6], big disc, const std::vector<PrimeSym> &symbols)
{
    big mass = 2 * disc;
    const big a = (big)q[5] * q[5] - 4 * (big)q[0] * q[1];
    const big b = -(big)q[0] * disc;
    for (const PrimeSym &s : symbols)
    {
        mass *= (s.p + hilbert_symbol(a, b, s.p));
        mass /= 2;
        mass /= s.p;
    }
    return mass;
}

// Proper automorphisms (det +1, identity excluded) of a reduced positive definite form.
// The reference looks them up in a table keyed by boundary conditions
// (QuadForm.h:289-528 / Isometry.h:359-644); here they are found directly: an automorphism
// maps the reduced basis to vectors of the same norms, and for an Eisenstein-reduced form
// those have coordinates in {-1, 0, 1}.
typedef std::array<int, 9> SmallIso;
static std::vector<SmallIso> proper_automorphisms(const int64_t q[6])
{
    const int64_t a = q[0], b = q[1], c = q[2], f = q[3], g = q[4], h = q[5];
    auto norm = [&](const int *v) { return a * v[0] * v[0] + b * v[1] * v[1] + c * v[2] * v[2] + f * v[1] * v[2] + g * v[0] * v[2] + h * v[0] * v[1]; };
    auto inner2 = [&](const int *u, const int *v) {          // 2 B(u, v)
        return 2 * a * u[0] * v[0] + 2 * b * u[1] * v[1] + 2 * c * u[2] * v[2] + f * (u[1] * v[2] + u[2] * v[1]) +
               g * (u[0] * v[2] + u[2] * v[0]) + h * (u[0] * v[1] + u[1] * v[0]);
    };
    std::vector<std::array<int, 3>> va, vb, vc;
    for (int x = -1; x <= 1; ++x) for (int y = -1; y <= 1; ++y) for (int z = -1; z <= 1; ++z)
    {
        const int v[3] = {x, y, z};
        const int64_t n = norm(v);
        if (n == a) va.push_back({x, y, z});
        if (n == b) vb.push_back({x, y, z});
        if (n == c) vc.push_back({x, y, z});
    }
    std::vector<SmallIso> out;
    for (const auto &u : va) for (const auto &v : vb)
    {
        if (inner2(u.data(), v.data()) != h) continue;
        for (const auto &w : vc)
        {
            if (inner2(u.data(), w.data()) != g || inner2(v.data(), w.data()) != f) continue;
            const int det = u[0] * (v[1] * w[2] - v[2] * w[1]) - v[0] * (u[1] * w[2] - u[2] * w[1]) + w[0] * (u[1] * v[2] - u[2] * v[1]);
            if (det != 1) continue;
            const SmallIso m = {u[0], v[0], w[0], u[1], v[1], w[1], u[2], v[2], w[2]};   // columns u, v, w
            if (m == SmallIso{1, 0, 0, 0, 1, 0, 0, 0, 1}) continue;
            out.push_back(m);
        }
    }
    return out;
}

// Spinor::norm with scalar 1 on a small automorphism (Spinor.h:19-66).
static u64 host_spinor_norm(const int64_t q[6], const SmallIso &s, const std::vector<int64_t> &primes)
{
    auto vals = [&](big x) {
        u64 val = 0, mask = 1;
        for (int64_t p : primes) { while (x != 0 && x % p == 0) { x /= p; val ^= mask; } mask <<= 1; }
        return val;
    };
    const u64 twist = (1ull << primes.size()) - 1;
    const big a = q[0], b = q[1], f = q[3], g = q[4], h = q[5];
    const big scalar = 1;
    const big tr = s[0] + s[4] + s[8];
    if (tr != -scalar) return vals(tr + scalar);
    const big delta = 2 * a * (s[4] + s[8]) - (g * s[6] + h * s[3]);
    if (delta != 0) return vals(delta) ^ twist;
    const big abh = 4 * a * b - h * h;
    const big abhm33 = abh * s[8] + s[7] * (g * h - 2 * a * f) + s[6] * (f * h - 2 * b * g);
    if (abhm33 != abh * scalar) return vals(abhm33 - abh * scalar) ^ vals(2 * a) ^ twist;
    return vals(abh * scalar);
}

struct BuiltGenus {
    int64_t N = 0, K = 0, disc = 0;
    uint64_t seed = 0;
    std::vector<int64_t> primes, conductors, dims, q, s, sinv, scalar, parent, p, num_auts;
    std::vector<int32_t> lut;
};

struct FormKey {
    int64_t v[6];
    bool operator==(const FormKey &o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct FormKeyHash {
    size_t operator()(const FormKey &k) const
    {
        u64 fnv = 0x811c9dc5ull;
        for (int i = 0; i < 6; ++i) fnv = (fnv ^ (u64)k.v[i]) * 0x01000193ull;
        return (size_t)fnv;
    }
};

static void mat3_mul(const big *x, const big *y, big *out)
{
    big t[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        t[3 * i + j] = x[3 * i] * y[j] + x[3 * i + 1] * y[3 + j] + x[3 * i + 2] * y[6 + j];
    for (int i = 0; i < 9; ++i) out[i] = t[i];
}

}  // namespace tb
