// tb_build.cuh -- genus construction (the caller side of the hot path).
//
//   quad_form_for        the mother form of a level (same form as QuadForm<Z>::get_quad_form, QuadForm.cpp:444-777)
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

// ---------------------------------------------------------------------------
// The mother form of a level: QuadForm<Z>::get_quad_form's result (QuadForm.cpp:444-777).
//
// Which form comes out is part of the contract -- it is representative 0, and the order of all
// the others follows from it -- so the construction has to make the reference's choices:
//   1. local data: the form <1, b, c> with b = -(product of a set S of "generators") has to be
//      ramified exactly at the flagged primes (and at the real place).  Generators are -1, 2,
//      the primes of the discriminant and, when those do not suffice, ONE auxiliary prime
//      (3, 5, 7, ... in turn).  Hilbert symbols are multiplicative, so S is a solution of a
//      linear system over GF(2); the reference takes the FIRST solution in its enumeration
//      order (subsets as binary numbers, first generator = top bit) at which the auxiliary
//      prime stays unramified -- candidate_sets() walks the same order;
//   2. c = radical of the discriminant with the primes of S toggled in or out; primes toggled
//      IN (and 2) now divide the discriminant of <1, b, c> to a square too many;
//   3. those squares are stripped: the first vector (in the order (0,0,1), (0,1,z), (1,y,z))
//      isotropic modulo p^2 becomes a basis vector and p^2 is divided out of its row;
//   4. reduce, scale up to the prescribed prime powers, reduce again (on the GPU:
//      reduce_form_kernel), and check the local data of the result.
// Written against that description; integers are __int128 (levels up to ~2^60).
// ---------------------------------------------------------------------------
struct LocalData {
    std::vector<int64_t> places;     // finite places, in order: 2 first when the level is odd, then the level's primes
    std::vector<char> ramified;      // per finite place
    big radical = 1, disc = 1;

    // bit pattern of the symbols (x, y)_v over (real place, places...), real place = top bit; 1 = symbol is -1
    u64 pattern(big x, big y) const
    {
        u64 bits = (x < 0 && y < 0) ? 1u : 0u;
        for (int64_t p : places) bits = (bits << 1) | (hilbert_symbol(x, y, p) == -1 ? 1u : 0u);
        return bits;
    }
    u64 wanted() const
    {
        u64 bits = 1;                // definite: ramified at the real place
        for (char r : ramified) bits = (bits << 1) | (r ? 1u : 0u);
        return bits;
    }
};

// Subsets of `rows` (as masks, rows[0] = top bit) whose XOR is `want`, in ascending mask order.
template <class Visit>
static bool candidate_sets(const std::vector<u64> &rows, u64 want, Visit visit)
{
    const size_t n = rows.size();
    for (u64 pick = 1; pick < (1ull << n); ++pick)
    {
        u64 x = 0;
        for (size_t j = 0; j < n; ++j)
            if (pick & (1ull << (n - 1 - j))) x ^= rows[j];
        if (x == want && visit(pick)) return true;
    }
    return false;
}

// Q(v) and the polar form B(v, e_j) = Q(v + e_j) - Q(v) - Q(e_j) of a ternary form
static big polar_with_basis(const HostForm &q, const big v[3], int j)
{
    const big diag[3] = {q.a, q.b, q.c};
    const big cross[3][3] = {{0, q.h, q.g}, {q.h, 0, q.f}, {q.g, q.f, 0}};
    big out = 2 * diag[j] * v[j];
    for (int i = 0; i < 3; ++i) out += cross[j][i] * v[i];
    return out;
}

// One p^2 out of the discriminant: the first vector isotropic modulo p^2 replaces the basis vector at its
// leading 1, whose row of the Gram matrix is then divisible by p (and its diagonal entry by p^2).
static bool strip_square(HostForm &q, big p)
{
    const big pp = p * p;
    big v[3] = {0, 0, 1};
    int lead = 2;
    bool found = q.c % pp == 0;
    for (big z = 0; z < pp && !found; ++z)
        if (q.eval(0, 1, z) % pp == 0) { v[0] = 0; v[1] = 1; v[2] = z; lead = 1; found = true; }
    for (big y = 0; y < pp && !found; ++y)
        for (big z = 0; z < pp && !found; ++z)
            if (q.eval(1, y, z) % pp == 0) { v[0] = 1; v[1] = y; v[2] = z; lead = 0; found = true; }
    if (!found) return false;
    const big norm = q.eval(v[0], v[1], v[2]);
    big with[3];
    for (int j = 0; j < 3; ++j) with[j] = polar_with_basis(q, v, j);
    // new basis: e_lead := v; only row `lead` of the Gram matrix changes
    if (lead == 2) { q.c = norm / pp; q.f = with[1] / p; q.g = with[0] / p; }
    else if (lead == 1) { q.b = norm / pp; q.f = with[2] / p; q.h = with[0] / p; }
    else { q.a = norm / pp; q.g = with[2] / p; q.h = with[1] / p; }
    return true;
}

static void quad_form_for(const std::vector<PrimeSym> &input, cudaStream_t st, int64_t out[6])
{
    // -- 0. what is asked for (the two argument errors are the reference's, QuadForm.cpp:456,478) --------
    LocalData L;
    int flagged = 0;
    bool has_two = false;
    for (const PrimeSym &s : input)
    {
        if (s.power % 2 == 0) throw std::invalid_argument("Prime powers must be odd.");
        has_two |= s.p == 2;
        flagged += s.ramified ? 1 : 0;
        L.radical *= s.p;
        for (int k = 0; k < s.power; ++k) L.disc *= s.p;
    }
    if (flagged % 2 == 0) throw std::invalid_argument("Must specify an odd number of ramified primes.");
    if (!has_two) { L.places.push_back(2); L.ramified.push_back(0); }
    for (const PrimeSym &s : input) { L.places.push_back(s.p); L.ramified.push_back(s.ramified ? 1 : 0); }

    // -- 1. the set S ---------------------------------------------------------------------------------------
    std::vector<big> gens;
    gens.push_back(-1);
    for (int64_t p : L.places) gens.push_back(p);
    std::vector<u64> rows;
    for (const big &x : gens) rows.push_back(L.pattern(x, -L.radical));
    const u64 want = L.wanted();
    const size_t fixed = gens.size();
    u64 chosen = 0;
    int64_t aux = 2;
    for (bool with_aux = false; chosen == 0; with_aux = true)
    {
        if (with_aux)
        {
            do { aux = next_prime64(aux); } while (L.disc % aux == 0);
            if (aux > 100000) throw std::runtime_error("get_quad_form: no auxiliary prime found");
            gens.resize(fixed); rows.resize(fixed);
            gens.push_back(aux);
            rows.push_back(L.pattern(aux, -L.radical));
        }
        candidate_sets(rows, want, [&](u64 pick) {
            big b = -1;
            for (size_t j = 0; j < gens.size(); ++j)
                if (pick & (1ull << (gens.size() - 1 - j))) b *= gens[j];
            if (with_aux && hilbert_symbol(-b, -L.disc, aux) == -1) return false;   // the helper prime must stay unramified
            chosen = pick;
            return true;
        });
    }

    // -- 2. the diagonal form <1, b, c> -----------------------------------------------------------------------
    big b = -1, c = L.radical;
    std::vector<big> squares;                 // primes whose square has to come out of the discriminant again
    squares.push_back(2);
    for (size_t j = 0; j < gens.size(); ++j)
    {
        if (!(chosen & (1ull << (gens.size() - 1 - j)))) continue;
        b *= gens[j];
        if (gens[j] < 0) continue;
        if (c % gens[j] == 0) c /= gens[j];
        else { c *= gens[j]; squares.push_back(gens[j]); }
    }
    HostForm q{1, b, c, 0, 0, 0};

    // -- 3. strip the extra squares -----------------------------------------------------------------------------
    for (const big &p : squares)
        while (q.disc() % (p * p) == 0 && strip_square(q, p)) {}

    // -- 4. reduce, scale to the prescribed powers, reduce ---------------------------------------------------------
    if (reduce_on_device(q, st)) throw std::runtime_error("get_quad_form: intermediate form out of range");
    for (const PrimeSym &s : input)
        for (int j = 3; j <= s.power; j += 2) { q.f *= s.p; q.g *= s.p; q.c *= (big)s.p * s.p; }
    if (reduce_on_device(q, st)) throw std::runtime_error("get_quad_form: intermediate form out of range");
    if (L.pattern(-(4 * q.a * q.b - q.h * q.h), -L.disc) != want || q.disc() != L.disc)
        throw std::runtime_error("get_quad_form: the constructed form does not have the requested local data");
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
