// birch.h -- facade twin of the reference's src/birch.h (types the binding sees).
//
// The headers in this directory re-create, over the C ABI of libtbirch.so
// (include/tbirch.h), exactly the C++ entities ternary_birch.pyx declares in its
// `cdef extern` blocks (pyx:45-118), so the Cython/Sage binding can be pointed at
// them unchanged.  R is `int64_t` (precise=False, wrap-around int64 on the GPU) or
// an arbitrary-precision integer class such as GMP's `mpz_class` (precise=True,
// overflow-checked 128-bit on the GPU).  No reference code is included or copied.
#ifndef TBIRCH_FACADE_BIRCH_H
#define TBIRCH_FACADE_BIRCH_H

#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/tbirch.h"

typedef uint16_t W16;
typedef uint32_t W32;
typedef uint64_t W64;
typedef int32_t Z32;
typedef int64_t Z64;
#ifdef __GMP_PLUSPLUS__
typedef mpz_class Z;
#endif

template<typename R> struct PrimeSymbol { R p; int power; bool ramified; };   // birch.h:107-112

template<typename R> class QuadForm;
template<typename R> class Isometry;
template<typename R> class Genus;
template<typename R> class Eigenvector;
template<typename R> class EigenvectorManager;
template<typename R, typename S, typename T> class IsometrySequence;

namespace tbirch {

// Integer conversions at the boundary: integers cross the C ABI as int64.
template<typename R, typename Enable = void> struct IntegerTraits;

template<typename R>
struct IntegerTraits<R, typename std::enable_if<std::is_integral<R>::value>::type> {
    static const bool precise = false;                 // Genus<Z64>: wrap-around int64
    static int64_t to_i64(const R& x) { return (int64_t)x; }
    static R from_i64(int64_t x) { return (R)x; }
};

#ifdef __GMP_PLUSPLUS__
template<>
struct IntegerTraits<mpz_class, void> {
    static const bool precise = true;                  // Genus<Z>: checked 128-bit replaces GMP
    static int64_t to_i64(const mpz_class& x)
    {
        if (!x.fits_slong_p()) throw std::overflow_error("Integer does not fit the accelerated path's 64-bit interface.");
        return x.get_si();
    }
    static mpz_class from_i64(int64_t x) { return mpz_class((long)x); }
};
#endif

// tb_status -> the exception type the reference throws, same what() text.
inline void check(int status)
{
    if (status == TB_OK) return;
    const std::string msg = tb_last_error();
    switch (status)
    {
        case TB_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
        case TB_ERR_OVERFLOW: throw std::overflow_error(msg);
        case TB_ERR_DOMAIN: throw std::domain_error(msg);
        default: throw std::runtime_error(msg);
    }
}

// Host copy of a genus table (the state of Genus<R> after construction) plus its
// device twin; shared between Genus<Z> and the Genus<Z64> made from it.
struct Table {
    int64_t N = 0, K = 0, disc = 0;
    uint64_t seed = 0;
    std::vector<int64_t> primes, conductors, dims, q, s, sinv, scalar;
    std::vector<int32_t> lut;
    tb_genus *dev = nullptr;

    Table() = default;
    Table(const Table&) = delete;
    Table& operator=(const Table&) = delete;
    ~Table() { if (dev) tb_genus_destroy(dev); }

    tb_genus_desc desc() const
    {
        tb_genus_desc d;
        d.N = N; d.K = K; d.seed = seed; d.disc = disc;
        d.primes = primes.data(); d.conductors = conductors.data(); d.dims = dims.data();
        d.q = q.data(); d.s = s.data(); d.sinv = sinv.data(); d.scalar = scalar.data(); d.lut = lut.data();
        return d;
    }
    tb_genus *device()
    {
        if (!dev) { tb_genus_desc d = desc(); check(tb_genus_create(&d, &dev)); }
        return dev;
    }
    static std::shared_ptr<Table> from_desc(const tb_genus_desc& d)
    {
        std::shared_ptr<Table> t = std::make_shared<Table>();
        t->N = d.N; t->K = d.K; t->seed = d.seed; t->disc = d.disc;
        const size_t N = (size_t)d.N, C = (size_t)1 << d.K;
        t->primes.assign(d.primes, d.primes + d.K);
        t->conductors.assign(d.conductors, d.conductors + C);
        t->dims.assign(d.dims, d.dims + C);
        t->q.assign(d.q, d.q + 6 * N);
        t->s.assign(d.s, d.s + 9 * N);
        t->sinv.assign(d.sinv, d.sinv + 9 * N);
        t->scalar.assign(d.scalar, d.scalar + N);
        t->lut.assign(d.lut, d.lut + C * N);
        return t;
    }
    // flat int64 stream, layout of ternary_birch_b200/table.py ("TBGENUS1")
    static std::shared_ptr<Table> from_stream(const int64_t *a, size_t len)
    {
        if (len < 5 || a[0] != 0x3153554e45474254LL) throw std::invalid_argument("not a genus table stream");
        std::shared_ptr<Table> t = std::make_shared<Table>();
        t->N = a[1]; t->K = a[2]; t->seed = (uint64_t)a[3]; t->disc = a[4];
        const size_t N = (size_t)t->N, C = (size_t)1 << t->K;
        size_t pos = 5;
        auto take = [&](std::vector<int64_t>& v, size_t n) {
            if (pos + n > len) throw std::invalid_argument("genus table stream truncated");
            v.assign(a + pos, a + pos + n); pos += n;
        };
        std::vector<int64_t> skip, lut64;
        take(t->primes, (size_t)t->K); take(t->conductors, C); take(t->dims, C);
        take(t->q, 6 * N); take(t->s, 9 * N); take(t->sinv, 9 * N); take(t->scalar, N);
        take(skip, N); take(skip, N); take(skip, N);          // parent, p, num_auts: construction-side data
        take(lut64, C * N);
        t->lut.assign(lut64.begin(), lut64.end());
        return t;
    }
};

}  // namespace tbirch

#endif  // TBIRCH_FACADE_BIRCH_H
