// oracle/shim/gmpxx.h -- TEST INFRASTRUCTURE ONLY.
//
// A plain value-semantics `mpz_class` over the GMP C ABI (see shim/gmp.h), with
// exactly the operator surface the unmodified reference sources use.  It exists
// because this image has libgmp.so.10 but neither gmpxx.h nor libgmpxx.  The
// semantics follow the GMP C++ interface: `/` and `%` truncate toward zero
// (mpz_tdiv_q / mpz_tdiv_r), `>>` is an arithmetic (floor) shift.
#ifndef TB_ORACLE_SHIM_GMPXX_H
#define TB_ORACLE_SHIM_GMPXX_H
#ifndef __GMP_PLUSPLUS__
#define __GMP_PLUSPLUS__ 1          /* what the real gmpxx.h defines; the facade keys IntegerTraits<mpz_class> on it */
#endif

#include <gmp.h>
#include <cstdlib>
#include <iostream>
#include <string>
#include <type_traits>
#include <utility>

class mpz_class
{
public:
    mpz_class() { mpz_init(v); }
    mpz_class(const mpz_class& o) { mpz_init_set(v, o.v); }
    mpz_class(mpz_class&& o) noexcept { mpz_init(v); mpz_swap(v, o.v); }
    mpz_class(mpz_srcptr z) { mpz_init_set(v, z); }

    template<typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    mpz_class(T x)
    {
        if (std::is_signed<T>::value) mpz_init_set_si(v, (long)x);
        else mpz_init_set_ui(v, (unsigned long)x);
    }

    ~mpz_class() { mpz_clear(v); }

    mpz_class& operator=(const mpz_class& o) { mpz_set(v, o.v); return *this; }
    mpz_class& operator=(mpz_class&& o) noexcept { mpz_swap(v, o.v); return *this; }

    mpz_ptr get_mpz_t() { return v; }
    mpz_srcptr get_mpz_t() const { return v; }
    long get_si() const { return mpz_get_si(v); }
    bool fits_slong_p() const { return mpz_fits_slong_p(v) != 0; }
    unsigned long get_ui() const { return mpz_get_ui(v); }
    double get_d() const { return mpz_get_d(v); }
    std::string get_str(int base = 10) const
    {
        char *s = mpz_get_str(nullptr, base, v);
        std::string out(s);
        std::free(s);
        return out;
    }

    friend mpz_class operator+(const mpz_class& a, const mpz_class& b) { mpz_class r; mpz_add(r.v, a.v, b.v); return r; }
    friend mpz_class operator-(const mpz_class& a, const mpz_class& b) { mpz_class r; mpz_sub(r.v, a.v, b.v); return r; }
    friend mpz_class operator*(const mpz_class& a, const mpz_class& b) { mpz_class r; mpz_mul(r.v, a.v, b.v); return r; }
    friend mpz_class operator/(const mpz_class& a, const mpz_class& b) { mpz_class r; mpz_tdiv_q(r.v, a.v, b.v); return r; }
    friend mpz_class operator%(const mpz_class& a, const mpz_class& b) { mpz_class r; mpz_tdiv_r(r.v, a.v, b.v); return r; }
    friend mpz_class operator-(const mpz_class& a) { mpz_class r; mpz_neg(r.v, a.v); return r; }
    friend mpz_class operator+(const mpz_class& a) { return a; }
    friend mpz_class operator<<(const mpz_class& a, unsigned long n) { mpz_class r; mpz_mul_2exp(r.v, a.v, n); return r; }
    friend mpz_class operator>>(const mpz_class& a, unsigned long n) { mpz_class r; mpz_fdiv_q_2exp(r.v, a.v, n); return r; }

    mpz_class& operator+=(const mpz_class& b) { mpz_add(v, v, b.v); return *this; }
    mpz_class& operator-=(const mpz_class& b) { mpz_sub(v, v, b.v); return *this; }
    mpz_class& operator*=(const mpz_class& b) { mpz_mul(v, v, b.v); return *this; }
    mpz_class& operator/=(const mpz_class& b) { mpz_tdiv_q(v, v, b.v); return *this; }
    mpz_class& operator%=(const mpz_class& b) { mpz_tdiv_r(v, v, b.v); return *this; }
    mpz_class& operator<<=(unsigned long n) { mpz_mul_2exp(v, v, n); return *this; }
    mpz_class& operator>>=(unsigned long n) { mpz_fdiv_q_2exp(v, v, n); return *this; }

    mpz_class& operator++() { mpz_add_ui(v, v, 1); return *this; }
    mpz_class operator++(int) { mpz_class old(*this); mpz_add_ui(v, v, 1); return old; }
    mpz_class& operator--() { mpz_sub_ui(v, v, 1); return *this; }
    mpz_class operator--(int) { mpz_class old(*this); mpz_sub_ui(v, v, 1); return old; }

    friend bool operator==(const mpz_class& a, const mpz_class& b) { return mpz_cmp(a.v, b.v) == 0; }
    friend bool operator!=(const mpz_class& a, const mpz_class& b) { return mpz_cmp(a.v, b.v) != 0; }
    friend bool operator<(const mpz_class& a, const mpz_class& b) { return mpz_cmp(a.v, b.v) < 0; }
    friend bool operator<=(const mpz_class& a, const mpz_class& b) { return mpz_cmp(a.v, b.v) <= 0; }
    friend bool operator>(const mpz_class& a, const mpz_class& b) { return mpz_cmp(a.v, b.v) > 0; }
    friend bool operator>=(const mpz_class& a, const mpz_class& b) { return mpz_cmp(a.v, b.v) >= 0; }

    friend mpz_class abs(const mpz_class& a) { mpz_class r; mpz_abs(r.v, a.v); return r; }

    friend std::ostream& operator<<(std::ostream& os, const mpz_class& a) { return os << a.get_str(); }

private:
    mpz_t v;
};

#endif // TB_ORACLE_SHIM_GMPXX_H
