// QuadForm.h -- facade twin of the reference's QuadForm<R> as the binding uses it
// (pyx:57-66): coefficient accessors and the static get_quad_form.
#ifndef TBIRCH_FACADE_QUADFORM_H
#define TBIRCH_FACADE_QUADFORM_H

#include "birch.h"

template<typename R>
class QuadForm
{
public:
    QuadForm() = default;
    QuadForm(const R& a, const R& b, const R& c, const R& f, const R& g, const R& h)
        : a_(a), b_(b), c_(c), f_(f), g_(g), h_(h) {}

    const R& a(void) const { return a_; }
    const R& b(void) const { return b_; }
    const R& c(void) const { return c_; }
    const R& f(void) const { return f_; }
    const R& g(void) const { return g_; }
    const R& h(void) const { return h_; }

    R discriminant(void) const        // QuadForm.h:27-32
    {
        return a_ * (4 * b_ * c_ - f_ * f_) - b_ * g_ * g_ + h_ * (f_ * g_ - c_ * h_);
    }
    bool operator==(const QuadForm<R>& q) const
    {
        return a_ == q.a_ && b_ == q.b_ && c_ == q.c_ && f_ == q.f_ && g_ == q.g_ && h_ == q.h_;
    }

    // QuadForm<Z>::get_quad_form (QuadForm.cpp:444-777): the positive definite ternary form of
    // discriminant prod p^power ramified at the flagged primes -- tb_quad_form in the library.
    static QuadForm<R> get_quad_form(const std::vector<PrimeSymbol<R>>& primes)
    {
        typedef tbirch::IntegerTraits<R> Traits;
        std::vector<tb_prime_symbol> symbols;
        for (const PrimeSymbol<R>& s : primes)
        {
            tb_prime_symbol t;
            t.p = Traits::to_i64(s.p); t.power = s.power; t.ramified = s.ramified ? 1 : 0;
            symbols.push_back(t);
        }
        int64_t f[6];
        tbirch::check(tb_quad_form(symbols.data(), (int64_t)symbols.size(), f));
        return QuadForm<R>(Traits::from_i64(f[0]), Traits::from_i64(f[1]), Traits::from_i64(f[2]),
                           Traits::from_i64(f[3]), Traits::from_i64(f[4]), Traits::from_i64(f[5]));
    }

protected:
    R a_, b_, c_, f_, g_, h_;
};

#endif  // TBIRCH_FACADE_QUADFORM_H
